// tree_jit.h — run-time specialisation of the tree interpreter (tree_jit.cu): internal interface.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "dwdf_kernels.h"

namespace dwdf
{
struct TreeJit; // one specialised circuit: generated source, cubin, per-device modules

bool tree_jit_supported (const TreeProgram& p);
std::string tree_jit_generate (const TreeProgram& p); // the circuit-specific source
std::string tree_jit_full_source (const TreeProgram& p); // + the fixed kernel skeleton (what NVRTC compiles, minus the two headers)
const char* tree_jit_header (int which); // 0: dwdf_math.cuh, 1: dwdf_tma.cuh, as embedded at build time
TreeJit* tree_jit_create (const TreeProgram& p, std::string& err); // generates and compiles (NVRTC, sm_100a); needs no device
bool tree_jit_load (TreeJit* j, std::string& err); // loads the cubin on the current device
void tree_jit_destroy (TreeJit* j);
const std::string& tree_jit_source (const TreeJit* j);
size_t tree_jit_cubin_bytes (const TreeJit* j);
// tmx == nullptr: the direct-global-access twin (any T, any alignment)
bool tree_jit_forward (TreeJit* j, const CUtensorMap* tmx, const CUtensorMap* tmy, const float* params, const float* x, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream, std::string& err);
bool tree_jit_adjoint (TreeJit* j, const CUtensorMap* tmx, const CUtensorMap* tmg, const float* params, const float* x, const float* g, const float* ckpt, bool target, int skip, double* partials, int64_t B, int64_t T, cudaStream_t stream,
                       std::string& err);
} // namespace dwdf
