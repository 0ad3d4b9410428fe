// dwdf_kernels.h — host-callable launchers of the CUDA kernels (internal; the public surface is
// include/dwdf.h). Each launcher enqueues on `stream` and returns the cudaError_t of the launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <type_traits>

#include "dwdf_math.cuh"

namespace dwdf
{

// Forward tiles: [rows x kFwdTileSamples samples]. 32 samples = 128-byte rows (128-byte swizzle), 16 samples = 64-byte
// rows (64-byte swizzle, half the shared memory per stage: more resident one-warp CTAs per SM). Build-time choice
// (-DDWDF_FWD_TILE_T / -DDWDF_FWD_STAGES; measured on the B200).
#ifndef DWDF_FWD_TILE_T
#define DWDF_FWD_TILE_T 32
#endif
#ifndef DWDF_FWD_STAGES
#define DWDF_FWD_STAGES 3
#endif
constexpr int kFwdTileSamples = DWDF_FWD_TILE_T;
constexpr int kSeg = 16; // capacitor-state checkpoint spacing [samples]; also the adjoint's segment
constexpr int kTimeChunk = 256; // samples per chunk of the neural root's time-parallel kernels
constexpr int kMapFloatsPerChunk = 12; // clipper adjoint with time chunks: floats per (sequence, chunk)
constexpr int kMaxResidentCtas = 148 * 32; // upper bound of one-warp CTAs a B200 holds at once: sizes the time-chunk scratch
constexpr int kPartialStride = 8; // doubles per sequence-group in the clipper partials buffer
constexpr int kTreePartialStride = 24; // ... in the tree interpreter partials buffer
// partial sums one warp (32 sequences) hands to the finalize kernel
enum : int
{
    kAccGamma = 0, // sum G dz'/dgamma
    kAccEll = 1, // sum G dz'/d ell,  ell = ln(Rp Is)
    kAccV = 2, // sum G dz'/dV
    kAccSse = 3, // sum (y - target)^2
    kAccSt2 = 4, // sum target^2
    kAccEllRp = 5 // resistance channel: sum G dz'/d ell Rp[n] (kAccGamma then holds sum G dz'/dgamma gamma (1 - gamma))
};

// tuning / A-B switches of the clipper kernels (dwdf_set_option); 0 = shipped behaviour
enum : int
{
    kOptNoFastStep = 1, // forward, approx: per-sample warp vote (clip_step) instead of clip_step_fast chunks
    kOptNoPair = 2, // forward, approx: one sequence per lane instead of the packed-fp32x2 pair kernel
    kOptNoChunks = 8, // never use the time-parallel kernels
    kOptForceChunks = 16, // use them whatever the batch size (crossover measurements)
    kOptWarm13 = 64, kOptWarm8 = 128, // time-chunk warm-up until the off-state decay is 1e-13 / 1e-8 instead of 1e-10 (longer: no repairs; shorter: many)
    kOptNoPdl = 256, // no programmatic dependent launches
    kOptL2Prefetch = 4 // cp.async.bulk.prefetch.tensor L2 run-ahead (measured: slower — forward 0.47 -> 0.58 ms, adjoint 0.63 -> 0.94 ms; off)
};
// Launch `kernel` as a programmatic dependent of whatever precedes it in `stream` (the kernel calls grid_dependency_wait()
// before it touches anything a predecessor wrote): the short kernels of a small-batch training step hide their launch latency
// behind the kernel before them. kOptNoPdl: plain launches (A/B).
template <class... KArgs, class... Args>
cudaError_t launch_dependent (bool pdl, void (*kernel) (KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args&&... args)
{
    cudaLaunchConfig_t cfg {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx (&cfg, kernel, KArgs (args)...);
}
extern std::atomic<int> g_clip_opts; // diagnostic switches (dwdf_set_option); read once per launch
extern std::atomic<int64_t> g_extra_launches; // kernels the launchers add on their own (the time-chunk stitch passes)

struct ClipVariant
{
    int mode; // kModeApprox / kModeExact
    bool general; // N_up / N_down law
    bool pyorder; // probe after tree.incident
};

// TMA descriptors of one launch (valid only when use_tma)
struct ClipTmaMaps
{
    CUtensorMap x, y, g; // forward (x, y): 32 x 32 tiles, 128-byte swizzle; adjoint (x, y, g): 16 x 32 tiles, 64-byte swizzle
    CUtensorMap x2, y2; // paired forward: [64 sequences x 32 samples] tiles
    CUtensorMap r; // per-sample resistance channel (the *_r kernels): same tile shape as x
    bool pair = false; // x2 / y2 are valid and the paired kernel is wanted
    // time chunks (fewer sequences than the SMs hold warps): the launcher proposes up to kcap chunks. Forward scratch zs / ze:
    // kcap_fwd * B floats each; adjoint scratch cmaps: kcap_adj * kMapFloatsPerChunk * B floats. nullptr / 1: no chunks.
    int kcap_fwd = 1, kcap_adj = 1;
    float *zs = nullptr, *ze = nullptr, *cmaps = nullptr;
    int* redone = nullptr; // optional counter of chunks the forward verification had to recompute
};

// per-(root mode, law) launchers, each specialised in its own translation unit (clipper_kernels.cu, 4 parts)
template <int MODE, bool GENERAL>
cudaError_t clipper_forward_part (bool py, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream);
template <int MODE, bool GENERAL>
cudaError_t clipper_adjoint_part (bool py, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* y, const float* ckpt, const float* g, bool target, int skip, float* gx, double* partials, int64_t B, int64_t T, cudaStream_t stream);
template <int MODE, bool GENERAL>
cudaError_t clipper_train_part (bool py, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* target, int skip, float* y, double* partials, int64_t B, int64_t T, cudaStream_t stream);

template <int MODE, bool GENERAL>
cudaError_t clipper_forward_r_part (bool py, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* r, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream);
template <int MODE, bool GENERAL>
cudaError_t clipper_adjoint_r_part (bool py, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* r, const float* y, const float* ckpt, const float* g, bool target, int skip, double* partials, int64_t B, int64_t T, cudaStream_t stream);

// the same circuit with the source resistance as a per-sample input channel r (clipper_pot.py:114-117); partials carry kAccEllRp too
cudaError_t launch_clipper_forward_r (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* r, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream);
cudaError_t launch_clipper_adjoint_r (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* r, const float* y, const float* ckpt, const float* g, bool target, int64_t skip, double* partials, int64_t B, int64_t T, cudaStream_t stream);
cudaError_t launch_clipper_finalize_r (const ClipDesc& desc, const float* params, const double* partials, int64_t n_groups, const double* raw_in, bool raw_only, bool target, int loss_kind, double count, double* out, cudaStream_t stream);

cudaError_t launch_clipper_forward (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream);

// adjoint (reverse sweep over x, the forward output y and g = dL/dy or the target): raw sums per group of 32 sequences into partials[(group0 + group) * kPartialStride + k]
cudaError_t launch_clipper_adjoint (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* y, const float* ckpt, const float* g, bool target, int64_t skip, float* gx, double* partials, int64_t B, int64_t T, cudaStream_t stream);

// fused training pass: forward + loss + parameter sensitivities in one sweep
cudaError_t launch_clipper_train (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* target, int64_t skip, float* y, double* partials, int64_t B, int64_t T, cudaStream_t stream);

// fixed-order reduction of the partials + chain rule to (Is, nabla, R, C) + loss -> out[DWDF_OUT_LEN]
cudaError_t launch_clipper_finalize (const ClipDesc& desc, const float* params, const double* partials, int64_t n_groups, const double* raw_in, bool raw_only, bool target, int loss_kind, double count, double* out, cudaStream_t stream);

// ---- multi-GPU exchange over peer memory (clipper_dispatch.cu explains the protocol) ----------------------------
constexpr int kDpMaxWorld = 16; // ranks of one node
constexpr int kDpSlotDoubles = 2048; // doubles per mailbox slot (16 bytes each on the wire: two tagged 8-byte words)
struct DpPeers // by-value kernel argument
{
    int rank, world;
    unsigned long long timeout_ns;
    char* mailbox[kDpMaxWorld]; // every rank's mailbox as mapped into this process; mailbox[rank] is this rank's own
};
inline size_t dp_mailbox_bytes (int world) { return (size_t) 2 * world * kDpSlotDoubles * 2 * sizeof (unsigned long long) + 256; } // two parities x world slots, then the epoch counter
cudaError_t launch_clipper_finalize_dp (const ClipDesc& desc, float* params, const double* partials, int64_t n_groups, bool target, int loss_kind, double count, double* out, const DpPeers& dp, float* m, float* v, int32_t* step, int n_params,
                                        float lr, const float* lr_vec, float beta1, float beta2, float eps, const float* lo, const float* hi, cudaStream_t stream);
cudaError_t launch_peer_allreduce (double* inout, int n, const DpPeers& dp, cudaStream_t stream);

// ---- the reference's loss as its training loop calls it (loss_kernels.cu): sums over (y, target), then dL/dy -------------
size_t loss_scratch_doubles (int64_t B); // scratch of launch_loss_sums; the sums (sse, sum y^2, sum target^2, count) land in scratch[0 .. 4)
cudaError_t launch_loss_sums (const float* y, const float* t, int64_t B, int64_t T, int skip, double* scratch, cudaStream_t stream);
cudaError_t launch_loss_ybar (const float* y, const float* t, int64_t B, int64_t T, int skip, const double* sums, float* ybar, cudaStream_t stream);
cudaError_t launch_loss_write (const double* sums, double* out, cudaStream_t stream);

cudaError_t launch_adam (float* params, const double* out, float* m, float* v, int32_t* step, int n_params, float lr, const float* lr_vec, float beta1, float beta2, float eps, double grad_scale, const float* lo, const float* hi, cudaStream_t stream);

// ---- neural diode-pair root (inference): clipper tree + b = -MLP(a, ln Rp), hidden width 4 / 8 / 16 ----------
cudaError_t launch_nn_forward (int hidden, int n_hidden, bool pyorder, const float* x, const float* r, float* y, const float* params, int slot_R, int slot_C, float fs, const float* weights,
                               int n_weights, float* state, float* ckpt, int64_t B, int64_t T, int K, float* scratch, int* redone, cudaStream_t stream);
int nn_time_chunks (int64_t T); // chunks of the time-parallel variant (K > 1 needs scratch of 4 * ceil(B/2) * K floats)
// reverse sweep: dL/d(weights) partials per warp (64 sequences): [n_groups][n_weights + 8] doubles (then sse, st2)
int64_t nn_ckpt_floats (int64_t B, int64_t T);
int64_t nn_groups (int64_t B);
cudaError_t launch_nn_adjoint (int hidden, int n_hidden, bool pyorder, bool target, const float* x, const float* r, const float* y, const float* g, const float* ckpt, const float* params, int slot_R, int slot_C,
                               float fs, const float* weights, int n_weights, double* partials, int skip, int64_t B, int64_t T, int K, float* scratch, float* gx, cudaStream_t stream);
int64_t nn_adjoint_ctas (int64_t B, int K);
// partial vectors -> raw sums (grad_w_raw[n_weights]; raw[kAccSse], raw[kAccSt2], raw[23] = count), then raw -> gradients + loss in place
cudaError_t launch_nn_reduce (const double* partials, int64_t n_groups, int n_weights, double count, double* grad_w_raw, double* raw, cudaStream_t stream);
cudaError_t launch_nn_scale (int n_weights, bool target, int loss_kind, double* grad_w_inout, double* raw_inout, cudaStream_t stream);
cudaError_t launch_adam_vec (float* w, const double* gw, float* m, float* v, int32_t* step, int64_t n, float lr, float beta1, float beta2, float eps, double grad_scale, cudaStream_t stream);

// ---- generic tree interpreter -----------------------------------------------------------------
struct TreeProgram // by-value kernel argument (fits the 4 KB parameter space comfortably)
{
    int n_nodes;
    int kind[16], c1[16], c2[16], param[16];
    int root_kind, root_mode, pyorder, probe, probe_current, source, r_node;
    int slot_Is, slot_nabla, n_params, n_iter;
    float fs, Vt, n_up, n_down, tol;
    int n_states;
    int state_of[16]; // reactive leaf -> (first) state index; alpha-transform leaves own two (z, previous reflected wave)
};

cudaError_t launch_tree_forward (const TreeProgram& p, const float* params, const float* x, const float* r, float* y, float* state, int64_t B, int64_t T, cudaStream_t stream);
// reverse-mode gradient of the interpreter (linear trees closed by an ideal voltage source or a
// diode pair): per-sequence tape in `tape` (device scratch), raw dL/dparam partials per group.
cudaError_t launch_tree_adjoint (const TreeProgram& p, const float* params, const float* x, const float* r, const float* g, bool target, int64_t skip, double* partials, float* tape, int64_t B, int64_t T, cudaStream_t stream);
cudaError_t launch_tree_finalize (const TreeProgram& p, const float* params, const double* partials, int64_t n_groups, const double* raw_in, bool raw_only, bool target, int loss_kind, double count, double* out, cudaStream_t stream);

} // namespace dwdf
