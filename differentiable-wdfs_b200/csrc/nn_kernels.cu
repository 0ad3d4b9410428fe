// nn_kernels.cu — the diode clipper closed by the reference's NEURAL diode-pair root (inference).
//
// Circuit: Parallel(ResistiveVoltageSource, Capacitor) + b = -MLP(a, ln Rp)
// (clipper_pot.py:94-127 with DenseRootModel, layers.py:42-82; plugin: DiodePairNeuralModel.h:62-75 with the
// RTNeural model types of :5-41: 2xH = 2 -> H -> H -> H -> 1, 4xH = 2 -> H x5 -> 1, tanh between layers).
// Optional per-sample source resistance channel r (clipper_pot.py:114-117): gamma, Rp and ln Rp are then
// recomputed every sample.
//
// Decomposition: the network dominates (2x16: 560 multiply-adds + 48 tanh per sample against ~10 for the
// tree), so the kernel is arranged around issue slots for the multiply-adds: each lane owns TWO sequences in
// packed fp32x2 registers and every weight (read once per pair from shared memory, 16-byte broadcast loads)
// feeds one FFMA2 with the weight as the broadcast scalar operand. tanh = 1 - 2 / (1 + 2^(2 log2(e) x)):
// MUFU.EX2 + MUFU.RCP, absolute error ~1e-7. At ~1000 issue slots per sample the 8 bytes of traffic are
// irrelevant: rows are read and written directly (each 128-byte line lives in L1 for its 32 samples).
#include "dwdf_kernels.h"

namespace dwdf
{
namespace
{
__device__ __forceinline__ float tanh_acc (float x)
{
    const float e = ex2_ (2.885390081777927f * x); // e^(2x)
    return fma_ (-2.0f, rcp (e + 1.0f), 1.0f);
}
__device__ __forceinline__ f2 tanhv (f2 x) { return f2 { tanh_acc (x.x), tanh_acc (x.y) }; }

// One dense layer IN -> OUT on a pair: out[j] = bias[j] + sum_i w[i][j] h[i]; w is (IN x OUT) row-major in shared memory
template <int IN, int OUT, bool TANH>
__device__ __forceinline__ void dense (const float* __restrict__ w, const float* __restrict__ bias, const f2 (&h)[IN], f2 (&out)[OUT])
{
    static_assert (OUT % 4 == 0 || OUT == 1, "layer widths are multiples of 4");
    if constexpr (OUT == 1)
    {
        f2 acc = bc (f2 {}, bias[0]);
#pragma unroll
        for (int i = 0; i < IN; i += 4)
        {
            const float4 w4 = *reinterpret_cast<const float4*> (w + i);
            acc = fmav (bc (f2 {}, w4.x), h[i], acc);
            acc = fmav (bc (f2 {}, w4.y), h[i + 1], acc);
            acc = fmav (bc (f2 {}, w4.z), h[i + 2], acc);
            acc = fmav (bc (f2 {}, w4.w), h[i + 3], acc);
        }
        out[0] = acc;
    }
    else
    {
#pragma unroll
        for (int j = 0; j < OUT; j += 4)
        {
            const float4 b4 = *reinterpret_cast<const float4*> (bias + j);
            out[j] = bc (f2 {}, b4.x), out[j + 1] = bc (f2 {}, b4.y), out[j + 2] = bc (f2 {}, b4.z), out[j + 3] = bc (f2 {}, b4.w);
        }
#pragma unroll
        for (int i = 0; i < IN; ++i)
#pragma unroll
            for (int j = 0; j < OUT; j += 4)
            {
                const float4 w4 = *reinterpret_cast<const float4*> (w + i * OUT + j);
                out[j] = fmav (bc (f2 {}, w4.x), h[i], out[j]);
                out[j + 1] = fmav (bc (f2 {}, w4.y), h[i], out[j + 1]);
                out[j + 2] = fmav (bc (f2 {}, w4.z), h[i], out[j + 2]);
                out[j + 3] = fmav (bc (f2 {}, w4.w), h[i], out[j + 3]);
            }
        if (TANH)
#pragma unroll
            for (int j = 0; j < OUT; ++j)
                out[j] = tanhv (out[j]);
    }
}

// model(a, ln Rp) for a pair; weights: [2 x H | H] ([H x H | H]) x n_hidden [H | 1]
template <int H>
__device__ __forceinline__ f2 mlp (const float* __restrict__ sw, int n_hidden, f2 a, f2 lr)
{
    f2 in[2] = { a, lr }, h[H], g[H];
    dense<2, H, true> (sw, sw + 2 * H, in, h);
    const float* w = sw + 3 * H;
    for (int l = 0; l < n_hidden; ++l)
    {
        dense<H, H, true> (w, w + H * H, h, g);
#pragma unroll
        for (int j = 0; j < H; ++j)
            h[j] = g[j];
        w += H * H + H;
    }
    f2 o[1];
    dense<H, 1, false> (w, w + H, h, o);
    return o[0];
}

template <int H, bool PY>
__global__ void __launch_bounds__ (128) nn_clipper_forward (const float* __restrict__ x, const float* __restrict__ r, float* __restrict__ y, const float* __restrict__ params, int slot_R, int slot_C, float fs,
                                                           const float* __restrict__ weights, int n_weights, int n_hidden, float* __restrict__ state, int64_t B, int T)
{
    extern __shared__ __align__ (16) float sw[];
    for (int i = threadIdx.x; i < n_weights; i += blockDim.x)
        sw[i] = __ldg (weights + i);
    __syncthreads ();
    const int64_t pair = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t rowA = 2 * pair, rowB = rowA + 1;
    if (rowA >= B)
        return;
    const bool validB = rowB < B;
    const float* xa = x + rowA * T;
    const float* xb = x + (validB ? rowB : rowA) * T;
    const float* ra = r != nullptr ? r + rowA * T : nullptr;
    const float* rb = r != nullptr ? r + (validB ? rowB : rowA) * T : nullptr;
    float* ya = y + rowA * T;
    float* yb = y + rowB * T;
    // tf_wdf.py:114-115 (Capacitor), :168-177 (Parallel): Gc = 2 C fs; G = Gv + Gc; Rp = 1/G; gamma = Gv/G
    const float Gc = 2.0f * __ldg (params + slot_C) * fs;
    const float Gv0 = 1.0f / __ldg (params + slot_R);
    const float Rp0 = 1.0f / (Gv0 + Gc);
    f2 gamma = bc (f2 {}, Gv0 * Rp0), lr = bc (f2 {}, logf (Rp0));
    f2 z { state != nullptr ? state[rowA] : 0.0f, (state != nullptr && validB) ? state[rowB] : 0.0f };
    for (int n = 0; n < T; ++n)
    {
        const f2 xv { __ldg (xa + n), __ldg (xb + n) };
        if (ra != nullptr)
        { // clipper_pot.py:116-117: set_resistance + calc_impedance every sample
            const float GvA = 1.0f / __ldg (ra + n), GvB = 1.0f / __ldg (rb + n);
            const float RpA = 1.0f / (GvA + Gc), RpB = 1.0f / (GvB + Gc);
            gamma = f2 { GvA * RpA, GvB * RpB };
            lr = f2 { logf (RpA), logf (RpB) };
        }
        const f2 t = mulv (gamma, addv (xv, negv (z))); // -p1R (b2 - b1), tf_wdf.py:185-192
        const f2 a = addv (z, t);
        const f2 b = negv (mlp<H> (sw, n_hidden, a, lr)); // clipper_pot.py:119-121 / DiodePairNeuralModel.h:70-75
        const f2 zn = addv (b, t);
        const f2 yo = PY ? mulv (bc (f2 {}, 0.5f), addv (zn, z)) : z;
        ya[n] = yo.x;
        if (validB)
            yb[n] = yo.y;
        z = zn;
    }
    if (state != nullptr)
    {
        state[rowA] = z.x;
        if (validB)
            state[rowB] = z.y;
    }
}
} // namespace

cudaError_t launch_nn_forward (int hidden, int n_hidden, bool pyorder, const float* x, const float* r, float* y, const float* params, int slot_R, int slot_C, float fs, const float* weights,
                               int n_weights, float* state, int64_t B, int64_t T, cudaStream_t stream)
{
    const int64_t pairs = (B + 1) / 2;
    const unsigned grid = (unsigned) ((pairs + 127) / 128);
    const size_t smem = (size_t) ((n_weights + 3) / 4 * 4) * sizeof (float);
#define DWDF_NN(HH) \
    if (hidden == HH) \
    { \
        if (pyorder) \
            nn_clipper_forward<HH, true><<<grid, 128, smem, stream>>> (x, r, y, params, slot_R, slot_C, fs, weights, n_weights, n_hidden, state, B, (int) T); \
        else \
            nn_clipper_forward<HH, false><<<grid, 128, smem, stream>>> (x, r, y, params, slot_R, slot_C, fs, weights, n_weights, n_hidden, state, B, (int) T); \
        return cudaGetLastError (); \
    }
    DWDF_NN (4)
    DWDF_NN (8)
    DWDF_NN (16)
#undef DWDF_NN
    return cudaErrorInvalidValue;
}

} // namespace dwdf
