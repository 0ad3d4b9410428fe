// loss_kernels.cu — the reference's training loss AS ITS LOOP CALLS IT (DWDF_LOSS_MSE_ESR_AS_CALLED).
//
// clipper_pot.py:148-156 defines esr_loss(target_y, predicted_y) = sqrt(sum (target - pred)^2 / (sum target^2 + eps) / N),
// and the training loop calls loss_func(outs, train_Y) (clipper_pot.py:248,177): the network OUTPUT lands in the
// `target_y` slot, so the energy the reference really divides by is the prediction's, sum y^2, and that term has a
// gradient of its own:
//     esr = sqrt(S / (E N)),  S = sum (y - t)^2,  E = sum y^2 + eps
//     d esr / dy = (y - t) / (esr N E)  -  (esr / E) y
// The adjoint kernels are linear in dL/dy, but the two coefficients depend on S and E of the WHOLE batch, which are only
// known after the forward pass. So this loss runs as a composition of what exists: one streaming reduction over (y, t)
// (8 B/sample), one pass that writes dL/dy (12 B/sample), then the ordinary reverse sweep in DWDF_GRAD_UPSTREAM mode.
// All roots (analytic clipper, tree interpreter, neural) share it; across GPUs the four sums are exchanged before the
// second pass. The textbook form (energy of the target, DWDF_LOSS_MSE_ESR) stays fused in the adjoint kernels.
#include "dwdf_kernels.h"
#include "../../include/dwdf.h"

namespace dwdf
{
namespace
{
constexpr int kLossThreads = 256;

__device__ __forceinline__ void block_sum4 (double (&v)[4], double (&sm)[4][kLossThreads])
{
    const int tid = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        sm[k][tid] = v[k];
    __syncthreads ();
    for (int o = kLossThreads / 2; o > 0; o >>= 1)
    {
        if (tid < o)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                sm[k][tid] += sm[k][tid + o];
        __syncthreads ();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        v[k] = sm[k][0];
}

// block b owns rows b, b + gridDim.x, ...: a fixed assignment and a fixed summation order (bit-reproducible)
__global__ void __launch_bounds__ (kLossThreads) loss_sums_kernel (const float* __restrict__ y, const float* __restrict__ t, int64_t B, int T, int skip, double* __restrict__ partials)
{
    __shared__ double sm[4][kLossThreads];
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x)
    {
        const float* yr = y + b * T;
        const float* tr = t + b * T;
        float s = 0.0f, ey = 0.0f, et = 0.0f; // fp32 within a row slice (at most T / 256 terms per thread), double across rows
        for (int n = skip + threadIdx.x; n < T; n += kLossThreads)
        {
            const float yv = __ldg (yr + n), tv = __ldg (tr + n), e = yv - tv;
            s = fmaf (e, e, s);
            ey = fmaf (yv, yv, ey);
            et = fmaf (tv, tv, et);
        }
        acc[0] += (double) s;
        acc[1] += (double) ey;
        acc[2] += (double) et;
    }
    block_sum4 (acc, sm);
    if (threadIdx.x == 0)
    {
        partials[4 * blockIdx.x + 0] = acc[0];
        partials[4 * blockIdx.x + 1] = acc[1];
        partials[4 * blockIdx.x + 2] = acc[2];
        partials[4 * blockIdx.x + 3] = 0.0;
    }
}

__global__ void __launch_bounds__ (kLossThreads) loss_reduce_kernel (const double* __restrict__ partials, int n_blocks, double count, double* __restrict__ sums)
{
    __shared__ double sm[4][kLossThreads];
    double acc[4] = { 0.0, 0.0, 0.0, 0.0 };
    for (int i = threadIdx.x; i < n_blocks; i += kLossThreads)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            acc[k] += partials[4 * i + k];
    block_sum4 (acc, sm);
    if (threadIdx.x == 0)
    {
        sums[0] = acc[0], sums[1] = acc[1], sums[2] = acc[2];
        sums[3] = count; // samples in the loss (this rank's; summed over ranks by the exchange)
    }
}

struct LossCoef
{
    double c1, c2, loss, mse, esr;
};
__device__ __forceinline__ LossCoef loss_coef (const double* __restrict__ sums)
{
    const double S = sums[0], E = sums[1] + 2.220446049250313e-16, N = sums[3] > 0.0 ? sums[3] : 1.0; // eps = np.finfo(float).eps, clipper_pot.py:145
    LossCoef c;
    c.mse = S / N;
    c.esr = sqrt (S / E / N);
    c.loss = c.mse + c.esr;
    c.c1 = 2.0 / N + (c.esr > 0.0 ? 1.0 / (c.esr * N * E) : 0.0);
    c.c2 = c.esr > 0.0 ? -c.esr / E : 0.0;
    return c;
}

__global__ void __launch_bounds__ (kLossThreads) loss_ybar_kernel (const float* __restrict__ y, const float* __restrict__ t, int64_t total, int T, int skip, const double* __restrict__ sums, float* __restrict__ ybar)
{
    const LossCoef c = loss_coef (sums);
    const float c1 = (float) c.c1, c12 = (float) (c.c1 + c.c2);
    for (int64_t i = (int64_t) blockIdx.x * kLossThreads + threadIdx.x; i < total; i += (int64_t) gridDim.x * kLossThreads)
    {
        const int n = (int) (i % T);
        ybar[i] = n >= skip ? fmaf (c12, __ldg (y + i), -c1 * __ldg (t + i)) : 0.0f; // c1 (y - t) + c2 y
    }
}

__global__ void loss_write_kernel (const double* __restrict__ sums, double* __restrict__ out)
{
    const LossCoef c = loss_coef (sums);
    out[DWDF_OUT_LOSS] = c.loss;
    out[DWDF_OUT_MSE] = c.mse;
    out[DWDF_OUT_ESR] = c.esr;
}
} // namespace

int loss_sum_blocks (int64_t B) { return (int) (B < 148 * 8 ? (B > 0 ? B : 1) : 148 * 8); }
size_t loss_scratch_doubles (int64_t B) { return (size_t) 4 * loss_sum_blocks (B) + 8; }

// scratch: loss_scratch_doubles(B) doubles; the four sums land in scratch[0 .. 4)
cudaError_t launch_loss_sums (const float* y, const float* t, int64_t B, int64_t T, int skip, double* scratch, cudaStream_t stream)
{
    const int nb = loss_sum_blocks (B);
    loss_sums_kernel<<<nb, kLossThreads, 0, stream>>> (y, t, B, (int) T, skip, scratch + 8);
    loss_reduce_kernel<<<1, kLossThreads, 0, stream>>> (scratch + 8, nb, (double) B * (double) (T - skip), scratch);
    return cudaGetLastError ();
}

cudaError_t launch_loss_ybar (const float* y, const float* t, int64_t B, int64_t T, int skip, const double* sums, float* ybar, cudaStream_t stream)
{
    const int64_t total = B * T;
    const int64_t want = (total + kLossThreads * 8 - 1) / (kLossThreads * 8);
    loss_ybar_kernel<<<(unsigned) (want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16), kLossThreads, 0, stream>>> (y, t, total, (int) T, skip, sums, ybar);
    return cudaGetLastError ();
}

cudaError_t launch_loss_write (const double* sums, double* out, cudaStream_t stream)
{
    loss_write_kernel<<<1, 1, 0, stream>>> (sums, out);
    return cudaGetLastError ();
}

} // namespace dwdf
