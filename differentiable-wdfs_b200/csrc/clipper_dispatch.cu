// clipper_dispatch.cu — variant dispatch of the diode-clipper kernels (compiled in four parts from
// clipper_kernels.cu), plus the two tiny single-block kernels of a training step: the fixed-order
// finalize (reduction, chain rule, loss) and Adam.
#include "dwdf_kernels.h"

namespace dwdf
{
std::atomic<int> g_clip_opts { 0 };
std::atomic<int64_t> g_extra_launches { 0 };

namespace
{
// =================================================================================================
// finalize: fixed-order reduction over groups, chain rule to (Is, nabla, R, C), loss
// =================================================================================================
// Loss: tf.keras.losses.MeanSquaredError (clipper_pot.py:176) [+ esr_loss, clipper_pot.py:148-156
// with eps = float64 eps (:145)]; loss = mse + esr (:177). In upstream mode the sums already carry
// the caller's dL/dy scale.
__global__ void __launch_bounds__ (256) clipper_finalize (const ClipDesc desc, const float* __restrict__ params, const double* __restrict__ partials, int64_t n_groups, const double* raw_in, int raw_only, int target, int loss_kind, double count, double* out)
{
    __shared__ double sm[5][256];
    const int tid = threadIdx.x;
    if (raw_in == nullptr)
    {
        double a[5] = { 0, 0, 0, 0, 0 };
        for (int64_t g = tid; g < n_groups; g += 256)
#pragma unroll
            for (int k = 0; k < 5; ++k)
                a[k] += partials[g * kPartialStride + k];
#pragma unroll
        for (int k = 0; k < 5; ++k)
            sm[k][tid] = a[k];
        __syncthreads ();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (tid < o)
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    sm[k][tid] += sm[k][tid + o];
            __syncthreads ();
        }
    }
    if (tid != 0)
        return;
    double acc_g, acc_l, acc_v, sse, st2;
    if (raw_in != nullptr)
    { // sums that were reduced (and possibly all-reduced over ranks) earlier
        acc_g = raw_in[kAccGamma], acc_l = raw_in[kAccEll], acc_v = raw_in[kAccV], sse = raw_in[kAccSse], st2 = raw_in[kAccSt2];
        count = raw_in[23];
    }
    else
        acc_g = sm[kAccGamma][0], acc_l = sm[kAccEll][0], acc_v = sm[kAccV][0], sse = sm[kAccSse][0], st2 = sm[kAccSt2][0];
    if (raw_only)
    {
        for (int k = 0; k < 24; ++k)
            out[k] = 0.0;
        out[kAccGamma] = acc_g, out[kAccEll] = acc_l, out[kAccV] = acc_v, out[kAccSse] = sse, out[kAccSt2] = st2;
        out[23] = count;
        return;
    }
    double alpha = 1.0, loss = 0.0, mse = 0.0, esr = 0.0;
    if (target)
    {
        const double N = count > 0.0 ? count : 1.0;
        mse = sse / N;
        alpha = 2.0 / N;
        loss = mse;
        if (loss_kind == 1)
        {
            const double energy = st2 + 2.220446049250313e-16;
            esr = sqrt (sse / energy / N);
            loss += esr;
            if (esr > 0.0)
                alpha += 1.0 / (esr * energy * N);
        }
    }
    const double R = params[desc.slot_R], C = params[desc.slot_C], Is = params[desc.slot_Is];
    const double fs = desc.fs, Vt = desc.Vt;
    const double Gv = 1.0 / R, Gc = 2.0 * C * fs, Rp = 1.0 / (Gv + Gc), gam = Gv * Rp;
    const double dgam_dR = -gam * (1.0 - gam) / R, dgam_dC = -gam * (1.0 - gam) / C;
    const double dell_dR = Rp / (R * R), dell_dC = -2.0 * fs * Rp;
    for (int k = 0; k < 24; ++k)
        out[k] = 0.0;
    out[desc.slot_Is] = alpha * acc_l / Is;
    out[desc.slot_nabla] = alpha * acc_v * Vt;
    out[desc.slot_R] = alpha * (acc_g * dgam_dR + acc_l * dell_dR);
    out[desc.slot_C] = alpha * (acc_g * dgam_dC + acc_l * dell_dC);
    out[16] = loss;
    out[17] = mse;
    out[18] = esr;
}

// Adam (clipper_pot.py:180: Adam(1e-4, beta_1=0.5)) + the Keras clip constraints of tf_wdf.py:74,104
__global__ void adam_kernel (float* __restrict__ params, const double* __restrict__ out, float* __restrict__ m, float* __restrict__ v, int32_t* __restrict__ step, int n_params, float lr, const float* __restrict__ lr_vec, float beta1, float beta2, float eps, double grad_scale, const float* __restrict__ lo, const float* __restrict__ hi)
{
    const int k = threadIdx.x;
    const int t = *step + 1;
    if (k < n_params)
    {
        const float g = (float) (out[k] * grad_scale);
        const float mk = beta1 * m[k] + (1.0f - beta1) * g;
        const float vk = beta2 * v[k] + (1.0f - beta2) * g * g;
        m[k] = mk;
        v[k] = vk;
        // Keras: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  p -= lr_t * m / (sqrt(v) + eps)
        const float lr_t = (lr_vec != nullptr ? lr_vec[k] : lr) * sqrtf (1.0f - powf (beta2, (float) t)) / (1.0f - powf (beta1, (float) t));
        float p = params[k] - lr_t * mk / (sqrtf (vk) + eps);
        if (lo != nullptr)
            p = fmaxf (p, lo[k]);
        if (hi != nullptr)
            p = fminf (p, hi[k]);
        params[k] = p;
    }
    __syncthreads ();
    if (k == 0)
        *step = t;
}

} // namespace

namespace
{
template <class F>
cudaError_t by_part (const ClipVariant& v, F&& f)
{
    if (v.mode == kModeApprox)
        return v.general ? f (std::integral_constant<int, kModeApprox> {}, std::true_type {}) : f (std::integral_constant<int, kModeApprox> {}, std::false_type {});
    if (v.mode == kModeExact)
        return v.general ? f (std::integral_constant<int, kModeExact> {}, std::true_type {}) : f (std::integral_constant<int, kModeExact> {}, std::false_type {});
    return cudaErrorInvalidValue;
}
int clamp_skip (int64_t skip, int64_t T) { return (int) (skip < 0 ? 0 : (skip > T ? T : skip)); }
} // namespace

cudaError_t launch_clipper_forward (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream)
{
    return by_part (v, [&] (auto M, auto G) { return clipper_forward_part<decltype (M)::value, decltype (G)::value> (v.pyorder, use_tma, maps, desc, params, x, y, ckpt, state, B, T, stream); });
}

cudaError_t launch_clipper_adjoint (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* y, const float* ckpt, const float* g, bool target, int64_t skip, float* gx, double* partials, int64_t B, int64_t T, cudaStream_t stream)
{
    return by_part (v, [&] (auto M, auto G) { return clipper_adjoint_part<decltype (M)::value, decltype (G)::value> (v.pyorder, use_tma, maps, desc, params, x, y, ckpt, g, target, clamp_skip (skip, T), gx, partials, B, T, stream); });
}

cudaError_t launch_clipper_train (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* target, int64_t skip, float* y, double* partials, int64_t B, int64_t T, cudaStream_t stream)
{
    return by_part (v, [&] (auto M, auto G) { return clipper_train_part<decltype (M)::value, decltype (G)::value> (v.pyorder, use_tma, maps, desc, params, x, target, clamp_skip (skip, T), y, partials, B, T, stream); });
}

cudaError_t launch_clipper_finalize (const ClipDesc& desc, const float* params, const double* partials, int64_t n_groups, const double* raw_in, bool raw_only, bool target, int loss_kind, double count, double* out, cudaStream_t stream)
{
    clipper_finalize<<<1, 256, 0, stream>>> (desc, params, partials, n_groups, raw_in, raw_only ? 1 : 0, target ? 1 : 0, loss_kind, count, out);
    return cudaGetLastError ();
}

cudaError_t launch_adam (float* params, const double* out, float* m, float* v, int32_t* step, int n_params, float lr, const float* lr_vec, float beta1, float beta2, float eps, double grad_scale, const float* lo, const float* hi, cudaStream_t stream)
{
    adam_kernel<<<1, 32, 0, stream>>> (params, out, m, v, step, n_params, lr, lr_vec, beta1, beta2, eps, grad_scale, lo, hi);
    return cudaGetLastError ();
}

} // namespace dwdf
