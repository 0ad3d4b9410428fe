// clipper_dispatch.cu — variant dispatch of the diode-clipper kernels (compiled in four parts from
// clipper_kernels.cu), plus the two tiny single-block kernels of a training step: the fixed-order
// finalize (reduction, chain rule, loss) and Adam.
#include "dwdf_kernels.h"
#include "dwdf_tma.cuh"

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

// block-wide fixed-order sum of the per-group partials; the totals land in sm[k][0]
template <int N>
__device__ __forceinline__ void reduce_partials (double (&sm)[N][256], const double* __restrict__ partials, int64_t n_groups)
{
    const int tid = threadIdx.x;
    double a[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
        a[k] = 0.0;
    for (int64_t g = tid; g < n_groups; g += 256)
#pragma unroll
        for (int k = 0; k < N; ++k)
            a[k] += partials[g * kPartialStride + k];
#pragma unroll
    for (int k = 0; k < N; ++k)
        sm[k][tid] = a[k];
    __syncthreads ();
    for (int o = 128; o > 0; o >>= 1)
    {
        if (tid < o)
#pragma unroll
            for (int k = 0; k < N; ++k)
                sm[k][tid] += sm[k][tid + o];
        __syncthreads ();
    }
}

// raw sums (kAcc* slots, raw[23] = number of samples in the loss) -> out[DWDF_OUT_LEN]: gradients per slot, loss, mse, esr
// the same with the source resistance as an input channel: gamma and Rp vary per sample, so the adjoint kernel has already
// folded them in (kAccGamma = sum G cg gamma (1 - gamma), kAccEllRp = sum G cl Rp); what is left is constant
__device__ __forceinline__ void finalize_math_r (const ClipDesc& desc, const float* __restrict__ params, const double* raw, int target, int loss_kind, double* out);

__device__ __forceinline__ void loss_scale (const double* raw, int target, int loss_kind, double& alpha, double& loss, double& mse, double& esr)
{
    const double sse = raw[kAccSse], st2 = raw[kAccSt2], count = raw[23];
    alpha = 1.0, loss = 0.0, mse = 0.0, esr = 0.0;
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
}

__device__ __forceinline__ void finalize_math_r (const ClipDesc& desc, const float* __restrict__ params, const double* raw, int target, int loss_kind, double* out)
{
    double alpha, loss, mse, esr;
    loss_scale (raw, target, loss_kind, alpha, loss, mse, esr);
    const double C = params[desc.slot_C], Is = params[desc.slot_Is];
    for (int k = 0; k < 24; ++k)
        out[k] = 0.0;
    out[desc.slot_Is] = alpha * raw[kAccEll] / Is; // ell = ln Rp + ln Is
    out[desc.slot_nabla] = alpha * raw[kAccV] * (double) desc.Vt;
    out[desc.slot_C] = alpha * (-raw[kAccGamma] / C - 2.0 * (double) desc.fs * raw[kAccEllRp]); // dgamma/dC = -gamma (1 - gamma) / C, d ell/dC = -2 fs Rp
    out[16] = loss; // (slot_R stays 0: the resistance is an input here, tf_wdf.py:51-52)
    out[17] = mse;
    out[18] = esr;
}

__device__ __forceinline__ void finalize_math (const ClipDesc& desc, const float* __restrict__ params, const double* raw, int target, int loss_kind, double* out)
{
    const double acc_g = raw[kAccGamma], acc_l = raw[kAccEll], acc_v = raw[kAccV], sse = raw[kAccSse], st2 = raw[kAccSt2], count = raw[23];
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

__global__ void __launch_bounds__ (256) clipper_finalize (const ClipDesc desc, const float* __restrict__ params, const double* __restrict__ partials, int64_t n_groups, const double* raw_in, int raw_only, int target, int loss_kind, double count, double* out)
{
    __shared__ double sm[5][256];
    const int tid = threadIdx.x;
    if (raw_in == nullptr)
        reduce_partials (sm, partials, n_groups);
    if (tid != 0)
        return;
    double raw[24];
    for (int k = 0; k < 24; ++k)
        raw[k] = 0.0;
    if (raw_in != nullptr)
    { // sums that were reduced (and possibly all-reduced over ranks) earlier
        raw[kAccGamma] = raw_in[kAccGamma], raw[kAccEll] = raw_in[kAccEll], raw[kAccV] = raw_in[kAccV], raw[kAccSse] = raw_in[kAccSse], raw[kAccSt2] = raw_in[kAccSt2];
        raw[23] = raw_in[23];
    }
    else
    {
        raw[kAccGamma] = sm[kAccGamma][0], raw[kAccEll] = sm[kAccEll][0], raw[kAccV] = sm[kAccV][0], raw[kAccSse] = sm[kAccSse][0], raw[kAccSt2] = sm[kAccSt2][0];
        raw[23] = count;
    }
    if (raw_only)
    {
        for (int k = 0; k < 24; ++k)
            out[k] = raw[k];
        return;
    }
    finalize_math (desc, params, raw, target, loss_kind, out);
}

__global__ void __launch_bounds__ (256) clipper_finalize_r (const ClipDesc desc, const float* __restrict__ params, const double* __restrict__ partials, int64_t n_groups, const double* raw_in, int raw_only, int target, int loss_kind, double count, double* out)
{
    __shared__ double sm[6][256];
    const int tid = threadIdx.x;
    if (raw_in == nullptr)
        reduce_partials (sm, partials, n_groups);
    if (tid != 0)
        return;
    double raw[24];
    for (int k = 0; k < 24; ++k)
        raw[k] = 0.0;
    for (int k = 0; k < 6; ++k)
        raw[k] = raw_in != nullptr ? raw_in[k] : sm[k][0];
    raw[23] = raw_in != nullptr ? raw_in[23] : count;
    if (raw_only)
    {
        for (int k = 0; k < 24; ++k)
            out[k] = raw[k];
        return;
    }
    finalize_math_r (desc, params, raw, target, loss_kind, out);
}

// Adam (clipper_pot.py:180: Adam(1e-4, beta_1=0.5)) + the Keras clip constraints of tf_wdf.py:74,104; slot k, step number t
__device__ __forceinline__ void adam_update (int k, int t, float* __restrict__ params, double grad, float* __restrict__ m, float* __restrict__ v, float lr, float beta1, float beta2, float eps, const float* __restrict__ lo, const float* __restrict__ hi)
{
    const float g = (float) grad;
    const float mk = beta1 * m[k] + (1.0f - beta1) * g;
    const float vk = beta2 * v[k] + (1.0f - beta2) * g * g;
    m[k] = mk;
    v[k] = vk;
    // Keras: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  p -= lr_t * m / (sqrt(v) + eps)
    const float lr_t = lr * sqrtf (1.0f - powf (beta2, (float) t)) / (1.0f - powf (beta1, (float) t));
    float p = params[k] - lr_t * mk / (sqrtf (vk) + eps);
    if (lo != nullptr)
        p = fmaxf (p, lo[k]);
    if (hi != nullptr)
        p = fminf (p, hi[k]);
    params[k] = p;
}

__global__ void adam_kernel (float* __restrict__ params, const double* __restrict__ out, float* __restrict__ m, float* __restrict__ v, int32_t* __restrict__ step, int n_params, float lr, const float* __restrict__ lr_vec, float beta1, float beta2, float eps, double grad_scale, const float* __restrict__ lo, const float* __restrict__ hi)
{
    const int k = threadIdx.x;
    const int t = *step + 1;
    if (k < n_params)
        adam_update (k, t, params, out[k] * grad_scale, m, v, lr_vec != nullptr ? lr_vec[k] : lr, beta1, beta2, eps, lo, hi);
    __syncthreads ();
    if (k == 0)
        *step = t;
}

// =================================================================================================
// multi-GPU: the step's one exchange, over peer memory
// =================================================================================================
// Sequences shard over GPUs with no data-path collective; what a training step exchanges is the sum
// over ranks of a handful of doubles (24 for the analytic root, the weight-gradient vector for the neural root). That
// message is pure latency, so instead of a library all-reduce between two tiny kernels, the reduction kernel does the
// exchange itself through NVLink peer memory: every rank owns a MAILBOX (device memory its peers map with CUDA IPC),
// with one slot per sender and epoch parity. A step's kernel
//     1. writes its vector into slot[my rank] of EVERY rank's mailbox as self-validating 8-byte words (epoch tag | half a
//        double: peer_allreduce below), plain remote stores, no fence,
//     2. polls the slots of its OWN mailbox (local memory) until every word carries this epoch's tag,
//     3. sums the slots in rank order — every rank adds the same numbers in the same order, so the results (and
//        after Adam the parameters) are bit-identical on all ranks without a broadcast.
// Slots alternate with the epoch's parity: a rank can only be one epoch ahead of a peer (it needs that peer's words of
// the current epoch to finish), so the slot it writes next is never one a peer is still reading. A peer that never
// arrives is reported after a timeout (the result block is filled with NaN) instead of hanging the GPU.
__device__ __forceinline__ void st_relaxed_sys (unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_relaxed_sys (const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns ()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// vals[0..n) (shared or global memory of this block) <- sum over ranks. All 256 threads of the single block call it.
// Returns false on timeout (some thread saw a peer missing); the caller decides what to report.
// Wire format: every double travels as two 8-byte words {epoch tag (high 32 bits) | half of the double (low 32 bits)}. An
// 8-byte store is atomic on NVLink, so a word whose tag equals this epoch IS its data: no fence between payload and flag, no
// separate flag round trip — the exchange costs one one-way trip plus the skew between the ranks. (Tag 0 = the zeroed mailbox.)
constexpr int kDpChunk = 64; // doubles received per round: kDpMaxWorld x 2 x kDpChunk halves staged in shared memory
__device__ __forceinline__ bool peer_allreduce (double* vals, int n, const DpPeers& dp, unsigned long long epoch, int* timed_out_sm)
{
    __shared__ unsigned int halves[kDpMaxWorld][2 * kDpChunk];
    const int tid = threadIdx.x;
    const unsigned long long tag = (epoch & 0xffffffffull) << 32;
    const size_t slot_bytes = (size_t) kDpSlotDoubles * 2 * sizeof (unsigned long long);
    const size_t parity_off = (size_t) (epoch & 1ull) * dp.world * slot_bytes;
    // send: my 2 n words into slot[my rank] of every rank's mailbox (my own included)
    for (int p = 0; p < dp.world; ++p)
    {
        unsigned long long* dst = reinterpret_cast<unsigned long long*> (dp.mailbox[p] + parity_off + (size_t) dp.rank * slot_bytes);
        for (int w = tid; w < 2 * n; w += blockDim.x)
        {
            const unsigned long long bits = (unsigned long long) __double_as_longlong (vals[w >> 1]);
            st_relaxed_sys (dst + w, tag | ((w & 1) ? (bits >> 32) : (bits & 0xffffffffull)));
        }
    }
    __syncthreads (); // every word of vals has been read before the sums overwrite it
    const unsigned long long t0 = global_timer_ns ();
    for (int c0 = 0; c0 < n; c0 += kDpChunk)
    {
        const int nc = min (kDpChunk, n - c0);
        // receive: all threads poll the words of this round in parallel (local memory: the peers' stores land here)
        for (int k = tid; k < dp.world * 2 * nc; k += blockDim.x)
        {
            const int p = k / (2 * nc), w = k % (2 * nc);
            const unsigned long long* src = reinterpret_cast<const unsigned long long*> (dp.mailbox[dp.rank] + parity_off + (size_t) p * slot_bytes) + 2 * c0 + w;
            unsigned long long v;
            while (((v = ld_relaxed_sys (src)) & 0xffffffff00000000ull) != tag)
                if (global_timer_ns () - t0 > dp.timeout_ns)
                {
                    *timed_out_sm = 1;
                    break;
                }
            halves[p][w] = (unsigned int) v;
        }
        __syncthreads ();
        const bool ok = *timed_out_sm == 0;
        for (int i = tid; i < nc; i += blockDim.x)
        { // summed in rank order: every rank adds the same numbers in the same order
            double sum = 0.0;
            for (int p = 0; p < dp.world; ++p)
                sum += __longlong_as_double ((long long) (((unsigned long long) halves[p][2 * i + 1] << 32) | halves[p][2 * i]));
            vals[c0 + i] = ok ? sum : __longlong_as_double (0x7ff8000000000000ll);
        }
        __syncthreads ();
    }
    return *timed_out_sm == 0;
}

__device__ __forceinline__ unsigned long long next_epoch (const DpPeers& dp, unsigned long long* epoch_sm)
{
    if (threadIdx.x == 0)
    {
        unsigned long long* ctr = reinterpret_cast<unsigned long long*> (dp.mailbox[dp.rank] + (size_t) 2 * dp.world * kDpSlotDoubles * 2 * sizeof (unsigned long long));
        *epoch_sm = *ctr + 1;
        *ctr = *epoch_sm;
    }
    __syncthreads ();
    return *epoch_sm;
}

// generic: inout[0..n) <- sum over ranks (n <= kDpSlotDoubles - 1)
__global__ void __launch_bounds__ (256) peer_allreduce_kernel (double* __restrict__ inout, int n, const DpPeers dp)
{
    __shared__ unsigned long long epoch_sm;
    __shared__ int timed_out;
    if (threadIdx.x == 0)
        timed_out = 0;
    const unsigned long long epoch = next_epoch (dp, &epoch_sm);
    peer_allreduce (inout, n, dp, epoch, &timed_out);
}

// The data-parallel training step's tail in ONE kernel: fixed-order reduction of this rank's partials -> exchange of
// the raw sums over peer memory -> chain rule and loss -> Adam. count = this rank's samples in the loss.
__global__ void __launch_bounds__ (256) clipper_finalize_dp (const ClipDesc desc, float* __restrict__ params, const double* __restrict__ partials, int64_t n_groups, int target, int loss_kind, double count, double* __restrict__ out, const DpPeers dp,
                                                            float* __restrict__ m, float* __restrict__ v, int32_t* __restrict__ step, int n_params, float lr, const float* __restrict__ lr_vec, float beta1, float beta2, float eps, const float* __restrict__ lo, const float* __restrict__ hi)
{
    __shared__ double sm[5][256];
    __shared__ double raw[24], res[24];
    __shared__ unsigned long long epoch_sm;
    __shared__ int timed_out;
    const int tid = threadIdx.x;
    grid_dependency_wait (); // (a programmatic dependent of the adjoint pass)
    reduce_partials (sm, partials, n_groups);
    if (tid == 0)
    {
        timed_out = 0;
        for (int k = 0; k < 24; ++k)
            raw[k] = 0.0;
        raw[kAccGamma] = sm[kAccGamma][0], raw[kAccEll] = sm[kAccEll][0], raw[kAccV] = sm[kAccV][0], raw[kAccSse] = sm[kAccSse][0], raw[kAccSt2] = sm[kAccSt2][0];
        raw[23] = count;
    }
    __syncthreads ();
    if (dp.world > 1)
    {
        const unsigned long long epoch = next_epoch (dp, &epoch_sm);
        peer_allreduce (raw, 24, dp, epoch, &timed_out);
    }
    if (tid == 0)
    {
        finalize_math (desc, params, raw, target, loss_kind, res);
        for (int k = 0; k < 24; ++k)
            out[k] = res[k];
    }
    __syncthreads ();
    if (m != nullptr)
    {
        const int t = *step + 1;
        __syncthreads ();
        if (tid < n_params)
            adam_update (tid, t, params, res[tid], m, v, lr_vec != nullptr ? lr_vec[tid] : lr, beta1, beta2, eps, lo, hi);
        if (tid == 0)
            *step = t;
    }
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

cudaError_t launch_clipper_forward_r (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* r, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream)
{
    return by_part (v, [&] (auto M, auto G) { return clipper_forward_r_part<decltype (M)::value, decltype (G)::value> (v.pyorder, use_tma, maps, desc, params, x, r, y, ckpt, state, B, T, stream); });
}

cudaError_t launch_clipper_adjoint_r (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* r, const float* y, const float* ckpt, const float* g, bool target, int64_t skip, double* partials, int64_t B, int64_t T, cudaStream_t stream)
{
    return by_part (v, [&] (auto M, auto G) { return clipper_adjoint_r_part<decltype (M)::value, decltype (G)::value> (v.pyorder, use_tma, maps, desc, params, x, r, y, ckpt, g, target, clamp_skip (skip, T), partials, B, T, stream); });
}

cudaError_t launch_clipper_finalize_r (const ClipDesc& desc, const float* params, const double* partials, int64_t n_groups, const double* raw_in, bool raw_only, bool target, int loss_kind, double count, double* out, cudaStream_t stream)
{
    clipper_finalize_r<<<1, 256, 0, stream>>> (desc, params, partials, n_groups, raw_in, raw_only ? 1 : 0, target ? 1 : 0, loss_kind, count, out);
    return cudaGetLastError ();
}

cudaError_t launch_clipper_finalize (const ClipDesc& desc, const float* params, const double* partials, int64_t n_groups, const double* raw_in, bool raw_only, bool target, int loss_kind, double count, double* out, cudaStream_t stream)
{
    clipper_finalize<<<1, 256, 0, stream>>> (desc, params, partials, n_groups, raw_in, raw_only ? 1 : 0, target ? 1 : 0, loss_kind, count, out);
    return cudaGetLastError ();
}

cudaError_t launch_clipper_finalize_dp (const ClipDesc& desc, float* params, const double* partials, int64_t n_groups, bool target, int loss_kind, double count, double* out, const DpPeers& dp, float* m, float* v, int32_t* step, int n_params,
                                        float lr, const float* lr_vec, float beta1, float beta2, float eps, const float* lo, const float* hi, cudaStream_t stream)
{
    return launch_dependent (! (g_clip_opts & kOptNoPdl), clipper_finalize_dp, dim3 (1), dim3 (256), stream, desc, params, partials, n_groups, target ? 1 : 0, loss_kind, count, out, dp, m, v, step, n_params, lr, lr_vec, beta1, beta2, eps, lo,
                             hi);
}

cudaError_t launch_peer_allreduce (double* inout, int n, const DpPeers& dp, cudaStream_t stream)
{
    peer_allreduce_kernel<<<1, 256, 0, stream>>> (inout, n, dp);
    return cudaGetLastError ();
}

cudaError_t launch_adam (float* params, const double* out, float* m, float* v, int32_t* step, int n_params, float lr, const float* lr_vec, float beta1, float beta2, float eps, double grad_scale, const float* lo, const float* hi, cudaStream_t stream)
{
    adam_kernel<<<1, 32, 0, stream>>> (params, out, m, v, step, n_params, lr, lr_vec, beta1, beta2, eps, grad_scale, lo, hi);
    return cudaGetLastError ();
}

} // namespace dwdf
