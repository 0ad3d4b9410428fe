// clipper_kernels.cu — the diode-clipper hot path on sm_100a.
//
// Circuit: Parallel(ResistiveVoltageSource, Capacitor) closed by an analytic DiodePair root
// (clipper_pot.py:94-127 with the root of diode_pretraining.py:39-60; DiodeClipperWDF.h:18-25).
// One sample is the explicit map  z' = f((1-g) z + g x) + g (x - z)  on the single capacitor state.
//
// Work decomposition (B200-first, not the reference's per-sample op graph):
//   * one warp owns 32 sequences (one lane per circuit instance) and is its own CTA: warps never
//     synchronise with each other, so the grid is simply ceil(B/32) one-warp CTAs, 13-14 resident
//     per SM at B = 65536 — every SM sub-partition holds 3-4 independent recurrences to interleave;
//   * inputs/outputs are (B, T) batch-major. A lane marching along its own row would touch 32
//     different 128-byte lines per load, so the warp moves [32 sequences x 32 samples] tiles with the
//     TMA unit (cp.async.bulk.tensor, 128-byte swizzle) into a private 3-slot shared-memory ring,
//     completion on an mbarrier; each lane then reads its row 4 samples at a time with conflict-free
//     LDS.128, writes the outputs back IN PLACE and one lane issues the TMA store of the tile;
//   * forward writes the capacitor state every 16 samples; the adjoint walks those checkpoints in
//     reverse, replays a 16-sample segment into a register tape and sweeps it backwards — no tape
//     in memory, no autodiff framework;
//   * parameter gradients are reduced lane -> warp (shuffles, fixed order) -> partials[group] ->
//     one finalize block (fixed order): bit-reproducible run to run.
#include "dwdf_kernels.h"
#include "dwdf_tma.cuh"

#include <type_traits>

namespace dwdf
{

namespace
{
constexpr int kLanes = 32;
constexpr int kFwdTileT = 32; // samples per forward tile: 32 x 4 B = one 128-byte swizzle row
constexpr int kFwdTileBytes = kLanes * kFwdTileT * 4; // 4 KB
constexpr int kFwdStages = 3;
constexpr int kAdjTileT = kSeg; // samples per adjoint tile = one checkpoint segment (64-byte rows)
constexpr int kAdjTileBytes = kLanes * kAdjTileT * 4; // 2 KB
constexpr int kAdjStages = 3;

// 16-byte chunk `c` (4 samples) of row `lane` inside a swizzled tile
__device__ __forceinline__ uint32_t chunk128 (uint32_t tile, int lane, int c) { return tile + lane * 128 + ((c ^ (lane & 7)) << 4); } // CU_TENSOR_MAP_SWIZZLE_128B
__device__ __forceinline__ uint32_t chunk64 (uint32_t tile, int lane, int c) { return tile + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4); } // CU_TENSOR_MAP_SWIZZLE_64B

__device__ __forceinline__ void load_consts (ClipConst& c, const ClipDesc& d, const float* __restrict__ params)
{
    clip_setup (c, d, __ldg (params + d.slot_R), __ldg (params + d.slot_C), __ldg (params + d.slot_Is), __ldg (params + d.slot_nabla));
}

__device__ __forceinline__ double warp_sum (double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync (0xffffffffu, v, o);
    return v;
}

// =================================================================================================
// forward
// =================================================================================================
template <int MODE, bool GENERAL, bool LSMALL, bool PY>
__device__ __forceinline__ void forward_tma_body (const ClipConst& c, const CUtensorMap* tmx, const CUtensorMap* tmy, uint32_t tiles, uint32_t bars, float* __restrict__ ckpt, float& z, int64_t B, int T, int lane, int b0)
{
    const int ntiles = (T + kFwdTileT - 1) / kFwdTileT;
    const bool valid = (int64_t) b0 + lane < B;
    if (lane == 0)
    {
        for (int s = 0; s < kFwdStages - 1 && s < ntiles; ++s)
        {
            mbar_expect_tx (bars + 8 * s, kFwdTileBytes);
            tma_load_2d (tiles + s * kFwdTileBytes, tmx, s * kFwdTileT, b0, bars + 8 * s);
        }
    }
    for (int i = 0; i < ntiles; ++i)
    {
        const int s = i % kFwdStages;
        const uint32_t tile = tiles + s * kFwdTileBytes;
        mbar_wait (bars + 8 * s, (i / kFwdStages) & 1);
        const int nch = min (8, (T - i * kFwdTileT) >> 2);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
        {
            if (cc < nch)
            {
                if ((cc & 3) == 0 && ckpt != nullptr && valid)
                    ckpt[(int64_t) (i * 2 + (cc >> 2)) * B + b0 + lane] = z;
                const uint32_t addr = chunk128 (tile, lane, cc);
                const float4 v = lds128 (addr);
                float4 o;
                o.x = clip_step<MODE, GENERAL, LSMALL, PY> (c, v.x, z);
                o.y = clip_step<MODE, GENERAL, LSMALL, PY> (c, v.y, z);
                o.z = clip_step<MODE, GENERAL, LSMALL, PY> (c, v.z, z);
                o.w = clip_step<MODE, GENERAL, LSMALL, PY> (c, v.w, z);
                sts128 (addr, o);
            }
        }
        fence_proxy_async (); // my st.shared results -> visible to the TMA unit
        __syncwarp ();
        if (lane == 0)
        {
            tma_store_2d (tmy, i * kFwdTileT, b0, tile);
            tma_commit ();
            const int j = i + kFwdStages - 1; // next tile to fetch goes into the slot tile i-1 was stored from
            if (j < ntiles)
            {
                tma_wait_read<1> ();
                const int sj = j % kFwdStages;
                mbar_expect_tx (bars + 8 * sj, kFwdTileBytes);
                tma_load_2d (tiles + sj * kFwdTileBytes, tmx, j * kFwdTileT, b0, bars + 8 * sj);
            }
        }
    }
    if (lane == 0)
        tma_wait_all<0> ();
}

template <int MODE, bool GENERAL, bool PY>
__global__ void __launch_bounds__ (kLanes) clipper_forward_tma (const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, const float* __restrict__ params, const ClipDesc desc, float* __restrict__ ckpt, float* __restrict__ state, int64_t B, int T)
{
    __shared__ __align__ (1024) uint8_t smem[kFwdStages * kFwdTileBytes];
    __shared__ __align__ (8) uint64_t bar_mem[kFwdStages];
    const int lane = threadIdx.x;
    const int b0 = blockIdx.x * kLanes;
    const uint32_t tiles = smem_u32 (smem), bars = smem_u32 (bar_mem);
    if (lane == 0)
    {
        tma_prefetch_desc (&tmx);
        tma_prefetch_desc (&tmy);
        for (int s = 0; s < kFwdStages; ++s)
            mbar_init (bars + 8 * s, 1);
        fence_mbar_init ();
    }
    __syncwarp ();
    ClipConst c;
    load_consts (c, desc, params);
    const bool valid = (int64_t) b0 + lane < B;
    float z = (state != nullptr && valid) ? state[b0 + lane] : 0.0f;
    if (MODE == kModeApprox && ! GENERAL && c.pair.L < kOmega3Zero)
        forward_tma_body<MODE, GENERAL, true, PY> (c, &tmx, &tmy, tiles, bars, ckpt, z, B, T, lane, b0);
    else
        forward_tma_body<MODE, GENERAL, false, PY> (c, &tmx, &tmy, tiles, bars, ckpt, z, B, T, lane, b0);
    if (state != nullptr && valid)
        state[b0 + lane] = z;
}

// Same recurrence with plain global accesses: any T, any alignment (the TMA path needs T % 4 == 0
// and 16-byte aligned rows). One lane per sequence.
template <int MODE, bool GENERAL, bool LSMALL, bool PY>
__device__ __forceinline__ void forward_direct_body (const ClipConst& c, const float* __restrict__ xr, float* __restrict__ yr, float* __restrict__ ckpt, float& z, int64_t B, int64_t b, int T)
{
#pragma unroll 4
    for (int n = 0; n < T; ++n)
    {
        if ((n & (kSeg - 1)) == 0 && ckpt != nullptr)
            ckpt[(int64_t) (n / kSeg) * B + b] = z;
        yr[n] = clip_step<MODE, GENERAL, LSMALL, PY> (c, __ldg (xr + n), z);
    }
}

template <int MODE, bool GENERAL, bool PY>
__global__ void __launch_bounds__ (kLanes) clipper_forward_direct (const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ params, const ClipDesc desc, float* __restrict__ ckpt, float* __restrict__ state, int64_t B, int T)
{
    const int64_t b = (int64_t) blockIdx.x * kLanes + threadIdx.x;
    if (b >= B)
        return;
    ClipConst c;
    load_consts (c, desc, params);
    float z = state != nullptr ? state[b] : 0.0f;
    if (MODE == kModeApprox && ! GENERAL && c.pair.L < kOmega3Zero)
        forward_direct_body<MODE, GENERAL, true, PY> (c, x + b * T, y + b * T, ckpt, z, B, b, T);
    else
        forward_direct_body<MODE, GENERAL, false, PY> (c, x + b * T, y + b * T, ckpt, z, B, b, T);
    if (state != nullptr)
        state[b] = z;
}

// =================================================================================================
// adjoint
// =================================================================================================
// Per-lane accumulators of the reverse sweep. fp32 inside a 16-sample segment, double across segments.
struct AdjAcc
{
    double g = 0.0, l = 0.0, v = 0.0, sse = 0.0, st2 = 0.0;
};

// One checkpoint segment: replay kSeg samples from state z into a register tape, then sweep it
// backwards with the running adjoint G of z[n+1].
//   IO::x4(cc) / IO::g4(cc): 4 samples of x / (dL/dy or target); IO::put_e / get_e: stash of the
//   residuals e = y - target (target mode); IO::put_gx: dL/dx (optional).
template <int MODE, bool GENERAL, bool LSMALL, bool PY, bool TARGET, bool WANT_GX, class IO>
__device__ __forceinline__ void adjoint_segment (const ClipConst& c, IO& io, float z, float& G, AdjAcc& acc, int n0, int nvalid, int skip)
{
    StepTape tape[kSeg];
    float sse = 0.0f, st2 = 0.0f;
#pragma unroll
    for (int cc = 0; cc < kSeg / 4; ++cc)
    {
        if (cc * 4 < nvalid)
        {
            const float4 xv = io.x4 (cc);
            const float xs[4] = { xv.x, xv.y, xv.z, xv.w };
            float ys[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                ys[k] = clip_step_tape<MODE, GENERAL, LSMALL, PY> (c, xs[k], z, tape[cc * 4 + k]);
            if (TARGET)
            {
                const float4 tv = io.g4 (cc);
                const float ts[4] = { tv.x, tv.y, tv.z, tv.w };
                float es[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const bool on = n0 + cc * 4 + k >= skip && cc * 4 + k < nvalid;
                    es[k] = on ? ys[k] - ts[k] : 0.0f;
                    sse = fma_ (es[k], es[k], sse);
                    st2 = on ? fma_ (ts[k], ts[k], st2) : st2;
                }
                io.put_e (cc, make_float4 (es[0], es[1], es[2], es[3]));
            }
        }
    }
    float ag = 0.0f, al = 0.0f, av = 0.0f;
    const float gxk = c.gamma / (1.0f - c.gamma); // dz'/dx = gamma (f'+1) = (A + 1) gamma / (1 - gamma)
#pragma unroll
    for (int cc = kSeg / 4 - 1; cc >= 0; --cc)
    {
        if (cc * 4 < nvalid)
        {
            const float4 gv = TARGET ? io.get_e (cc) : io.g4 (cc);
            const float gs[4] = { gv.x, gv.y, gv.z, gv.w };
            float gxs[4];
#pragma unroll
            for (int k = 3; k >= 0; --k)
            {
                const StepTape& tp = tape[cc * 4 + k];
                const float gy = (TARGET || cc * 4 + k < nvalid) ? gs[k] : 0.0f;
                if (PY)
                    G = fma_ (0.5f, gy, G); // y[n] = (z[n+1] + z[n]) / 2
                ag = fma_ (G, tp.cg, ag);
                al = fma_ (G, tp.cl, al);
                av = fma_ (G, tp.cv, av);
                if (WANT_GX)
                    gxs[k] = G * (tp.A + 1.0f) * gxk;
                G = fma_ (G, tp.A, PY ? 0.5f * gy : gy);
            }
            if (WANT_GX)
                io.put_gx (cc, make_float4 (gxs[0], gxs[1], gxs[2], gxs[3]));
        }
    }
    acc.g += (double) ag;
    acc.l += (double) al;
    acc.v += (double) av;
    if (TARGET)
    {
        acc.sse += (double) sse;
        acc.st2 += (double) st2;
    }
}

__device__ __forceinline__ void write_partials (AdjAcc& acc, double* __restrict__ partials, int group, int lane)
{
    const double g = warp_sum (acc.g), l = warp_sum (acc.l), v = warp_sum (acc.v), sse = warp_sum (acc.sse), st2 = warp_sum (acc.st2);
    if (lane == 0)
    {
        double* p = partials + (int64_t) group * kPartialStride;
        p[kAccGamma] = g;
        p[kAccEll] = l;
        p[kAccV] = v;
        p[kAccSse] = sse;
        p[kAccSt2] = st2;
    }
}

// shared-memory tile IO of the TMA adjoint (x tile and g tile of one 16-sample segment, 64-byte swizzle)
struct TileIO
{
    uint32_t xt, gt;
    int lane;
    __device__ __forceinline__ float4 x4 (int cc) const { return lds128 (chunk64 (xt, lane, cc)); }
    __device__ __forceinline__ float4 g4 (int cc) const { return lds128 (chunk64 (gt, lane, cc)); }
    __device__ __forceinline__ void put_e (int cc, float4 e) const { sts128 (chunk64 (gt, lane, cc), e); } // residuals overwrite the target in place
    __device__ __forceinline__ float4 get_e (int cc) const { return lds128 (chunk64 (gt, lane, cc)); }
    __device__ __forceinline__ void put_gx (int, float4) const {}
};

template <int MODE, bool GENERAL, bool LSMALL, bool PY, bool TARGET>
__device__ __forceinline__ void adjoint_tma_body (const ClipConst& c, const CUtensorMap* tmx, const CUtensorMap* tmg, uint32_t tiles, uint32_t bars, const float* __restrict__ ckpt, AdjAcc& acc, int64_t B, int T, int skip, int lane, int b0)
{
    const int nseg = (T + kSeg - 1) / kSeg;
    const bool valid = (int64_t) b0 + lane < B;
    constexpr int kStageBytes = 2 * kAdjTileBytes;
    // segments are processed last to first; the k-th processed segment is i = nseg - 1 - k
    if (lane == 0)
    {
        for (int k = 0; k < kAdjStages - 1 && k < nseg; ++k)
        {
            const int i = nseg - 1 - k;
            mbar_expect_tx (bars + 8 * k, kStageBytes);
            tma_load_2d (tiles + k * kStageBytes, tmx, i * kSeg, b0, bars + 8 * k);
            tma_load_2d (tiles + k * kStageBytes + kAdjTileBytes, tmg, i * kSeg, b0, bars + 8 * k);
        }
    }
    float G = 0.0f;
    float znext = valid ? __ldg (ckpt + (int64_t) (nseg - 1) * B + b0 + lane) : 0.0f;
    for (int k = 0; k < nseg; ++k)
    {
        const int i = nseg - 1 - k;
        const int s = k % kAdjStages;
        const float z = znext;
        if (i > 0 && valid)
            znext = __ldg (ckpt + (int64_t) (i - 1) * B + b0 + lane); // in flight while this segment is processed
        mbar_wait (bars + 8 * s, (k / kAdjStages) & 1);
        TileIO io { tiles + s * kStageBytes, tiles + s * kStageBytes + kAdjTileBytes, lane };
        adjoint_segment<MODE, GENERAL, LSMALL, PY, TARGET, false> (c, io, z, G, acc, i * kSeg, min (kSeg, T - i * kSeg), skip);
        const int kn = k + kAdjStages - 1; // refill the slot freed by the previous segment
        if (kn < nseg)
        {
            fence_proxy_async ();
            __syncwarp ();
            if (lane == 0)
            {
                const int in = nseg - 1 - kn, sn = kn % kAdjStages;
                mbar_expect_tx (bars + 8 * sn, kStageBytes);
                tma_load_2d (tiles + sn * kStageBytes, tmx, in * kSeg, b0, bars + 8 * sn);
                tma_load_2d (tiles + sn * kStageBytes + kAdjTileBytes, tmg, in * kSeg, b0, bars + 8 * sn);
            }
        }
    }
}

template <int MODE, bool GENERAL, bool PY, bool TARGET>
__global__ void __launch_bounds__ (kLanes, 16) clipper_adjoint_tma (const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmg, const float* __restrict__ params, const ClipDesc desc, const float* __restrict__ ckpt, double* __restrict__ partials, int64_t B, int T, int skip)
{
    __shared__ __align__ (1024) uint8_t smem[kAdjStages * 2 * kAdjTileBytes];
    __shared__ __align__ (8) uint64_t bar_mem[kAdjStages];
    const int lane = threadIdx.x;
    const int b0 = blockIdx.x * kLanes;
    const uint32_t tiles = smem_u32 (smem), bars = smem_u32 (bar_mem);
    if (lane == 0)
    {
        tma_prefetch_desc (&tmx);
        tma_prefetch_desc (&tmg);
        for (int s = 0; s < kAdjStages; ++s)
            mbar_init (bars + 8 * s, 1);
        fence_mbar_init ();
    }
    __syncwarp ();
    ClipConst c;
    load_consts (c, desc, params);
    AdjAcc acc;
    if (MODE == kModeApprox && ! GENERAL && c.pair.L < kOmega3Zero)
        adjoint_tma_body<MODE, GENERAL, true, PY, TARGET> (c, &tmx, &tmg, tiles, bars, ckpt, acc, B, T, skip, lane, b0);
    else
        adjoint_tma_body<MODE, GENERAL, false, PY, TARGET> (c, &tmx, &tmg, tiles, bars, ckpt, acc, B, T, skip, lane, b0);
    write_partials (acc, partials, blockIdx.x, lane);
}

// direct-access IO (any T / alignment, optional dL/dx)
struct GlobalIO
{
    const float* __restrict__ xr;
    const float* __restrict__ gr;
    float* __restrict__ gxr;
    int n0, T;
    float4 e[kSeg / 4];
    __device__ __forceinline__ float at (const float* __restrict__ p, int n) const { return n < T ? __ldg (p + n) : 0.0f; }
    __device__ __forceinline__ float4 x4 (int cc) const
    {
        const int n = n0 + cc * 4;
        return make_float4 (at (xr, n), at (xr, n + 1), at (xr, n + 2), at (xr, n + 3));
    }
    __device__ __forceinline__ float4 g4 (int cc) const
    {
        const int n = n0 + cc * 4;
        return make_float4 (at (gr, n), at (gr, n + 1), at (gr, n + 2), at (gr, n + 3));
    }
    __device__ __forceinline__ void put_e (int cc, float4 v) { e[cc] = v; }
    __device__ __forceinline__ float4 get_e (int cc) const { return e[cc]; }
    __device__ __forceinline__ void put_gx (int cc, float4 v) const
    {
        const int n = n0 + cc * 4;
        const float vs[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (n + k < T)
                gxr[n + k] = vs[k];
    }
};

template <int MODE, bool GENERAL, bool LSMALL, bool PY, bool TARGET, bool WANT_GX>
__device__ __forceinline__ void adjoint_direct_body (const ClipConst& c, const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx, const float* __restrict__ ckpt, AdjAcc& acc, int64_t B, int64_t b, int T, int skip)
{
    const int nseg = (T + kSeg - 1) / kSeg;
    float G = 0.0f;
    GlobalIO io { x + b * T, g + b * T, WANT_GX ? gx + b * T : nullptr, 0, T, {} };
    for (int i = nseg - 1; i >= 0; --i)
    {
        io.n0 = i * kSeg;
        const float z = __ldg (ckpt + (int64_t) i * B + b);
        adjoint_segment<MODE, GENERAL, LSMALL, PY, TARGET, WANT_GX> (c, io, z, G, acc, i * kSeg, min (kSeg, T - i * kSeg), skip);
    }
}

template <int MODE, bool GENERAL, bool PY, bool TARGET, bool WANT_GX>
__global__ void __launch_bounds__ (kLanes, 12) clipper_adjoint_direct (const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx, const float* __restrict__ params, const ClipDesc desc, const float* __restrict__ ckpt, double* __restrict__ partials, int64_t B, int T, int skip)
{
    const int lane = threadIdx.x;
    const int64_t b = (int64_t) blockIdx.x * kLanes + lane;
    ClipConst c;
    load_consts (c, desc, params);
    AdjAcc acc;
    if (b < B)
    {
        if (MODE == kModeApprox && ! GENERAL && c.pair.L < kOmega3Zero)
            adjoint_direct_body<MODE, GENERAL, true, PY, TARGET, WANT_GX> (c, x, g, gx, ckpt, acc, B, b, T, skip);
        else
            adjoint_direct_body<MODE, GENERAL, false, PY, TARGET, WANT_GX> (c, x, g, gx, ckpt, acc, B, b, T, skip);
    }
    write_partials (acc, partials, blockIdx.x, lane);
}

// =================================================================================================
// fused training pass: forward + loss + parameter sensitivities in one sweep
// =================================================================================================
// With only three raw parameters (gamma, ell, V) the tangents s_k[n] = dz[n]/dk ride along the
// recurrence:  s_k' = A s_k + c_k,  dL/dk = sum_n e[n] dy[n]/dk,  dy/dk = (s_k' + s_k)/2 (python
// ordering) or s_k (plugin ordering). Same raw sums as the adjoint, one pass, no checkpoints.
template <int MODE, bool GENERAL, bool LSMALL, bool PY>
struct TrainState
{
    float z = 0.0f, sg = 0.0f, sl = 0.0f, sv = 0.0f;
    float ag = 0.0f, al = 0.0f, av = 0.0f, sse = 0.0f, st2 = 0.0f;
    __device__ __forceinline__ float step (const ClipConst& c, float x, float t, bool on)
    {
        StepTape tp;
        const float y = clip_step_tape<MODE, GENERAL, LSMALL, PY> (c, x, z, tp);
        const float e = on ? y - t : 0.0f;
        const float ng = fma_ (tp.A, sg, tp.cg), nl = fma_ (tp.A, sl, tp.cl), nv = fma_ (tp.A, sv, tp.cv);
        if (PY)
        {
            const float h = 0.5f * e;
            ag = fma_ (h, ng + sg, ag);
            al = fma_ (h, nl + sl, al);
            av = fma_ (h, nv + sv, av);
        }
        else
        {
            ag = fma_ (e, sg, ag);
            al = fma_ (e, sl, al);
            av = fma_ (e, sv, av);
        }
        sse = fma_ (e, e, sse);
        st2 = on ? fma_ (t, t, st2) : st2;
        sg = ng;
        sl = nl;
        sv = nv;
        return y;
    }
    __device__ __forceinline__ void flush (AdjAcc& acc)
    {
        acc.g += (double) ag;
        acc.l += (double) al;
        acc.v += (double) av;
        acc.sse += (double) sse;
        acc.st2 += (double) st2;
        ag = al = av = sse = st2 = 0.0f;
    }
};

template <int MODE, bool GENERAL, bool LSMALL, bool PY>
__device__ __forceinline__ void train_tma_body (const ClipConst& c, const CUtensorMap* tmx, const CUtensorMap* tmt, const CUtensorMap* tmy, bool want_y, uint32_t tiles, uint32_t bars, AdjAcc& acc, int T, int skip, int lane, int b0)
{
    constexpr int kStageBytes = 2 * kFwdTileBytes; // x tile (outputs written in place) + target tile
    constexpr int kStages = 2;
    const int ntiles = (T + kFwdTileT - 1) / kFwdTileT;
    if (lane == 0)
    {
        mbar_expect_tx (bars, kStageBytes);
        tma_load_2d (tiles, tmx, 0, b0, bars);
        tma_load_2d (tiles + kFwdTileBytes, tmt, 0, b0, bars);
    }
    TrainState<MODE, GENERAL, LSMALL, PY> st;
    for (int i = 0; i < ntiles; ++i)
    {
        const int s = i % kStages;
        const uint32_t xt = tiles + s * kStageBytes, tt = xt + kFwdTileBytes;
        mbar_wait (bars + 8 * s, (i / kStages) & 1);
        const int nch = min (8, (T - i * kFwdTileT) >> 2);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
        {
            if (cc == 2 && lane == 0 && i + 1 < ntiles)
            { // fetch tile i+1 into the other slot; the store issued from it at i-1 has had a quarter tile to drain
                const int sn = (i + 1) % kStages;
                if (want_y)
                    tma_wait_read<0> ();
                mbar_expect_tx (bars + 8 * sn, kStageBytes);
                tma_load_2d (tiles + sn * kStageBytes, tmx, (i + 1) * kFwdTileT, b0, bars + 8 * sn);
                tma_load_2d (tiles + sn * kStageBytes + kFwdTileBytes, tmt, (i + 1) * kFwdTileT, b0, bars + 8 * sn);
            }
            if (cc < nch)
            {
                const uint32_t addr = chunk128 (xt, lane, cc);
                const float4 v = lds128 (addr), t = lds128 (chunk128 (tt, lane, cc));
                const int n = i * kFwdTileT + cc * 4;
                float4 o;
                o.x = st.step (c, v.x, t.x, n >= skip);
                o.y = st.step (c, v.y, t.y, n + 1 >= skip);
                o.z = st.step (c, v.z, t.z, n + 2 >= skip);
                o.w = st.step (c, v.w, t.w, n + 3 >= skip);
                if (want_y)
                    sts128 (addr, o);
                if ((cc & 3) == 3)
                    st.flush (acc);
            }
        }
        st.flush (acc);
        fence_proxy_async ();
        __syncwarp ();
        if (lane == 0 && want_y)
        {
            tma_store_2d (tmy, i * kFwdTileT, b0, xt);
            tma_commit ();
        }
    }
    if (lane == 0 && want_y)
        tma_wait_all<0> ();
}

template <int MODE, bool GENERAL, bool PY>
__global__ void __launch_bounds__ (kLanes, 16) clipper_train_tma (const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmt, const __grid_constant__ CUtensorMap tmy, const int want_y, const float* __restrict__ params, const ClipDesc desc, double* __restrict__ partials, int T, int skip)
{
    __shared__ __align__ (1024) uint8_t smem[2 * 2 * kFwdTileBytes];
    __shared__ __align__ (8) uint64_t bar_mem[2];
    const int lane = threadIdx.x;
    const int b0 = blockIdx.x * kLanes;
    const uint32_t tiles = smem_u32 (smem), bars = smem_u32 (bar_mem);
    if (lane == 0)
    {
        mbar_init (bars, 1);
        mbar_init (bars + 8, 1);
        fence_mbar_init ();
    }
    __syncwarp ();
    ClipConst c;
    load_consts (c, desc, params);
    AdjAcc acc;
    if (MODE == kModeApprox && ! GENERAL && c.pair.L < kOmega3Zero)
        train_tma_body<MODE, GENERAL, true, PY> (c, &tmx, &tmt, &tmy, want_y != 0, tiles, bars, acc, T, skip, lane, b0);
    else
        train_tma_body<MODE, GENERAL, false, PY> (c, &tmx, &tmt, &tmy, want_y != 0, tiles, bars, acc, T, skip, lane, b0);
    write_partials (acc, partials, blockIdx.x, lane);
}

template <int MODE, bool GENERAL, bool PY>
__global__ void __launch_bounds__ (kLanes) clipper_train_direct (const float* __restrict__ x, const float* __restrict__ target, float* __restrict__ y, const float* __restrict__ params, const ClipDesc desc, double* __restrict__ partials, int64_t B, int T, int skip)
{
    const int lane = threadIdx.x;
    const int64_t b = (int64_t) blockIdx.x * kLanes + lane;
    ClipConst c;
    load_consts (c, desc, params);
    AdjAcc acc;
    if (b < B)
    {
        const float* xr = x + b * T;
        const float* tr = target + b * T;
        float* yr = y != nullptr ? y + b * T : nullptr;
        const bool lsmall = MODE == kModeApprox && ! GENERAL && c.pair.L < kOmega3Zero;
        TrainState<MODE, GENERAL, true, PY> sa;
        TrainState<MODE, GENERAL, false, PY> sb;
        for (int n = 0; n < T; ++n)
        {
            const float yv = lsmall ? sa.step (c, __ldg (xr + n), __ldg (tr + n), n >= skip) : sb.step (c, __ldg (xr + n), __ldg (tr + n), n >= skip);
            if (yr != nullptr)
                yr[n] = yv;
            if ((n & (kSeg - 1)) == kSeg - 1)
            {
                sa.flush (acc);
                sb.flush (acc);
            }
        }
        sa.flush (acc);
        sb.flush (acc);
    }
    write_partials (acc, partials, blockIdx.x, lane);
}

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

// ---- dispatch over the template variants ------------------------------------------------------
template <class F>
cudaError_t dispatch_variant (const ClipVariant& v, F&& f)
{
#define DWDF_CASE(M, G, P) \
    if (v.mode == M && v.general == G && v.pyorder == P) \
        return f (std::integral_constant<int, M> {}, std::integral_constant<bool, G> {}, std::integral_constant<bool, P> {});
    DWDF_CASE (kModeApprox, false, false)
    DWDF_CASE (kModeApprox, false, true)
    DWDF_CASE (kModeApprox, true, false)
    DWDF_CASE (kModeApprox, true, true)
    DWDF_CASE (kModeExact, false, false)
    DWDF_CASE (kModeExact, false, true)
    DWDF_CASE (kModeExact, true, false)
    DWDF_CASE (kModeExact, true, true)
#undef DWDF_CASE
    return cudaErrorInvalidValue;
}
} // namespace

cudaError_t launch_clipper_forward (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream)
{
    const unsigned grid = (unsigned) ((B + kLanes - 1) / kLanes);
    return dispatch_variant (v, [&] (auto M, auto G, auto P) {
        if (use_tma)
            clipper_forward_tma<decltype (M)::value, decltype (G)::value, decltype (P)::value><<<grid, kLanes, 0, stream>>> (maps->x, maps->y, params, desc, ckpt, state, B, (int) T);
        else
            clipper_forward_direct<decltype (M)::value, decltype (G)::value, decltype (P)::value><<<grid, kLanes, 0, stream>>> (x, y, params, desc, ckpt, state, B, (int) T);
        return cudaGetLastError ();
    });
}

cudaError_t launch_clipper_adjoint (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* ckpt, const float* g, bool target, int64_t skip, float* gx, double* partials, int64_t B, int64_t T, cudaStream_t stream)
{
    const unsigned grid = (unsigned) ((B + kLanes - 1) / kLanes);
    const int sk = (int) (skip < 0 ? 0 : (skip > T ? T : skip));
    return dispatch_variant (v, [&] (auto M, auto G, auto P) {
        if (use_tma && gx == nullptr)
        {
            if (target)
                clipper_adjoint_tma<decltype (M)::value, decltype (G)::value, decltype (P)::value, true><<<grid, kLanes, 0, stream>>> (maps->x, maps->y, params, desc, ckpt, partials, B, (int) T, sk);
            else
                clipper_adjoint_tma<decltype (M)::value, decltype (G)::value, decltype (P)::value, false><<<grid, kLanes, 0, stream>>> (maps->x, maps->y, params, desc, ckpt, partials, B, (int) T, sk);
        }
        else if (gx != nullptr)
        {
            if (target)
                clipper_adjoint_direct<decltype (M)::value, decltype (G)::value, decltype (P)::value, true, true><<<grid, kLanes, 0, stream>>> (x, g, gx, params, desc, ckpt, partials, B, (int) T, sk);
            else
                clipper_adjoint_direct<decltype (M)::value, decltype (G)::value, decltype (P)::value, false, true><<<grid, kLanes, 0, stream>>> (x, g, gx, params, desc, ckpt, partials, B, (int) T, sk);
        }
        else
        {
            if (target)
                clipper_adjoint_direct<decltype (M)::value, decltype (G)::value, decltype (P)::value, true, false><<<grid, kLanes, 0, stream>>> (x, g, gx, params, desc, ckpt, partials, B, (int) T, sk);
            else
                clipper_adjoint_direct<decltype (M)::value, decltype (G)::value, decltype (P)::value, false, false><<<grid, kLanes, 0, stream>>> (x, g, gx, params, desc, ckpt, partials, B, (int) T, sk);
        }
        return cudaGetLastError ();
    });
}

cudaError_t launch_clipper_train (const ClipVariant& v, bool use_tma, const ClipTmaMaps* maps, const ClipDesc& desc, const float* params, const float* x, const float* target, int64_t skip, float* y, double* partials, int64_t B, int64_t T, cudaStream_t stream)
{
    const unsigned grid = (unsigned) ((B + kLanes - 1) / kLanes);
    const int sk = (int) (skip < 0 ? 0 : (skip > T ? T : skip));
    return dispatch_variant (v, [&] (auto M, auto G, auto P) {
        if (use_tma)
            clipper_train_tma<decltype (M)::value, decltype (G)::value, decltype (P)::value><<<grid, kLanes, 0, stream>>> (maps[0].x, maps[0].y, maps[1].y, y != nullptr ? 1 : 0, params, desc, partials, (int) T, sk);
        else
            clipper_train_direct<decltype (M)::value, decltype (G)::value, decltype (P)::value><<<grid, kLanes, 0, stream>>> (x, target, y, params, desc, partials, B, (int) T, sk);
        return cudaGetLastError ();
    });
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
