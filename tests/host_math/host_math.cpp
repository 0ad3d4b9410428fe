// TEST INFRASTRUCTURE ONLY. Compiles the product's device math header (csrc/dwdf_math.cuh) with
// g++ for the HOST so that `-m "not gpu"` tests can check its logic against the oracle on a
// machine without a GPU (MUFU/round-down intrinsics are replaced by their IEEE host equivalents,
// see the #if blocks in the header). Never shipped, never loaded by the product.
#include "dwdf_math.cuh"
#include <vector>

using namespace dwdf;

// kind: 0 omega3, 1 omega4, 2 exp_approx, 3 log_approx (x > 0), 4 omega_exact(2 iterations), 5 omega_exact(1 iteration)
extern "C" void hm_scalar (int kind, const float* x, float* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
    {
        switch (kind)
        {
            case 0: out[i] = omega3_approx (x[i]); break;
            case 1: out[i] = omega4_approx (x[i]); break;
            case 2: out[i] = exp_approx (x[i]); break;
            case 3: out[i] = log_approx_pos (x[i]); break;
            case 4: out[i] = omega_exact (x[i], 2, 0.0f); break;
            default: out[i] = omega_exact (x[i], 1, 0.0f); break;
        }
    }
}

template <int MODE, bool GENERAL, bool LSMALL>
static void pair_run (const PairConst& c, const float* a, float* b, float* dv, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
    {
        PairDeriv d;
        b[i] = pair_reflect<MODE, GENERAL, true, LSMALL> (c, a[i], &d);
        if (dv)
        {
            dv[3 * i] = d.S1;
            dv[3 * i + 1] = d.M1;
            dv[3 * i + 2] = d.dV;
        }
    }
}

// root law at a fixed port impedance; dv (optional) receives (S1, M1, dV) per sample
extern "C" void hm_pair (int mode, int general, int lsmall, float Rp, float Is, float Vt, float nabla, float n_up, float n_down, const float* a, float* b, float* dv, int64_t n)
{
    PairConst c;
    pair_setup (c, Rp, Is, Vt, nabla, n_up, n_down, 2, 0.0f);
    if (mode == kModeApproxGood)
    {
        for (int64_t i = 0; i < n; ++i)
            b[i] = pair_reflect<kModeApproxGood, false, false, false> (c, a[i], nullptr);
        return;
    }
    if (mode == kModeApprox)
    {
        if (general) pair_run<kModeApprox, true, false> (c, a, b, dv, n);
        else if (lsmall) pair_run<kModeApprox, false, true> (c, a, b, dv, n);
        else pair_run<kModeApprox, false, false> (c, a, b, dv, n);
    }
    else
    {
        if (general) pair_run<kModeExact, true, false> (c, a, b, dv, n);
        else pair_run<kModeExact, false, false> (c, a, b, dv, n);
    }
}

template <int MODE, bool GENERAL, bool LSMALL, bool PY>
static void clip_run (const ClipConst& c, const float* x, const float* g, float* y, double* acc, int64_t B, int64_t T)
{
    std::vector<StepTape> tape ((size_t) T);
    for (int64_t s = 0; s < B; ++s)
    {
        float z = 0.0f;
        for (int64_t n = 0; n < T; ++n)
        {
            if (g)
                y[s * T + n] = clip_step_tape<MODE, GENERAL, LSMALL, PY> (c, x[s * T + n], z, tape[(size_t) n]);
            else
                y[s * T + n] = clip_step<MODE, GENERAL, LSMALL, PY> (c, x[s * T + n], z);
        }
        if (! g)
            continue;
        double G = 0.0; // adjoint of z[n+1]
        for (int64_t n = T - 1; n >= 0; --n)
        {
            const StepTape& tp = tape[(size_t) n];
            const double gy = g[s * T + n];
            if (PY) G += 0.5 * gy;
            acc[0] += G * tp.cg;
            acc[1] += G * tp.cl;
            acc[2] += G * tp.cv;
            G = (PY ? 0.5 * gy : gy) + G * tp.A;
        }
    }
}

// the clipper recurrence on the host; g == NULL: forward only, else also the raw adjoint sums
// acc[0..2] = sum G dz'/d{gamma, ell, V} for upstream gradient g
extern "C" void hm_clipper (int mode, int general, int pyorder, float fs, float R, float C, float Is, float Vt, float nabla, float n_up, float n_down, const float* x, const float* g, float* y, double* acc, int64_t B, int64_t T)
{
    ClipDesc d { fs, Vt, n_up, n_down, 0.0f, 2, 0, 1, 2, 3 };
    ClipConst c;
    clip_setup (c, d, R, C, Is, nabla);
    const bool lsmall = mode == kModeApprox && ! general && lsmall_ok (c.pair.L);
#define RUN(M, G, L) (pyorder ? clip_run<M, G, L, true> (c, x, g, y, acc, B, T) : clip_run<M, G, L, false> (c, x, g, y, acc, B, T))
    if (mode == kModeApprox)
    {
        if (general) RUN (kModeApprox, true, false);
        else if (lsmall) RUN (kModeApprox, false, true);
        else RUN (kModeApprox, false, false);
    }
    else
    {
        if (general) RUN (kModeExact, true, false);
        else RUN (kModeExact, false, false);
    }
#undef RUN
}

// The adjoint's way (clipper_kernels.cu adjoint_segment): forward once for y, then a reverse sweep that
// recovers the states from y (re-anchored every 16 samples at a checkpoint) and each step's
// linearisation with clip_step_recover. acc as in hm_clipper.
// FROMY: the hot variants' step, clip_step_recover_yv — the linearisation from (y, z) alone, x never read.
template <int MODE, bool GENERAL, bool LSMALL, bool PY, bool FROMY = false>
static void recover_run (const ClipConst& c, const float* x, const float* g, float* y, double* acc, int64_t B, int64_t T)
{
    constexpr int SEG = 16;
    const float inv_gamma = 1.0f / c.gamma;
    std::vector<float> ck ((size_t) ((T + SEG - 1) / SEG)), zs ((size_t) T + 1);
    for (int64_t s = 0; s < B; ++s)
    {
        float z = 0.0f;
        for (int64_t n = 0; n < T; ++n)
        {
            if (n % SEG == 0)
                ck[(size_t) (n / SEG)] = z;
            y[s * T + n] = clip_step<MODE, GENERAL, LSMALL, PY> (c, x[s * T + n], z);
        }
        for (int64_t n = 0; n < T; ++n) // states from the output only
        {
            if (n % SEG == 0)
                zs[(size_t) n] = ck[(size_t) (n / SEG)];
            if (PY)
                zs[(size_t) n + 1] = std::fmaf (2.0f, y[s * T + n], -zs[(size_t) n]);
            else
                zs[(size_t) n] = y[s * T + n];
        }
        // python ordering: the segment's last z[n+1] is overwritten by the next checkpoint in the loop above
        // (n % SEG == 0), exactly like the kernel, which starts every segment from its checkpoint
        double G = 0.0;
        for (int64_t n = T - 1; n >= 0; --n)
        {
            const double gy = g[s * T + n];
            if (! PY && n == T - 1)
            {
                G = gy;
                continue;
            }
            float zn = zs[(size_t) n + 1];
            if (PY && (n + 1) % SEG == 0)
                zn = std::fmaf (2.0f, y[s * T + n], -zs[(size_t) n]); // within-segment reconstruction, not the next checkpoint
            StepTape tp;
            if (FROMY)
            {
                StepTapeY<f1> ty;
                const float v = PY ? y[s * T + n] : 0.5f * (zs[(size_t) n] + zn);
                clip_step_recover_yv<f1> (c, f1 { v }, f1 { zs[(size_t) n] }, ty);
                tp.A = ty.A.x;
                tape_scale (c, inv_gamma, ty.cg.x, ty.m1.x, ty.as.x, ty.ww.x, tp.cg, tp.cl, tp.cv);
            }
            else
                clip_step_recover<MODE, GENERAL, LSMALL> (c, x[s * T + n], zs[(size_t) n], zn, tp);
            if (PY) G += 0.5 * gy;
            acc[0] += G * tp.cg;
            acc[1] += G * tp.cl;
            acc[2] += G * tp.cv;
            G = (PY ? 0.5 * gy : gy) + G * tp.A;
        }
    }
}

extern "C" void hm_clipper_recover (int mode, int general, int pyorder, float fs, float R, float C, float Is, float Vt, float nabla, float n_up, float n_down, const float* x, const float* g, float* y, double* acc, int64_t B, int64_t T)
{
    ClipDesc d { fs, Vt, n_up, n_down, 0.0f, 2, 0, 1, 2, 3 };
    ClipConst c;
    clip_setup (c, d, R, C, Is, nabla);
    const bool lsmall = mode == kModeApprox && ! general && lsmall_ok (c.pair.L);
#define RUN(M, G, L) (pyorder ? recover_run<M, G, L, true> (c, x, g, y, acc, B, T) : recover_run<M, G, L, false> (c, x, g, y, acc, B, T))
    if (mode == kModeApprox)
    {
        if (general) RUN (kModeApprox, true, false);
        else if (lsmall) RUN (kModeApprox, false, true);
        else RUN (kModeApprox, false, false);
    }
    else
    {
        if (general) RUN (kModeExact, true, false);
        else RUN (kModeExact, false, false);
    }
#undef RUN
}

// symmetric pair, fromy_ok parameters only (returns 1 otherwise): the reverse sweep that never reads x
extern "C" int hm_clipper_recover_y (int mode, int pyorder, float fs, float R, float C, float Is, float Vt, float nabla, const float* x, const float* g, float* y, double* acc, int64_t B, int64_t T)
{
    ClipDesc d { fs, Vt, 1.0f, 1.0f, 0.0f, 1, 0, 1, 2, 3 };
    ClipConst c;
    clip_setup (c, d, R, C, Is, nabla);
    if (! fromy_ok (c.pair))
        return 1;
    if (mode == kModeApprox)
        pyorder ? recover_run<kModeApprox, false, true, true, true> (c, x, g, y, acc, B, T) : recover_run<kModeApprox, false, true, false, true> (c, x, g, y, acc, B, T);
    else
        pyorder ? recover_run<kModeExact, false, true, true, true> (c, x, g, y, acc, B, T) : recover_run<kModeExact, false, true, false, true> (c, x, g, y, acc, B, T);
    return 0;
}

// The forward kernels' fast path exactly as clipper_kernels.cu arranges it: clip_chunk_fastv on 4-sample
// chunks, V = f1 (one sequence) or f2 (two per lane: rows s, s + 1), the instances that crossed omega3's
// log branch redone with the general step; a tail of T % 4 samples the general way (direct kernel).
// fallbacks (optional) counts the redone chunk-instances.
template <class V, bool PY>
static void fastv_run (const ClipConst& c, const float* x, float* y, int64_t* fallbacks, int64_t B, int64_t T)
{
    constexpr int W = sizeof (V) / sizeof (float);
    int64_t mismatches = 0;
    for (int64_t s = 0; s < B; s += W)
    {
        int64_t row[2] = { s, s + 1 < B ? s + 1 : s };
        float z[2] = { 0.0f, 0.0f };
        bool hint = false;
        int64_t n = 0;
        for (; n + 4 <= T; n += 4)
        {
            V xv[4], ov[4], zz, um;
            float* zp = (float*) &zz;
            float* up = (float*) &um;
            for (int w = 0; w < W; ++w)
                zp[w] = z[w], up[w] = -1.0e30f;
            for (int k = 0; k < 4; ++k)
                for (int w = 0; w < W; ++w)
                    ((float*) &xv[k])[w] = x[row[w] * T + n + k];
            // the kernels' scheme (clipper_kernels.cu forward_chunk / forward_chunk2): the plain fast step, and a chunk with a
            // loud instance again with the LOUD step (packed, both instances); a chunk after a loud one goes
            // straight to the LOUD step. An instance below the branch must get the same bits from either: checked here on
            // every chunk (mismatches are reported through *fallbacks as a negative count).
            bool ahead = hint, loud = false; // hint: the previous chunk of these rows was loud
            V zf = zz, of[4];
            clip_chunk_fastv<V, PY> (c, xv, zf, of, um);
            for (int w = 0; w < W; ++w)
                loud = loud || up[w] >= kFastLoud;
            V zl = zz, ol[4], uml;
            for (int w = 0; w < W; ++w)
                ((float*) &uml)[w] = -1.0e30f;
            clip_chunk_loudv<V, PY> (c, xv, zl, ol, uml); // (always evaluated here, for the bit check)
            for (int w = 0; w < W; ++w)
            {
                const bool inst_loud = up[w] >= kFastLoud;
                if (! inst_loud)
                    for (int k = 0; k < 4; ++k)
                        if (std::memcmp (&((float*) &of[k])[w], &((float*) &ol[k])[w], 4) != 0 || std::memcmp (&((float*) &zf)[w], &((float*) &zl)[w], 4) != 0)
                            ++mismatches;
                const bool use_loud = ahead || loud;
                z[w] = use_loud ? ((float*) &zl)[w] : ((float*) &zf)[w];
                for (int k = 0; k < 4; ++k)
                    y[row[w] * T + n + k] = use_loud ? ((float*) &ol[k])[w] : ((float*) &of[k])[w];
                if (inst_loud && fallbacks)
                    ++*fallbacks;
            }
            hint = loud;
        }
        for (; n < T; ++n)
            for (int w = 0; w < W; ++w)
                y[row[w] * T + n] = clip_step<kModeApprox, false, false, PY> (c, x[row[w] * T + n], z[w]);
    }
    if (mismatches != 0 && fallbacks)
        *fallbacks = -mismatches;
}

// pairs = 0: one sequence per lane (f1), 1: two (f2, packed fp32x2 on the device)
extern "C" int hm_clipper_fast (int pairs, int pyorder, float fs, float R, float C, float Is, float Vt, float nabla, const float* x, float* y, int64_t* fallbacks, int64_t B, int64_t T)
{
    ClipDesc d { fs, Vt, 1.0f, 1.0f, 0.0f, 2, 0, 1, 2, 3 };
    ClipConst c;
    clip_setup (c, d, R, C, Is, nabla);
    if (! fast_ok (c.pair.L))
        return 1;
    if (fallbacks)
        *fallbacks = 0;
    if (pairs)
        pyorder ? fastv_run<f2, true> (c, x, y, fallbacks, B, T) : fastv_run<f2, false> (c, x, y, fallbacks, B, T);
    else
        pyorder ? fastv_run<f1, true> (c, x, y, fallbacks, B, T) : fastv_run<f1, false> (c, x, y, fallbacks, B, T);
    return 0;
}

// clip_step_exactv — the exact root's forward sample (V = f2: packed pairs, V = f1: one sequence) — over whole sequences
template <class V, bool PY>
static void exactv_run (const ClipConst& c, const float* x, float* y, int64_t B, int64_t T)
{
    constexpr int W = sizeof (V) / sizeof (float);
    for (int64_t s = 0; s < B; s += W)
    {
        const int64_t row[2] = { s, s + 1 < B ? s + 1 : s };
        V z;
        for (int w = 0; w < W; ++w)
            ((float*) &z)[w] = 0.0f;
        for (int64_t n = 0; n < T; ++n)
        {
            V xv;
            for (int w = 0; w < W; ++w)
                ((float*) &xv)[w] = x[row[w] * T + n];
            const V o = clip_step_exactv<V, PY> (c, xv, z);
            for (int w = 0; w < W; ++w)
                y[row[w] * T + n] = ((const float*) &o)[w];
        }
    }
}
extern "C" int hm_clipper_exactv (int pairs, int pyorder, float fs, float R, float C, float Is, float Vt, float nabla, const float* x, float* y, int64_t B, int64_t T)
{
    ClipDesc d { fs, Vt, 1.0f, 1.0f, 0.0f, 1, 0, 1, 2, 3 };
    ClipConst c;
    clip_setup (c, d, R, C, Is, nabla);
    if (! exact_fast_ok (c.pair))
        return 1;
    if (pairs)
        pyorder ? exactv_run<f2, true> (c, x, y, B, T) : exactv_run<f2, false> (c, x, y, B, T);
    else
        pyorder ? exactv_run<f1, true> (c, x, y, B, T) : exactv_run<f1, false> (c, x, y, B, T);
    return 0;
}

// clip_step_recoverv<f2> (the adjoint kernel's packed pair step) against the scalar clip_step_recover on
// the same (x, z, z') triples: out[8 i ...] = { A, cg, cl, cv } packed, then scalar.
template <int MODE>
static void recover_pairs_run (const ClipConst& c, const float* x, const float* z, const float* zn, float* out, int64_t n)
{
    for (int64_t i = 0; i + 1 < n; i += 2)
    {
        StepTapeY<f2> tp; // unscaled: tape_scale applies the constant factors (the sweep does it once per segment)
        clip_step_recoverv<f2, MODE> (c, f2 { x[i], x[i + 1] }, f2 { z[i], z[i + 1] }, f2 { zn[i], zn[i + 1] }, tp);
        float pk[2][4] = { { tp.A.x, 0, 0, 0 }, { tp.A.y, 0, 0, 0 } };
        tape_scale (c, 1.0f, tp.cg.x, tp.m1.x, tp.as.x, tp.ww.x, pk[0][1], pk[0][2], pk[0][3]);
        tape_scale (c, 1.0f, tp.cg.y, tp.m1.y, tp.as.y, tp.ww.y, pk[1][1], pk[1][2], pk[1][3]);
        for (int k = 0; k < 2; ++k)
        {
            StepTape ts;
            clip_step_recover<MODE, false, true> (c, x[i + k], z[i + k], zn[i + k], ts);
            const float sc[4] = { ts.A, ts.cg, ts.cl, ts.cv };
            for (int j = 0; j < 4; ++j)
            {
                out[8 * (i + k) + j] = pk[k][j];
                out[8 * (i + k) + 4 + j] = sc[j];
            }
        }
    }
}

extern "C" int hm_recover_pairs_mode (int exact, float fs, float R, float C, float Is, float Vt, float nabla, const float* x, const float* z, const float* zn, float* out, int64_t n)
{
    ClipDesc d { fs, Vt, 1.0f, 1.0f, 0.0f, 2, 0, 1, 2, 3 };
    ClipConst c;
    clip_setup (c, d, R, C, Is, nabla);
    if (! rev_small_ok (c.pair))
        return 1;
    exact ? recover_pairs_run<kModeExact> (c, x, z, zn, out, n) : recover_pairs_run<kModeApprox> (c, x, z, zn, out, n);
    return 0;
}

extern "C" int hm_recover_pairs (float fs, float R, float C, float Is, float Vt, float nabla, const float* x, const float* z, const float* zn, float* out, int64_t n)
{
    ClipDesc d { fs, Vt, 1.0f, 1.0f, 0.0f, 2, 0, 1, 2, 3 };
    ClipConst c;
    clip_setup (c, d, R, C, Is, nabla);
    if (! lsmall_ok (c.pair.L))
        return 1;
    for (int64_t i = 0; i + 1 < n; i += 2)
    {
        StepTapeY<f2> tp;
        clip_step_recoverv<f2> (c, f2 { x[i], x[i + 1] }, f2 { z[i], z[i + 1] }, f2 { zn[i], zn[i + 1] }, tp);
        float pk[2][4] = { { tp.A.x, 0, 0, 0 }, { tp.A.y, 0, 0, 0 } };
        tape_scale (c, 1.0f, tp.cg.x, tp.m1.x, tp.as.x, tp.ww.x, pk[0][1], pk[0][2], pk[0][3]);
        tape_scale (c, 1.0f, tp.cg.y, tp.m1.y, tp.as.y, tp.ww.y, pk[1][1], pk[1][2], pk[1][3]);
        for (int k = 0; k < 2; ++k)
        {
            StepTape ts;
            clip_step_recover<kModeApprox, false, true> (c, x[i + k], z[i + k], zn[i + k], ts);
            const float sc[4] = { ts.A, ts.cg, ts.cl, ts.cv };
            for (int j = 0; j < 4; ++j)
            {
                out[8 * (i + k) + j] = pk[k][j];
                out[8 * (i + k) + 4 + j] = sc[j];
            }
        }
    }
    return 0;
}
