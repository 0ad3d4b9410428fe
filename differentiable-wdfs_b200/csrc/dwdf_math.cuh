// dwdf_math.cuh — scalar arithmetic of the WDF hot path, written for the sm_100a FMA/ALU pipes.
//
// What is computed follows the reference (cited per function); how it is computed does not:
// floor/ldexp are done with round-down magic-number adds and integer adds on the exponent field
// instead of float<->int conversions (the conversion/MUFU pipe issues at 1/8 the FMA rate),
// divisions are MUFU.RCP + FMUL, every select is branch-free, polynomials are FMA Horner chains.
//
// The functions are __host__ __device__ so that tests/ can compile the same source with g++ and
// check the logic against the oracle on a machine without a GPU; the product only ever calls them
// from kernels.
#pragma once
#if defined(__CUDACC_RTC__) // runtime-specialised tree kernels (tree_jit.cu): NVRTC has no host headers
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#ifndef INFINITY
#define INFINITY __int_as_float (0x7f800000)
#endif
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#endif

#if defined(__CUDACC__)
#define DWDF_HD __host__ __device__ __forceinline__
#else
#define DWDF_HD inline
#endif

namespace dwdf
{

// ---- bit casts / primitives -----------------------------------------------------------------
DWDF_HD float i2f (int32_t i)
{
#if defined(__CUDA_ARCH__)
    return __int_as_float (i);
#else
    float f;
    std::memcpy (&f, &i, 4);
    return f;
#endif
}
DWDF_HD int32_t f2i (float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int (f);
#else
    int32_t i;
    std::memcpy (&i, &f, 4);
    return i;
#endif
}
// reciprocal: one MUFU.RCP (<= 1 ulp) on the device
DWDF_HD float rcp (float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
// x + y rounded toward -inf (FADD.RM)
DWDF_HD float add_rd (float x, float y)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rd (x, y);
#else
    double s = (double) x + (double) y; // exact in double for the magnitudes used here
    float f = (float) s;
    if ((double) f > s)
        f = std::nextafterf (f, -INFINITY);
    return f;
#endif
}
// products / sums the compiler must not contract into a neighbouring FMA (where two evaluations of the
// same function have to agree bit for bit whatever code surrounds them)
DWDF_HD float mul_ (float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn (a, b);
#else
    return a * b;
#endif
}
DWDF_HD float add_ (float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn (a, b);
#else
    return a + b;
#endif
}
DWDF_HD float fma_ (float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn (a, b, c);
#else
    return std::fmaf (a, b, c);
#endif
}

// ---- omega.h: the "approx" Wright-omega ----------------------------------------------------
// exp_approx, omega.h:99-116 (with pow2_approx :83-92): 2^floor(x') * cubic(x' - floor(x')),
// x' = max(-126, x * log2(e)).
//   floor: x' + 1.5*2^23 rounded DOWN leaves floor(x') in the low mantissa bits (one FADD.RM),
//   ldexp: (bits << 23) added to the polynomial's exponent field (the magic constant's own bits
//          shift out).
// Differs from the reference only at negative integer x', where omega.h's truncation quirk picks
// (l = x'-1, f = 1) and this picks (l = x', f = 0): both are 2^x' to 1.3e-5 (cubic end-point error).
template <bool CLAMP = true>
DWDF_HD float exp_approx_scaled (float xp) // takes x' = x * log2(e); CLAMP = false: the caller guarantees x' >= -126
{
    const float kMagic = 12582912.0f; // 1.5 * 2^23
    if (CLAMP)
        xp = fmaxf (xp, -126.0f);
    const float t = add_rd (xp, kMagic);
    const float l = t - kMagic;
    const float f = xp - l;
    const float p = fma_ (f, fma_ (f, fma_ (f, 0.07944154167983575f, 0.2274112777602189f), 0.6931471805599453f), 1.0f);
    return i2f (f2i (p) + (int32_t) ((uint32_t) f2i (t) << 23));
}
DWDF_HD float exp_approx (float x) { return exp_approx_scaled<true> (1.442695040888963f * x); }

// a with the sign bit of s XOR-ed in: a * sign(s) for s != 0 (one LOP3)
DWDF_HD float xor_sign (float a, float s) { return i2f (f2i (a) ^ (f2i (s) & (int32_t) 0x80000000)); }

// log_approx, omega.h:49-63 (with log2_approx :33-42), for x > 0: ln2 * (exponent + cubic(mantissa)).
//   exponent as float without I2F: (bits >> 23) OR'ed into the mantissa of 2^23, minus (2^23 + 127).
DWDF_HD float log_approx_pos (float x)
{
    const int32_t bits = f2i (x);
    const float m = i2f ((bits & 0x007fffff) | 0x3f800000);
    const float e = i2f ((bits >> 23) | 0x4b000000) - 8388735.0f;
    const float p = fma_ (m, fma_ (m, fma_ (m, 0.1640425613334452f, -1.098865286222744f), 3.148297929334117f), -2.213475204444817f);
    return 0.693147180559945f * (e + p);
}

// omega3, omega.h:159-169.
// WARP (device only): the x >= 8 branch (x - log_approx(x), ~11 instructions) is skipped when no lane of
// the warp that is active here needs it — one vote instead of computing both sides of the select for
// every sample. Same value either way (the vote only gates work; a lane that needs the branch always
// sees its own vote), so the call is legal under divergence.
template <bool WARP = false>
DWDF_HD float omega3_approx (float x)
{
    float y = fma_ (x, fma_ (x, fma_ (x, -1.314293149877800e-3f, 4.775931364975583e-2f), 3.631952663804445e-1f), 6.313183464296682e-1f);
#if defined(__CUDA_ARCH__)
    if (! WARP || __any_sync (__activemask (), x >= 8.0f))
#endif
        y = x < 8.0f ? y : x - log_approx_pos (x);
    y = x < -3.341459552768620f ? 0.0f : y;
    return y;
}

// omega4, omega.h:172-177: omega3 + one Newton step on  w - exp(x - w).
// FAST: WARP as above, and the caller guarantees x * log2(e) >= -126 (no clamp in exp_approx).
template <bool FAST = false>
DWDF_HD float omega4_approx (float x)
{
    const float y = omega3_approx<FAST> (x);
    const float e = exp_approx_scaled<! FAST> (1.442695040888963f * (x - y));
    return y - (y - e) * rcp (y + 1.0f);
}

constexpr float kOmega3Zero = -3.341459552768620f; // below this omega3 == 0 and omega4(x) == exp_approx(x)
constexpr float kLog2e = 1.442695040888963f;
// The clipper kernels' fast path ("LSMALL") applies when  -87 < L < kOmega3Zero  (every physical
// diode: Rp*Is << V, and Rp*Is/V a normal float): the reverse-biased omega4(L - |a|/V) is exactly
// exp_approx(L - |a|/V), and the forward-biased argument never needs exp_approx's -126 clamp.
DWDF_HD bool lsmall_ok (float L) { return L < kOmega3Zero && L > -87.0f; }

// ---- "exact" Wright-omega in fp32 ---------------------------------------------------------------
// Replaces Toms917DiodePair.h:64-67 (float -> complex<double> TOMS-917 -> float) and
// scipy.special.wrightomega (diode_pretraining.py:57-58). Same published algorithm (Lawrence,
// Corless & Jeffrey, ACM TOMS 917) restricted to the real axis, where only three of its regions are
// reachable (toms917.cpp:240-261,290-296): a series start, then Fritsch-Shafer-Crowley iterations
// (toms917.cpp:345-364). Arranged for a SIMT machine whose lanes sit in different regions at the same
// time: all three starts are evaluated branch-free (MUFU.EX2 / MUFU.LG2 / MUFU.RCP instead of libm
// calls) and selected, then one common FSC iteration.
//   x <= -2      w = e^x s with s the root of  s = exp(-e^x s)  — the same equation divided by e^x, all
//                terms O(1): the residual x - w - ln(w) of the original form cancels catastrophically in
//                fp32 here. The series in e^x (toms917.cpp:240-248) as a degree-6 fit of W(E)/E; no refinement needed.
//   -2 < x <= 1+pi  series about x = 1 (toms917.cpp:253-261), one FSC iteration
//   else            asymptotic series in ln(x) (toms917.cpp:290-296), one FSC iteration
// Measured against scipy.special.wrightomega on [-60, 300]: <= 3.5e-7 relative after ONE iteration
// (fp32 round-off), so n_iter defaults to 1; TOMS-917's second iteration is conditional on a
// double-precision residual test that cannot fire at this accuracy. n_iter > 1 and tol ("Newton
// tolerance": stop when |residual| <= tol) bound further refinement.
DWDF_HD float ex2_ (float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return exp2f (x);
#endif
}
DWDF_HD float lg2_ (float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return log2f (x);
#endif
}
// e^x for x <= 0 to fp32 round-off: x log2(e) as a compensated product (the rounding of a single product
// alone costs |x| * 4e-8 relative)
DWDF_HD float exp_nonpos (float x)
{
    const float kHi = 1.44269502162933349609375f, kLo = 1.925963033500011e-8f; // log2(e) = kHi + kLo
    const float t = mul_ (x, kHi);
    const float lo = fma_ (x, kLo, fma_ (x, kHi, -t));
    const float e0 = ex2_ (t);
    return fma_ (e0, mul_ (0.693147180559945f, lo), e0);
}
DWDF_HD float ln_ (float x) { return 0.693147180559945f * lg2_ (x); }

DWDF_HD float fsc_step (float w, float r)
{
    const float wp1 = w + 1.0f;
    const float q = 2.0f * wp1 * fma_ (0.66666666666666667f, r, wp1);
    const float e = (r * (q - r)) * rcp (wp1 * fma_ (-2.0f, r, q));
    return fma_ (w, e, w);
}

// omega(x) for x <= -2 (any x: the value is only meaningful there): w = E s(E), E = e^x <= e^-2, s = W(E) / E the root of
// s = exp(-E s). TOMS-917 starts from the series 1 - E + 3/2 E^2 - 8/3 E^3 + 125/24 E^4 (toms917.cpp:240-248) and refines; here s
// is its degree-6 Chebyshev fit on [0, e^-2] (1.9e-9 from W(E)/E, constant term exactly 1) and nothing is refined: six FMAs
// and the one MUFU.EX2 of E, where series + Newton step took two more MUFU ops per evaluation (the exact root's forward
// pass is bound by the MUFU pipe: 59 % busy at 10 per sample, profiles/r01_e_ncu_exact_packed_summary.txt).
constexpr float kOmLow1 = -0.999998602f, kOmLow2 = 1.49982983f, kOmLow3 = -2.65880158f, kOmLow4 = 5.03137976f, kOmLow5 = -8.66441333f, kOmLow6 = 9.14379624f;
DWDF_HD float omega_exact_low (float x)
{
    // (explicit mul_/fma_: the forward-biased and the reverse-biased branch evaluate this at the same
    //  argument when a == 0 and must then cancel exactly — silence in, silence out)
    const float E = exp_nonpos (fminf (x, 0.0f));
    const float Ec = fminf (E, 0.1353352832f);
    const float s = fma_ (Ec, fma_ (Ec, fma_ (Ec, fma_ (Ec, fma_ (Ec, fma_ (Ec, kOmLow6, kOmLow5), kOmLow4), kOmLow3), kOmLow2), kOmLow1), 1.0f);
    return mul_ (E, s);
}

DWDF_HD float omega_exact (float x, int n_iter, float tol)
{
    const float w_lo = omega_exact_low (x);
    // -2 < x <= 1 + pi
    const float p = x - 1.0f;
    const float ser = fma_ (p, fma_ (p, fma_ (p, 2.1158854166666667e-4f, -3.2552083333333333e-4f), -5.2083333333333333e-3f), 0.0625f);
    const float w_mid = fma_ (p * p, ser, fma_ (0.5f, x, 0.5f));
    // x > 1 + pi
    const float xc = fmaxf (x, 1.0f);
    const float l = ln_ (xc);
    const float it = rcp (xc);
    const float li = l * it;
    const float w_hi = (xc - l) + li * (1.0f + it * (fma_ (0.5f, l, -1.0f) + it * fma_ (l, fma_ (l, 0.33333333333333333f, -1.5f), 1.0f)));
    float w = fmaxf (x <= 4.141592653589793f ? w_mid : w_hi, 0.05f); // (the floor only ever acts on lanes that end up taking w_lo)
    w = fsc_step (w, (x - w) - ln_ (w));
    for (int k = 1; k < n_iter; ++k)
    {
        const float r = (x - w) - ln_ (w);
        if (fabsf (r) <= tol)
            break;
        w = fsc_step (w, r);
    }
    return x <= -2.0f ? w_lo : w;
}

// ---- diode-pair root ---------------------------------------------------------------------------
enum : int
{
    kModeApprox = 0,
    kModeExact = 1,
    kModeApproxGood = 2
};

// Constants of one root at one port impedance: wdf_t.h:875-882 (setDiodeParameters) and :928-933
// (calcImpedanceInternal); Toms917DiodePair.h:28-42 is identical. The general law adds the
// per-branch logs of diode_pretraining.py:49-50.
struct PairConst
{
    float V, twoV, invV; // V = nDiodes * Vt
    float L; // ln(Rp * Is / V)
    float Ll2e, invVl2e; // L * log2(e), log2(e) / V: the reverse-biased exp_approx argument in one FMA
    float inv2V; // 1 / (2 V)
    float RIs, RIs_overV; // "Good" law only
    float n_up, n_dn, L_up, L_dn, inv_up, inv_dn; // general law: mu, ln(Rp Is / (V mu)), 1 / (mu V)
    float rn_up, rn_dn; // 1 / mu
    int n_iter;
    float tol;
};

// ln of a launch constant, correctly rounded (through double on the device: CUDA's logf may miss by an ulp). The N_up != N_down
// law uses TWO such constants, one per branch; an ulp of imbalance between them (5e-7 of L) is a constant offset of the
// reflected wave — every sample, the same sign — which a circuit with a memory of thousands of samples (gamma ~ 1e-4)
// integrates into 1.1 … 1.3e-5 of the output's peak (3 of 8000 random circuits of the end-of-round-2 sweep; the same source
// built for the host, with libm's logf, agrees with the reference to 3e-7). Once per thread and launch, two-constant law only.
DWDF_HD float log_setup (float x)
{
#if defined(__CUDA_ARCH__)
    return (float) log ((double) x);
#else
    return logf (x);
#endif
}

DWDF_HD void pair_setup (PairConst& c, float Rp, float Is, float Vt, float nabla, float n_up, float n_down, int n_iter, float tol)
{
    c.V = nabla * Vt;
    c.twoV = 2.0f * c.V;
    c.invV = 1.0f / c.V;
    c.RIs = Rp * Is;
    c.RIs_overV = c.RIs * c.invV;
    // (a symmetric pair uses ONE constant for both branches, its last ulp cancels: the hot kernels keep logf)
    const bool two_constants = n_up != n_down;
    c.L = two_constants ? log_setup (c.RIs_overV) : logf (c.RIs_overV);
    c.Ll2e = kLog2e * c.L;
    c.invVl2e = kLog2e * c.invV;
    c.inv2V = 0.5f * c.invV;
    c.n_up = n_up;
    c.n_dn = n_down;
    c.L_up = two_constants ? log_setup (c.RIs_overV / n_up) : logf (c.RIs_overV / n_up);
    c.L_dn = two_constants ? log_setup (c.RIs_overV / n_down) : logf (c.RIs_overV / n_down);
    c.inv_up = 1.0f / (n_up * c.V);
    c.inv_dn = 1.0f / (n_down * c.V);
    c.rn_up = 1.0f / n_up;
    c.rn_dn = 1.0f / n_down;
    c.n_iter = n_iter <= 0 ? 1 : n_iter;
    c.tol = tol;
}

template <int MODE, bool FAST = false>
DWDF_HD float root_omega (const PairConst& c, float u)
{
    if (MODE == kModeExact)
        return omega_exact (u, c.n_iter, c.tol);
    return omega4_approx<FAST> (u);
}

// The reverse-biased branch, omega(ln(Rp Is / (mu V)) - |a| / (mu V)). REVSMALL: the caller guarantees that
// every such argument stays below kOmega3Zero (rev_small_ok: true for every physical diode, Rp*Is << V);
// then omega4 is exactly exp_approx (omega3 == 0 there) and the exact omega needs only its x <= -2 region.
template <int MODE, bool REVSMALL>
DWDF_HD float root_omega_rev (const PairConst& c, float u)
{
    if (REVSMALL)
        return MODE == kModeExact ? omega_exact_low (u) : exp_approx (u);
    return root_omega<MODE> (c, u);
}
DWDF_HD bool rev_small_ok (const PairConst& c) { return lsmall_ok (c.L) && lsmall_ok (c.L_up) && lsmall_ok (c.L_dn); }

// Derivative pieces of b = f(a; ell, V) with ell = ln(Rp Is), omega' = omega / (1 + omega):
//   S1 = w0' + w1'                       df/da      = 1 - 2 S1
//   M1 = lambda (mu0 w0' - mu1 w1')      df/d ell   = -2 V M1
//   dV = (b - a)/V + 2 M1 + 2 (a/V) S1   df/dV at fixed ell
struct PairDeriv
{
    float S1, M1, dV;
};

// (S1, M1, dV) from the two omegas of one sample.
// dV is written without cancellation: with w - w' = w w' the textbook form (b - a)/V + 2 M1 + 2 (a/V) S1 equals
//   -2 lambda (mu0 w0 w0' - mu1 w1 w1') + 2 (a/V) S1,
// and while the diodes are off (w << 1) its first two terms cancel to second order in w — far below fp32's
// resolution of b - a (dL/dnF of a quiet signal was off by tens of percent; found by tests/test_gpu_fuzz.py).
template <bool GENERAL>
DWDF_HD void pair_deriv (const PairConst& c, float a, float b, float w0, float w1, float mu0, float mu1, PairDeriv* d)
{
    (void) b;
    const float wp0 = w0 * rcp (1.0f + w0), wp1 = w1 * rcp (1.0f + w1);
    d->S1 = wp0 + wp1;
    float ww;
    if (GENERAL)
    {
        const float lam = a == 0.0f ? 0.0f : copysignf (1.0f, a);
        d->M1 = lam * fma_ (mu0, wp0, -(mu1 * wp1));
        ww = lam * fma_ (mu0 * w0, wp0, -(mu1 * w1 * wp1));
    }
    else
    {
        d->M1 = xor_sign (wp0 - wp1, a); // wp0 == wp1 at a == 0
        ww = xor_sign (fma_ (w0, wp0, -(w1 * wp1)), a);
    }
    d->dV = fma_ (2.0f * a * c.invV, d->S1, -2.0f * ww);
}

// b = f(a).  Symmetric (eq. 39): wdf_t.h:917-924 == Toms917DiodePair.h:51-59.
//            General   (eq. 45): diode_pretraining.py:39-60 with N_up / N_down.
//            Good      (eq. 18): wdf_t.h:907-913.
// LSMALL: the caller guarantees rev_small_ok (reverse-biased arguments below kOmega3Zero) and, for the
// approx symmetric law, a converged warp on the device (omega3's log branch is skipped by a warp vote).
template <int MODE, bool GENERAL, bool DERIV, bool LSMALL>
DWDF_HD float pair_reflect (const PairConst& c, float a, PairDeriv* d)
{
    const float aa = fabsf (a);
    float w0, w1, mu0 = 1.0f, mu1 = 1.0f, b;
    if (MODE == kModeApproxGood)
    {
        w0 = omega4_approx (c.L + aa * c.invV + c.RIs_overV);
        const float lam = (0.0f < a) - (a < 0.0f);
        return a + 2.0f * lam * (c.RIs - c.V * w0);
    }
    if (GENERAL)
    {
        const bool pos = a >= 0.0f; // diode_pretraining.py:46-47
        mu0 = pos ? c.n_dn : c.n_up;
        mu1 = pos ? c.n_up : c.n_dn;
        const float l0 = pos ? c.L_dn : c.L_up, l1 = pos ? c.L_up : c.L_dn;
        const float q0 = aa * (pos ? c.inv_dn : c.inv_up), q1 = aa * (pos ? c.inv_up : c.inv_dn);
        w0 = root_omega<MODE> (c, l0 + q0);
        w1 = root_omega_rev<MODE, LSMALL> (c, l1 - q1);
        const float s = a == 0.0f ? 0.0f : copysignf (c.twoV, a); // 2 V lambda, lambda = signum(a) (signum.h:5-9; 0 at a == 0)
        b = fma_ (-s, fma_ (mu0, w0, -(mu1 * w1)), a);
    }
    else
    {
        if (MODE == kModeApprox && LSMALL)
        {
            w0 = omega4_approx<true> (fma_ (aa, c.invV, c.L));
            w1 = exp_approx_scaled<true> (fma_ (aa, -c.invVl2e, c.Ll2e));
        }
        else
        {
            const float q = aa * c.invV;
            w0 = root_omega<MODE> (c, c.L + q);
            w1 = root_omega_rev<MODE, LSMALL> (c, c.L - q);
        }
        // a == 0 makes both branches the same computation, w0 - w1 == 0 and b == a: signum's zero needs no select
        b = fma_ (-c.twoV, xor_sign (w0 - w1, a), a);
    }
    if (DERIV)
        pair_deriv<GENERAL> (c, a, b, w0, w1, mu0, mu1, d);
    return b;
}

// ---- the diode clipper: Parallel(ResistiveVoltageSource, Capacitor) closed by a DiodePair ------
// Constants of tf_wdf.py:168-177 / wdf_t.h:465-470 with P1 = Vs, P2 = C; Capacitor tf_wdf.py:114-115.
struct ClipConst
{
    float gamma; // p1R = Gv / G
    float one_m_gamma, two_gamma;
    float Rp; // port resistance seen by the root
    PairConst pair;
};

struct ClipDesc // by-value kernel argument: everything that is not a trainable parameter
{
    float fs, Vt, n_up, n_down, tol;
    int n_iter;
    int slot_R, slot_C, slot_Is, slot_nabla; // positions in the parameter vector
};

DWDF_HD void clip_setup (ClipConst& c, const ClipDesc& d, float R, float C, float Is, float nabla)
{
    const float Gv = 1.0f / R;
    const float Rc = 1.0f / (2.0f * C * d.fs);
    const float Gc = 1.0f / Rc;
    const float G = Gv + Gc;
    c.Rp = 1.0f / G;
    c.gamma = Gv / G;
    c.one_m_gamma = 1.0f - c.gamma;
    c.two_gamma = 2.0f * c.gamma;
    pair_setup (c.pair, c.Rp, Is, d.Vt, nabla, d.n_up, d.n_down, d.n_iter, d.tol);
}

// ---- the source resistance as a per-sample input channel -----------------------------------------------
// clipper_pot.py:114-117 feeds the reference's training batches as (B, T, 2) = (x, R): every sample calls
// Vs.set_resistance(R[n]) and P1.calc_impedance(), so the adaptor coefficient and the root's port resistance change
// from sample to sample. ClipRBase holds what does not (capacitor conductance, diode constants); clip_set_r derives the
// rest for one sample (tf_wdf.py:168-177, wdf_t.h:928-933) on the MUFU pipe: two reciprocals (1 / r, 1 / G; p1R = Gv Rp)
// and one log2 — which IS the scaled constant the approx root's exponentials take (L log2(e) = log2(Rp Is / V)).
struct ClipRBase
{
    float Gc; // 2 C fs
    float Is;
    float ln_up, ln_dn; // ln N_up, ln N_down (general law: ln(Rp Is / (V mu)) = L - ln mu)
};

DWDF_HD void clip_setup_r (ClipConst& c, ClipRBase& b, const ClipDesc& d, float R0, float C, float Is, float nabla)
{
    clip_setup (c, d, R0, C, Is, nabla); // everything that depends on V only (and a valid set for R0)
    const float Rc = 1.0f / (2.0f * C * d.fs);
    b.Gc = 1.0f / Rc;
    b.Is = Is;
    b.ln_up = logf (d.n_up);
    b.ln_dn = logf (d.n_down);
}

template <bool GENERAL>
DWDF_HD void clip_set_r (ClipConst& c, const ClipRBase& b, float r)
{
    const float Gv = rcp (r);
    c.Rp = rcp (Gv + b.Gc);
    c.gamma = Gv * c.Rp;
    c.one_m_gamma = 1.0f - c.gamma;
    c.two_gamma = 2.0f * c.gamma;
    PairConst& p = c.pair;
    p.RIs = c.Rp * b.Is;
    p.RIs_overV = p.RIs * p.invV;
    p.Ll2e = lg2_ (p.RIs_overV);
    p.L = 0.693147180559945f * p.Ll2e;
    if (GENERAL)
    {
        p.L_up = p.L - b.ln_up;
        p.L_dn = p.L - b.ln_dn;
    }
}

// One sample of clipper_pot.py:113-124 / DiodeClipperWDF.cpp:24-29:
//   up-sweep   b_diff = z - x; b_temp = -p1R b_diff; a = z + b_temp        (tf_wdf.py:185-192)
//   root       b = f(a)
//   down-sweep z' = b + b_temp  (Capacitor.incident, tf_wdf.py:120-122,179-183)
//   probe      voltage(C): python ordering (z' + z)/2, plugin ordering z
template <int MODE, bool GENERAL, bool LSMALL, bool PYORDER>
DWDF_HD float clip_step (const ClipConst& c, float x, float& z); // defined at the end: dispatches the exact root's common case to clip_step_exactv<f1>
template <int MODE, bool GENERAL, bool LSMALL, bool PYORDER>
DWDF_HD float clip_step_scalar (const ClipConst& c, float x, float& z)
{
    const float t = -c.gamma * (z - x);
    const float a = z + t;
    const float b = pair_reflect<MODE, GENERAL, false, LSMALL> (c.pair, a, nullptr);
    const float zn = b + t;
    const float y = PYORDER ? 0.5f * (zn + z) : z;
    z = zn;
    return y;
}

// ---- the forward kernel's sample: approx root, symmetric pair, fast-path parameters ---------------------
// ncu on the first version of the forward kernel (one lane per sequence, omega4 evaluated as omega.h
// writes it): ~47 issue slots per sample, issue ports ~80 % busy on the schedulers that hold 4 warps,
// fma pipe ~40 %, DRAM ~45 %: bound by instruction ISSUE. The sample below is the same arithmetic
// arranged for this machine:
//   * two circuit instances per lane in packed fp32x2 registers: every add / mul / fma is ONE FFMA2 /
//     FADD2 / FMUL2 issue for both (sm_100a). Only the sign/exponent bit operations, the selects, the
//     reciprocal (MUFU) and min/max stay per element. Inputs enter through a per-element subtraction
//     (x - z) and outputs leave through a per-element add, so no register-pairing moves are needed;
//   * fewer operations: the cubic of omega3 is evaluated in the log2(e)-scaled argument the exp needs
//     anyway, z' = b + gamma (x - z) is taken as  -2V lambda (w0 - w1) + (2a - z),  and the Newton step
//     y - (y - e)/(y + 1)  as  e r + y (1 - r),  r = 1/(y + 1);
//   * omega3's x >= 8 branch (x - log_approx(x), ~11 instructions) is not evaluated: max(u0) is tracked
//     per instance and the caller redoes a chunk with a loud instance (|a| > ~0.78 V for the 1N4148
//     clipper) with the LOUD form of this step, which gives every instance below the branch the same
//     bits — so a sequence's result never depends on which other sequences share its warp.
// ~23 issue slots and ~14 fma-pipe cycles per sample instead of ~47 and ~17.
// V = f2 (two instances, packed) or f1 (one instance; same operations in the same order, hence
// bit-identical per sequence — the direct-access kernel and sub-warp batches use it).
// Preconditions (fast_ok): lsmall_ok(L) and L > -39, so that wherever a chunk is accepted (u0 < 8)
// the reverse-biased argument (2L - u0) log2(e) stays above exp_approx's -126 clamp.
struct f1
{
    float x;
};
struct f2
{
    float x, y;
};
DWDF_HD f1 bc (f1, float a) { return f1 { a }; }
DWDF_HD f2 bc (f2, float a) { return f2 { a, a }; }
DWDF_HD f1 fmav (f1 a, f1 b, f1 c) { return f1 { fma_ (a.x, b.x, c.x) }; }
DWDF_HD f1 addv (f1 a, f1 b) { return f1 { add_ (a.x, b.x) }; } // (never contracted into an fma: the packed forms are not either)
DWDF_HD f1 mulv (f1 a, f1 b) { return f1 { mul_ (a.x, b.x) }; }
DWDF_HD f1 addv_rd (f1 a, f1 b) { return f1 { add_rd (a.x, b.x) }; }
#if defined(__CUDA_ARCH__)
DWDF_HD float2 as_float2 (f2 a) { return make_float2 (a.x, a.y); }
DWDF_HD f2 from_float2 (float2 a) { return f2 { a.x, a.y }; }
DWDF_HD f2 fmav (f2 a, f2 b, f2 c) { return from_float2 (__ffma2_rn (as_float2 (a), as_float2 (b), as_float2 (c))); }
DWDF_HD f2 addv (f2 a, f2 b) { return from_float2 (__fadd2_rn (as_float2 (a), as_float2 (b))); }
DWDF_HD f2 mulv (f2 a, f2 b) { return from_float2 (__fmul2_rn (as_float2 (a), as_float2 (b))); }
DWDF_HD f2 addv_rd (f2 a, f2 b) { return from_float2 (__fadd2_rd (as_float2 (a), as_float2 (b))); }
#else
DWDF_HD f2 fmav (f2 a, f2 b, f2 c) { return f2 { fma_ (a.x, b.x, c.x), fma_ (a.y, b.y, c.y) }; }
DWDF_HD f2 addv (f2 a, f2 b) { return f2 { a.x + b.x, a.y + b.y }; }
DWDF_HD f2 mulv (f2 a, f2 b) { return f2 { a.x * b.x, a.y * b.y }; }
DWDF_HD f2 addv_rd (f2 a, f2 b) { return f2 { add_rd (a.x, b.x), add_rd (a.y, b.y) }; }
#endif
// per-element helpers (negation and |.| fold into the consumer's operand modifiers)
DWDF_HD f1 negv (f1 a) { return f1 { -a.x }; }
DWDF_HD f2 negv (f2 a) { return f2 { -a.x, -a.y }; }
DWDF_HD f1 absv (f1 a) { return f1 { fabsf (a.x) }; }
DWDF_HD f2 absv (f2 a) { return f2 { fabsf (a.x), fabsf (a.y) }; }
DWDF_HD f1 maxv (f1 a, f1 b) { return f1 { fmaxf (a.x, b.x) }; }
DWDF_HD f2 maxv (f2 a, f2 b) { return f2 { fmaxf (a.x, b.x), fmaxf (a.y, b.y) }; }
DWDF_HD f1 rcpv (f1 a) { return f1 { rcp (a.x) }; }
DWDF_HD f2 rcpv (f2 a) { return f2 { rcp (a.x), rcp (a.y) }; }
DWDF_HD f1 ldexp_bits (f1 p, f1 t) { return f1 { i2f (f2i (p.x) + (int32_t) ((uint32_t) f2i (t.x) << 23)) }; }
DWDF_HD f2 ldexp_bits (f2 p, f2 t) { return f2 { i2f (f2i (p.x) + (int32_t) ((uint32_t) f2i (t.x) << 23)), i2f (f2i (p.y) + (int32_t) ((uint32_t) f2i (t.y) << 23)) }; }
DWDF_HD f1 xor_signv (f1 a, f1 s) { return f1 { xor_sign (a.x, s.x) }; }
DWDF_HD f2 xor_signv (f2 a, f2 s) { return f2 { xor_sign (a.x, s.x), xor_sign (a.y, s.y) }; }
DWDF_HD f1 zero_below (f1 y, f1 u, float thr) { return f1 { u.x < thr ? 0.0f : y.x }; }
DWDF_HD f2 zero_below (f2 y, f2 u, float thr) { return f2 { u.x < thr ? 0.0f : y.x, u.y < thr ? 0.0f : y.y }; }
// element-wise, deliberately NOT packed: the way data enters and leaves the packed registers
DWDF_HD f1 sub_in (f1 x, f1 z) { return f1 { add_ (x.x, -z.x) }; }
DWDF_HD f2 sub_in (f2 x, f2 z) { return f2 { add_ (x.x, -z.x), add_ (x.y, -z.y) }; }
DWDF_HD f1 add_out (f1 a, f1 b) { return f1 { a.x + b.x }; }
DWDF_HD f2 add_out (f2 a, f2 b) { return f2 { a.x + b.x, a.y + b.y }; }

DWDF_HD f1 select_below (f1 lo, f1 hi, f1 u, float thr) { return f1 { u.x < thr ? lo.x : hi.x }; }
DWDF_HD f2 select_below (f2 lo, f2 hi, f2 u, float thr) { return f2 { u.x < thr ? lo.x : hi.x, u.y < thr ? lo.y : hi.y }; }
DWDF_HD f1 clamp_exp_arg (f1 a) { return f1 { fmaxf (a.x, -126.0f) }; }
DWDF_HD f2 clamp_exp_arg (f2 a) { return f2 { fmaxf (a.x, -126.0f), fmaxf (a.y, -126.0f) }; }
// log_approx_pos over V (omega.h:49-63): mantissa and exponent per element (bit operations), the cubic and the sum packed
DWDF_HD f1 mant12 (f1 a) { return f1 { i2f ((f2i (a.x) & 0x007fffff) | 0x3f800000) }; }
DWDF_HD f2 mant12 (f2 a) { return f2 { i2f ((f2i (a.x) & 0x007fffff) | 0x3f800000), i2f ((f2i (a.y) & 0x007fffff) | 0x3f800000) }; }
DWDF_HD f1 expo_magic (f1 a) { return f1 { i2f ((f2i (a.x) >> 23) | 0x4b000000) }; }
DWDF_HD f2 expo_magic (f2 a) { return f2 { i2f ((f2i (a.x) >> 23) | 0x4b000000), i2f ((f2i (a.y) >> 23) | 0x4b000000) }; }
template <class V>
DWDF_HD V log_approx_posv (V x)
{
    const V m = mant12 (x);
    const V e = addv (expo_magic (x), bc (V {}, -8388735.0f));
    const V pl = fmav (m, fmav (m, fmav (m, bc (V {}, 0.1640425613334452f), bc (V {}, -1.098865286222744f)), bc (V {}, 3.148297929334117f)), bc (V {}, -2.213475204444817f));
    return mulv (bc (V {}, 0.693147180559945f), addv (e, pl));
}

DWDF_HD bool fast_ok (float L) { return lsmall_ok (L) && L > -39.0f; }
constexpr float kFastLoud = 8.0f * kLog2e; // omega3's log branch starts at u0 = 8, i.e. here in the scaled argument

// 2^floor(x') * cubic(frac(x')) of exp_approx (omega.h:83-116), x' already scaled by log2(e) and >= -126
template <class V>
DWDF_HD V exp_approx_scaledv (V xp)
{
    const V kMagic = bc (V {}, 12582912.0f);
    const V t = addv_rd (xp, kMagic);
    const V f = addv (xp, negv (addv (t, negv (kMagic))));
    const V p = fmav (f, fmav (f, fmav (f, bc (V {}, 0.07944154167983575f), bc (V {}, 0.2274112777602189f)), bc (V {}, 0.6931471805599453f)), bc (V {}, 1.0f));
    return ldexp_bits (p, t);
}

// One sample. State: z = z[n] and hz = z[n]/2 (python ordering's output is hz[n+1] + hz[n]).
// umax = max(umax, u0 log2(e)). Returns y[n].
template <class V>
struct StepTapeV;
template <class V>
DWDF_HD void fast_step_tape (const ClipConst& c, V xz, V a, V w0, V w1, StepTapeV<V>* tp);
// TAPE: also the step's linearisation (A, cg, cl, cv as in clip_step_tape), for the fused forward-mode training pass
// LOUD: the same step with omega3's x >= 8 branch (x - log_approx(x), omega.h:165) selected per element and the reverse-biased
// exponential's argument clamped — what a loud chunk is redone with, still two instances per lane in packed registers. For an
// element below the branch every operation and operand is the one of the plain step (the select keeps the cubic, the clamp
// does not act), so its result is the same bits: which of the two ran never shows in a sequence's output.
template <class V>
DWDF_HD V log_approx_posv (V x);
template <class V, bool PY, bool TAPE, bool LOUD = false>
DWDF_HD V clip_step_fastv_impl (const ClipConst& c, V x, V& z, V& hz, V& umax, StepTapeV<V>* tp)
{
    const PairConst& p = c.pair;
    // a = z + gamma (x - z) exactly as the adaptor writes it (tf_wdf.py:185-192): the form (1 - gamma) z + gamma x
    // rounds the pole 1 - 2 gamma by half an ulp of ONE, which a long RC memory amplifies by 1 / (2 gamma) into the
    // DC gain (1e-5 relative at gamma = 1e-3; found by the randomised sweep, tests/test_gpu_fuzz.py)
    const V xz = sub_in (x, z);
    const V a = fmav (bc (V {}, c.gamma), xz, z);
    const V a2z = fmav (bc (V {}, c.two_gamma), xz, z); // a + gamma (x - z)
    const V aa = absv (a);
    const V us = fmav (aa, bc (V {}, p.invVl2e), bc (V {}, p.Ll2e)); // u0 log2(e), u0 = L + |a| / V
    const V u1s = fmav (aa, bc (V {}, -p.invVl2e), bc (V {}, p.Ll2e));
    const V w1 = exp_approx_scaledv (LOUD ? clamp_exp_arg (u1s) : u1s); // omega4(L - |a|/V) = exp_approx
    umax = maxv (umax, us);
    // omega3 below 8 (omega.h:160-167) in the scaled argument; exact 0 below the cubic's root (silence in, silence out)
    constexpr float i1 = 1.0f / kLog2e, i2 = i1 * i1, i3 = i2 * i1;
    V y = fmav (us, fmav (us, fmav (us, bc (V {}, -1.314293149877800e-3f * i3), bc (V {}, 4.775931364975583e-2f * i2)), bc (V {}, 3.631952663804445e-1f * i1)), bc (V {}, 6.313183464296682e-1f));
    y = zero_below (y, us, kOmega3Zero * kLog2e);
    if (LOUD)
    {
        const V u0 = mulv (us, bc (V {}, i1)); // u0 = L + |a| / V
        y = select_below (y, addv (u0, negv (log_approx_posv (maxv (u0, bc (V {}, 1.0f))))), us, kFastLoud);
    }
    const V r = rcpv (addv (y, bc (V {}, 1.0f)));
    const V q = fmav (negv (y), r, y); // y (1 - r)
    const V e = exp_approx_scaledv (fmav (y, bc (V {}, -kLog2e), us));
    const V w0 = fmav (e, r, q); // omega4 = y - (y - e) / (y + 1)
    const V ds = xor_signv (addv (w0, negv (w1)), a); // lambda (w0 - w1); w0 == w1 bit for bit at a == 0
    const V zn = fmav (bc (V {}, -p.twoV), ds, a2z); // b + gamma (x - z),  b = a - 2 V lambda (w0 - w1)
    const V hzn = mulv (bc (V {}, 0.5f), zn);
    const V yo = PY ? add_out (hzn, hz) : z;
    if (TAPE)
        fast_step_tape (c, xz, a, w0, w1, tp);
    z = zn;
    hz = hzn;
    return yo;
}
template <class V, bool PY>
DWDF_HD V clip_step_fastv (const ClipConst& c, V x, V& z, V& hz, V& umax)
{
    return clip_step_fastv_impl<V, PY, false> (c, x, z, hz, umax, nullptr);
}

// Four consecutive samples (one 16-byte chunk of a row) through the fast step.
template <class V, bool PY>
DWDF_HD void clip_chunk_fastv (const ClipConst& c, const V (&x)[4], V& z, V (&o)[4], V& umax)
{
    V hz = mulv (bc (V {}, 0.5f), z);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        o[k] = clip_step_fastv<V, PY> (c, x[k], z, hz, umax);
}
// ... and a loud chunk's redo: the same four samples through the LOUD step
template <class V, bool PY>
DWDF_HD void clip_chunk_loudv (const ClipConst& c, const V (&x)[4], V& z, V (&o)[4], V& umax)
{
    V hz = mulv (bc (V {}, 0.5f), z);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        o[k] = clip_step_fastv_impl<V, PY, false, true> (c, x[k], z, hz, umax, nullptr);
}
// ... and the general way (omega3's log branch included), one instance: parameters outside the packed paths' range
template <bool PY>
DWDF_HD void clip_chunk_general (const ClipConst& c, const float (&x)[4], float& z, float (&o)[4]);

// Same step, also returning what the adjoint sweep needs:
//   A  = dz'/dz     = f'(a)(1 - gamma) - gamma
//   cg = dz'/dgamma = (x - z)(f'(a) + 1)
//   cl = dz'/d ell  = -2 V M1
//   cv = dz'/dV     (at fixed ell)
struct StepTape
{
    float A, cg, cl, cv;
};
template <int MODE, bool GENERAL, bool LSMALL, bool PYORDER>
DWDF_HD float clip_step_tape (const ClipConst& c, float x, float& z, StepTape& tp)
{
    const float xz = x - z;
    const float t = c.gamma * xz;
    const float a = z + t;
    PairDeriv d;
    const float b = pair_reflect<MODE, GENERAL, true, LSMALL> (c.pair, a, &d);
    const float zn = b + t;
    const float y = PYORDER ? 0.5f * (zn + z) : z;
    const float fp1 = fma_ (-2.0f, d.S1, 2.0f); // f'(a) + 1
    tp.A = fma_ (fp1, c.one_m_gamma, -1.0f); // (fp1 - 1)(1 - gamma) - gamma = fp1 (1 - gamma) - 1
    tp.cg = xz * fp1;
    tp.cl = -c.pair.twoV * d.M1;
    tp.cv = d.dV;
    z = zn;
    return y;
}

// The same linearisation recovered from BOTH end points of a step — the adjoint's way: it knows the
// state before (z) and after (zn) each sample from the forward pass's output, so nothing is replayed.
//   a = z + gamma (x - z),  b = zn - gamma (x - z)            (the step's own equations, solved for b)
//   b = a - 2 V lambda (mu0 w0 - mu1 w1)  =>  mu0 w0 = mu1 w1 + lambda (a - b) / (2 V)
// Only the reverse-biased omega w1 is evaluated (argument <= L: the cheap branch; under LSMALL one
// exp_approx); the forward-biased w0 — the expensive one, and the one whose value fixes f'(a) — is
// read off the reflected wave the forward pass produced. (x, z, zn) -> (A, cg, cl, cv).
template <int MODE, bool GENERAL, bool LSMALL>
DWDF_HD void clip_step_recover (const ClipConst& c, float x, float z, float zn, StepTape& tp)
{
    const float xz = x - z;
    const float a = fma_ (c.gamma, xz, z);
    const float b = fma_ (-c.gamma, xz, zn);
    const float aa = fabsf (a);
    const float dl = xor_sign (a - b, a); // lambda (a - b) >= 0
    float w0, w1, mu0 = 1.0f, mu1 = 1.0f, u0;
    if (GENERAL)
    {
        const bool pos = a >= 0.0f;
        mu0 = pos ? c.pair.n_dn : c.pair.n_up;
        mu1 = pos ? c.pair.n_up : c.pair.n_dn;
        w1 = root_omega_rev<MODE, LSMALL> (c.pair, (pos ? c.pair.L_up : c.pair.L_dn) - aa * (pos ? c.pair.inv_up : c.pair.inv_dn));
        w0 = fma_ (dl, c.pair.inv2V, mu1 * w1) * (pos ? c.pair.rn_dn : c.pair.rn_up);
        u0 = (pos ? c.pair.L_dn : c.pair.L_up) + aa * (pos ? c.pair.inv_dn : c.pair.inv_up);
        if (a == 0.0f) // lambda = 0 hides w0 from b; rare, evaluate it
            w0 = root_omega<MODE> (c.pair, c.pair.L_dn);
    }
    else
    {
        if (MODE == kModeApprox && LSMALL)
            w1 = exp_approx_scaled<true> (fma_ (aa, -c.pair.invVl2e, c.pair.Ll2e));
        else
            w1 = root_omega_rev<MODE, LSMALL> (c.pair, c.pair.L - aa * c.pair.invV);
        w0 = fma_ (dl, c.pair.inv2V, w1);
        u0 = fma_ (aa, c.pair.invV, c.pair.L);
    }
    // While the forward-biased diode is still off its omega is below fp32's resolution of a - b (2 V w0 against
    // ulp(a)): read off the wave it would be noise — harmless for the state recurrence, but dL/dIs of a quiet signal is
    // the sum of exactly these terms. There the omega is one cheap evaluation (omega4 = exp_approx below
    // kOmega3Zero; the exact omega's x <= -2 region), so it is taken directly.
    {
        const float thr = MODE == kModeExact ? -2.0f : kOmega3Zero;
        const float w0_lo = MODE == kModeExact ? omega_exact_low (u0) : exp_approx (u0);
        w0 = u0 < thr ? w0_lo : w0;
    }
    PairDeriv d;
    pair_deriv<GENERAL> (c.pair, a, b, w0, w1, mu0, mu1, &d);
    const float fp1 = fma_ (-2.0f, d.S1, 2.0f); // f'(a) + 1
    tp.A = fma_ (fp1, c.one_m_gamma, -1.0f);
    tp.cg = xz * fp1;
    tp.cl = -c.pair.twoV * d.M1;
    tp.cv = d.dV;
}

// clip_step_recover for the fast-path parameters (approx root, symmetric pair, lsmall_ok(L)), written
// over V so that the adjoint kernel can run it on PAIRS OF CONSECUTIVE SAMPLES of one sequence in packed
// fp32x2 registers: with the trajectory known, neighbouring samples are independent work (only the
// two-FMA recurrence of the running adjoint stays serial), and a 16-byte shared-memory read delivers them
// as aligned register pairs. (x, z, z') -> (A, cg, cl, cv) as in clip_step_recover.
template <class V>
struct StepTapeV
{
    V A, cg, cl, cv;
};
// (A, cg, cl, cv) from the two omegas of a fast step: pair_deriv's formulas (cancellation-free dV) over V
template <class V>
DWDF_HD void fast_step_tape (const ClipConst& c, V xz, V a, V w0, V w1, StepTapeV<V>* tp)
{
    const PairConst& p = c.pair;
    const V wp0 = mulv (w0, rcpv (addv (w0, bc (V {}, 1.0f))));
    const V wp1 = mulv (w1, rcpv (addv (w1, bc (V {}, 1.0f))));
    const V S1 = addv (wp0, wp1);
    const V M1 = xor_signv (addv (wp0, negv (wp1)), a);
    const V fp1 = fmav (bc (V {}, -2.0f), S1, bc (V {}, 2.0f)); // f'(a) + 1
    tp->A = fmav (fp1, bc (V {}, c.one_m_gamma), bc (V {}, -1.0f));
    tp->cg = mulv (xz, fp1);
    tp->cl = mulv (bc (V {}, -p.twoV), M1);
    const V ww = xor_signv (fmav (w0, wp0, negv (mulv (w1, wp1))), a);
    tp->cv = fmav (mulv (a, bc (V {}, 2.0f * p.invV)), S1, mulv (bc (V {}, -2.0f), ww));
}
// The packed intrinsics (__fmul2_rn, __fadd2_rn) ARE contracted into FFMA2 by the compiler where a product feeds a sum
// (the scalar __fmul_rn / __fadd_rn never are), so V-form code that must agree bit for bit between V = f1 and V = f2
// spells every fused multiply-add out as fmav and never adds a bare product; where a product must be subtracted
// unfused (omega(x0) - omega(x1), which has to cancel exactly at a == 0), opaque() hides it from the optimiser.
DWDF_HD f1 opaque (f1 a)
{
#if defined(__CUDA_ARCH__)
    asm volatile ("" : "+f"(a.x));
#endif
    return a;
}
DWDF_HD f2 opaque (f2 a)
{
#if defined(__CUDA_ARCH__)
    asm volatile ("" : "+f"(a.x), "+f"(a.y));
#endif
    return a;
}
DWDF_HD f1 minv (f1 a, float b) { return f1 { fminf (a.x, b) }; }
DWDF_HD f2 minv (f2 a, float b) { return f2 { fminf (a.x, b), fminf (a.y, b) }; }
DWDF_HD f1 ex2v (f1 a) { return f1 { ex2_ (a.x) }; }
DWDF_HD f2 ex2v (f2 a) { return f2 { ex2_ (a.x), ex2_ (a.y) }; }
// exp_nonpos / omega_exact_low over V: the same operations in the same order
template <class V>
DWDF_HD V exp_nonposv (V x)
{
    const V kHi = bc (V {}, 1.44269502162933349609375f);
    const V t = mulv (x, kHi);
    const V lo = fmav (x, bc (V {}, 1.925963033500011e-8f), fmav (x, kHi, negv (t)));
    const V e0 = ex2v (t);
    return fmav (e0, mulv (bc (V {}, 0.693147180559945f), lo), e0);
}
template <class V>
DWDF_HD V omega_exact_lowv (V x)
{
    const V E = exp_nonposv (minv (x, 0.0f));
    const V Ec = minv (E, 0.1353352832f);
    const V s = fmav (Ec, fmav (Ec, fmav (Ec, fmav (Ec, fmav (Ec, fmav (Ec, bc (V {}, kOmLow6), bc (V {}, kOmLow5)), bc (V {}, kOmLow4)), bc (V {}, kOmLow3)), bc (V {}, kOmLow2)), bc (V {}, kOmLow1)), bc (V {}, 1.0f));
    return mulv (E, s);
}

DWDF_HD f1 lg2v (f1 a) { return f1 { lg2_ (a.x) }; }
DWDF_HD f2 lg2v (f2 a) { return f2 { lg2_ (a.x), lg2_ (a.y) }; }
DWDF_HD f1 select_le (f1 u, float thr, f1 lo, f1 hi) { return f1 { u.x <= thr ? lo.x : hi.x }; }
DWDF_HD f2 select_le (f2 u, float thr, f2 lo, f2 hi) { return f2 { u.x <= thr ? lo.x : hi.x, u.y <= thr ? lo.y : hi.y }; }
constexpr float kLn2f = 0.693147180559945f;
template <class V>
DWDF_HD V fsc_stepv (V w, V r)
{
    const V wp1 = addv (w, bc (V {}, 1.0f));
    const V w2 = mulv (bc (V {}, 2.0f), wp1), f = fmav (bc (V {}, 0.66666666666666667f), r, wp1);
    const V q = mulv (w2, f), qmr = fmav (w2, f, negv (r)); // q and q - r
    const V e = mulv (mulv (r, qmr), rcpv (mulv (wp1, fmav (bc (V {}, -2.0f), r, q))));
    return fmav (w, e, w);
}
// omega_exact with one FSC iteration (n_iter == 1, the default) over V: every kernel's exact-root forward step goes
// through this one function (V = f1: one sequence per lane, V = f2: two), so they agree bit for bit per sequence.
template <class V>
DWDF_HD V omega_exact1v (V x)
{
    const V one = bc (V {}, 1.0f);
    const V w_lo = omega_exact_lowv (x);
    // -2 < x <= 1 + pi
    const V p = addv (x, bc (V {}, -1.0f));
    const V ser = fmav (p, fmav (p, fmav (p, bc (V {}, 2.1158854166666667e-4f), bc (V {}, -3.2552083333333333e-4f)), bc (V {}, -5.2083333333333333e-3f)), bc (V {}, 0.0625f));
    const V w_mid = fmav (mulv (p, p), ser, fmav (bc (V {}, 0.5f), x, bc (V {}, 0.5f)));
    // x > 1 + pi
    const V xc = maxv (x, one);
    const V l2 = lg2v (xc);
    const V l = mulv (bc (V {}, kLn2f), l2);
    const V it = rcpv (xc);
    const V li = mulv (l, it);
    const V inner = fmav (it, fmav (l, fmav (l, bc (V {}, 0.33333333333333333f), bc (V {}, -1.5f)), one), fmav (bc (V {}, 0.5f), l, bc (V {}, -1.0f)));
    const V w_hi = fmav (li, fmav (it, inner, one), fmav (bc (V {}, -kLn2f), l2, xc));
    V w = maxv (select_le (x, 4.141592653589793f, w_mid, w_hi), bc (V {}, 0.05f)); // (the floor only ever acts on elements that end up taking w_lo)
    w = fsc_stepv (w, fmav (bc (V {}, -kLn2f), lg2v (w), addv (x, negv (w)))); // residual (x - w) - ln w
    return select_le (x, -2.0f, w_lo, w);
}
DWDF_HD bool exact_fast_ok (const PairConst& c) { return rev_small_ok (c) && c.n_iter == 1; }
// One sample of the clipper with the exact root and the symmetric pair (clip_step's equations), exact_fast_ok parameters.
template <class V, bool PYORDER>
DWDF_HD V clip_step_exactv (const ClipConst& c, V x, V& z)
{
    const PairConst& p = c.pair;
    const V xz = addv (x, negv (z));
    const V a = fmav (bc (V {}, c.gamma), xz, z); // the adaptor's own form (see clip_step_fastv)
    const V aa = absv (a);
    const V w0 = opaque (omega_exact1v (fmav (aa, bc (V {}, p.invV), bc (V {}, p.L))));
    const V w1 = opaque (omega_exact_lowv (fmav (aa, bc (V {}, -p.invV), bc (V {}, p.L))));
    // a == 0 makes both branches the same computation (L <= -2: both take omega_exact_low(L)), w0 - w1 == 0 and b == a
    const V b = fmav (bc (V {}, -p.twoV), xor_signv (addv (w0, negv (w1)), a), a);
    const V zn = fmav (bc (V {}, c.gamma), xz, b);
    const V y = PYORDER ? mulv (bc (V {}, 0.5f), addv (zn, z)) : z;
    z = zn;
    return y;
}

// The same step with its linearisation (the one-sweep tangent pass on the exact root): w0 and w1 are at hand, so the tape costs
// two reciprocals and a dozen FMAs on top of the forward step. Output and state bit-identical to clip_step_exactv.
template <class V, bool PYORDER>
DWDF_HD V clip_step_exact_tapev (const ClipConst& c, V x, V& z, StepTapeV<V>& tp)
{
    const PairConst& p = c.pair;
    const V xz = addv (x, negv (z));
    const V a = fmav (bc (V {}, c.gamma), xz, z);
    const V aa = absv (a);
    const V w0 = opaque (omega_exact1v (fmav (aa, bc (V {}, p.invV), bc (V {}, p.L))));
    const V w1 = opaque (omega_exact_lowv (fmav (aa, bc (V {}, -p.invV), bc (V {}, p.L))));
    const V b = fmav (bc (V {}, -p.twoV), xor_signv (addv (w0, negv (w1)), a), a);
    const V zn = fmav (bc (V {}, c.gamma), xz, b);
    const V y = PYORDER ? mulv (bc (V {}, 0.5f), addv (zn, z)) : z;
    z = zn;
    const V wp0 = mulv (w0, rcpv (addv (w0, bc (V {}, 1.0f))));
    const V wp1 = mulv (w1, rcpv (addv (w1, bc (V {}, 1.0f))));
    const V S1 = addv (wp0, wp1);
    const V M1 = xor_signv (addv (wp0, negv (wp1)), a);
    const V fp1 = fmav (bc (V {}, -2.0f), S1, bc (V {}, 2.0f)); // f'(a) + 1
    tp.A = fmav (fp1, bc (V {}, c.one_m_gamma), bc (V {}, -1.0f));
    tp.cg = mulv (xz, fp1);
    tp.cl = mulv (bc (V {}, -p.twoV), M1);
    const V ww = xor_signv (fmav (w0, wp0, negv (mulv (w1, wp1))), a);
    tp.cv = fmav (mulv (a, bc (V {}, 2.0f * p.invV)), S1, mulv (bc (V {}, -2.0f), ww));
    return y;
}

// The reverse sweep's tape, UNSCALED: the sweep accumulates G cg, G m1, G as, G ww over a segment and applies the constant
// factors once (tape_scale) instead of once per sample.
template <class V>
struct StepTapeY
{
    V A; // dz'/dz
    V cg; // dz'/dgamma / 2  (clip_step_recover_yv: times gamma)
    V m1; // dz'/d ell / (-2 V)
    V as, ww; // dz'/dV = (2 / V) as - 2 ww
};
// sums of G cg, G m1, G as, G ww over a segment -> the sums of G dz'/dgamma, G dz'/d ell, G dz'/dV
// (inv_gamma: 1 / gamma for clip_step_recover_yv's sums, 1 for clip_step_recoverv's)
DWDF_HD void tape_scale (const ClipConst& c, float inv_gamma, float sg, float sm, float sas, float sww, float& g, float& l, float& v)
{
    g = 2.0f * inv_gamma * sg;
    l = -c.pair.twoV * sm;
    v = fma_ (2.0f * c.pair.invV, sas, -2.0f * sww);
}

// MODE kModeApprox: fast-path parameters (lsmall_ok(L)); kModeExact: rev_small_ok (the reverse-biased argument stays
// in TOMS-917's x <= -2 region, and so does the forward-biased one wherever it is evaluated directly)
template <class V, int MODE = kModeApprox>
DWDF_HD void clip_step_recoverv (const ClipConst& c, V x, V z, V zn, StepTapeY<V>& tp)
{
    const PairConst& p = c.pair;
    const V xz = addv (x, negv (z));
    const V a = fmav (bc (V {}, c.gamma), xz, z);
    const V b = fmav (bc (V {}, -c.gamma), xz, zn);
    const V aa = absv (a);
    const V d = addv (a, negv (b));
    V w0, w1;
    if (MODE == kModeExact)
    {
        w1 = omega_exact_lowv (fmav (aa, bc (V {}, -p.invV), bc (V {}, p.L)));
        w0 = fmav (xor_signv (d, a), bc (V {}, p.inv2V), w1); // mu0 w0 = mu1 w1 + lambda (a - b) / (2 V)
        const V u0 = fmav (aa, bc (V {}, p.invV), bc (V {}, p.L));
        w0 = select_below (omega_exact_lowv (u0), w0, u0, -2.0f); // diode still off: evaluated directly (see clip_step_recover)
    }
    else
    {
        w1 = exp_approx_scaledv (clamp_exp_arg (fmav (aa, bc (V {}, -p.invVl2e), bc (V {}, p.Ll2e))));
        w0 = fmav (xor_signv (d, a), bc (V {}, p.inv2V), w1);
        // diode still off (u0 < kOmega3Zero): omega4 = exp_approx there, taken directly (see clip_step_recover);
        // its argument is at least L log2(e) > -126 under lsmall_ok: no clamp
        const V us = fmav (aa, bc (V {}, p.invVl2e), bc (V {}, p.Ll2e));
        w0 = select_below (exp_approx_scaledv (us), w0, us, kOmega3Zero * kLog2e);
    }
    const V wp0 = mulv (w0, rcpv (addv (w0, bc (V {}, 1.0f))));
    const V wp1 = mulv (w1, rcpv (addv (w1, bc (V {}, 1.0f))));
    const V S1 = addv (wp0, wp1);
    tp.m1 = xor_signv (addv (wp0, negv (wp1)), a);
    tp.A = fmav (S1, bc (V {}, -2.0f * c.one_m_gamma), bc (V {}, c.one_m_gamma - c.gamma)); // (f' + 1)(1 - gamma) - 1,  f' + 1 = 2 - 2 S1
    tp.cg = fmav (negv (S1), xz, xz); // (x - z)(f' + 1) / 2
    tp.as = mulv (a, S1);
    tp.ww = xor_signv (fmav (w0, wp0, negv (mulv (w1, wp1))), a); // lambda (w0 w0' - w1 w1'): the cancellation-free form of pair_deriv
}

// The same linearisation from the forward OUTPUT ALONE — the reverse sweep of the exact root (symmetric pair, fromy_ok
// parameters) reads y and the target and never touches x again: 8 bytes per sample instead of 12.
// The two step equations a = z + t, b = z' - t (t = gamma (x - z)) give a + b = z + z' = 2 v, the diode voltage — in the
// python ordering that IS the stored output y[n] — and the root's law b = a - 2 V lambda (w0 - w1) gives a - b. With
// w + ln w = u (Wright omega) and a = v + V lambda (w0 - w1) the two omegas become explicit in v:
//     w0 = E exp(-w1),  w1 = W exp(-w0),   E = k e^{|v|/V},  W = k e^{-|v|/V},  k = Rp Is / V = e^L.
// fromy_ok asks for k < e^-6 (every diode of diode_config.py at any clipper impedance: k ~ 1e-4): then w1 <= k, w0 w1 <= k^2 <
// 6.2e-6, and first order in w1 is fp32-exact:  w1 = W exp(-E),  w0 = E (1 - w1),  1 / (1 + w1) = 1 - w1.
// Everything the tape needs follows: gamma (x - z) = v - z + V lambda delta with delta = w0 - w1, f'(a) from S1 = w0' + w1',
//     M1 = lambda delta r0 r1,   lambda (w0 w0' - w1 w1') = M1 (w0 + w1 + w0 w1),   r = 1 / (1 + w).
// While the diodes are off (|v| < V / 8) delta = w0 - w1 would cancel (dL/dIs of a quiet signal is the sum of exactly these
// terms), so there it is k 2 sinh(|v|/V) with the sinh as its odd series (neglected: k^2 / 2, (|v|/V)^6 / 5040).
// Three MUFU.EX2 and one MUFU.RCP per sample and no omega evaluation (clip_step_recoverv<exact>: four EX2, four RCP).
// The tape comes back UNSCALED (StepTapeY): the sweep accumulates G cg, G m1, G as, G ww and applies the constant factors
// once per segment (tape_scale).
// EXACT ROOT ONLY. Recovering a from v is ill-conditioned by 1 + w0 (a conducting diode pins its voltage), which fp32's
// resolution of v survives (sums within 7e-6 of the fp64 oracle at +-10 V) but the approx root's own error does not:
// omega4 misses omega by up to ~1e-3 w0, that miss is exponentiated on the way back (w0 e^{-(omega4 - omega)}), and the
// sums drift by 1e-3 ... 2e-2 from the oracle's (measured, tests/test_host_math.py) — the approx root's sweep keeps reading x.
// (Repairing it takes one omega4 evaluation per sample, w0~ = w0 + (omega4(c + w0) - w0)(1 + w0) with c = ln E - w1: more
// issue slots than the 4 bytes cost.)
DWDF_HD bool fromy_ok (const PairConst& c) { return rev_small_ok (c) && c.L < -6.0f; }
// v: the step's diode voltage (z + z')/2; z: the state before the step.
template <class V>
DWDF_HD void clip_step_recover_yv (const ClipConst& c, V v, V z, StepTapeY<V>& tp)
{
    const PairConst& p = c.pair;
    const V one = bc (V {}, 1.0f);
    const V av = absv (v);
    const V E = ex2v (fmav (av, bc (V {}, p.invVl2e), bc (V {}, p.Ll2e)));
    const V W = ex2v (fmav (av, bc (V {}, -p.invVl2e), bc (V {}, p.Ll2e)));
    const V w1 = mulv (W, ex2v (mulv (E, bc (V {}, -kLog2e)))); // W exp(-E)
    const V w0 = fmav (negv (E), w1, E); // E (1 - w1)
    const V r0 = rcpv (addv (w0, one));
    const V wp0 = mulv (w0, r0), wp1 = fmav (negv (w1), w1, w1);
    const V S1 = addv (wp0, wp1);
    const V q = mulv (av, bc (V {}, p.invV)), q2 = mulv (q, q);
    const V sh = mulv (mulv (q, bc (V {}, 2.0f * p.RIs_overV)), fmav (q2, fmav (q2, bc (V {}, 1.0f / 120.0f), bc (V {}, 1.0f / 6.0f)), one)); // k 2 sinh q
    const V ld = xor_signv (select_below (sh, addv (w0, negv (w1)), q, 0.125f), v); // lambda (w0 - w1)
    const V t = mulv (ld, r0);
    tp.m1 = fmav (negv (t), w1, t); // lambda delta r0 (1 - w1)
    tp.ww = mulv (tp.m1, fmav (w0, w1, addv (w0, w1)));
    const V xzg = fmav (bc (V {}, p.V), ld, addv (v, negv (z))); // gamma (x - z)
    tp.A = fmav (S1, bc (V {}, -2.0f * c.one_m_gamma), bc (V {}, c.one_m_gamma - c.gamma)); // (f' + 1)(1 - gamma) - 1,  f' + 1 = 2 - 2 S1
    tp.cg = fmav (negv (S1), xzg, xzg);
    tp.as = mulv (fmav (bc (V {}, p.V), ld, v), S1);
}
// The forward step every kernel calls. Exact root, symmetric pair, rev_small_ok parameters, one FSC iteration: the
// V-form step (so that the packed two-sequences-per-lane kernel and the one-per-lane kernels agree bit for bit).
template <int MODE, bool GENERAL, bool LSMALL, bool PYORDER>
DWDF_HD float clip_step (const ClipConst& c, float x, float& z)
{
    if (MODE == kModeExact && ! GENERAL && LSMALL && c.pair.n_iter == 1)
    {
        f1 zz { z };
        const f1 y = clip_step_exactv<f1, PYORDER> (c, f1 { x }, zz);
        z = zz.x;
        return y.x;
    }
    return clip_step_scalar<MODE, GENERAL, LSMALL, PYORDER> (c, x, z);
}

template <bool PY>
DWDF_HD void clip_chunk_general (const ClipConst& c, const float (&x)[4], float& z, float (&o)[4])
{
    o[0] = clip_step<kModeApprox, false, false, PY> (c, x[0], z);
    o[1] = clip_step<kModeApprox, false, false, PY> (c, x[1], z);
    o[2] = clip_step<kModeApprox, false, false, PY> (c, x[2], z);
    o[3] = clip_step<kModeApprox, false, false, PY> (c, x[3], z);
}

// four samples of two sequences, exact root (exact_fast_ok parameters)
template <bool PY>
DWDF_HD void clip_chunk_exact2 (const ClipConst& c, const f2 (&x)[4], f2& z, f2 (&o)[4])
{
    o[0] = clip_step_exactv<f2, PY> (c, x[0], z);
    o[1] = clip_step_exactv<f2, PY> (c, x[1], z);
    o[2] = clip_step_exactv<f2, PY> (c, x[2], z);
    o[3] = clip_step_exactv<f2, PY> (c, x[3], z);
}
// four samples of one sequence the general way, any root mode (parameters outside the packed paths' range)
template <int MODE, bool PY>
DWDF_HD void clip_chunk_any (const ClipConst& c, const float (&x)[4], float& z, float (&o)[4])
{
    o[0] = clip_step<MODE, false, false, PY> (c, x[0], z);
    o[1] = clip_step<MODE, false, false, PY> (c, x[1], z);
    o[2] = clip_step<MODE, false, false, PY> (c, x[2], z);
    o[3] = clip_step<MODE, false, false, PY> (c, x[3], z);
}

} // namespace dwdf
