/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of the reference's algorithm for the differentiable-WDF hot path:
 * the wdf_py elements/adaptors (wdf_py/lib/tf_wdf.py), the chowdsp_wdf diode-pair roots
 * (wdf_t.h:859-985, Toms917DiodePair.h), the Wright-omega approximations (omega.h), TOMS-917
 * (modules/toms917/toms917.cpp) and the per-sample loops that drive them (lpf.py:38-46,
 * clipper_pot.py:103-127, DiodeClipperWDF.cpp:18-30), plus the reverse-mode gradient the
 * reference obtains from tf.GradientTape (clipper_pot.py:246-269).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load this
 * library — as the checker or as the timed CPU baseline, never as the product path. The product
 * (differentiable-wdfs_b200/) is CUDA only and does not link, import or call anything in oracle/.
 *
 * Pinning: tests/test_oracle.py checks this file against (1) every golden vector the reference's
 * own tests hold for the path (omega table OmegaTest.cpp:6-48, wdf_standalone_test
 * 4.77, StaticWDFTest fixture, RC low-pass magnitudes, divider) and (2) outputs of the reference
 * itself compiled here (oracle/_ref/libdwdf_ref.so, fixtures under tests/golden/).
 * Gradients w.r.t. (Is, nabla, R, C): PARITY UNPINNED by the reference (it never computes or asserts
 * them); pinned instead against torch.autograd over oracle/torch_wdf.py and fp64 central differences.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, like the plugin's build).
 */
#define _GNU_SOURCE
#include <complex.h>
#include <fenv.h>
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { OW_RESISTOR = 0, OW_CAPACITOR = 1, OW_RESVS = 2, OW_SERIES = 3, OW_PARALLEL = 4, OW_INVERTER = 5, OW_INDUCTOR = 6, OW_CAPACITOR_ALPHA = 7, OW_INDUCTOR_ALPHA = 8, OW_RESCS = 9, OW_YPARAM = 10 };
enum { OW_ROOT_IDEAL_VS = 0, OW_ROOT_DIODE_PAIR = 1, OW_ROOT_IDEAL_CS = 3, OW_ROOT_DIODE = 4, OW_ROOT_SWITCH = 5 };
enum { OW_ORDER_PLUGIN = 0, OW_ORDER_PYTHON = 1 };
#define OW_MAX_NODES 64
#define OW_MAX_THREADS 256

/* ---- TOMS-917 (modules/toms917/toms917.cpp:72-375), complex double ---------------------------- */

/* asymptotic / negative-log series shared by regions 5, 6, 7 (toms917.cpp:269-296) */
static double complex toms_log_series (double complex t, double complex lt)
{
    double complex p = lt;
    return ((1.0 + (-3.0 / 2.0 + 1.0 / 3.0 * p) * p) * p + ((-1.0 + 1.0 / 2.0 * p) * p + (p + (-p + t) * t) * t) * t) / (t * t * t);
}

/* one Fritsch-Shafer-Crowley step (toms917.cpp:347-352 and :358-363) */
static double complex toms_fsc (double complex z, double complex w, double s, double complex* r_out, double complex* wp1_out)
{
    double complex r = z - s * w - clog (w);
    double complex wp1 = s * w + 1.0;
    double complex e = r / wp1 * (2.0 * wp1 * (wp1 + 2.0 / 3.0 * r) - r) / (2.0 * wp1 * (wp1 + 2.0 / 3.0 * r) - 2.0 * r);
    *r_out = r;
    *wp1_out = wp1;
    return w * (1.0 + e);
}

double complex ow_toms917 (double complex z)
{
    const double pi = M_PI, near = 0.01;
    double x = creal (z), y = cimag (z);
    double ympi = y - pi, yppi = y + pi, s = 1.0;
    double complex w, pz, t, r, wp1;
    const double complex I1 = CMPLX (0.0, 1.0);

    /* special values (toms917.cpp:150-207) */
    if (isnan (x) || isnan (y))
        return CMPLX (NAN, NAN);
    if (isinf (x) && x < 0.0 && -pi < y && y <= pi)
        return CMPLX (fabs (y) <= pi / 2.0 ? 0.0 : -0.0, 0.0 <= y ? 0.0 : -0.0);
    if (isinf (x) || isinf (y))
        return CMPLX (x, y);
    if (x == -1.0 && fabs (y) == pi)
        return CMPLX (-1.0, 0.0);

    /* initial approximation by region (toms917.cpp:215-296) */
    if (-2.0 < x && x <= 1.0 && 1.0 < y && y < 2.0 * pi)
    { /* region 1: series about -1 + i*pi */
        pz = conj (csqrt (conj (2.0 * (z + CMPLX (1.0, -pi)))));
        w = -1.0 + (I1 + (1.0 / 3.0 + (-1.0 / 36.0 * I1 + (1.0 / 270.0 + 1.0 / 4320.0 * I1 * pz) * pz) * pz) * pz) * pz;
    }
    else if (-2.0 < x && x <= 1.0 && -2.0 * pi < y && y < -1.0)
    { /* region 2: series about -1 - i*pi */
        pz = conj (csqrt (conj (2.0 * (z + 1.0 + CMPLX (0.0, pi)))));
        w = -1.0 + (-I1 + (1.0 / 3.0 + (1.0 / 36.0 * I1 + (1.0 / 270.0 - 1.0 / 4320.0 * I1 * pz) * pz) * pz) * pz) * pz;
    }
    else if (x <= -2.0 && -pi < y && y <= pi)
    { /* region 3: series in exp(z) */
        pz = cexp (z);
        w = (1.0 + (-1.0 + (3.0 / 2.0 + (-8.0 / 3.0 + 125.0 / 24.0 * pz) * pz) * pz) * pz) * pz;
    }
    else if ((-2.0 < x && x <= 1.0 && -1.0 <= y && y <= 1.0) || (-2.0 < x && (x - 1.0) * (x - 1.0) + y * y <= pi * pi))
    { /* region 4: series about z = 1 */
        pz = z - 1.0;
        w = 1.0 / 2.0 + 1.0 / 2.0 * z + (1.0 / 16.0 + (-1.0 / 192.0 + (-1.0 / 3072.0 + 13.0 / 61440.0 * pz) * pz) * pz) * pz * pz;
    }
    else if (x <= -1.05 && pi < y && y - pi <= -0.75 * (x + 1.0))
    { /* region 5: top wing */
        t = z - CMPLX (0.0, pi);
        w = toms_log_series (t, clog (-t));
    }
    else if (x <= -1.05 && 0.75 * (x + 1.0) < y + pi && y + pi <= 0.0)
    { /* region 6: bottom wing */
        t = z + CMPLX (0.0, pi);
        w = toms_log_series (t, clog (-t));
    }
    else
    { /* region 7: series about infinity */
        w = toms_log_series (z, clog (z));
    }

    /* regularisation near the branch cuts (toms917.cpp:300-343) */
    if (x <= -1.0 + near && (fabs (ympi) <= near || fabs (yppi) <= near))
    {
        s = -1.0;
        if (fabs (ympi) <= near)
        {
            fesetround (FE_UPWARD);
            volatile double v = y - pi;
            if (v <= 0.0)
            {
                fesetround (FE_DOWNWARD);
                v = y - pi;
            }
            z = CMPLX (x, v);
            fesetround (FE_TONEAREST);
        }
        else
        {
            fesetround (FE_UPWARD);
            volatile double v = y + pi;
            if (v <= 0.0)
            {
                fesetround (FE_DOWNWARD);
                v = y + pi;
            }
            z = CMPLX (x, v);
            fesetround (FE_TONEAREST);
        }
    }

    /* iteration one, always (toms917.cpp:347-352) */
    w = s * w;
    w = toms_fsc (z, w, s, &r, &wp1);
    /* iteration two, if the estimated error is not yet below DBL_EPSILON (toms917.cpp:356-364) */
    if (cabs ((2.0 * w * w - 8.0 * w - 1.0) * pow (cabs (r), 4.0)) >= DBL_EPSILON * 72.0 * pow (cabs (wp1), 6.0))
        w = toms_fsc (z, w, s, &r, &wp1);
    return s * w;
}

/* Toms917DiodePair.h:64-67: real( wrightomega( complex<double>(x) ) ) */
double ow_toms917_real (double x) { return creal (ow_toms917 (CMPLX (x, 0.0))); }

void ow_toms917_real_n (const double* x, double* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
        out[i] = ow_toms917_real (x[i]);
}

/* ---- loss + chain rule shared by both precisions ---------------------------------------------- */
/* acc = {g_gamma, g_ell, g_V, sse, st2, count};  out = {dIs, dnabla, dR, dC, loss, mse, esr}
 * MSE: tf.keras.losses.MeanSquaredError (clipper_pot.py:176); ESR: clipper_pot.py:148-156 with
 * eps = np.finfo(float).eps (:145); loss = mse + esr (:177). */
void ow_finalize_grads (const double* acc, int mode, int loss_kind, double fs, double R, double C, double Is, double Vt, double nabla, double* out)
{
    double alpha = 1.0, loss = 0.0, mse = 0.0, esr = 0.0;
    if (mode == 1)
    {
        double N = acc[5] > 0 ? acc[5] : 1.0;
        mse = acc[3] / N;
        alpha = 2.0 / N;
        loss = mse;
        if (loss_kind == 1)
        {
            double energy = acc[4] + DBL_EPSILON;
            esr = sqrt (acc[3] / energy / N);
            loss += esr;
            if (esr > 0.0)
                alpha += 1.0 / (esr * energy * N);
        }
    }
    double Gv = 1.0 / R, Gc = 2.0 * C * fs, Rp = 1.0 / (Gv + Gc), gam = Gv * Rp;
    double dgam_dR = -gam * (1.0 - gam) / R, dgam_dC = -gam * (1.0 - gam) / C;
    double dell_dR = Rp / (R * R), dell_dC = -2.0 * fs * Rp;
    out[0] = alpha * acc[1] / Is;
    out[1] = alpha * acc[2] * Vt;
    out[2] = alpha * (acc[0] * dgam_dR + acc[1] * dell_dR);
    out[3] = alpha * (acc[0] * dgam_dC + acc[1] * dell_dC);
    out[4] = loss;
    out[5] = mse;
    out[6] = esr;
    (void) nabla;
}

/* ---- precision-generic part ------------------------------------------------------------------- */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_ (a, b)

#define REAL float
#define IS_F32 1
#define SFX(name) CAT (name, _f32)
#define LOG(x) logf (x)
#include "wdf_oracle_impl.h"
#undef REAL
#undef IS_F32
#undef SFX
#undef LOG

#define REAL double
#define IS_F32 0
#define SFX(name) CAT (name, _f64)
#define LOG(x) log (x)
#include "wdf_oracle_impl.h"
#undef REAL
#undef IS_F32
#undef SFX
#undef LOG

/* ---- array entry points for the scalar functions (ctypes) ------------------------------------- */
/* kind: 0 omega1, 1 omega2, 2 omega3, 3 omega4, 4 log_approx, 5 exp_approx, 6 log2_approx, 7 pow2_approx */
void ow_omega_f32 (int kind, const float* x, float* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
    {
        float v = x[i];
        switch (kind)
        {
            case 0: out[i] = omega1_f32 (v); break;
            case 1: out[i] = omega2_f32 (v); break;
            case 2: out[i] = omega3_f32 (v); break;
            case 3: out[i] = omega4_f32 (v); break;
            case 4: out[i] = log_approx_f32 (v); break;
            case 5: out[i] = exp_approx_f32 (v); break;
            case 6: out[i] = log2_approx_f32 (v); break;
            default: out[i] = pow2_approx_f32 (v); break;
        }
    }
}

void ow_omega_f64 (int kind, const double* x, double* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
    {
        double v = x[i];
        switch (kind)
        {
            case 0: out[i] = omega1_f64 (v); break;
            case 1: out[i] = omega2_f64 (v); break;
            case 2: out[i] = omega3_f64 (v); break;
            case 3: out[i] = omega4_f64 (v); break;
            case 4: out[i] = log_approx_f64 (v); break;
            case 5: out[i] = exp_approx_f64 (v); break;
            case 6: out[i] = log2_approx_f64 (v); break;
            default: out[i] = pow2_approx_f64 (v); break;
        }
    }
}

int ow_signum_f32 (float v) { return signum_f32 (v); }

/* root law at a fixed port impedance: par = {exact, good, Is, Vt, nDiodes, N_up, N_down} */
void ow_diode_pair_f32 (const float* par, float Rp, const float* a, float* b, int64_t n)
{
    pair_t_f32 d;
    pair_setup_f32 (&d, (int) par[0], (int) par[1], par[2], par[3], par[4], par[5], par[6], Rp);
    for (int64_t i = 0; i < n; ++i)
        b[i] = pair_reflect_f32 (&d, a[i], NULL);
}

void ow_diode_pair_f64 (const double* par, double Rp, const double* a, double* b, int64_t n)
{
    pair_t_f64 d;
    pair_setup_f64 (&d, (int) par[0], (int) par[1], par[2], par[3], par[4], par[5], par[6], Rp);
    for (int64_t i = 0; i < n; ++i)
        b[i] = pair_reflect_f64 (&d, a[i], NULL);
}
