/* TEST INFRASTRUCTURE ONLY — see wdf_oracle.c. Included twice (REAL = float, REAL = double).
 *
 * Required macros: REAL, SFX(name), LOG(x) (libm log of REAL), IS_F32.
 * Every function cites the reference file:line whose arithmetic it restates, operation for
 * operation and in the reference's evaluation order, so that with -ffp-contract=off the fp32
 * instance reproduces the reference's fp32 results and the fp64 instance its fp64 results.
 */

/* ---- omega.h ------------------------------------------------------------------------------- */

/* omega.h:33-42  cubic fit of log2 on [1,2] */
static inline REAL SFX (log2_approx) (REAL x)
{
    const REAL alpha = (REAL) 0.1640425613334452;
    const REAL beta = (REAL) -1.098865286222744;
    const REAL gamma = (REAL) 3.148297929334117;
    const REAL zeta = (REAL) -2.213475204444817;
    return zeta + x * (gamma + x * (beta + x * alpha));
}

/* omega.h:83-92  cubic fit of 2^x on [0,1] */
static inline REAL SFX (pow2_approx) (REAL x)
{
    const REAL alpha = (REAL) 0.07944154167983575;
    const REAL beta = (REAL) 0.2274112777602189;
    const REAL gamma = (REAL) 0.6931471805599453;
    const REAL zeta = (REAL) 1.0;
    return zeta + x * (gamma + x * (beta + x * alpha));
}

#if IS_F32
/* omega.h:49-63  exponent field extraction + cubic on the mantissa (32-bit) */
static inline float SFX (log_approx) (float x)
{
    union { int32_t i; float f; } v;
    v.f = x;
    int32_t ex = v.i & 0x7f800000;
    int32_t e = (ex >> 23) - 127;
    v.i = (v.i - ex) | 0x3f800000;
    return 0.693147180559945f * ((float) e + SFX (log2_approx) (v.f));
}

/* omega.h:99-116  2^floor * cubic(frac) (32-bit). Note the truncating cast and the x<0 adjustment:
 * for a negative integer x this gives l = x-1, f = 1 (not l = x, f = 0). */
static inline float SFX (exp_approx) (float x)
{
    x = 1.442695040888963f * x;
    if (! (x > -126.0f)) /* std::max(-126.0f, x): returns -126 unless -126 < x */
        x = -126.0f;
    union { int32_t i; float f; } v;
    int32_t xi = (int32_t) x;
    int32_t l = x < 0.0f ? xi - 1 : xi;
    float f = x - (float) l;
    v.i = (int32_t) ((uint32_t) (l + 127) << 23);
    return v.f * SFX (pow2_approx) (f);
}
#else
/* omega.h:66-80  (64-bit). The reference's shift/bias constants (>> 53, - 510) are reproduced as
 * written; they are what the reference computes. */
static inline double SFX (log_approx) (double x)
{
    union { int64_t i; double d; } v;
    v.d = x;
    int64_t ex = v.i & 0x7ff0000000000000LL;
    int64_t e = (ex >> 53) - 510;
    v.i = (v.i - ex) | 0x3ff0000000000000LL;
    return 0.693147180559945 * ((double) e + SFX (log2_approx) (v.d));
}

/* omega.h:119-136 (64-bit) */
static inline double SFX (exp_approx) (double x)
{
    x = 1.442695040888963 * x;
    if (! (x > -126.0))
        x = -126.0;
    union { int64_t i; double d; } v;
    int64_t xi = (int64_t) x;
    int64_t l = x < 0.0 ? xi - 1 : xi;
    double d = x - (double) l;
    v.i = (int64_t) ((uint64_t) (l + 1023) << 52);
    return v.d * SFX (pow2_approx) (d);
}
#endif

/* omega.h:139-143 */
static inline REAL SFX (omega1) (REAL x) { return x > (REAL) 0 ? x : (REAL) 0; }

/* omega.h:146-156 */
static inline REAL SFX (omega2) (REAL x)
{
    const REAL x1 = (REAL) -3.684303659906469;
    const REAL x2 = (REAL) 1.972967391708859;
    const REAL a = (REAL) 9.451797158780131e-3;
    const REAL b = (REAL) 1.126446405111627e-1;
    const REAL c = (REAL) 4.451353886588814e-1;
    const REAL d = (REAL) 5.836596684310648e-1;
    return x < x1 ? (REAL) 0 : (x > x2 ? x : d + x * (c + x * (b + x * a)));
}

/* omega.h:159-169 */
static inline REAL SFX (omega3) (REAL x)
{
    const REAL x1 = (REAL) -3.341459552768620;
    const REAL x2 = (REAL) 8.0;
    const REAL a = (REAL) -1.314293149877800e-3;
    const REAL b = (REAL) 4.775931364975583e-2;
    const REAL c = (REAL) 3.631952663804445e-1;
    const REAL d = (REAL) 6.313183464296682e-1;
    return x < x1 ? (REAL) 0 : (x < x2 ? d + x * (c + x * (b + x * a)) : x - SFX (log_approx) (x));
}

/* omega.h:172-177  omega3 + one Newton step on w - exp(x - w) */
static inline REAL SFX (omega4) (REAL x)
{
    const REAL y = SFX (omega3) (x);
    return y - (y - SFX (exp_approx) (x - y)) / (y + (REAL) 1);
}

/* signum.h:5-9 */
static inline int SFX (signum) (REAL v) { return ((REAL) 0 < v) - (v < (REAL) 0); }

/* Wright omega used by a root: kind 0 = omega4 in REAL (wdf_t.h:923), kind 1 = TOMS-917 in double,
 * cast to REAL (Toms917DiodePair.h:64-67). */
static inline REAL SFX (root_omega) (int exact, REAL x)
{
    return exact ? (REAL) ow_toms917_real ((double) x) : SFX (omega4) (x);
}

/* ---- diode pair roots ----------------------------------------------------------------------- */

typedef struct
{
    int exact; /* 0: omega4 (wdf_t.h DiodePairT), 1: TOMS-917 (Toms917DiodePair.h) */
    int good; /* 1: eq. (18) "Good" (wdf_t.h:907-913), only with exact == 0 */
    int general; /* 1: eq. (45) with N_up / N_down (diode_pretraining.py:39-60) */
    REAL Is, Vt, twoVt, oneOverVt; /* Vt here = nDiodes * Vt  (wdf_t.h:875-882) */
    REAL R_Is, R_Is_overVt, logR_Is_overVt; /* wdf_t.h:928-933 */
    REAL n_up, n_down, log_up, log_down; /* general law */
} SFX (pair_t);

/* wdf_t.h:875-882 setDiodeParameters + :928-933 calcImpedanceInternal
 * (Toms917DiodePair.h:28-42 is identical). */
static void SFX (pair_setup) (SFX (pair_t) * d, int exact, int good, REAL Is, REAL Vt, REAL nDiodes, REAL n_up, REAL n_down, REAL Rp)
{
    d->exact = exact;
    d->good = good;
    d->general = ! (n_up == (REAL) 1 && n_down == (REAL) 1);
    d->Is = Is;
    d->Vt = nDiodes * Vt;
    d->twoVt = (REAL) 2 * d->Vt;
    d->oneOverVt = (REAL) 1 / d->Vt;
    d->R_Is = Rp * Is;
    d->R_Is_overVt = d->R_Is * d->oneOverVt;
    d->logR_Is_overVt = LOG (d->R_Is_overVt);
    d->n_up = n_up;
    d->n_down = n_down;
    /* diode_pretraining.py:49-50 */
    d->log_up = LOG (d->R_Is_overVt / n_up);
    d->log_down = LOG (d->R_Is_overVt / n_down);
}

/* b = f(a).  Symmetric "Best": wdf_t.h:917-924 == Toms917DiodePair.h:51-59 (eq. 39);
 * "Good": wdf_t.h:907-913 (eq. 18); general: diode_pretraining.py:39-60 (eq. 45).
 * If w != NULL it receives {omega(u0), omega(u1), mu0, mu1, lambda} for the gradient oracle. */
static inline REAL SFX (pair_reflect) (const SFX (pair_t) * d, REAL a, REAL* w)
{
    REAL lambda = (REAL) SFX (signum) (a);
    if (d->general)
    {
        REAL mu0 = a >= (REAL) 0 ? d->n_down : d->n_up; /* diode_pretraining.py:46 */
        REAL mu1 = a >= (REAL) 0 ? d->n_up : d->n_down; /* :47 */
        REAL l0 = a >= (REAL) 0 ? d->log_down : d->log_up;
        REAL l1 = a >= (REAL) 0 ? d->log_up : d->log_down;
        REAL q0 = lambda * a / (mu0 * d->Vt); /* :53 */
        REAL q1 = lambda * a / (mu1 * d->Vt); /* :54 */
        REAL w0 = SFX (root_omega) (d->exact, l0 + q0);
        REAL w1 = SFX (root_omega) (d->exact, l1 - q1);
        if (w)
        {
            w[0] = w0; w[1] = w1; w[2] = mu0; w[3] = mu1; w[4] = lambda;
        }
        return a - (REAL) 2 * d->Vt * lambda * (mu0 * w0 - mu1 * w1); /* :56-59 */
    }
    if (d->good)
    {
        REAL w0 = SFX (omega4) (d->logR_Is_overVt + lambda * a * d->oneOverVt + d->R_Is_overVt);
        if (w)
        {
            w[0] = w0; w[1] = 0; w[2] = 1; w[3] = 1; w[4] = lambda;
        }
        return a + (REAL) 2 * lambda * (d->R_Is - d->Vt * w0);
    }
    REAL lambda_a_over_vt = lambda * a * d->oneOverVt;
    REAL w0 = SFX (root_omega) (d->exact, d->logR_Is_overVt + lambda_a_over_vt);
    REAL w1 = SFX (root_omega) (d->exact, d->logR_Is_overVt - lambda_a_over_vt);
    if (w)
    {
        w[0] = w0; w[1] = w1; w[2] = 1; w[3] = 1; w[4] = lambda;
    }
    return a - d->twoVt * lambda * (w0 - w1);
}

/* ---- WDF elements and adaptors: a tree interpreter over the port protocol -------------------- */

/* One node of the element tree in post-order (children before parents). Fields a, b, R follow
 * tf_wdf.py; G is the C++ twin's admittance (wdf_t.h:53-64). */
typedef struct
{
    int kind, c1, c2;
    REAL value; /* R of Resistor / ResistiveVoltageSource, C of Capacitor */
    REAL a, b, R, G;
    REAL z; /* Capacitor state (tf_wdf.py:112) */
    REAL Vs; /* source voltage (tf_wdf.py:48-49) */
    REAL p1R, b_diff, b_temp; /* adaptor coefficients / carried temporaries (tf_wdf.py:146,189-190) */
    REAL aux[3]; /* alpha-transform leaves: alpha; Y-parameter: y12, y21, y22 (value = y11) */
    REAL cA, cB, cC; /* alpha leaves: a_coef, b_coef (wdf_t.h:205-206); Y-parameter: A, B, C (wdf_t.h:621-626) */
} SFX (node_t);

static void SFX (calc_impedance) (SFX (node_t) * t, int i, REAL fs)
{
    SFX (node_t)* n = &t[i];
    switch (n->kind)
    {
        case OW_RESISTOR: /* tf_wdf.py:77-78 (R is the variable itself); wdf_t.h:91-95 */
        case OW_RESVS: /* tf_wdf.py:42-43; wdf_t.h:719-723 */
            n->R = n->value;
            n->G = (REAL) 1 / n->R;
            break;
        case OW_CAPACITOR: /* tf_wdf.py:114-115  R = 1/(C*(2*FS)); wdf_t.h:166-170 */
            n->R = (REAL) 1 / ((REAL) 2 * n->value * fs);
            n->G = (REAL) 1 / n->R;
            break;
        case OW_RESCS: /* wdf_t.h:809-813 */
            n->R = n->value;
            n->G = (REAL) 1 / n->R;
            break;
        case OW_INDUCTOR: /* wdf_t.h:320-324 */
            n->R = (REAL) 2 * n->value * fs;
            n->G = (REAL) 1 / n->R;
            break;
        case OW_CAPACITOR_ALPHA: /* wdf_t.h:247-251, coefficients :205-206 */
            n->R = (REAL) 1 / (((REAL) 1 + n->aux[0]) * n->value * fs);
            n->G = (REAL) 1 / n->R;
            n->cB = ((REAL) 1 - n->aux[0]) / (REAL) 2;
            n->cA = ((REAL) 1 + n->aux[0]) / (REAL) 2;
            break;
        case OW_INDUCTOR_ALPHA: /* wdf_t.h:411-415 */
            n->R = ((REAL) 1 + n->aux[0]) * n->value * fs;
            n->G = (REAL) 1 / n->R;
            n->cB = ((REAL) 1 - n->aux[0]) / (REAL) 2;
            n->cA = ((REAL) 1 + n->aux[0]) / (REAL) 2;
            break;
        case OW_YPARAM: /* wdf_t.h:614-627 */
        {
            SFX (calc_impedance) (t, n->c1, fs);
            REAL y11 = n->value, y12 = n->aux[0], y21 = n->aux[1], y22 = n->aux[2], R1 = t[n->c1].R;
            REAL den = y22 + R1 * y11 * y22 - R1 * y12 * y21;
            n->R = (R1 * y11 + (REAL) 1) / den;
            n->G = (REAL) 1 / n->R;
            REAL rSq = R1 * R1;
            REAL num1A = -y22 * rSq * y11 * y11;
            REAL num2A = y12 * y21 * rSq * y11;
            n->cA = (num1A + num2A + y22) / (den * (R1 * y11 + (REAL) 1));
            n->cB = -R1 * y12 / (R1 * y11 + (REAL) 1);
            n->cC = -y21 / den;
            break;
        }
        case OW_SERIES: /* tf_wdf.py:139-145; wdf_t.h:525-530 */
            SFX (calc_impedance) (t, n->c1, fs);
            SFX (calc_impedance) (t, n->c2, fs);
            n->R = t[n->c1].R + t[n->c2].R;
            n->G = (REAL) 1 / n->R;
            n->p1R = t[n->c1].R / n->R;
            break;
        case OW_PARALLEL: /* tf_wdf.py:168-177; wdf_t.h:465-470 */
            SFX (calc_impedance) (t, n->c1, fs);
            SFX (calc_impedance) (t, n->c2, fs);
            n->G = t[n->c1].G + t[n->c2].G;
            n->R = (REAL) 1 / n->G;
            n->p1R = t[n->c1].G / n->G;
            break;
        case OW_INVERTER: /* tf_wdf.py:204-206; wdf_t.h:573-577 */
            SFX (calc_impedance) (t, n->c1, fs);
            n->R = t[n->c1].R;
            n->G = (REAL) 1 / n->R;
            break;
        default: break;
    }
}

static REAL SFX (reflected) (SFX (node_t) * t, int i)
{
    SFX (node_t)* n = &t[i];
    switch (n->kind)
    {
        case OW_RESISTOR: n->b = (REAL) 0; break; /* tf_wdf.py:86-88 */
        case OW_RESVS: n->b = n->Vs; break; /* tf_wdf.py:57-59 */
        case OW_CAPACITOR: n->b = n->z; break; /* tf_wdf.py:124-126 */
        case OW_RESCS: n->b = n->R * n->Vs; break; /* wdf_t.h:827-831 (Vs holds the source current) */
        case OW_INDUCTOR: n->b = -n->z; break; /* wdf_t.h:334-338 */
        case OW_CAPACITOR_ALPHA: n->b = n->cB * n->b + n->cA * n->z; break; /* wdf_t.h:262-266 */
        case OW_INDUCTOR_ALPHA: n->b = n->cB * n->b - n->cA * n->z; break; /* wdf_t.h:426-430 */
        case OW_YPARAM: n->b = n->cC * SFX (reflected) (t, n->c1); break; /* wdf_t.h:637-641 */
        case OW_SERIES: /* tf_wdf.py:153-155 */
        {
            REAL r1 = SFX (reflected) (t, n->c1);
            REAL r2 = SFX (reflected) (t, n->c2);
            n->b = (REAL) 0 - (r1 + r2);
            break;
        }
        case OW_PARALLEL: /* tf_wdf.py:185-192 */
        {
            REAL b1 = SFX (reflected) (t, n->c1);
            REAL b2 = SFX (reflected) (t, n->c2);
            n->b_diff = b2 - b1;
            n->b_temp = (REAL) 0 - n->p1R * n->b_diff;
            n->b = b2 + n->b_temp;
            break;
        }
        case OW_INVERTER: n->b = (REAL) 0 - SFX (reflected) (t, n->c1); break; /* tf_wdf.py:212-214 */
        default: break;
    }
    return n->b;
}

static void SFX (incident) (SFX (node_t) * t, int i, REAL x)
{
    SFX (node_t)* n = &t[i];
    switch (n->kind)
    {
        case OW_RESISTOR:
        case OW_RESVS: n->a = x; break; /* tf_wdf.py:83-84, 54-55 */
        case OW_CAPACITOR: n->a = x; n->z = n->a; break; /* tf_wdf.py:120-122 */
        case OW_RESCS: n->a = x; break; /* wdf_t.h:821-824 */
        case OW_INDUCTOR: /* wdf_t.h:327-331 */
        case OW_CAPACITOR_ALPHA: /* wdf_t.h:254-258 */
        case OW_INDUCTOR_ALPHA: n->a = x; n->z = n->a; break; /* wdf_t.h:418-422 */
        case OW_YPARAM: n->a = x; SFX (incident) (t, n->c1, n->cA * t[n->c1].b + n->cB * x); break; /* wdf_t.h:630-634 */
        case OW_SERIES: /* tf_wdf.py:147-151 */
        {
            REAL b1 = t[n->c1].b - n->p1R * (x + t[n->c1].b + t[n->c2].b);
            SFX (incident) (t, n->c1, b1);
            SFX (incident) (t, n->c2, (REAL) 0 - (x + b1));
            n->a = x;
            break;
        }
        case OW_PARALLEL: /* tf_wdf.py:179-183 */
        {
            REAL b2 = x + n->b_temp;
            SFX (incident) (t, n->c1, n->b_diff + b2);
            SFX (incident) (t, n->c2, b2);
            n->a = x;
            break;
        }
        case OW_INVERTER: SFX (incident) (t, n->c1, (REAL) 0 - x); n->a = x; break; /* tf_wdf.py:208-210 */
        default: break;
    }
}

/* Runs a whole circuit: tree (post-order, top = n_nodes-1) closed by a root.
 *   root_kind: OW_ROOT_IDEAL_VS (tf_wdf.py:13-28; the input x drives the root) or OW_ROOT_DIODE_PAIR
 *              (the input x drives ResistiveVoltageSource node `source`).
 *   root_par:  {exact, good, Is, Vt, nDiodes, N_up, N_down}
 *   r_in:      optional per-sample resistance channel for node `r_node`, with calc_impedance every
 *              sample (clipper_pot.py:114-117); NULL => impedances computed once (lpf.py:38).
 *   ordering:  OW_ORDER_PYTHON probe after tree.incident (lpf.py:42-45, clipper_pot.py:121-123),
 *              OW_ORDER_PLUGIN probe between root.incident and tree.incident (DiodeClipperWDF.cpp:26-28).
 *   x, y:      (B, T) batch-major; every sequence starts from reset state (clipper_pot.py:110-111).
 */
int SFX (ow_tree_run) (int n_nodes, const int* kind, const int* c1, const int* c2, const REAL* value, REAL fs, int root_kind, const REAL* root_par, int source, int probe, int ordering, int r_node, const REAL* x, const REAL* r_in, REAL* y, int64_t nB, int64_t nT)
{
    if (n_nodes <= 0 || n_nodes > OW_MAX_NODES)
        return 1;
    SFX (node_t) t[OW_MAX_NODES];
    const int top = n_nodes - 1;
    for (int64_t s = 0; s < nB; ++s)
    {
        memset (t, 0, sizeof (t));
        for (int i = 0; i < n_nodes; ++i)
        {
            t[i].kind = kind[i];
            t[i].c1 = c1[i];
            t[i].c2 = c2[i];
            t[i].value = value[i];
        }
        SFX (calc_impedance) (t, top, fs);
        SFX (pair_t) dp;
        if (root_kind == OW_ROOT_DIODE_PAIR)
            SFX (pair_setup) (&dp, (int) root_par[0], (int) root_par[1], root_par[2], root_par[3], root_par[4], root_par[5], root_par[6], t[top].R);
        REAL root_a = 0, root_b = 0;
        for (int64_t n = 0; n < nT; ++n)
        {
            REAL xin = x[s * nT + n];
            if (r_in != NULL)
            {
                t[r_node].value = r_in[s * nT + n];
                SFX (calc_impedance) (t, top, fs);
                if (root_kind == OW_ROOT_DIODE_PAIR)
                    SFX (pair_setup) (&dp, (int) root_par[0], (int) root_par[1], root_par[2], root_par[3], root_par[4], root_par[5], root_par[6], t[top].R);
            }
            if (root_kind == OW_ROOT_DIODE_PAIR)
                t[source].Vs = xin;
            root_a = SFX (reflected) (t, top); /* root.incident(tree.reflected()) */
            if (root_kind == OW_ROOT_IDEAL_VS)
                root_b = (REAL) 0 - root_a + (REAL) 2 * xin; /* tf_wdf.py:26-28 */
            else
                root_b = SFX (pair_reflect) (&dp, root_a, NULL);
            if (ordering == OW_ORDER_PLUGIN)
                y[s * nT + n] = (t[probe].a + t[probe].b) * (REAL) 0.5; /* tf_wdf.py:8-10 */
            SFX (incident) (t, top, root_b);
            if (ordering != OW_ORDER_PLUGIN)
                y[s * nT + n] = (t[probe].a + t[probe].b) * (REAL) 0.5;
        }
    }
    return 0;
}

/* The same executor for the full chowdsp_wdf element set: aux = 3 extra constants per node (alpha; y12, y21, y22),
 * root kinds OW_ROOT_IDEAL_CS (wdf_t.h:746-784), OW_ROOT_DIODE (:987-1072, eq. 10, always omega4) and OW_ROOT_SWITCH
 * (:1076-1106; root_par[0] != 0: closed), a ResistiveCurrentSource as the driven leaf, and a current probe
 * (a - b) / (2 R), wdf_t.h:1119-1123. root_par for the diode: {_, _, Is, Vt, nDiodes}. */
int SFX (ow_tree_run_ext) (int n_nodes, const int* kind, const int* c1, const int* c2, const REAL* value, const REAL* aux, REAL fs, int root_kind, const REAL* root_par, int source, int probe, int probe_current, int ordering, const REAL* x, REAL* y, int64_t nB, int64_t nT)
{
    if (n_nodes <= 0 || n_nodes > OW_MAX_NODES)
        return 1;
    SFX (node_t) t[OW_MAX_NODES];
    const int top = n_nodes - 1;
    for (int64_t s = 0; s < nB; ++s)
    {
        memset (t, 0, sizeof (t));
        for (int i = 0; i < n_nodes; ++i)
        {
            t[i].kind = kind[i];
            t[i].c1 = c1[i];
            t[i].c2 = c2[i];
            t[i].value = value[i];
            for (int k = 0; k < 3; ++k)
                t[i].aux[k] = aux[i * 3 + k];
        }
        SFX (calc_impedance) (t, top, fs);
        SFX (pair_t) dp;
        REAL twoR_Is = 0;
        if (root_kind == OW_ROOT_DIODE_PAIR)
            SFX (pair_setup) (&dp, (int) root_par[0], (int) root_par[1], root_par[2], root_par[3], root_par[4], root_par[5], root_par[6], t[top].R);
        if (root_kind == OW_ROOT_DIODE)
        { /* wdf_t.h:1003-1010,1047-1052: Vt = nDiodes Vt; twoR_Is, R_Is_overVt, log */
            SFX (pair_setup) (&dp, 0, 0, root_par[2], root_par[3], root_par[4], (REAL) 1, (REAL) 1, t[top].R);
            twoR_Is = (REAL) 2 * t[top].R * root_par[2];
        }
        for (int64_t n = 0; n < nT; ++n)
        {
            REAL xin = x[s * nT + n];
            if (source >= 0)
                t[source].Vs = xin;
            REAL root_a = SFX (reflected) (t, top), root_b;
            switch (root_kind)
            {
                case OW_ROOT_IDEAL_VS: root_b = (REAL) 0 - root_a + (REAL) 2 * xin; break;
                case OW_ROOT_IDEAL_CS: root_b = (REAL) 2 * t[top].R * xin + root_a; break; /* wdf_t.h:777-781 */
                case OW_ROOT_SWITCH: root_b = root_par[0] != (REAL) 0 ? -root_a : root_a; break; /* wdf_t.h:1094-1098 */
                case OW_ROOT_DIODE: root_b = root_a + twoR_Is - dp.twoVt * SFX (omega4) (dp.logR_Is_overVt + root_a * dp.oneOverVt + dp.R_Is_overVt); break; /* wdf_t.h:1027-1032 */
                default: root_b = SFX (pair_reflect) (&dp, root_a, NULL); break;
            }
            #define OW_PROBE() (probe_current ? (t[probe].a - t[probe].b) * ((REAL) 0.5 * t[probe].G) : (t[probe].a + t[probe].b) * (REAL) 0.5)
            if (ordering == OW_ORDER_PLUGIN)
                y[s * nT + n] = OW_PROBE ();
            SFX (incident) (t, top, root_b);
            if (ordering != OW_ORDER_PLUGIN)
                y[s * nT + n] = OW_PROBE ();
            #undef OW_PROBE
        }
    }
    return 0;
}

/* ---- the diode clipper: Parallel(ResistiveVoltageSource, Capacitor) + DiodePair root --------- */

typedef struct
{
    int exact, good, ordering;
    REAL fs, R, C, Is, Vt, nabla, n_up, n_down;
} SFX (clip_par_t);

/* Forward only, one range of sequences. Element code paths: the Parallel adaptor of tf_wdf.py:158-192
 * over ResistiveVoltageSource (:31-59) and Capacitor (:91-126), loop of clipper_pot.py:103-127 /
 * DiodeClipperWDF.cpp:22-29, written out for this one topology (same operations, same order as the
 * tree interpreter above performs for it — tests assert the two agree bit for bit). */
static void SFX (clipper_rows) (const SFX (clip_par_t) * p, const REAL* x, REAL* y, REAL* z_final, int64_t b0, int64_t b1, int64_t nT)
{
    /* calc_impedance: tf_wdf.py:168-177 with P1 = Vs, P2 = C */
    REAL Rv = p->R, Gv = (REAL) 1 / Rv;
    REAL Rc = (REAL) 1 / ((REAL) 2 * p->C * p->fs), Gc = (REAL) 1 / Rc;
    REAL G = Gv + Gc, Rp = (REAL) 1 / G, p1R = Gv / G;
    SFX (pair_t) dp;
    SFX (pair_setup) (&dp, p->exact, p->good, p->Is, p->Vt, p->nabla, p->n_up, p->n_down, Rp);
    for (int64_t s = b0; s < b1; ++s)
    {
        REAL z = 0, ca = 0; /* Capacitor.z, Capacitor.a after reset */
        const REAL* xs = x + s * nT;
        REAL* ys = y + s * nT;
        for (int64_t n = 0; n < nT; ++n)
        {
            REAL b_diff = z - xs[n]; /* b2 - b1 */
            REAL b_temp = (REAL) 0 - p1R * b_diff;
            REAL a = z + b_temp; /* P1.b -> root.a */
            REAL b = SFX (pair_reflect) (&dp, a, NULL);
            if (p->ordering == OW_ORDER_PLUGIN)
                ys[n] = (ca + z) * (REAL) 0.5; /* voltage(C) before tree.incident: C.a = previous incident, C.b = z */
            REAL zn = b + b_temp; /* P2.incident(b2): Capacitor.z = a */
            ca = zn;
            if (p->ordering != OW_ORDER_PLUGIN)
                ys[n] = (ca + z) * (REAL) 0.5; /* voltage(C) after: C.a = new, C.b = z */
            z = zn;
        }
        if (z_final)
            z_final[s] = z;
    }
}

typedef struct
{
    const SFX (clip_par_t) * p;
    const REAL* x;
    REAL* y;
    int64_t b0, b1, nT;
} SFX (clip_job_t);

static void* SFX (clip_thread) (void* arg)
{
    SFX (clip_job_t)* j = (SFX (clip_job_t)*) arg;
    SFX (clipper_rows) (j->p, j->x, j->y, NULL, j->b0, j->b1, j->nT);
    return NULL;
}

int SFX (ow_clipper_forward) (int exact, int good, int ordering, REAL fs, REAL R, REAL C, REAL Is, REAL Vt, REAL nabla, REAL n_up, REAL n_down, const REAL* x, REAL* y, int64_t nB, int64_t nT, int n_threads)
{
    SFX (clip_par_t) p = { exact, good, ordering, fs, R, C, Is, Vt, nabla, n_up, n_down };
    if (n_threads <= 1)
    {
        SFX (clipper_rows) (&p, x, y, NULL, 0, nB, nT);
        return 0;
    }
    if (n_threads > OW_MAX_THREADS)
        n_threads = OW_MAX_THREADS;
    pthread_t th[OW_MAX_THREADS];
    SFX (clip_job_t) job[OW_MAX_THREADS];
    for (int t = 0; t < n_threads; ++t)
    {
        job[t].p = &p; job[t].x = x; job[t].y = y; job[t].nT = nT;
        job[t].b0 = nB * t / n_threads;
        job[t].b1 = nB * (t + 1) / n_threads;
        pthread_create (&th[t], NULL, SFX (clip_thread), &job[t]);
    }
    for (int t = 0; t < n_threads; ++t)
        pthread_join (th[t], NULL);
    return 0;
}

/* ---- forward + reverse-mode gradient (what tf.GradientTape does for clipper_pot.py:246-269) --- */
/* Tape-based: the forward sweep records per-sample intermediates, the reverse sweep walks the tape.
 * Differentiated parameters: (Is, nabla, R, C) — new surface (the reference differentiates only R, C
 * and MLP weights), so the derivative of omega is DEFINED as w/(1+w) applied to the
 * omega value the root mode produced (custom-gradient convention), for both root modes.
 *
 * acc[0..2] += sum_n G[n+1] * dz'/d{gamma, ell, V}   with ell = ln(Rp*Is), V = nabla*Vt
 * acc[3] += sum (y-t)^2, acc[4] += sum t^2, acc[5] += count   (only when target mode)
 * gy_or_target: mode 0 => upstream gradient gy (B,T); mode 1 => target, gy := (y - target) [unit scale],
 *               samples n < skip excluded (clipper_pot.py:232,248).
 * gx (optional) receives dL/dx. Always accumulates in double. */
static void SFX (clipper_grad_rows) (const SFX (clip_par_t) * p, const REAL* x, const REAL* gy_or_target, int mode, int64_t skip, REAL* y, REAL* gx, double* acc, int64_t b0, int64_t b1, int64_t nT)
{
    REAL Rv = p->R, Gv = (REAL) 1 / Rv;
    REAL Rc = (REAL) 1 / ((REAL) 2 * p->C * p->fs), Gc = (REAL) 1 / Rc;
    REAL G = Gv + Gc, Rp = (REAL) 1 / G, p1R = Gv / G;
    SFX (pair_t) dp;
    SFX (pair_setup) (&dp, p->exact, 0, p->Is, p->Vt, p->nabla, p->n_up, p->n_down, Rp);
    const double gam = (double) p1R, V = (double) dp.Vt;
    double* tape = (double*) malloc (sizeof (double) * (size_t) nT * 6);
    for (int64_t s = b0; s < b1; ++s)
    {
        const REAL* xs = x + s * nT;
        REAL z = 0;
        for (int64_t n = 0; n < nT; ++n)
        {
            REAL w[5];
            REAL b_diff = z - xs[n];
            REAL b_temp = (REAL) 0 - p1R * b_diff;
            REAL a = z + b_temp;
            REAL b = SFX (pair_reflect) (&dp, a, w);
            REAL zn = b + b_temp;
            REAL yn = p->ordering == OW_ORDER_PLUGIN ? z : (zn + z) * (REAL) 0.5;
            if (y)
                y[s * nT + n] = yn;
            double wp0 = (double) w[0] / (1.0 + (double) w[0]), wp1 = (double) w[1] / (1.0 + (double) w[1]);
            double S1 = wp0 + wp1;
            double M1 = (double) w[4] * ((double) w[2] * wp0 - (double) w[3] * wp1);
            double* tp = tape + 6 * n;
            tp[0] = S1; /* f'(a) = 1 - 2 S1 */
            tp[1] = M1; /* df/d ell = -2 V M1 */
            tp[2] = ((double) b - (double) a) / V + 2.0 * M1 + 2.0 * (double) a / V * S1; /* df/dV at fixed ell */
            tp[3] = (double) xs[n] - (double) z; /* x - z */
            tp[4] = (double) yn;
            tp[5] = 0;
            z = zn;
        }
        /* reverse sweep */
        double Gn1 = 0; /* adjoint of z[n+1] */
        const REAL* gt = gy_or_target + s * nT;
        for (int64_t n = nT - 1; n >= 0; --n)
        {
            const double* tp = tape + 6 * n;
            double gyn;
            if (mode == 0)
                gyn = (double) gt[n];
            else if (n >= skip)
            {
                gyn = tp[4] - (double) gt[n];
                acc[3] += gyn * gyn;
                acc[4] += (double) gt[n] * (double) gt[n];
                acc[5] += 1.0;
            }
            else
                gyn = 0.0;
            double Gz = 0; /* direct contribution of y[n] to z[n] */
            if (p->ordering == OW_ORDER_PLUGIN)
                Gz = gyn; /* y[n] = z[n] */
            else
            {
                Gn1 += 0.5 * gyn; /* y[n] = (z[n+1] + z[n]) / 2 */
                Gz = 0.5 * gyn;
            }
            double fp = 1.0 - 2.0 * tp[0];
            acc[0] += Gn1 * tp[3] * (fp + 1.0);
            acc[1] += Gn1 * (-2.0 * V * tp[1]);
            acc[2] += Gn1 * tp[2];
            if (gx)
                gx[s * nT + n] = (REAL) (Gn1 * gam * (fp + 1.0));
            Gn1 = Gz + Gn1 * (fp * (1.0 - gam) - gam);
        }
    }
    free (tape);
}

typedef struct
{
    const SFX (clip_par_t) * p;
    const REAL *x, *gt;
    REAL *y, *gx;
    int mode;
    int64_t skip, b0, b1, nT;
    double acc[8];
} SFX (grad_job_t);

static void* SFX (grad_thread) (void* arg)
{
    SFX (grad_job_t)* j = (SFX (grad_job_t)*) arg;
    SFX (clipper_grad_rows) (j->p, j->x, j->gt, j->mode, j->skip, j->y, j->gx, j->acc, j->b0, j->b1, j->nT);
    return NULL;
}

/* out[0..3] = dL/d(Is, nabla, R, C); out[4] = loss; out[5] = mse; out[6] = esr;
 * raw[0..5] (optional) = the six raw sums before the chain rule (gamma, ell, V, sse, st2, count).
 * loss_kind (mode 1 only): 0 = MSE, 1 = MSE + ESR (clipper_pot.py:148-156,176-177, eps = DBL_EPSILON :145). */
int SFX (ow_clipper_grad) (int exact, int ordering, REAL fs, REAL R, REAL C, REAL Is, REAL Vt, REAL nabla, REAL n_up, REAL n_down, const REAL* x, const REAL* gy_or_target, int mode, int loss_kind, int64_t skip, REAL* y, REAL* gx, double* out, double* raw, int64_t nB, int64_t nT, int n_threads)
{
    SFX (clip_par_t) p = { exact, 0, ordering, fs, R, C, Is, Vt, nabla, n_up, n_down };
    if (n_threads < 1)
        n_threads = 1;
    if (n_threads > OW_MAX_THREADS)
        n_threads = OW_MAX_THREADS;
    if ((int64_t) n_threads > nB)
        n_threads = nB > 0 ? (int) nB : 1;
    pthread_t th[OW_MAX_THREADS];
    SFX (grad_job_t) job[OW_MAX_THREADS];
    for (int t = 0; t < n_threads; ++t)
    {
        memset (&job[t], 0, sizeof (job[t]));
        job[t].p = &p; job[t].x = x; job[t].gt = gy_or_target; job[t].y = y; job[t].gx = gx;
        job[t].mode = mode; job[t].skip = skip; job[t].nT = nT;
        job[t].b0 = nB * t / n_threads;
        job[t].b1 = nB * (t + 1) / n_threads;
        if (n_threads > 1)
            pthread_create (&th[t], NULL, SFX (grad_thread), &job[t]);
        else
            SFX (grad_thread) (&job[t]);
    }
    double acc[6] = { 0, 0, 0, 0, 0, 0 };
    for (int t = 0; t < n_threads; ++t)
    {
        if (n_threads > 1)
            pthread_join (th[t], NULL);
        for (int k = 0; k < 6; ++k)
            acc[k] += job[t].acc[k];
    }
    if (raw)
        memcpy (raw, acc, sizeof (acc));
    ow_finalize_grads (acc, mode, loss_kind, (double) fs, (double) R, (double) C, (double) Is, (double) Vt, (double) nabla, out);
    return 0;
}
