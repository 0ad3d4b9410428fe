// tree_kernels.cu — generic WDF tree interpreter (forward and reverse-mode) on the GPU.
//
// Runs ANY binary tree of the reference's elements (tf_wdf.py:31-214 / wdf_t.h) closed by an
// IdealVoltageSource or a DiodePair root: the RC low-pass of lpf.py:23-28, the voltage divider of
// voltage_divider.py, the plugin's HPF clipper (HPFDiodeClipper.h:25-37), ... The flat post-order
// program (TreeProgram) is a by-value kernel argument; one lane runs one sequence and sweeps the
// node list up (reflected waves, children before parents) and down (incident waves, parents before
// children) once per sample, exactly the order the reference's recursive calls produce.
// The diode clipper itself never comes here: it has the specialised kernels of clipper_kernels.cu.
//
// Reverse mode: pass 1 replays the forward recurrence and records only the capacitor states (and,
// for the plugin probe ordering, the probe's previous incident wave) per sample; pass 2 walks time
// backwards, recomputes the waves of each sample from the recorded state and applies the adjoint of
// every adaptor equation. Per-sample adjoints are collected on the adaptor coefficients (p1R) and
// the root's (ell, V); the chain rule through calc_impedance down to the leaf values R, C runs once,
// in double, in the finalize kernel after the fixed-order batch reduction.
#include "dwdf_kernels.h"
#include "../../include/dwdf.h"

namespace dwdf
{
namespace
{
constexpr int kMaxN = 16;
constexpr int kTreeStride = 24; // doubles per group: [0,16) d/dp1R per node, 16 ell, 17 V, 18 sse, 19 st2

struct TreeWaves
{
    float a[kMaxN], b[kMaxN], bdiff[kMaxN], btemp[kMaxN];
};
struct TreeImp
{
    float R[kMaxN], G[kMaxN], p1R[kMaxN]; // p1R doubles as the Y-parameter's coefficient C
    float ca[kMaxN], cb[kMaxN]; // alpha-transform leaves: a_coef, b_coef; Y-parameter: A, B
};

// calc_impedance, children before parents: tf_wdf.py:77-78,114-115,139-145,168-177,204-206
__device__ __forceinline__ void tree_impedance (const TreeProgram& p, const float* __restrict__ val, const float (&aux)[3][kMaxN], TreeImp& m)
{
    for (int i = 0; i < p.n_nodes; ++i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        switch (p.kind[i])
        {
            case DWDF_RESISTOR:
            case DWDF_RESISTIVE_VS:
                m.R[i] = val[i];
                m.G[i] = 1.0f / m.R[i];
                break;
            case DWDF_CAPACITOR:
                m.R[i] = 1.0f / (2.0f * val[i] * p.fs);
                m.G[i] = 1.0f / m.R[i];
                break;
            case DWDF_RESISTIVE_CS: // wdf_t.h:809-813
                m.R[i] = val[i];
                m.G[i] = 1.0f / m.R[i];
                break;
            case DWDF_INDUCTOR: // wdf_t.h:320-324: Z_L = 2 fs L
                m.R[i] = 2.0f * val[i] * p.fs;
                m.G[i] = 1.0f / m.R[i];
                break;
            case DWDF_CAPACITOR_ALPHA: // wdf_t.h:247-251: 1 / ((1 + alpha) fs C); b_coef, a_coef :205-206
                m.R[i] = 1.0f / ((1.0f + aux[0][i]) * val[i] * p.fs);
                m.G[i] = 1.0f / m.R[i];
                m.cb[i] = (1.0f - aux[0][i]) / 2.0f;
                m.ca[i] = (1.0f + aux[0][i]) / 2.0f;
                break;
            case DWDF_INDUCTOR_ALPHA: // wdf_t.h:411-415: (1 + alpha) fs L
                m.R[i] = (1.0f + aux[0][i]) * val[i] * p.fs;
                m.G[i] = 1.0f / m.R[i];
                m.cb[i] = (1.0f - aux[0][i]) / 2.0f;
                m.ca[i] = (1.0f + aux[0][i]) / 2.0f;
                break;
            case DWDF_Y_PARAMETER: // wdf_t.h:614-627; val = y11, aux = y12, y21, y22
            {
                const float y11 = val[i], y12 = aux[0][i], y21 = aux[1][i], y22 = aux[2][i], R1 = m.R[c1];
                const float den = y22 + R1 * y11 * y22 - R1 * y12 * y21;
                m.R[i] = (R1 * y11 + 1.0f) / den;
                m.G[i] = 1.0f / m.R[i];
                const float rSq = R1 * R1;
                const float num1A = -y22 * rSq * y11 * y11, num2A = y12 * y21 * rSq * y11;
                m.ca[i] = (num1A + num2A + y22) / (den * (R1 * y11 + 1.0f)); // A
                m.cb[i] = -R1 * y12 / (R1 * y11 + 1.0f); // B
                m.p1R[i] = -y21 / den; // C
                break;
            }
            case DWDF_SERIES:
                m.R[i] = m.R[c1] + m.R[c2];
                m.G[i] = 1.0f / m.R[i];
                m.p1R[i] = m.R[c1] / m.R[i];
                break;
            case DWDF_PARALLEL:
                m.G[i] = m.G[c1] + m.G[c2];
                m.R[i] = 1.0f / m.G[i];
                m.p1R[i] = m.G[c1] / m.G[i];
                break;
            default: // DWDF_INVERTER
                m.R[i] = m.R[c1];
                m.G[i] = 1.0f / m.R[i];
                break;
        }
    }
}

__device__ __forceinline__ float tree_root_pair (const TreeProgram& p, const PairConst& pc, float a, PairDeriv* d)
{
    const bool general = ! (p.n_up == 1.0f && p.n_down == 1.0f);
    PairDeriv unused;
    if (d == nullptr)
        d = &unused;
    if (p.root_mode == DWDF_MODE_APPROX_GOOD)
        return pair_reflect<kModeApproxGood, false, false, false> (pc, a, nullptr);
    if (p.root_mode == DWDF_MODE_EXACT)
        return general ? pair_reflect<kModeExact, true, true, false> (pc, a, d) : pair_reflect<kModeExact, false, true, false> (pc, a, d);
    return general ? pair_reflect<kModeApprox, true, true, false> (pc, a, d) : pair_reflect<kModeApprox, false, true, false> (pc, a, d);
}

// voltage (a + b) / 2 (tf_wdf.py:8-10, wdf_t.h:1112-1115) or current (a - b) / (2 R) (wdf_t.h:1119-1123) of the probe node
__device__ __forceinline__ float tree_probe (const TreeProgram& p, const TreeImp& m, const TreeWaves& w)
{
    if (p.probe_current)
        return (w.a[p.probe] - w.b[p.probe]) * (0.5f * m.G[p.probe]);
    return (w.a[p.probe] + w.b[p.probe]) * 0.5f;
}

// One sample. z: reactive-element states by node index (updated). Returns the probe value.
__device__ __forceinline__ float tree_sample (const TreeProgram& p, const TreeImp& m, const PairConst& pc, float x, float* __restrict__ z, TreeWaves& w, PairDeriv* d, float* a_root_out)
{
    const int top = p.n_nodes - 1;
    // up-sweep: reflected(), tf_wdf.py:57-59,86-88,124-126,153-155,185-192,212-214
    for (int i = 0; i <= top; ++i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        switch (p.kind[i])
        {
            case DWDF_RESISTOR: w.b[i] = 0.0f; break;
            case DWDF_RESISTIVE_VS: w.b[i] = (i == p.source) ? x : 0.0f; break;
            case DWDF_RESISTIVE_CS: w.b[i] = (i == p.source) ? m.R[i] * x : 0.0f; break; // wdf_t.h:827-831: b = R Is
            case DWDF_CAPACITOR: w.b[i] = z[i]; break;
            case DWDF_INDUCTOR: w.b[i] = 0.0f - z[i]; break; // wdf_t.h:334-338
            case DWDF_CAPACITOR_ALPHA: w.b[i] = m.cb[i] * w.b[i] + m.ca[i] * z[i]; break; // wdf_t.h:262-266 (b on the right is the previous sample's)
            case DWDF_INDUCTOR_ALPHA: w.b[i] = m.cb[i] * w.b[i] - m.ca[i] * z[i]; break; // wdf_t.h:426-430
            case DWDF_Y_PARAMETER: w.b[i] = m.p1R[i] * w.b[c1]; break; // wdf_t.h:637-641: b = C port1.b
            case DWDF_SERIES: w.b[i] = 0.0f - (w.b[c1] + w.b[c2]); break;
            case DWDF_PARALLEL:
                w.bdiff[i] = w.b[c2] - w.b[c1];
                w.btemp[i] = 0.0f - m.p1R[i] * w.bdiff[i];
                w.b[i] = w.b[c2] + w.btemp[i];
                break;
            default: w.b[i] = 0.0f - w.b[c1]; break;
        }
    }
    // root: tf_wdf.py:23-28 or the diode pair
    const float a_root = w.b[top];
    float b_root;
    if (p.root_kind == DWDF_ROOT_IDEAL_VS)
        b_root = 0.0f - a_root + 2.0f * x;
    else if (p.root_kind == DWDF_ROOT_DIODE_PAIR)
        b_root = tree_root_pair (p, pc, a_root, d);
    else if (p.root_kind == DWDF_ROOT_IDEAL_CS)
        b_root = 2.0f * m.R[top] * x + a_root; // wdf_t.h:777-781: b = 2 R Is + a
    else if (p.root_kind == DWDF_ROOT_SWITCH)
        b_root = p.root_mode != 0 ? 0.0f - a_root : a_root; // wdf_t.h:1094-1098 (root_mode: 1 closed, 0 open)
    else // DWDF_ROOT_DIODE, wdf_t.h:1027-1032 (eq. 10): b = a + 2 R Is - 2 Vt omega4(ln(R Is / Vt) + a / Vt + R Is / Vt)
        b_root = a_root + 2.0f * pc.RIs - pc.twoV * omega4_approx (pc.L + a_root * pc.invV + pc.RIs_overV);
    if (a_root_out != nullptr)
        *a_root_out = a_root;
    float y = 0.0f;
    if (! p.pyorder)
        y = tree_probe (p, m, w); // probe between the sweeps sees the previous incident wave
    // down-sweep: incident(), tf_wdf.py:147-151,179-183,208-210,120-122
    w.a[top] = b_root;
    for (int i = top; i >= 0; --i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        const float xin = w.a[i];
        switch (p.kind[i])
        {
            case DWDF_SERIES:
            {
                const float b1 = w.b[c1] - m.p1R[i] * (xin + w.b[c1] + w.b[c2]);
                w.a[c1] = b1;
                w.a[c2] = 0.0f - (xin + b1);
                break;
            }
            case DWDF_PARALLEL:
            {
                const float b2 = xin + w.btemp[i];
                w.a[c1] = w.bdiff[i] + b2;
                w.a[c2] = b2;
                break;
            }
            case DWDF_INVERTER: w.a[c1] = 0.0f - xin; break;
            case DWDF_Y_PARAMETER: w.a[c1] = m.ca[i] * w.b[c1] + m.cb[i] * xin; break; // wdf_t.h:630-634: port1.incident(A port1.b + B x)
            case DWDF_CAPACITOR:
            case DWDF_INDUCTOR:
            case DWDF_CAPACITOR_ALPHA:
            case DWDF_INDUCTOR_ALPHA: z[i] = xin; break;
            default: break;
        }
    }
    if (p.pyorder)
        y = tree_probe (p, m, w);
    return y;
}

// leaf values; elements with more than one constant keep the others in the slots after their value
// (alpha-transform leaves: alpha; Y-parameter: y12, y21, y22 after y11)
__device__ __forceinline__ void tree_load_values (const TreeProgram& p, const float* __restrict__ params, float* __restrict__ val, float (&aux)[3][kMaxN])
{
    for (int i = 0; i < p.n_nodes; ++i)
    {
        val[i] = p.param[i] >= 0 ? __ldg (params + p.param[i]) : 0.0f;
        const int extra = (p.kind[i] == DWDF_CAPACITOR_ALPHA || p.kind[i] == DWDF_INDUCTOR_ALPHA) ? 1 : (p.kind[i] == DWDF_Y_PARAMETER ? 3 : 0);
        for (int k = 0; k < 3; ++k)
            aux[k][i] = k < extra ? __ldg (params + p.param[i] + 1 + k) : 0.0f;
    }
}

__device__ __forceinline__ void tree_pair_setup (const TreeProgram& p, const float* __restrict__ params, float Rp, PairConst& pc)
{
    if (p.root_kind == DWDF_ROOT_DIODE_PAIR || p.root_kind == DWDF_ROOT_DIODE) // (the single diode shares the constants: V = nDiodes Vt, R Is, ln(R Is / V))
        pair_setup (pc, Rp, __ldg (params + p.slot_Is), p.Vt, __ldg (params + p.slot_nabla), p.n_up, p.n_down, p.n_iter, p.tol);
}

__global__ void __launch_bounds__ (32) tree_forward (const TreeProgram p, const float* __restrict__ params, const float* __restrict__ x, const float* __restrict__ r, float* __restrict__ y, float* __restrict__ state, int64_t B, int T)
{
    const int64_t b = (int64_t) blockIdx.x * 32 + threadIdx.x;
    if (b >= B)
        return;
    float val[kMaxN], z[kMaxN], aux[3][kMaxN];
    TreeImp m;
    TreeWaves w;
    PairConst pc;
    tree_load_values (p, params, val, aux);
    for (int i = 0; i < kMaxN; ++i)
    {
        w.a[i] = w.b[i] = w.bdiff[i] = w.btemp[i] = 0.0f;
        z[i] = (state != nullptr && i < p.n_nodes && p.state_of[i] >= 0) ? state[(int64_t) p.state_of[i] * B + b] : 0.0f;
    }
    if (state != nullptr)
        for (int i = 0; i < p.n_nodes; ++i)
            if (p.kind[i] == DWDF_CAPACITOR_ALPHA || p.kind[i] == DWDF_INDUCTOR_ALPHA)
                w.b[i] = state[(int64_t) (p.state_of[i] + 1) * B + b]; // alpha-transform leaves also carry their previous reflected wave
    if (state != nullptr && ! p.pyorder)
        w.a[p.probe] = state[(int64_t) p.n_states * B + b]; // streaming: the probe's previous incident wave
    tree_impedance (p, val, aux, m);
    tree_pair_setup (p, params, m.R[p.n_nodes - 1], pc);
    const float* xr = x + b * T;
    const float* rr = r != nullptr ? r + b * T : nullptr;
    float* yr = y + b * T;
    for (int n = 0; n < T; ++n)
    {
        if (rr != nullptr)
        { // per-sample resistance channel: set_resistance + calc_impedance every sample (clipper_pot.py:114-117)
            val[p.r_node] = __ldg (rr + n);
            tree_impedance (p, val, aux, m);
            tree_pair_setup (p, params, m.R[p.n_nodes - 1], pc);
        }
        yr[n] = tree_sample (p, m, pc, __ldg (xr + n), z, w, nullptr, nullptr);
    }
    if (state != nullptr)
    {
        for (int i = 0; i < p.n_nodes; ++i)
            if (p.state_of[i] >= 0)
            {
                state[(int64_t) p.state_of[i] * B + b] = z[i];
                if (p.kind[i] == DWDF_CAPACITOR_ALPHA || p.kind[i] == DWDF_INDUCTOR_ALPHA)
                    state[(int64_t) (p.state_of[i] + 1) * B + b] = w.b[i];
            }
        if (! p.pyorder)
            state[(int64_t) p.n_states * B + b] = w.a[p.probe];
    }
}

// Chain rule of ONE sample through calc_impedance: adjoints on the adaptor coefficients (gp[i] = dL/dp1R_i) and
// on ell = ln(Rp Is) -> adjoints on the leaf values (R, C), added to gval[]. Needed per sample when a resistance
// is an input channel (clipper_pot.py:116-117: the impedances differ from sample to sample); with constant
// impedances the same chain runs once, in double, in tree_finalize.
__device__ __forceinline__ void tree_impedance_adjoint (const TreeProgram& p, const float* __restrict__ val, const TreeImp& m, const float* __restrict__ gp, float gell, float* __restrict__ gval)
{
    const int top = p.n_nodes - 1;
    float gR[kMaxN], gG[kMaxN];
    for (int i = 0; i <= top; ++i)
        gR[i] = gG[i] = 0.0f;
    if (p.root_kind == DWDF_ROOT_DIODE_PAIR)
        gR[top] += gell / m.R[top];
    for (int i = top; i >= 0; --i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        switch (p.kind[i])
        {
            case DWDF_SERIES: // R = R1 + R2, G = 1/R, p1R = R1 / R
                gR[i] += gG[i] * (-1.0f / (m.R[i] * m.R[i])) + gp[i] * (-m.R[c1] / (m.R[i] * m.R[i]));
                gR[c1] += gR[i] + gp[i] / m.R[i];
                gR[c2] += gR[i];
                break;
            case DWDF_PARALLEL: // G = G1 + G2, R = 1/G, p1R = G1 / G
                gG[i] += gR[i] * (-1.0f / (m.G[i] * m.G[i])) + gp[i] * (-m.G[c1] / (m.G[i] * m.G[i]));
                gG[c1] += gG[i] + gp[i] / m.G[i];
                gG[c2] += gG[i];
                break;
            case DWDF_INVERTER: // R = R1, G = 1/R
                gR[c1] += gR[i] + gG[i] * (-1.0f / (m.R[i] * m.R[i]));
                break;
            case DWDF_CAPACITOR: // R = 1/(2 C fs), G = 1/R
                gval[i] += (gR[i] + gG[i] * (-1.0f / (m.R[i] * m.R[i]))) * (-m.R[i] / val[i]);
                break;
            case DWDF_INDUCTOR: // R = 2 L fs
                gval[i] += (gR[i] + gG[i] * (-1.0f / (m.R[i] * m.R[i]))) * (m.R[i] / val[i]);
                break;
            default: // Resistor / ResistiveVoltageSource: R = value, G = 1/R
                gval[i] += gR[i] + gG[i] * (-1.0f / (m.R[i] * m.R[i]));
                break;
        }
    }
}

// ---- reverse mode --------------------------------------------------------------------------------
// r != nullptr: per-sample resistance channel on node p.r_node; partials[0..16) then hold dL/d(leaf value) per
// node (already chained, see above) and partials[20] = 1 tells tree_finalize so.
__global__ void __launch_bounds__ (32) tree_adjoint (const TreeProgram p, const float* __restrict__ params, const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ g, int target, int skip, double* __restrict__ partials, float* __restrict__ tape, int64_t B, int T)
{
    const int lane = threadIdx.x;
    const int64_t b = (int64_t) blockIdx.x * 32 + lane;
    const int top = p.n_nodes - 1;
    const int ns1 = p.n_states + 1;
    double acc[kTreeStride];
    for (int k = 0; k < kTreeStride; ++k)
        acc[k] = 0.0;
    if (b < B)
    {
        float val[kMaxN], z[kMaxN], aux[3][kMaxN];
        TreeImp m;
        TreeWaves w;
        PairConst pc;
        tree_load_values (p, params, val, aux);
        for (int i = 0; i < kMaxN; ++i)
            w.a[i] = w.b[i] = w.bdiff[i] = w.btemp[i] = z[i] = 0.0f;
        tree_impedance (p, val, aux, m);
        tree_pair_setup (p, params, m.R[top], pc);
        const float* xr = x + b * T;
        const float* gr = g + b * T;
        const float* rr = r != nullptr ? r + b * T : nullptr;
        // pass 1: forward, recording the state each sample starts from
        for (int n = 0; n < T; ++n)
        {
            if (rr != nullptr)
            {
                val[p.r_node] = __ldg (rr + n);
                tree_impedance (p, val, aux, m);
                tree_pair_setup (p, params, m.R[top], pc);
            }
            for (int i = 0; i <= top; ++i)
                if (p.state_of[i] >= 0)
                    tape[((int64_t) n * ns1 + p.state_of[i]) * B + b] = z[i];
            tape[((int64_t) n * ns1 + p.n_states) * B + b] = w.a[p.probe];
            tree_sample (p, m, pc, __ldg (xr + n), z, w, nullptr, nullptr);
        }
        // pass 2: backwards in time
        float gz[kMaxN]; // adjoint of the state a sample hands to the next one
        for (int i = 0; i < kMaxN; ++i)
            gz[i] = 0.0f;
        float carry = 0.0f; // plugin ordering: adjoint of the probe's previous incident wave
        float facc[kMaxN], fl = 0.0f, fv = 0.0f, fsse = 0.0f, fst2 = 0.0f;
        for (int i = 0; i < kMaxN; ++i)
            facc[i] = 0.0f;
        for (int n = T - 1; n >= 0; --n)
        {
            float pf[kMaxN], pl = 0.0f; // this sample's adjoints on p1R per node and on ell
            for (int i = 0; i < kMaxN; ++i)
                pf[i] = 0.0f;
            if (rr != nullptr)
            {
                val[p.r_node] = __ldg (rr + n);
                tree_impedance (p, val, aux, m);
                tree_pair_setup (p, params, m.R[top], pc);
            }
            for (int i = 0; i <= top; ++i)
                if (p.state_of[i] >= 0)
                    z[i] = tape[((int64_t) n * ns1 + p.state_of[i]) * B + b];
            w.a[p.probe] = tape[((int64_t) n * ns1 + p.n_states) * B + b];
            PairDeriv d { 0.0f, 0.0f, 0.0f };
            float a_root;
            const float yv = tree_sample (p, m, pc, __ldg (xr + n), z, w, &d, &a_root);
            float gy = __ldg (gr + n);
            if (target)
            {
                const bool on = n >= skip;
                const float t = gy;
                gy = on ? yv - t : 0.0f;
                fsse = fma_ (gy, gy, fsse);
                fst2 = on ? fma_ (t, t, fst2) : fst2;
            }
            float aa[kMaxN], ab[kMaxN], abdiff[kMaxN], abtemp[kMaxN];
            for (int i = 0; i < kMaxN; ++i)
                aa[i] = ab[i] = abdiff[i] = abtemp[i] = 0.0f;
            // state hand-over z' = a (Capacitor.incident) and the probe
            for (int i = 0; i <= top; ++i)
                if (p.kind[i] == DWDF_CAPACITOR || p.kind[i] == DWDF_INDUCTOR)
                    aa[i] += gz[i];
            ab[p.probe] += 0.5f * gy;
            if (p.pyorder)
                aa[p.probe] += 0.5f * gy;
            else
            {
                aa[p.probe] += carry; // this sample's incident wave is what the NEXT sample's probe read
                carry = 0.5f * gy;
            }
            // adjoint of the down-sweep, children first
            for (int i = 0; i <= top; ++i)
            {
                const int c1 = p.c1[i], c2 = p.c2[i];
                switch (p.kind[i])
                {
                    case DWDF_SERIES:
                    {
                        const float g2 = aa[c2], gb1 = aa[c1] - g2;
                        ab[c1] += gb1 * (1.0f - m.p1R[i]);
                        ab[c2] -= m.p1R[i] * gb1;
                        aa[i] -= m.p1R[i] * gb1 + g2;
                        pf[i] -= gb1 * (w.a[i] + w.b[c1] + w.b[c2]);
                        break;
                    }
                    case DWDF_PARALLEL:
                    {
                        const float gb2 = aa[c1] + aa[c2];
                        abdiff[i] += aa[c1];
                        abtemp[i] += gb2;
                        aa[i] += gb2;
                        break;
                    }
                    case DWDF_INVERTER: aa[i] -= aa[c1]; break;
                    default: break;
                }
            }
            // root
            const float gbroot = aa[top];
            if (p.root_kind == DWDF_ROOT_IDEAL_VS)
                ab[top] -= gbroot;
            else
            {
                ab[top] += gbroot * fma_ (-2.0f, d.S1, 1.0f);
                pl = gbroot * (-pc.twoV * d.M1);
                fl += pl;
                fv = fma_ (gbroot, d.dV, fv);
            }
            // adjoint of the up-sweep, parents first
            for (int i = top; i >= 0; --i)
            {
                const int c1 = p.c1[i], c2 = p.c2[i];
                switch (p.kind[i])
                {
                    case DWDF_SERIES:
                        ab[c1] -= ab[i];
                        ab[c2] -= ab[i];
                        break;
                    case DWDF_PARALLEL:
                    {
                        const float gbt = abtemp[i] + ab[i];
                        const float gbd = abdiff[i] - m.p1R[i] * gbt;
                        pf[i] -= w.bdiff[i] * gbt;
                        ab[c2] += ab[i] + gbd;
                        ab[c1] -= gbd;
                        break;
                    }
                    case DWDF_INVERTER: ab[c1] -= ab[i]; break;
                    case DWDF_CAPACITOR: gz[i] = ab[i]; break;
                    case DWDF_INDUCTOR: gz[i] = 0.0f - ab[i]; break; // b = -z
                    default: break;
                }
            }
            if (rr == nullptr)
                for (int i = 0; i <= top; ++i)
                    facc[i] += pf[i];
            else
                tree_impedance_adjoint (p, val, m, pf, pl, facc); // facc: dL/d(leaf value)
            if ((n & 15) == 0)
            { // fp32 inside a 16-sample block, double across blocks
                for (int i = 0; i <= top; ++i)
                {
                    acc[i] += (double) facc[i];
                    facc[i] = 0.0f;
                }
                acc[16] += (double) fl;
                acc[17] += (double) fv;
                acc[18] += (double) fsse;
                acc[19] += (double) fst2;
                fl = fv = fsse = fst2 = 0.0f;
            }
        }
    }
    for (int k = 0; k < 20; ++k)
    {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_xor_sync (0xffffffffu, v, o);
        if (lane == 0)
            partials[(int64_t) blockIdx.x * kTreeStride + k] = v;
    }
    if (lane == 0)
        partials[(int64_t) blockIdx.x * kTreeStride + 20] = r != nullptr ? 1.0 : 0.0;
}

// fixed-order reduction over groups, then the chain rule through calc_impedance to the leaf values
__global__ void __launch_bounds__ (256) tree_finalize (const TreeProgram p, const float* __restrict__ params, const double* __restrict__ partials, int64_t n_groups, const double* raw_in, int raw_only, int target, int loss_kind, double count, double* out)
{
    __shared__ double sm[20][256];
    const int tid = threadIdx.x;
    if (raw_in == nullptr)
    {
        double a[20];
        for (int k = 0; k < 20; ++k)
            a[k] = 0.0;
        for (int64_t g = tid; g < n_groups; g += 256)
            for (int k = 0; k < 20; ++k)
                a[k] += partials[g * kTreeStride + k];
        for (int k = 0; k < 20; ++k)
            sm[k][tid] = a[k];
        __syncthreads ();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (tid < o)
                for (int k = 0; k < 20; ++k)
                    sm[k][tid] += sm[k][tid + o];
            __syncthreads ();
        }
    }
    if (tid != 0)
        return;
    double raw[20];
    for (int k = 0; k < 20; ++k)
        raw[k] = raw_in != nullptr ? raw_in[k] : sm[k][0];
    const bool chained = raw_in != nullptr ? raw_in[20] != 0.0 : (n_groups > 0 && partials[20] != 0.0); // leaf gradients already (resistance channel)
    if (raw_in != nullptr)
        count = raw_in[23];
    if (raw_only)
    {
        for (int k = 0; k < DWDF_OUT_LEN; ++k)
            out[k] = k < 20 ? raw[k] : 0.0;
        out[20] = chained ? 1.0 : 0.0;
        out[23] = count;
        return;
    }
    const double sse = raw[18], st2 = raw[19];
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
    const int top = p.n_nodes - 1;
    double R[kMaxN], G[kMaxN], gR[kMaxN], gG[kMaxN];
    for (int i = 0; i <= top; ++i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        const double v = p.param[i] >= 0 ? (double) params[p.param[i]] : 0.0;
        gR[i] = gG[i] = 0.0;
        switch (p.kind[i])
        {
            case DWDF_RESISTOR:
            case DWDF_RESISTIVE_VS: R[i] = v; break;
            case DWDF_CAPACITOR: R[i] = 1.0 / (2.0 * v * (double) p.fs); break;
            case DWDF_INDUCTOR: R[i] = 2.0 * v * (double) p.fs; break;
            case DWDF_SERIES: R[i] = R[c1] + R[c2]; break;
            case DWDF_PARALLEL: R[i] = 1.0 / (1.0 / R[c1] + 1.0 / R[c2]); break;
            default: R[i] = R[c1]; break;
        }
        G[i] = 1.0 / R[i];
    }
    for (int k = 0; k < DWDF_OUT_LEN; ++k)
        out[k] = 0.0;
    if (chained)
    { // per-sample impedances: the kernel has already chained to the leaf values; the resistance channel's node has no parameter gradient
        for (int i = 0; i <= top; ++i)
            if (p.param[i] >= 0 && i != p.r_node)
                out[p.param[i]] += alpha * raw[i];
        if (p.root_kind == DWDF_ROOT_DIODE_PAIR)
        {
            out[p.slot_Is] += alpha * raw[16] / (double) params[p.slot_Is];
            out[p.slot_nabla] += alpha * raw[17] * (double) p.Vt;
        }
        out[DWDF_OUT_LOSS] = loss;
        out[DWDF_OUT_MSE] = mse;
        out[DWDF_OUT_ESR] = esr;
        return;
    }
    if (p.root_kind == DWDF_ROOT_DIODE_PAIR)
    {
        const double acc_l = raw[16], acc_v = raw[17];
        gR[top] += acc_l / R[top]; // ell = ln(Rp Is)
        out[p.slot_Is] += alpha * acc_l / (double) params[p.slot_Is];
        out[p.slot_nabla] += alpha * acc_v * (double) p.Vt;
    }
    for (int i = top; i >= 0; --i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        const double gp = raw[i];
        switch (p.kind[i])
        {
            case DWDF_SERIES: // R = R1 + R2, G = 1/R, p1R = R1 / R
                gR[i] += gG[i] * (-1.0 / (R[i] * R[i])) + gp * (-R[c1] / (R[i] * R[i]));
                gR[c1] += gR[i] + gp / R[i];
                gR[c2] += gR[i];
                break;
            case DWDF_PARALLEL: // G = G1 + G2, R = 1/G, p1R = G1 / G
                gG[i] += gR[i] * (-1.0 / (G[i] * G[i])) + gp * (-G[c1] / (G[i] * G[i]));
                gG[c1] += gG[i] + gp / G[i];
                gG[c2] += gG[i];
                break;
            case DWDF_INVERTER: // R = R1, G = 1/R
                gR[c1] += gR[i] + gG[i] * (-1.0 / (R[i] * R[i]));
                break;
            case DWDF_CAPACITOR: // R = 1/(2 C fs), G = 1/R
            {
                const double gRt = gR[i] + gG[i] * (-1.0 / (R[i] * R[i]));
                out[p.param[i]] += alpha * gRt * (-R[i] / (double) params[p.param[i]]);
                break;
            }
            case DWDF_INDUCTOR: // R = 2 L fs, G = 1/R
            {
                const double gRt = gR[i] + gG[i] * (-1.0 / (R[i] * R[i]));
                out[p.param[i]] += alpha * gRt * (R[i] / (double) params[p.param[i]]);
                break;
            }
            default: // Resistor / ResistiveVoltageSource: R = value, G = 1/R
                out[p.param[i]] += alpha * (gR[i] + gG[i] * (-1.0 / (R[i] * R[i])));
                break;
        }
    }
    out[DWDF_OUT_LOSS] = loss;
    out[DWDF_OUT_MSE] = mse;
    out[DWDF_OUT_ESR] = esr;
}
} // namespace

cudaError_t launch_tree_forward (const TreeProgram& p, const float* params, const float* x, const float* r, float* y, float* state, int64_t B, int64_t T, cudaStream_t stream)
{
    tree_forward<<<(unsigned) ((B + 31) / 32), 32, 0, stream>>> (p, params, x, r, y, state, B, (int) T);
    return cudaGetLastError ();
}

cudaError_t launch_tree_adjoint (const TreeProgram& p, const float* params, const float* x, const float* r, const float* g, bool target, int64_t skip, double* partials, float* tape, int64_t B, int64_t T, cudaStream_t stream)
{
    tree_adjoint<<<(unsigned) ((B + 31) / 32), 32, 0, stream>>> (p, params, x, r, g, target ? 1 : 0, (int) skip, partials, tape, B, (int) T);
    return cudaGetLastError ();
}

cudaError_t launch_tree_finalize (const TreeProgram& p, const float* params, const double* partials, int64_t n_groups, const double* raw_in, bool raw_only, bool target, int loss_kind, double count, double* out, cudaStream_t stream)
{
    tree_finalize<<<1, 256, 0, stream>>> (p, params, partials, n_groups, raw_in, raw_only ? 1 : 0, target ? 1 : 0, loss_kind, count, out);
    return cudaGetLastError ();
}

} // namespace dwdf
