// tree_jit.cu — run-time specialisation of the tree interpreter (dwdf_program_specialize).
//
// tree_kernels.cu runs ANY tree by sweeping a node list per sample: wave arrays indexed by node number live in local
// memory and every node costs a dispatch, so even the three-node RC low-pass of lpf.py:23-28 runs at 2 % of HBM
// bandwidth, and its reverse mode writes a per-sample tape to HBM. The circuit, however, is fixed for the life of a
// program. This file turns the flat post-order program into straight-line CUDA source — one named scalar per wave,
// impedances and adaptor coefficients computed once per launch, the reverse-mode step derived node by node in the
// same pass — compiles it with NVRTC for sm_100a and loads the cubin next to the built-in kernels:
//   * forward: one lane per sequence, one-warp CTAs, [32 sequences x 32 samples] TMA tiles through a 3-slot
//     shared-memory ring, results written in place (the data movement of the diode-clipper kernels); all states
//     in registers; state checkpoints every 16 samples for the reverse sweep;
//   * reverse mode: NO tape. Segments of 16 samples, last to first: reload the segment's checkpoint, replay the 16
//     samples keeping each sample's start states and the root's derivative pieces, then walk the segment backwards,
//     recomputing a sample's (linear) waves from its start state and applying the adjoint of every adaptor equation. Reads x and the target
//     (or dL/dy): 8 B/sample, the algorithmic minimum. Same partial-sum layout as tree_adjoint, so tree_finalize
//     (fixed-order reduction, chain rule through calc_impedance) is shared;
//   * direct-global-access twins of both kernels for ragged T / unaligned rows.
// Covered: every element and root of the interpreter (Resistor, Capacitor, Inductor, alpha-transform C / L, resistive and
// ideal voltage / current sources, Series, Parallel, Inverter, Y-parameter; diode pair of every law and mode, single diode,
// switch), voltage or current probe, both probe orderings; reverse mode where the interpreter has it (the wdf_py elements and
// the inductor, ideal-voltage-source or diode-pair root, voltage probe). Only the per-sample resistance channel stays on the
// interpreter. NVRTC is loaded with dlopen: without it dwdf_program_specialize reports DWDF_ERR_UNSUPPORTED and the
// program keeps running on the interpreter (still on the GPU — there is no CPU path anywhere).
#include "dwdf_kernels.h"
#include "tree_jit.h"
#include "../../include/dwdf.h"

#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace dwdf
{
namespace
{
// the device headers, embedded at build time (csrc/Makefile: *.inc = the header as a raw string literal)
const char kJitMathSrc[] =
#include "jit_math.inc"
    ;
const char kJitTmaSrc[] =
#include "jit_tma.inc"
    ;

struct Gen
{
    std::string s;
    void f (const char* fmt, ...)
    {
        char buf[1024];
        va_list ap;
        va_start (ap, fmt);
        vsnprintf (buf, sizeof (buf), fmt, ap);
        va_end (ap);
        s += buf;
    }
};

std::string hexf (float v)
{
    char b[64];
    snprintf (b, sizeof (b), "%af", (double) v); // hexadecimal float literal: exact
    return b;
}

bool is_adaptor (int k) { return k == DWDF_SERIES || k == DWDF_PARALLEL; }
bool is_reactive (int k) { return k == DWDF_CAPACITOR || k == DWDF_INDUCTOR; }

// index of node i among the adaptors (its slot in JC::p / pf[])
int adaptor_slot (const TreeProgram& p, int i)
{
    int j = 0;
    for (int k = 0; k < i; ++k)
        j += is_adaptor (p.kind[k]) ? 1 : 0;
    return j;
}

bool is_alpha (int k) { return k == DWDF_CAPACITOR_ALPHA || k == DWDF_INDUCTOR_ALPHA; }
// reverse mode exists for the wdf_py element set and the inductor, closed by an ideal voltage source or a diode pair, voltage probe
bool jit_differentiable (const TreeProgram& p)
{
    if (p.probe_current != 0 || (p.root_kind != DWDF_ROOT_IDEAL_VS && p.root_kind != DWDF_ROOT_DIODE_PAIR))
        return false;
    for (int i = 0; i < p.n_nodes; ++i)
        if (p.kind[i] > DWDF_INDUCTOR)
            return false;
    return true;
}

const char* pair_mode (const TreeProgram& p) { return p.root_mode == DWDF_MODE_EXACT ? "kModeExact" : (p.root_mode == DWDF_MODE_APPROX_GOOD ? "kModeApproxGood" : "kModeApprox"); }
const char* pair_general (const TreeProgram& p) { return (p.n_up == 1.0f && p.n_down == 1.0f) ? "false" : "true"; }

// reflected(): children before parents (tf_wdf.py:57-59,86-88,124-126,153-155,185-192,212-214), then the root
// root: 0 = forward only, 1 = with the derivative pieces, recorded in `rec` (the reverse sweep's replay), 2 = taken from `rec`
void emit_up_and_root (Gen& g, const TreeProgram& p, int root)
{
    const int top = p.n_nodes - 1;
    for (int i = 0; i <= top; ++i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        switch (p.kind[i])
        {
            case DWDF_RESISTOR: g.f ("    const float b%d = 0.0f;\n", i); break;
            case DWDF_RESISTIVE_VS: g.f ("    const float b%d = %s;\n", i, i == p.source ? "x" : "0.0f"); break;
            case DWDF_CAPACITOR: g.f ("    const float b%d = z[%d];\n", i, p.state_of[i]); break;
            case DWDF_INDUCTOR: g.f ("    const float b%d = 0.0f - z[%d];\n", i, p.state_of[i]); break; // wdf_t.h:334-338
            case DWDF_RESISTIVE_CS: // wdf_t.h:827-831: b = R Is
                if (i == p.source)
                    g.f ("    const float b%d = c.R[%d] * x;\n", i, i);
                else
                    g.f ("    const float b%d = 0.0f;\n", i);
                break;
            case DWDF_CAPACITOR_ALPHA: g.f ("    const float b%d = c.cb[%d] * z[%d] + c.ca[%d] * z[%d];\n", i, i, p.state_of[i] + 1, i, p.state_of[i]); break; // wdf_t.h:262-266 (z[s + 1]: the previous reflected wave)
            case DWDF_INDUCTOR_ALPHA: g.f ("    const float b%d = c.cb[%d] * z[%d] - c.ca[%d] * z[%d];\n", i, i, p.state_of[i] + 1, i, p.state_of[i]); break; // wdf_t.h:426-430
            case DWDF_Y_PARAMETER: g.f ("    const float b%d = c.cc[%d] * b%d;\n", i, i, c1); break; // wdf_t.h:637-641: b = C port1.b
            case DWDF_SERIES: g.f ("    const float b%d = 0.0f - (b%d + b%d);\n", i, c1, c2); break;
            case DWDF_PARALLEL:
                g.f ("    const float bd%d = b%d - b%d;\n", i, c2, c1);
                g.f ("    const float bt%d = 0.0f - c.p[%d] * bd%d;\n", i, adaptor_slot (p, i), i);
                g.f ("    const float b%d = b%d + bt%d;\n", i, c2, i);
                break;
            default: g.f ("    const float b%d = 0.0f - b%d;\n", i, c1); break; // inverter
        }
    }
    g.f ("    const float a_root = b%d;\n", top);
    if (p.root_kind == DWDF_ROOT_IDEAL_VS)
        g.f ("    const float b_root = 0.0f - a_root + 2.0f * x;\n"); // tf_wdf.py:23-28
    else if (p.root_kind == DWDF_ROOT_IDEAL_CS)
        g.f ("    const float b_root = 2.0f * c.R[%d] * x + a_root;\n", top); // wdf_t.h:777-781: b = 2 R Is + a
    else if (p.root_kind == DWDF_ROOT_SWITCH)
        g.f ("    const float b_root = %s;\n", p.root_mode != 0 ? "0.0f - a_root" : "a_root"); // wdf_t.h:1094-1098 (closed: -a, open: a)
    else if (p.root_kind == DWDF_ROOT_DIODE) // wdf_t.h:1027-1032 (eq. 10): b = a + 2 R Is - 2 Vt omega4(ln(R Is / Vt) + a / Vt + R Is / Vt)
        g.f ("    const float b_root = a_root + 2.0f * c.pc.RIs - c.pc.twoV * omega4_approx (c.pc.L + a_root * c.pc.invV + c.pc.RIs_overV);\n");
    else if (root == 2)
        g.f ("    const float b_root = rec.b;\n    const PairDeriv d { rec.S1, rec.M1, rec.dV };\n    (void) a_root;\n");
    else if (p.root_mode == DWDF_MODE_APPROX_GOOD)
        g.f ("    const float b_root = pair_reflect<kModeApproxGood, false, false, false> (c.pc, a_root, nullptr);\n");
    else if (root == 1)
    {
        g.f ("    PairDeriv d { 0.0f, 0.0f, 0.0f };\n");
        g.f ("    const float b_root = pair_reflect<%s, %s, true, false> (c.pc, a_root, &d);\n", pair_mode (p), pair_general (p));
        g.f ("    rec.b = b_root, rec.S1 = d.S1, rec.M1 = d.M1, rec.dV = d.dV;\n");
    }
    else
        g.f ("    const float b_root = pair_reflect<%s, %s, false, false> (c.pc, a_root, nullptr);\n", pair_mode (p), pair_general (p));
}

// incident(): parents before children (tf_wdf.py:147-151,179-183,208-210,120-122); reactive leaves hand their state on in zn[]
void emit_down (Gen& g, const TreeProgram& p, bool states)
{
    const int top = p.n_nodes - 1;
    g.f ("    const float a%d = b_root;\n", top);
    for (int i = top; i >= 0; --i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        switch (p.kind[i])
        {
            case DWDF_SERIES:
                g.f ("    const float t%d = b%d - c.p[%d] * (a%d + b%d + b%d);\n", i, c1, adaptor_slot (p, i), i, c1, c2);
                g.f ("    const float a%d = t%d;\n", c1, i);
                g.f ("    const float a%d = 0.0f - (a%d + t%d);\n", c2, i, i);
                break;
            case DWDF_PARALLEL:
                g.f ("    const float t%d = a%d + bt%d;\n", i, i, i);
                g.f ("    const float a%d = bd%d + t%d;\n", c1, i, i);
                g.f ("    const float a%d = t%d;\n", c2, i);
                break;
            case DWDF_INVERTER: g.f ("    const float a%d = 0.0f - a%d;\n", c1, i); break;
            case DWDF_Y_PARAMETER: g.f ("    const float a%d = c.ca[%d] * b%d + c.cb[%d] * a%d;\n", c1, i, c1, i, i); break; // wdf_t.h:630-634: port1.incident(A port1.b + B x)
            case DWDF_CAPACITOR:
            case DWDF_INDUCTOR:
                if (states)
                    g.f ("    zn[%d] = a%d;\n", p.state_of[i], i);
                break;
            case DWDF_CAPACITOR_ALPHA:
            case DWDF_INDUCTOR_ALPHA:
                if (states)
                    g.f ("    zn[%d] = a%d;\n    zn[%d] = b%d;\n", p.state_of[i], i, p.state_of[i] + 1, i);
                break;
            default: break;
        }
    }
}

} // namespace

bool tree_jit_supported (const TreeProgram& p)
{
    // every element and root of the interpreter; only the per-sample resistance channel (calc_impedance every sample) stays there
    if (p.n_nodes < 1 || p.n_nodes > 16 || p.r_node >= 0)
        return false;
    if (p.root_kind != DWDF_ROOT_IDEAL_VS && p.root_kind != DWDF_ROOT_DIODE_PAIR && p.root_kind != DWDF_ROOT_IDEAL_CS && p.root_kind != DWDF_ROOT_DIODE && p.root_kind != DWDF_ROOT_SWITCH)
        return false;
    for (int i = 0; i < p.n_nodes; ++i)
        if (p.kind[i] < DWDF_RESISTOR || p.kind[i] > DWDF_Y_PARAMETER)
            return false;
    return true;
}

// The circuit-specific part: constants, one forward sample, one reverse-mode sample.
std::string tree_jit_generate (const TreeProgram& p)
{
    Gen g;
    const int top = p.n_nodes - 1;
    int n_adapt = 0;
    for (int i = 0; i <= top; ++i)
        n_adapt += is_adaptor (p.kind[i]) ? 1 : 0;
    const bool diode = p.root_kind == DWDF_ROOT_DIODE_PAIR;
    g.f ("// generated by libdwdf (tree_jit.cu) for a %d-node tree, root kind %d mode %d, %s probe ordering\n", p.n_nodes, p.root_kind, p.root_mode, p.pyorder ? "python" : "plugin");
    g.f ("#include \"dwdf_math.cuh\"\n#include \"dwdf_tma.cuh\"\nnamespace dwdf\n{\n");
    g.f ("constexpr int kNS = %d, kNS1 = kNS + 1; // reactive states, + the probe's previous incident wave (plugin ordering reads it)\n", p.n_states);
    g.f ("constexpr int kNP = %d, kNPreal = %d; // adaptor coefficients (p1R): one gradient accumulator each\n", n_adapt > 0 ? n_adapt : 1, n_adapt);
    g.f ("constexpr bool kUnrollSegment = %s; // a diode root is too much code to unroll 16 samples of\n", diode ? "false" : "true");
    g.f ("constexpr int kNN = %d; // nodes\n", p.n_nodes);
    g.f ("struct JC\n{\n    float p[kNP]; // adaptor coefficients p1R\n    float R[kNN], ca[kNN], cb[kNN], cc[kNN]; // port resistances; alpha-transform a_coef / b_coef, Y-parameter A / B / C\n    float Gprobe;\n    PairConst pc;\n};\n");
    // ---- constants: calc_impedance, children before parents (tf_wdf.py:77-78,114-115,139-145,168-177,204-206)
    g.f ("__device__ __forceinline__ void jit_consts (const float* __restrict__ params, JC& c)\n{\n    const float fs = %s;\n    c.p[0] = 0.0f;\n", hexf (p.fs).c_str ());
    for (int i = 0; i <= top; ++i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i];
        switch (p.kind[i])
        {
            case DWDF_RESISTOR:
            case DWDF_RESISTIVE_VS: g.f ("    const float R%d = __ldg (params + %d), G%d = 1.0f / R%d;\n", i, p.param[i], i, i); break;
            case DWDF_CAPACITOR: g.f ("    const float R%d = 1.0f / (2.0f * __ldg (params + %d) * fs), G%d = 1.0f / R%d;\n", i, p.param[i], i, i); break;
            case DWDF_INDUCTOR: g.f ("    const float R%d = 2.0f * __ldg (params + %d) * fs, G%d = 1.0f / R%d;\n", i, p.param[i], i, i); break;
            case DWDF_SERIES:
                g.f ("    const float R%d = R%d + R%d, G%d = 1.0f / R%d;\n    c.p[%d] = R%d / R%d;\n", i, c1, c2, i, i, adaptor_slot (p, i), c1, i);
                break;
            case DWDF_PARALLEL:
                g.f ("    const float G%d = G%d + G%d, R%d = 1.0f / G%d;\n    c.p[%d] = G%d / G%d;\n", i, c1, c2, i, i, adaptor_slot (p, i), c1, i);
                break;
            case DWDF_RESISTIVE_CS: g.f ("    const float R%d = __ldg (params + %d), G%d = 1.0f / R%d;\n", i, p.param[i], i, i); break; // wdf_t.h:809-813
            case DWDF_CAPACITOR_ALPHA: // wdf_t.h:247-251: 1 / ((1 + alpha) fs C); b_coef, a_coef :205-206
                g.f ("    const float al%d = __ldg (params + %d);\n", i, p.param[i] + 1);
                g.f ("    const float R%d = 1.0f / ((1.0f + al%d) * __ldg (params + %d) * fs), G%d = 1.0f / R%d;\n", i, i, p.param[i], i, i);
                g.f ("    c.cb[%d] = (1.0f - al%d) / 2.0f;\n    c.ca[%d] = (1.0f + al%d) / 2.0f;\n", i, i, i, i);
                break;
            case DWDF_INDUCTOR_ALPHA: // wdf_t.h:411-415: (1 + alpha) fs L
                g.f ("    const float al%d = __ldg (params + %d);\n", i, p.param[i] + 1);
                g.f ("    const float R%d = (1.0f + al%d) * __ldg (params + %d) * fs, G%d = 1.0f / R%d;\n", i, i, p.param[i], i, i);
                g.f ("    c.cb[%d] = (1.0f - al%d) / 2.0f;\n    c.ca[%d] = (1.0f + al%d) / 2.0f;\n", i, i, i, i);
                break;
            case DWDF_Y_PARAMETER: // wdf_t.h:614-627
                g.f ("    const float y11_%d = __ldg (params + %d), y12_%d = __ldg (params + %d), y21_%d = __ldg (params + %d), y22_%d = __ldg (params + %d);\n", i, p.param[i], i, p.param[i] + 1, i, p.param[i] + 2, i,
                     p.param[i] + 3);
                g.f ("    const float den%d = y22_%d + R%d * y11_%d * y22_%d - R%d * y12_%d * y21_%d;\n", i, i, c1, i, i, c1, i, i);
                g.f ("    const float R%d = (R%d * y11_%d + 1.0f) / den%d, G%d = 1.0f / R%d;\n", i, c1, i, i, i, i);
                g.f ("    const float rSq%d = R%d * R%d;\n", i, c1, c1);
                g.f ("    const float n1A%d = -y22_%d * rSq%d * y11_%d * y11_%d, n2A%d = y12_%d * y21_%d * rSq%d * y11_%d;\n", i, i, i, i, i, i, i, i, i, i);
                g.f ("    c.ca[%d] = (n1A%d + n2A%d + y22_%d) / (den%d * (R%d * y11_%d + 1.0f));\n", i, i, i, i, i, c1, i);
                g.f ("    c.cb[%d] = -R%d * y12_%d / (R%d * y11_%d + 1.0f);\n    c.cc[%d] = -y21_%d / den%d;\n", i, c1, i, c1, i, i, i, i);
                break;
            default: g.f ("    const float R%d = R%d, G%d = 1.0f / R%d;\n", i, c1, i, i); break;
        }
        g.f ("    c.R[%d] = R%d;\n    (void) G%d;\n", i, i, i);
    }
    g.f ("    (void) fs;\n    c.Gprobe = G%d;\n", p.probe);
    if (diode || p.root_kind == DWDF_ROOT_DIODE)
        g.f ("    pair_setup (c.pc, R%d, __ldg (params + %d), %s, __ldg (params + %d), %s, %s, %d, %s);\n", top, p.slot_Is, hexf (p.Vt).c_str (), p.slot_nabla, hexf (p.n_up).c_str (), hexf (p.n_down).c_str (), p.n_iter,
             hexf (p.tol).c_str ());
    g.f ("}\n\n");
    // ---- one sample: root.incident(tree.reflected()); tree.incident(root.reflected()); y = voltage(probe)
    g.f ("struct JRec\n{\n    float b, S1, M1, dV; // the root's reflected wave and derivative pieces of one sample (diode-pair root)\n};\n");
    const bool differentiable = jit_differentiable (p);
    for (int rec = 0; rec < 2; ++rec)
    {
        if (rec == 1 && ! differentiable)
        { // forward-only circuit: reverse mode is refused at the API (the kernels of the skeleton still have to compile)
            g.f ("__device__ __forceinline__ float jit_step_rec (const JC& c, float x, float (&z)[kNS1], JRec&) { return jit_step (c, x, z); }\n\n");
            break;
        }
        if (rec == 0)
            g.f ("__device__ __forceinline__ float jit_step (const JC& c, float x, float (&z)[kNS1])\n{\n    float zn[kNS1];\n");
        else // the reverse sweep's replay: the same sample, keeping what the adjoint of the root needs
            g.f ("__device__ __forceinline__ float jit_step_rec (const JC& c, float x, float (&z)[kNS1], JRec& rec)\n{\n    float zn[kNS1];\n    (void) rec;\n");
        emit_up_and_root (g, p, rec);
        emit_down (g, p, true);
        // probe after tree.incident (clipper_pot.py:113-124) or between the sweeps (DiodeClipperWDF.cpp:22-29: the previous incident
        // wave); voltage (a + b) / 2 (tf_wdf.py:8-10) or current (a - b) / (2 R) (wdf_t.h:1119-1123)
        {
            char ap[32];
            if (p.pyorder)
                snprintf (ap, sizeof (ap), "a%d", p.probe);
            else
                snprintf (ap, sizeof (ap), "z[kNS]");
            if (p.probe_current)
                g.f ("    const float y = (%s - b%d) * (0.5f * c.Gprobe);\n", ap, p.probe);
            else
                g.f ("    const float y = (%s + b%d) * 0.5f;\n", ap, p.probe);
        }
        g.f ("    zn[kNS] = a%d;\n", p.probe);
        g.f ("#pragma unroll\n    for (int k = 0; k < kNS1; ++k)\n        z[k] = zn[k];\n    return y;\n}\n\n");
    }
    // ---- reverse mode of one sample: waves recomputed from the sample's start state, then the adjoint of every equation.
    // gz: in = dL/d(states handed to the next sample), out = dL/d(states this sample started from).
    if (! differentiable)
    {
        g.f ("__device__ __forceinline__ void jit_step_adj (const JC&, float, const float (&)[kNS1], const JRec&, float, float (&)[kNS1], float (&)[kNP], float&, float&) {}\n\n");
        g.f ("__device__ __forceinline__ int jit_pf_node (int) { return 0; }\n} // namespace dwdf\n\n");
        return g.s;
    }
    g.f ("__device__ __forceinline__ void jit_step_adj (const JC& c, float x, const float (&z)[kNS1], const JRec& rec, float gy, float (&gz)[kNS1], float (&pf)[kNP], float& fl, float& fv)\n{\n");
    g.f ("    (void) rec;\n");
    emit_up_and_root (g, p, 2);
    emit_down (g, p, false);
    for (int i = 0; i <= top; ++i)
    {
        g.f ("    float aa%d = 0.0f, ab%d = 0.0f;\n", i, i);
        if (p.kind[i] == DWDF_PARALLEL)
            g.f ("    float abd%d = 0.0f, abt%d = 0.0f;\n", i, i);
    }
    for (int i = 0; i <= top; ++i)
        if (is_reactive (p.kind[i]))
            g.f ("    aa%d += gz[%d];\n", i, p.state_of[i]); // z' = a (Capacitor.incident)
    g.f ("    ab%d += 0.5f * gy;\n", p.probe);
    if (p.pyorder)
        g.f ("    aa%d += 0.5f * gy;\n    const float gprobe = 0.0f;\n", p.probe);
    else
        g.f ("    aa%d += gz[kNS];\n    const float gprobe = 0.5f * gy;\n", p.probe); // this sample's incident wave is what the NEXT sample's probe read
    // adjoint of the down-sweep, children first
    for (int i = 0; i <= top; ++i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i], j = adaptor_slot (p, i);
        switch (p.kind[i])
        {
            case DWDF_SERIES:
                g.f ("    {\n        const float g2 = aa%d, gb1 = aa%d - g2;\n", c2, c1);
                g.f ("        ab%d += gb1 * (1.0f - c.p[%d]);\n        ab%d -= c.p[%d] * gb1;\n        aa%d -= c.p[%d] * gb1 + g2;\n", c1, j, c2, j, i, j);
                g.f ("        pf[%d] -= gb1 * (a%d + b%d + b%d);\n    }\n", j, i, c1, c2);
                break;
            case DWDF_PARALLEL:
                g.f ("    {\n        const float gb2 = aa%d + aa%d;\n        abd%d += aa%d;\n        abt%d += gb2;\n        aa%d += gb2;\n    }\n", c1, c2, i, c1, i, i);
                break;
            case DWDF_INVERTER: g.f ("    aa%d -= aa%d;\n", i, c1); break;
            default: break;
        }
    }
    // root
    g.f ("    const float gbroot = aa%d;\n", top);
    if (! diode)
        g.f ("    ab%d -= gbroot;\n    (void) fl;\n    (void) fv;\n", top);
    else
        g.f ("    ab%d += gbroot * fma_ (-2.0f, d.S1, 1.0f);\n    fl += gbroot * (-c.pc.twoV * d.M1);\n    fv = fma_ (gbroot, d.dV, fv);\n", top);
    // adjoint of the up-sweep, parents first
    for (int i = top; i >= 0; --i)
    {
        const int c1 = p.c1[i], c2 = p.c2[i], j = adaptor_slot (p, i);
        switch (p.kind[i])
        {
            case DWDF_SERIES: g.f ("    ab%d -= ab%d;\n    ab%d -= ab%d;\n", c1, i, c2, i); break;
            case DWDF_PARALLEL:
                g.f ("    {\n        const float gbt = abt%d + ab%d;\n        const float gbd = abd%d - c.p[%d] * gbt;\n        pf[%d] -= bd%d * gbt;\n        ab%d += ab%d + gbd;\n        ab%d -= gbd;\n    }\n", i, i, i, j, j, i, c2,
                     i, c1);
                break;
            case DWDF_INVERTER: g.f ("    ab%d -= ab%d;\n", c1, i); break;
            case DWDF_CAPACITOR: g.f ("    gz[%d] = ab%d;\n", p.state_of[i], i); break;
            case DWDF_INDUCTOR: g.f ("    gz[%d] = 0.0f - ab%d;\n", p.state_of[i], i); break; // b = -z
            default: break;
        }
    }
    g.f ("    gz[kNS] = gprobe;\n}\n\n");
    // which node each coefficient accumulator belongs to (the partials' layout is tree_adjoint's: one slot per node)
    g.f ("__device__ __forceinline__ int jit_pf_node (int j)\n{\n    switch (j)\n    {\n");
    for (int i = 0; i <= top; ++i)
        if (is_adaptor (p.kind[i]))
            g.f ("        case %d: return %d;\n", adaptor_slot (p, i), i);
    g.f ("        default: return 0;\n    }\n}\n} // namespace dwdf\n\n");
    return g.s;
}

namespace
{
// The circuit-independent part: tile movement, segment replay, reductions.
const char kJitSkeleton[] = R"DWDFJIT(
namespace dwdf
{
constexpr int kSeg = 16; // checkpoint spacing = reverse-mode segment (dwdf_kernels.h)
__device__ __forceinline__ uint32_t chunk128 (uint32_t tile, int lane, int c) { return tile + lane * 128 + ((c ^ (lane & 7)) << 4); } // CU_TENSOR_MAP_SWIZZLE_128B
__device__ __forceinline__ uint32_t chunk64 (uint32_t tile, int lane, int c) { return tile + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4); } // CU_TENSOR_MAP_SWIZZLE_64B
__device__ __forceinline__ double warp_sum (double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync (0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void load_state (float (&z)[kNS1], const float* __restrict__ state, long long B, long long b, bool valid)
{
#pragma unroll
    for (int k = 0; k < kNS1; ++k)
        z[k] = (state != nullptr && valid) ? state[k * B + b] : 0.0f;
}
__device__ __forceinline__ void store_state (const float (&z)[kNS1], float* __restrict__ state, long long B, long long b, bool valid)
{
    if (state != nullptr && valid)
#pragma unroll
        for (int k = 0; k < kNS1; ++k)
            state[k * B + b] = z[k];
}
__device__ __forceinline__ void store_ckpt (const float (&z)[kNS1], float* __restrict__ ckpt, int seg, long long B, long long b)
{
    float* ck = ckpt + ((long long) seg * kNS1) * B + b;
#pragma unroll
    for (int k = 0; k < kNS1; ++k)
        ck[k * B] = z[k];
}

// ---- forward: [32 sequences x 32 samples] TMA tiles, 3-slot ring, results written in place ------------------------
extern "C" __global__ void __launch_bounds__ (32) jit_tree_forward_tma (const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, const float* __restrict__ params, float* __restrict__ ckpt, float* __restrict__ state, long long B, int T)
{
    constexpr int kStages = 3, kTileBytes = 32 * 32 * 4;
    __shared__ __align__ (1024) unsigned char smem[kStages * kTileBytes];
    __shared__ __align__ (8) unsigned long long bar_mem[kStages];
    const int lane = threadIdx.x;
    const int b0 = blockIdx.x * 32;
    const long long b = (long long) b0 + lane;
    const bool valid = b < B;
    JC c;
    jit_consts (params, c);
    const uint32_t tiles = smem_u32 (smem), bars = smem_u32 (bar_mem);
    if (lane == 0)
    {
        tma_prefetch_desc (&tmx);
        tma_prefetch_desc (&tmy);
        for (int s = 0; s < kStages; ++s)
            mbar_init (bars + 8 * s, 1);
        fence_mbar_init ();
    }
    __syncwarp ();
    float z[kNS1];
    load_state (z, state, B, b, valid);
    const int ntiles = (T + 31) / 32;
    auto load = [&] (int j) {
        const int s = j % kStages;
        mbar_expect_tx (bars + 8 * s, kTileBytes);
        tma_load_2d (tiles + s * kTileBytes, &tmx, j * 32, b0, bars + 8 * s);
    };
    if (lane == 0)
        for (int j = 0; j < kStages - 1 && j < ntiles; ++j)
            load (j);
    for (int i = 0; i < ntiles; ++i)
    {
        const int s = i % kStages;
        mbar_wait (bars + 8 * s, (i / kStages) & 1);
        const uint32_t tile = tiles + s * kTileBytes;
        const int nch = min (8, (T - i * 32) >> 2);
#pragma unroll 2
        for (int cc = 0; cc < 8; ++cc)
        {
            if (cc < nch)
            {
                if ((cc & 3) == 0 && ckpt != nullptr && valid)
                    store_ckpt (z, ckpt, i * 2 + (cc >> 2), B, b);
                const uint32_t addr = chunk128 (tile, lane, cc);
                const float4 v = lds128 (addr);
                float4 o;
                o.x = jit_step (c, v.x, z);
                o.y = jit_step (c, v.y, z);
                o.z = jit_step (c, v.z, z);
                o.w = jit_step (c, v.w, z);
                sts128 (addr, o);
            }
        }
        fence_proxy_async (); // my st.shared results -> visible to the TMA unit
        __syncwarp ();
        if (lane == 0)
        {
            tma_store_2d (&tmy, i * 32, b0, tile);
            tma_commit ();
            const int jn = i + kStages - 1; // goes into the slot tile i-1 was stored from
            if (jn < ntiles)
            {
                tma_wait_read<1> ();
                load (jn);
            }
        }
    }
    if (lane == 0)
        tma_wait_all<0> ();
    store_state (z, state, B, b, valid);
}

// any T, any alignment: a lane walks its own row
extern "C" __global__ void __launch_bounds__ (32) jit_tree_forward_direct (const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ params, float* __restrict__ ckpt, float* __restrict__ state, long long B, int T)
{
    const long long b = (long long) blockIdx.x * 32 + threadIdx.x;
    if (b >= B)
        return;
    JC c;
    jit_consts (params, c);
    float z[kNS1];
    load_state (z, state, B, b, true);
    const float* xr = x + b * T;
    float* yr = y + b * T;
#pragma unroll 4
    for (int n = 0; n < T; ++n)
    {
        if ((n & (kSeg - 1)) == 0 && ckpt != nullptr)
            store_ckpt (z, ckpt, n / kSeg, B, b);
        yr[n] = jit_step (c, __ldg (xr + n), z);
    }
    store_state (z, state, B, b, true);
}

// ---- reverse mode -------------------------------------------------------------------------------------------------
struct JitAcc
{
    double pf[kNP], l = 0.0, v = 0.0, sse = 0.0, st2 = 0.0;
};

struct TileIO2
{
    uint32_t xt, gt;
    int lane;
    __device__ __forceinline__ float4 x4 (int cc) const { return lds128 (chunk64 (xt, lane, cc)); }
    __device__ __forceinline__ float4 g4 (int cc) const { return lds128 (chunk64 (gt, lane, cc)); }
};
struct GlobalIO2
{
    const float* __restrict__ xr;
    const float* __restrict__ gr;
    int n0, T;
    __device__ __forceinline__ float at (const float* __restrict__ p, int n) const { return n < T ? __ldg (p + n) : 0.0f; }
    __device__ __forceinline__ float4 row4 (const float* __restrict__ p, int cc) const
    {
        const int n = n0 + cc * 4;
        return make_float4 (at (p, n), at (p, n + 1), at (p, n + 2), at (p, n + 3));
    }
    __device__ __forceinline__ float4 x4 (int cc) const { return row4 (xr, cc); }
    __device__ __forceinline__ float4 g4 (int cc) const { return row4 (gr, cc); }
};

// One segment [n0, n0 + 16): replay from its checkpoint keeping each sample's start state, then backwards.
template <bool UNROLL, class IO>
__device__ __forceinline__ void jit_adjoint_segment (const JC& c, const IO& io, const float (&z0)[kNS1], float (&gz)[kNS1], int n0, int T, int skip, int target, JitAcc& acc)
{
    float zs[kSeg][kNS1], ys[kSeg], xs[kSeg];
    JRec rs[kSeg];
    float z[kNS1];
#pragma unroll
    for (int k = 0; k < kNS1; ++k)
        z[k] = z0[k];
    auto replay4 = [&] (int cc) {
        const float4 v = io.x4 (cc);
        const float x4[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const int s = cc * 4 + k;
            xs[s] = x4[k];
#pragma unroll
            for (int q = 0; q < kNS1; ++q)
                zs[s][q] = z[q];
            ys[s] = jit_step_rec (c, x4[k], z, rs[s]);
        }
    };
    float pf[kNP], fl = 0.0f, fv = 0.0f, fsse = 0.0f, fst2 = 0.0f;
#pragma unroll
    for (int j = 0; j < kNP; ++j)
        pf[j] = 0.0f;
    auto reverse4 = [&] (int cc) {
        const float4 gv = io.g4 (cc);
        const float g4[4] = { gv.x, gv.y, gv.z, gv.w };
#pragma unroll
        for (int k = 3; k >= 0; --k)
        {
            const int s = cc * 4 + k, n = n0 + s;
            if (n < T)
            {
                float gy = g4[k];
                if (target)
                {
                    const bool on = n >= skip;
                    const float t = gy;
                    gy = on ? ys[s] - t : 0.0f;
                    fsse = fma_ (gy, gy, fsse);
                    fst2 = on ? fma_ (t, t, fst2) : fst2;
                }
                jit_step_adj (c, xs[s], zs[s], rs[s], gy, gz, pf, fl, fv);
            }
        }
    };
    if (UNROLL)
    {
#pragma unroll
        for (int cc = 0; cc < kSeg / 4; ++cc)
            replay4 (cc);
#pragma unroll
        for (int cc = kSeg / 4 - 1; cc >= 0; --cc)
            reverse4 (cc);
    }
    else
    {
#pragma unroll 1
        for (int cc = 0; cc < kSeg / 4; ++cc)
            replay4 (cc);
#pragma unroll 1
        for (int cc = kSeg / 4 - 1; cc >= 0; --cc)
            reverse4 (cc);
    }
    // fp32 inside a segment, double across segments
#pragma unroll
    for (int j = 0; j < kNP; ++j)
        acc.pf[j] += (double) pf[j];
    acc.l += (double) fl;
    acc.v += (double) fv;
    acc.sse += (double) fsse;
    acc.st2 += (double) fst2;
}

// partials[group * 24 + k]: k < 16 d/dp1R of node k, 16 ell, 17 V, 18 sse, 19 st2, 20 "already chained" flag (tree_kernels.cu)
__device__ __forceinline__ void jit_write_partials (JitAcc& acc, double* __restrict__ partials, int lane)
{
    double* p = partials + (long long) blockIdx.x * 24;
    if (lane < 24)
        p[lane] = 0.0;
    __syncwarp ();
#pragma unroll
    for (int j = 0; j < kNP; ++j)
    {
        const double v = warp_sum (acc.pf[j]);
        if (lane == 0 && j < kNPreal)
            p[jit_pf_node (j)] = v;
    }
    const double l = warp_sum (acc.l), v = warp_sum (acc.v), sse = warp_sum (acc.sse), st2 = warp_sum (acc.st2);
    if (lane == 0)
        p[16] = l, p[17] = v, p[18] = sse, p[19] = st2;
}

__device__ __forceinline__ void load_ckpt (float (&z)[kNS1], const float* __restrict__ ckpt, int seg, long long B, long long b, bool valid)
{
    const float* ck = ckpt + ((long long) seg * kNS1) * B + b;
#pragma unroll
    for (int k = 0; k < kNS1; ++k)
        z[k] = valid ? __ldg (ck + k * B) : 0.0f;
}

// x and g (target or dL/dy) tiles [32 x 16], 64-byte swizzle, 3-slot ring; segments last to first
extern "C" __global__ void __launch_bounds__ (32) jit_tree_adjoint_tma (const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmg, const float* __restrict__ params, const float* __restrict__ ckpt, double* __restrict__ partials, long long B, int T, int skip, int target)
{
    constexpr int kStages = 3, kTile = 32 * kSeg * 4, kStage = 2 * kTile;
    __shared__ __align__ (1024) unsigned char smem[kStages * kStage];
    __shared__ __align__ (8) unsigned long long bar_mem[kStages];
    const int lane = threadIdx.x;
    const int b0 = blockIdx.x * 32;
    const long long b = (long long) b0 + lane;
    const bool valid = b < B;
    const uint32_t tiles = smem_u32 (smem), bars = smem_u32 (bar_mem);
    if (lane == 0)
    {
        tma_prefetch_desc (&tmx);
        tma_prefetch_desc (&tmg);
        for (int s = 0; s < kStages; ++s)
            mbar_init (bars + 8 * s, 1);
        fence_mbar_init ();
    }
    __syncwarp ();
    JC c;
    jit_consts (params, c);
    const int nseg = (T + kSeg - 1) / kSeg;
    auto fetch = [&] (int k) { // the k-th processed segment is i = nseg - 1 - k
        const int i = nseg - 1 - k, s = k % kStages;
        const uint32_t dst = tiles + s * kStage, bar = bars + 8 * s;
        mbar_expect_tx (bar, kStage);
        tma_load_2d (dst, &tmx, i * kSeg, b0, bar);
        tma_load_2d (dst + kTile, &tmg, i * kSeg, b0, bar);
    };
    if (lane == 0)
        for (int k = 0; k < kStages - 1 && k < nseg; ++k)
            fetch (k);
    JitAcc acc;
#pragma unroll
    for (int j = 0; j < kNP; ++j)
        acc.pf[j] = 0.0;
    float gz[kNS1], z0[kNS1], znext[kNS1];
#pragma unroll
    for (int k = 0; k < kNS1; ++k)
        gz[k] = 0.0f;
    load_ckpt (znext, ckpt, nseg - 1, B, b, valid);
    for (int k = 0; k < nseg; ++k)
    {
        const int i = nseg - 1 - k, s = k % kStages;
#pragma unroll
        for (int q = 0; q < kNS1; ++q)
            z0[q] = znext[q];
        if (i > 0)
            load_ckpt (znext, ckpt, i - 1, B, b, valid); // in flight while this segment is processed
        if (k + kStages - 1 < nseg)
        { // refill the slot the previous segment was read from
            fence_proxy_async ();
            __syncwarp ();
            if (lane == 0)
                fetch (k + kStages - 1);
        }
        mbar_wait (bars + 8 * s, (k / kStages) & 1);
        const TileIO2 io { tiles + s * kStage, tiles + s * kStage + kTile, lane };
        if (valid)
            jit_adjoint_segment<kUnrollSegment> (c, io, z0, gz, i * kSeg, T, skip, target, acc);
    }
    jit_write_partials (acc, partials, lane);
}

extern "C" __global__ void __launch_bounds__ (32) jit_tree_adjoint_direct (const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ params, const float* __restrict__ ckpt, double* __restrict__ partials, long long B, int T, int skip, int target)
{
    const int lane = threadIdx.x;
    const long long b = (long long) blockIdx.x * 32 + lane;
    JC c;
    jit_consts (params, c);
    JitAcc acc;
#pragma unroll
    for (int j = 0; j < kNP; ++j)
        acc.pf[j] = 0.0;
    if (b < B)
    {
        float gz[kNS1], z0[kNS1];
#pragma unroll
        for (int k = 0; k < kNS1; ++k)
            gz[k] = 0.0f;
        const int nseg = (T + kSeg - 1) / kSeg;
        GlobalIO2 io { x + b * T, g + b * T, 0, T };
        for (int i = nseg - 1; i >= 0; --i)
        {
            io.n0 = i * kSeg;
            load_ckpt (z0, ckpt, i, B, b, true);
            jit_adjoint_segment<kUnrollSegment> (c, io, z0, gz, i * kSeg, T, skip, target, acc);
        }
    }
    jit_write_partials (acc, partials, lane);
}
} // namespace dwdf
)DWDFJIT";

// ---- NVRTC (dlopen) and the driver's module API (entry points through the runtime) ---------------------------------
struct Nvrtc
{
    void* handle = nullptr;
    int (*CreateProgram) (void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram) (void*, int, const char* const*) = nullptr;
    int (*GetCUBINSize) (void*, size_t*) = nullptr;
    int (*GetCUBIN) (void*, char*) = nullptr;
    int (*GetProgramLogSize) (void*, size_t*) = nullptr;
    int (*GetProgramLog) (void*, char*) = nullptr;
    int (*DestroyProgram) (void**) = nullptr;
    bool ok = false;
};

const Nvrtc& nvrtc ()
{
    static Nvrtc n = [] {
        Nvrtc r;
        for (const char* name : { "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so" })
            if ((r.handle = dlopen (name, RTLD_NOW | RTLD_LOCAL)) != nullptr)
                break;
        if (r.handle == nullptr)
            return r;
        auto sym = [&] (const char* s) { return dlsym (r.handle, s); };
        r.CreateProgram = (decltype (r.CreateProgram)) sym ("nvrtcCreateProgram");
        r.CompileProgram = (decltype (r.CompileProgram)) sym ("nvrtcCompileProgram");
        r.GetCUBINSize = (decltype (r.GetCUBINSize)) sym ("nvrtcGetCUBINSize");
        r.GetCUBIN = (decltype (r.GetCUBIN)) sym ("nvrtcGetCUBIN");
        r.GetProgramLogSize = (decltype (r.GetProgramLogSize)) sym ("nvrtcGetProgramLogSize");
        r.GetProgramLog = (decltype (r.GetProgramLog)) sym ("nvrtcGetProgramLog");
        r.DestroyProgram = (decltype (r.DestroyProgram)) sym ("nvrtcDestroyProgram");
        r.ok = r.CreateProgram && r.CompileProgram && r.GetCUBINSize && r.GetCUBIN && r.GetProgramLogSize && r.GetProgramLog && r.DestroyProgram;
        return r;
    }();
    return n;
}

struct Driver
{
    CUresult (*ModuleLoadData) (CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction) (CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*ModuleUnload) (CUmodule) = nullptr;
    CUresult (*LaunchKernel) (CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    bool ok = false;
};

const Driver& driver ()
{
    static Driver d = [] {
        Driver r;
        auto get = [] (const char* name) -> void* {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint (name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
                return nullptr;
            return p;
        };
        r.ModuleLoadData = (decltype (r.ModuleLoadData)) get ("cuModuleLoadData");
        r.ModuleGetFunction = (decltype (r.ModuleGetFunction)) get ("cuModuleGetFunction");
        r.ModuleUnload = (decltype (r.ModuleUnload)) get ("cuModuleUnload");
        r.LaunchKernel = (decltype (r.LaunchKernel)) get ("cuLaunchKernel");
        r.ok = r.ModuleLoadData && r.ModuleGetFunction && r.ModuleUnload && r.LaunchKernel;
        (void) cudaGetLastError ();
        return r;
    }();
    return d;
}

struct DeviceModule
{
    CUmodule mod = nullptr;
    CUfunction fwd_tma = nullptr, fwd_direct = nullptr, adj_tma = nullptr, adj_direct = nullptr;
};
} // namespace

struct TreeJit
{
    std::string source; // the generated part (the headers and the skeleton are fixed)
    std::vector<char> cubin;
    std::mutex mu;
    std::map<int, DeviceModule> modules; // per device: the module lives in that device's primary context
};

std::string tree_jit_full_source (const TreeProgram& p) { return tree_jit_generate (p) + kJitSkeleton; }
const char* tree_jit_header (int which) { return which == 0 ? kJitMathSrc : kJitTmaSrc; }

TreeJit* tree_jit_create (const TreeProgram& p, std::string& err)
{
    if (! tree_jit_supported (p))
    {
        err = "the circuit uses elements the specialiser does not cover (it stays on the tree interpreter)";
        return nullptr;
    }
    const Nvrtc& rt = nvrtc ();
    if (! rt.ok)
    {
        err = "libnvrtc.so.12 could not be loaded (dlopen): no run-time specialisation, the program stays on the tree interpreter";
        return nullptr;
    }
    TreeJit* j = new TreeJit;
    j->source = tree_jit_generate (p);
    const std::string full = j->source + kJitSkeleton;
    void* prog = nullptr;
    const char* hdr_src[2] = { kJitMathSrc, kJitTmaSrc };
    const char* hdr_name[2] = { "dwdf_math.cuh", "dwdf_tma.cuh" };
    if (rt.CreateProgram (&prog, full.c_str (), "dwdf_tree_jit.cu", 2, hdr_src, hdr_name) != 0)
    {
        err = "nvrtcCreateProgram failed";
        delete j;
        return nullptr;
    }
    // --fmad=false: the generated wave arithmetic is evaluated as written, one rounding per operation like the reference's own
    // (where the compiler contracts a product into an FMA on one sweep and not on the other, an open switch is no longer exactly
    // silent); dwdf_math.cuh spells out the FMAs it wants (fma_), so the roots are unaffected
    const char* opts[] = { "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--fmad=false" };
    const int rc = rt.CompileProgram (prog, 4, opts);
    if (rc != 0)
    {
        size_t n = 0;
        rt.GetProgramLogSize (prog, &n);
        std::string log (n, '\0');
        if (n > 0)
            rt.GetProgramLog (prog, &log[0]);
        err = "NVRTC: " + log.substr (0, 400);
        rt.DestroyProgram (&prog);
        delete j;
        return nullptr;
    }
    size_t n = 0;
    rt.GetCUBINSize (prog, &n);
    j->cubin.resize (n);
    rt.GetCUBIN (prog, j->cubin.data ());
    rt.DestroyProgram (&prog);
    if (n == 0)
    {
        err = "NVRTC produced no cubin";
        delete j;
        return nullptr;
    }
    return j;
}

void tree_jit_destroy (TreeJit* j)
{
    if (j == nullptr)
        return;
    // a module lives in its device's primary context: make that device current for the unload, then put the caller's back
    if (! j->modules.empty () && driver ().ok)
    {
        int prev = -1;
        const bool have_prev = cudaGetDevice (&prev) == cudaSuccess;
        for (auto& dm : j->modules)
            if (dm.second.mod != nullptr && cudaSetDevice (dm.first) == cudaSuccess)
                (void) driver ().ModuleUnload (dm.second.mod);
        if (have_prev)
            (void) cudaSetDevice (prev);
        (void) cudaGetLastError ();
    }
    delete j;
}

const std::string& tree_jit_source (const TreeJit* j) { return j->source; }
size_t tree_jit_cubin_bytes (const TreeJit* j) { return j->cubin.size (); }

static bool device_module (TreeJit* j, DeviceModule& out, std::string& err)
{
    int dev = -1;
    if (cudaGetDevice (&dev) != cudaSuccess)
    {
        err = "no CUDA device";
        return false;
    }
    std::lock_guard<std::mutex> lock (j->mu);
    auto it = j->modules.find (dev);
    if (it != j->modules.end ())
    {
        out = it->second;
        return true;
    }
    const Driver& d = driver ();
    if (! d.ok)
    {
        err = "the driver's module entry points are unavailable";
        return false;
    }
    (void) cudaFree (nullptr); // make sure the device's primary context exists and is current on this thread
    DeviceModule m;
    CUresult r = d.ModuleLoadData (&m.mod, j->cubin.data ());
    if (r == CUDA_SUCCESS)
        r = d.ModuleGetFunction (&m.fwd_tma, m.mod, "jit_tree_forward_tma");
    if (r == CUDA_SUCCESS)
        r = d.ModuleGetFunction (&m.fwd_direct, m.mod, "jit_tree_forward_direct");
    if (r == CUDA_SUCCESS)
        r = d.ModuleGetFunction (&m.adj_tma, m.mod, "jit_tree_adjoint_tma");
    if (r == CUDA_SUCCESS)
        r = d.ModuleGetFunction (&m.adj_direct, m.mod, "jit_tree_adjoint_direct");
    if (r != CUDA_SUCCESS)
    {
        char buf[96];
        snprintf (buf, sizeof (buf), "loading the specialised module failed (CUresult %d)", (int) r);
        err = buf;
        return false;
    }
    j->modules[dev] = m;
    out = m;
    return true;
}

bool tree_jit_load (TreeJit* j, std::string& err)
{
    DeviceModule m;
    return device_module (j, m, err);
}

static bool launch (CUfunction f, int64_t B, void** args, cudaStream_t stream, std::string& err)
{
    const CUresult r = driver ().LaunchKernel (f, (unsigned) ((B + 31) / 32), 1, 1, 32, 1, 1, 0, (CUstream) stream, args, nullptr);
    if (r != CUDA_SUCCESS)
    {
        char buf[96];
        snprintf (buf, sizeof (buf), "cuLaunchKernel failed (CUresult %d)", (int) r);
        err = buf;
        return false;
    }
    return true;
}

bool tree_jit_forward (TreeJit* j, const CUtensorMap* tmx, const CUtensorMap* tmy, const float* params, const float* x, float* y, float* ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream, std::string& err)
{
    DeviceModule m;
    if (! device_module (j, m, err))
        return false;
    long long b = B;
    int t = (int) T;
    if (tmx != nullptr)
    {
        void* args[] = { (void*) tmx, (void*) tmy, &params, &ckpt, &state, &b, &t };
        return launch (m.fwd_tma, B, args, stream, err);
    }
    void* args[] = { &x, &y, &params, &ckpt, &state, &b, &t };
    return launch (m.fwd_direct, B, args, stream, err);
}

bool tree_jit_adjoint (TreeJit* j, const CUtensorMap* tmx, const CUtensorMap* tmg, const float* params, const float* x, const float* g, const float* ckpt, bool target, int skip, double* partials, int64_t B, int64_t T, cudaStream_t stream,
                       std::string& err)
{
    DeviceModule m;
    if (! device_module (j, m, err))
        return false;
    long long b = B;
    int t = (int) T, sk = skip, tg = target ? 1 : 0;
    if (tmx != nullptr)
    {
        void* args[] = { (void*) tmx, (void*) tmg, &params, &ckpt, &partials, &b, &t, &sk, &tg };
        return launch (m.adj_tma, B, args, stream, err);
    }
    void* args[] = { &x, &g, &params, &ckpt, &partials, &b, &t, &sk, &tg };
    return launch (m.adj_direct, B, args, stream, err);
}

} // namespace dwdf
