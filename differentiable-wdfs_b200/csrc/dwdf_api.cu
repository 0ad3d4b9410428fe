// dwdf_api.cu — host side of libdwdf.so: validates and lowers the element tree to the flat program
// the kernels consume, builds the TMA descriptors, sizes and launches the kernels, and implements
// the extern "C" surface of include/dwdf.h. No torch, no Python: plain C++ over the CUDA runtime.
#include "../../include/dwdf.h"
#include "dwdf_kernels.h"
#include "tree_jit.h"

#include <cstdlib>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include <cudaTypedefs.h>

using namespace dwdf;

struct dwdf_program
{
    std::vector<dwdf_node> nodes;
    dwdf_circuit_desc desc;
    bool is_clipper = false;
    bool clip_r = false; // the clipper with its source resistance as a per-sample input channel (clipper_pot.py:114-117)
    bool is_neural = false;
    bool differentiable = true; // every node and the root have a reverse-mode implementation
    dwdf_mlp_desc mlp {};
    ClipDesc clip {};
    ClipVariant variant {};
    TreeProgram tree {};
    int n_states = 0;
    // run-time specialised kernels of a tree program (dwdf_program_specialize), else the interpreter. Published once, fully built,
    // with release / acquire ordering: a launch on another thread sees either none or a complete one
    std::atomic<TreeJit*> jit { nullptr };
    std::mutex jit_mu;
    ~dwdf_program () { tree_jit_destroy (jit.load ()); }
};

namespace
{
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches { 0 };
std::atomic<int> g_use_tma { 1 };
// ---- per-device state -------------------------------------------------------------------------------
// Created eagerly by dwdf_program_create for the device current at that moment (never lazily in a launch path: a
// cudaMalloc there would break a stream capture) and keyed by device, so that circuits on different GPUs of one
// process do not share device pointers:
//   * a diagnostics counter (chunks the time-parallel forward had to recompute);
//   * a LIBRARY-OWNED memory pool for the stream-ordered scratch, with the release threshold raised so that freed
//     scratch stays in the pool across synchronisations (with the default of 0 a training loop that reads its loss
//     each step would re-allocate from the driver each step). The host application's default pool is left alone.
constexpr int kMaxDevices = 64;
struct DeviceState
{
    std::once_flag once;
    int* redone = nullptr;
    cudaMemPool_t pool = nullptr;
};
DeviceState g_dev[kMaxDevices];

DeviceState* device_state ()
{
    int dev = -1;
    if (cudaGetDevice (&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices)
    {
        (void) cudaGetLastError (); // no device: the compute entry points report it; creating a program needs none
        return nullptr;
    }
    DeviceState& st = g_dev[dev];
    std::call_once (st.once, [&st, dev] {
        if (cudaMalloc ((void**) &st.redone, sizeof (int)) != cudaSuccess || cudaMemset (st.redone, 0, sizeof (int)) != cudaSuccess)
            st.redone = nullptr;
        cudaMemPoolProps props {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        uint64_t keep = UINT64_MAX;
        if (cudaMemPoolCreate (&st.pool, &props) != cudaSuccess || cudaMemPoolSetAttribute (st.pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess)
            st.pool = nullptr;
        (void) cudaGetLastError ();
    });
    return &st;
}

int* redone_counter ()
{
    DeviceState* st = device_state ();
    return st != nullptr ? st->redone : nullptr;
}

// stream-ordered scratch that is released on every way out of the calling function
struct AsyncScratch
{
    float* p = nullptr;
    cudaStream_t stream = nullptr;
    cudaError_t alloc (size_t bytes, cudaStream_t s)
    {
        stream = s;
        DeviceState* st = device_state ();
        if (st != nullptr && st->pool != nullptr)
            return cudaMallocFromPoolAsync ((void**) &p, bytes, st->pool, s);
        return cudaMallocAsync ((void**) &p, bytes, s);
    }
    ~AsyncScratch ()
    {
        if (p != nullptr)
            cudaFreeAsync (p, stream);
    }
};

// ---- per-phase timing of the training step (dwdf_profile_begin / _end) ------------------------------------------------
// CUDA events recorded on the caller's stream at the kernel boundaries of dwdf_train_step / dwdf_train_step_dp: start,
// after the forward pass (with its verification kernel), after the adjoint (with its composition kernel), after the tail
// (reduction + exchange + chain rule + Adam). bench.py reads the dominant kernel's duration inside its timed region this way.
struct Profiler
{
    std::mutex mu;
    bool on = false;
    int cap = 0, n = 0;
    std::vector<cudaEvent_t> ev;
    void mark (int k, cudaStream_t s)
    {
        if (! on)
            return;
        std::lock_guard<std::mutex> lock (mu);
        if (on && n < cap)
            cudaEventRecord (ev[(size_t) 4 * n + k], s);
        if (on && k == 3 && n < cap)
            ++n;
    }
};
Profiler g_prof;

int fail (int code, const char* fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (g_err, sizeof (g_err), fmt, ap);
    va_end (ap);
    return code;
}

int cuda_fail (cudaError_t e, const char* what)
{
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return fail (DWDF_ERR_NO_DEVICE, "%s: %s (libdwdf has no CPU fallback)", what, cudaGetErrorString (e));
    return fail (DWDF_ERR_CUDA, "%s: %s", what, cudaGetErrorString (e));
}

#define DWDF_CUDA(call) \
    do \
    { \
        cudaError_t e__ = (call); \
        if (e__ != cudaSuccess) \
            return cuda_fail (e__, #call); \
    } while (0)

// ---- TMA descriptors ----------------------------------------------------------------------------
constexpr int kDefaultL2Promotion = 256; // measured at 65536 x 4096 (approx root): forward 0.462 ms at 128 B (or none) -> 0.434 ms at 256 B (the next tile of every row arrives in L2 with the current one); the reverse sweep is indifferent to 128 / 256 / none (0.66 ms) and takes 1.12 ms at 64 B
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder ()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = [] () -> PFN_cuTensorMapEncodeTiled_v12000 {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (PFN_cuTensorMapEncodeTiled_v12000) p;
    }();
    return fn;
}

// (B, T) fp32 batch-major tensor, tiles of [32 sequences x tile_t samples]; rows of a tile are
// tile_t*4 = 128 or 64 bytes and use the matching swizzle so that 32 lanes reading the same
// 16-byte chunk index of their own rows hit 32 distinct banks.
bool make_map (CUtensorMap* m, const float* base, int64_t B, int64_t T, int tile_t, int rows = 32)
{
    auto enc = tensor_map_encoder ();
    if (enc == nullptr)
        return false;
    const cuuint64_t dims[2] = { (cuuint64_t) T, (cuuint64_t) B };
    const cuuint64_t strides[1] = { (cuuint64_t) T * 4 };
    const cuuint32_t box[2] = { (cuuint32_t) tile_t, (cuuint32_t) rows };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUtensorMapSwizzle sw = tile_t == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    // L2 fetch granularity of a tile's row segments. Development switch DWDF_L2_PROMOTION = 0 / 64 / 128 / 256 (read once).
    static const CUtensorMapL2promotion promo = [] {
        const char* e = std::getenv ("DWDF_L2_PROMOTION");
        const int v = e != nullptr ? std::atoi (e) : kDefaultL2Promotion;
        return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    return enc (m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*> (base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tma_usable (const void* a, const void* b, const void* c, int64_t B, int64_t T)
{
    if (! g_use_tma.load () || T % 4 != 0 || T < 4 || B < 1)
        return false;
    for (const void* p : { a, b, c })
        if (p != nullptr && ((uintptr_t) p & 15u) != 0)
            return false;
    return tensor_map_encoder () != nullptr;
}

// ---- tree validation / lowering ---------------------------------------------------------------------
bool is_leaf (int k) { return k == DWDF_RESISTOR || k == DWDF_CAPACITOR || k == DWDF_RESISTIVE_VS || k == DWDF_INDUCTOR || k == DWDF_CAPACITOR_ALPHA || k == DWDF_INDUCTOR_ALPHA || k == DWDF_RESISTIVE_CS; }
int extra_params (int k) { return (k == DWDF_CAPACITOR_ALPHA || k == DWDF_INDUCTOR_ALPHA) ? 1 : (k == DWDF_Y_PARAMETER ? 3 : 0); } // constants kept in the slots after the value
int states_of (int k) { return (k == DWDF_CAPACITOR || k == DWDF_INDUCTOR) ? 1 : ((k == DWDF_CAPACITOR_ALPHA || k == DWDF_INDUCTOR_ALPHA) ? 2 : 0); }
bool has_adjoint (int k) { return k <= DWDF_INDUCTOR; } // reverse mode: the wdf_py element set and the inductor

int64_t n_groups (int64_t B) { return (B + 31) / 32; }
// Time chunks (fewer sequences than the SMs hold warps; clipper_kernels.cu explains the scheme): upper bounds of the
// chunk counts the launchers may propose, for sizing the scratch. The launchers pick the actual count from the kernel's
// measured residency, the kernels make the final plan from gamma.
int chunk_cap (int64_t groups, int64_t units)
{
    if (g_clip_opts & kOptNoChunks)
        return 1;
    int64_t k = kMaxResidentCtas / (groups > 0 ? groups : 1);
    if (g_clip_opts & kOptForceChunks)
        k = k > 4 ? k : 4;
    k = k < units ? k : units;
    return (int) (k > 1 ? k : 1);
}
int64_t n_fwd_tiles16 (int64_t T) { return (T + 15) / 16; }
int fwd_chunk_cap (int64_t B, int64_t T) { return T >= 128 ? chunk_cap ((B + 63) / 64, n_fwd_tiles16 (T)) : 1; } // (the one-sequence-per-lane kernel launches twice the groups: its proposal is clamped to this)
int adj_chunk_cap (int64_t B, int64_t T)
{
    const int k = T >= 128 ? chunk_cap ((B + 31) / 32, (T + 63) / 64) : 1;
    return (k < 8 && T >= 1024 && ! (g_clip_opts & kOptNoChunks)) ? 8 : k; // (the resistance-channel adjoint balances large batches with 8 chunks)
}
constexpr int64_t kNnChunkedMaxB = 16384; // neural root: the network makes every sample ~20x heavier, so lanes stay scarce longer
size_t partials_bytes (int64_t B) { return ((size_t) n_groups (B) * kTreePartialStride * sizeof (double) + 255) / 256 * 256 + 256; }
int64_t n_segments (int64_t T) { return (T + kSeg - 1) / kSeg; }
size_t adj_maps_bytes (int64_t B, int64_t T) { const int k = adj_chunk_cap (B, T); return k > 1 ? (size_t) k * kMapFloatsPerChunk * (size_t) B * sizeof (float) : 0; }

// DWDF_LOSS_MSE_ESR_AS_CALLED (loss_kernels.cu): batch sums over (y, target) -> [summed over ranks] -> dL/dy into scratch. The
// struct keeps the stream-ordered scratch alive until the caller has enqueued the sweep that reads it.
struct AsCalledLoss
{
    AsyncScratch buf;
    double* sums = nullptr;
    float* ybar = nullptr;
};
int as_called_prepare (AsCalledLoss& a, const float* y, const float* target, int64_t B, int64_t T, int64_t sk, const DpPeers* dp, cudaStream_t stream)
{
    const size_t head = (loss_scratch_doubles (B) * sizeof (double) + 255) / 256 * 256;
    DWDF_CUDA (a.buf.alloc (head + (size_t) B * (size_t) T * sizeof (float), stream));
    a.sums = reinterpret_cast<double*> (a.buf.p);
    a.ybar = reinterpret_cast<float*> (reinterpret_cast<char*> (a.buf.p) + head);
    DWDF_CUDA (launch_loss_sums (y, target, B, T, (int) sk, a.sums, stream));
    g_launches.fetch_add (2);
    if (dp != nullptr && dp->world > 1)
    {
        DWDF_CUDA (launch_peer_allreduce (a.sums, 4, *dp, stream));
        g_launches.fetch_add (1);
    }
    DWDF_CUDA (launch_loss_ybar (y, target, B, T, (int) sk, a.sums, a.ybar, stream));
    g_launches.fetch_add (1);
    return DWDF_OK;
}
bool loss_kind_ok (int32_t k) { return k == DWDF_LOSS_MSE || k == DWDF_LOSS_MSE_ESR || k == DWDF_LOSS_MSE_ESR_AS_CALLED; }
} // namespace

extern "C" {
static int check_batch (const dwdf_program* prog, const void* params, const void* x, int64_t B, int64_t T);

const char* dwdf_last_error (void) { return g_err; }
const char* dwdf_build_info (void) { return "libdwdf v4 sm_100a cuda-12.9 tma+mbarrier fp32x2 nvrtc-specialiser no-cpu-fallback"; }
int64_t dwdf_launch_count (void) { return g_launches.load () + g_extra_launches.load (); }
int dwdf_set_tma (int enable) { return g_use_tma.exchange (enable ? 1 : 0); }
int64_t dwdf_time_parallel_redone (void)
{
    int v = 0;
    int* counter = redone_counter (); // of the current device
    if (counter != nullptr && cudaMemcpy (&v, counter, sizeof (int), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    return v;
}
int dwdf_set_option (int bits)
{
    const int prev = g_clip_opts;
    g_clip_opts = bits;
    return prev;
}

int dwdf_program_create (const dwdf_node* nodes, int32_t n_nodes, const dwdf_circuit_desc* d, dwdf_program** out)
{
    if (out == nullptr || nodes == nullptr || d == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_nodes < 1 || n_nodes > DWDF_MAX_NODES)
        return fail (DWDF_ERR_INVALID, "n_nodes %d outside [1, %d]", n_nodes, DWDF_MAX_NODES);
    if (d->n_params < 1 || d->n_params > DWDF_MAX_PARAMS)
        return fail (DWDF_ERR_INVALID, "n_params %d outside [1, %d]", d->n_params, DWDF_MAX_PARAMS);
    if (! (d->fs > 0.0f))
        return fail (DWDF_ERR_INVALID, "sample rate must be positive");
    std::vector<int> parents (n_nodes, 0);
    int n_states = 0;
    for (int i = 0; i < n_nodes; ++i)
    {
        const dwdf_node& n = nodes[i];
        if (n.kind < DWDF_RESISTOR || n.kind > DWDF_Y_PARAMETER)
            return fail (DWDF_ERR_INVALID, "node %d: unknown kind %d", i, n.kind);
        if (is_leaf (n.kind) || n.kind == DWDF_Y_PARAMETER)
        {
            if (n.param < 0 || n.param + extra_params (n.kind) >= d->n_params)
                return fail (DWDF_ERR_INVALID, "node %d: parameter slots %d .. %d outside [0, %d)", i, n.param, n.param + extra_params (n.kind), d->n_params);
            n_states += states_of (n.kind);
        }
        if (! is_leaf (n.kind))
        {
            const int need = (n.kind == DWDF_INVERTER || n.kind == DWDF_Y_PARAMETER) ? 1 : 2;
            const int ch[2] = { n.child1, n.child2 };
            for (int k = 0; k < need; ++k)
            {
                if (ch[k] < 0 || ch[k] >= i)
                    return fail (DWDF_ERR_INVALID, "node %d: child %d must precede it (post-order)", i, ch[k]);
                ++parents[ch[k]];
            }
        }
    }
    for (int i = 0; i < n_nodes - 1; ++i)
        if (parents[i] != 1)
            return fail (DWDF_ERR_INVALID, "node %d has %d parents (the circuit must be a tree whose top is the last node)", i, parents[i]);
    if (d->probe < 0 || d->probe >= n_nodes)
        return fail (DWDF_ERR_INVALID, "probe node %d out of range", d->probe);
    if (d->ordering != DWDF_ORDER_PLUGIN && d->ordering != DWDF_ORDER_PYTHON)
        return fail (DWDF_ERR_INVALID, "unknown ordering %d", d->ordering);
    if (d->r_node >= n_nodes || (d->r_node >= 0 && nodes[d->r_node].kind != DWDF_RESISTOR && nodes[d->r_node].kind != DWDF_RESISTIVE_VS))
        return fail (DWDF_ERR_INVALID, "r_node must be a Resistor or ResistiveVoltageSource leaf");
    if (d->probe_current != 0 && d->probe_current != 1)
        return fail (DWDF_ERR_INVALID, "probe_current is 0 (voltage) or 1 (current)");
    const bool source_leaf = d->source >= 0 && d->source < n_nodes && (nodes[d->source].kind == DWDF_RESISTIVE_VS || nodes[d->source].kind == DWDF_RESISTIVE_CS);
    if (d->root_kind == DWDF_ROOT_DIODE || d->root_kind == DWDF_ROOT_SWITCH)
    {
        if (! source_leaf)
            return fail (DWDF_ERR_INVALID, "diode / switch circuits are driven through a ResistiveVoltageSource or ResistiveCurrentSource leaf (source = %d)", d->source);
        if (d->root_kind == DWDF_ROOT_DIODE && (d->param_Is < 0 || d->param_Is >= d->n_params || d->param_nabla < 0 || d->param_nabla >= d->n_params || ! (d->Vt > 0.0f)))
            return fail (DWDF_ERR_INVALID, "diode parameter slots out of range");
        if (d->root_kind == DWDF_ROOT_SWITCH && d->root_mode != 0 && d->root_mode != 1)
            return fail (DWDF_ERR_INVALID, "switch: root_mode is 1 (closed) or 0 (open)");
    }
    else if (d->root_kind == DWDF_ROOT_IDEAL_CS)
    {
        if (d->source >= 0)
            return fail (DWDF_ERR_INVALID, "the ideal current source is driven by x[n] itself (source = -1)");
    }
    else if (d->root_kind == DWDF_ROOT_DIODE_PAIR)
    {
        if (! source_leaf)
            return fail (DWDF_ERR_INVALID, "diode-pair circuits are driven through a ResistiveVoltageSource leaf (source = %d)", d->source);
        if (d->root_mode < DWDF_MODE_APPROX || d->root_mode > DWDF_MODE_APPROX_GOOD)
            return fail (DWDF_ERR_INVALID, "unknown root mode %d", d->root_mode);
        if (d->param_Is < 0 || d->param_Is >= d->n_params || d->param_nabla < 0 || d->param_nabla >= d->n_params)
            return fail (DWDF_ERR_INVALID, "diode parameter slots out of range");
        if (! (d->Vt > 0.0f) || ! (d->n_up > 0.0f) || ! (d->n_down > 0.0f))
            return fail (DWDF_ERR_INVALID, "Vt, n_up, n_down must be positive");
        if (d->root_mode == DWDF_MODE_APPROX_GOOD && (d->n_up != 1.0f || d->n_down != 1.0f))
            return fail (DWDF_ERR_UNSUPPORTED, "the eq.18 'Good' law is defined for a symmetric pair only");
    }
    else if (d->root_kind == DWDF_ROOT_NEURAL)
    {
        const bool shape = n_nodes == 3 && nodes[0].kind == DWDF_RESISTIVE_VS && nodes[1].kind == DWDF_CAPACITOR && nodes[2].kind == DWDF_PARALLEL && nodes[2].child1 == 0 && nodes[2].child2 == 1
                           && d->source == 0 && d->probe == 1 && (d->r_node == -1 || d->r_node == 0) && nodes[0].param != nodes[1].param;
        if (! shape)
            return fail (DWDF_ERR_UNSUPPORTED, "the neural root closes the clipper tree Parallel(ResistiveVoltageSource, Capacitor) with the probe on the capacitor");
    }
    else if (d->root_kind != DWDF_ROOT_IDEAL_VS)
        return fail (DWDF_ERR_INVALID, "unknown root kind %d", d->root_kind);

    dwdf_program* p = new (std::nothrow) dwdf_program;
    if (p == nullptr)
        return fail (DWDF_ERR_INVALID, "out of memory");
    (void) device_state (); // per-device counter and scratch pool, created here rather than in a launch path
    p->nodes.assign (nodes, nodes + n_nodes);
    p->desc = *d;
    p->n_states = n_states;
    p->differentiable = d->root_kind == DWDF_ROOT_IDEAL_VS || d->root_kind == DWDF_ROOT_DIODE_PAIR || d->root_kind == DWDF_ROOT_NEURAL;
    for (int i = 0; i < n_nodes; ++i)
        p->differentiable = p->differentiable && has_adjoint (nodes[i].kind);
    p->differentiable = p->differentiable && d->probe_current == 0;

    // flat program for the interpreter
    TreeProgram& t = p->tree;
    t.n_nodes = n_nodes;
    int st = 0;
    for (int i = 0; i < 16; ++i)
    {
        t.kind[i] = i < n_nodes ? nodes[i].kind : 0;
        t.c1[i] = i < n_nodes ? nodes[i].child1 : -1;
        t.c2[i] = i < n_nodes ? nodes[i].child2 : -1;
        t.param[i] = i < n_nodes ? nodes[i].param : -1;
        t.state_of[i] = -1;
        if (i < n_nodes && states_of (nodes[i].kind) > 0)
        {
            t.state_of[i] = st;
            st += states_of (nodes[i].kind);
        }
    }
    t.probe_current = d->probe_current;
    t.root_kind = d->root_kind;
    t.root_mode = d->root_mode;
    t.pyorder = d->ordering == DWDF_ORDER_PYTHON;
    t.probe = d->probe;
    t.source = d->source;
    t.r_node = d->r_node;
    t.slot_Is = d->param_Is;
    t.slot_nabla = d->param_nabla;
    t.n_params = d->n_params;
    t.n_iter = d->newton_max_iter;
    t.fs = d->fs;
    t.Vt = d->Vt;
    t.n_up = d->n_up;
    t.n_down = d->n_down;
    t.tol = d->newton_tol;
    t.n_states = n_states;

    // the diode clipper: Parallel(P1 = ResistiveVs, P2 = Capacitor) + DiodePair, probe on the capacitor
    const bool clip_shape = d->root_kind == DWDF_ROOT_DIODE_PAIR && n_nodes == 3 && nodes[0].kind == DWDF_RESISTIVE_VS && nodes[1].kind == DWDF_CAPACITOR && nodes[2].kind == DWDF_PARALLEL
                            && nodes[2].child1 == 0 && nodes[2].child2 == 1 && d->source == 0 && d->probe == 1 && (d->r_node < 0 || d->r_node == 0) && d->root_mode != DWDF_MODE_APPROX_GOOD;
    if (clip_shape)
    {
        const int s[4] = { nodes[0].param, nodes[1].param, d->param_Is, d->param_nabla };
        bool distinct = true;
        for (int a = 0; a < 4; ++a)
            for (int b = a + 1; b < 4; ++b)
                distinct = distinct && s[a] != s[b];
        if (distinct)
        {
            p->is_clipper = true;
            p->clip_r = d->r_node == 0;
            p->clip = ClipDesc { d->fs, d->Vt, d->n_up, d->n_down, d->newton_tol, d->newton_max_iter, s[0], s[1], s[2], s[3] };
            p->variant = ClipVariant { d->root_mode == DWDF_MODE_EXACT ? kModeExact : kModeApprox, ! (d->n_up == 1.0f && d->n_down == 1.0f), d->ordering == DWDF_ORDER_PYTHON };
        }
    }
    *out = p;
    return DWDF_OK;
}

size_t dwdf_mlp_weight_count (const dwdf_mlp_desc* m)
{
    if (m == nullptr || m->hidden <= 0 || m->n_hidden < 0)
        return 0;
    const size_t H = (size_t) m->hidden;
    return 3 * H + (size_t) m->n_hidden * (H * H + H) + H + 1;
}

int dwdf_program_create_neural (const dwdf_node* nodes, int32_t n_nodes, const dwdf_circuit_desc* d, const dwdf_mlp_desc* mlp, dwdf_program** out)
{
    if (mlp == nullptr || d == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    if (d->root_kind != DWDF_ROOT_NEURAL)
        return fail (DWDF_ERR_INVALID, "dwdf_program_create_neural needs root_kind = DWDF_ROOT_NEURAL");
    if (mlp->hidden != 4 && mlp->hidden != 8 && mlp->hidden != 16)
        return fail (DWDF_ERR_UNSUPPORTED, "hidden width %d: the kernels are built for 4, 8 and 16 (the reference's model sizes)", mlp->hidden);
    if (mlp->n_hidden < 1 || mlp->n_hidden > 6)
        return fail (DWDF_ERR_UNSUPPORTED, "%d hidden layers outside [1, 6]", mlp->n_hidden);
    if (int rc = dwdf_program_create (nodes, n_nodes, d, out))
        return rc;
    (*out)->is_neural = true;
    (*out)->mlp = *mlp;
    return DWDF_OK;
}

size_t dwdf_neural_ckpt_bytes (const dwdf_program* prog, int64_t B, int64_t T)
{
    return (prog == nullptr || ! prog->is_neural || B <= 0 || T <= 0) ? 0 : (size_t) nn_ckpt_floats (B, T) * sizeof (float);
}
static int nn_adjoint_chunks (int64_t B, int64_t T)
{
    return (! (g_clip_opts & kOptNoChunks) && (B <= kNnChunkedMaxB || (g_clip_opts & kOptForceChunks)) && T >= 2 * kTimeChunk) ? nn_time_chunks (T) : 1;
}
size_t dwdf_neural_workspace_bytes (const dwdf_program* prog, int64_t B, int64_t T)
{
    if (prog == nullptr || ! prog->is_neural || B <= 0 || T <= 0)
        return 0;
    // worst case over the option switches: the time-parallel layout (more partial vectors, plus (P, Q) and G per chunk)
    const int K = T >= 2 * kTimeChunk ? nn_time_chunks (T) : 1;
    const size_t partials = ((size_t) nn_adjoint_ctas (B, K) * (dwdf_mlp_weight_count (&prog->mlp) + 8) * sizeof (double) + 255) / 256 * 256;
    return partials + (size_t) ((B + 1) / 2) * K * 6 * sizeof (float) + 256;
}

int dwdf_forward_neural (const dwdf_program* prog, const float* params, const float* weights, const float* x, const float* r, float* y, float* state, float* z_ckpt, int64_t B, int64_t T, void* stream)
{
    if (int rc = check_batch (prog, params, x, B, T))
        return rc;
    if (! prog->is_neural)
        return fail (DWDF_ERR_INVALID, "not a neural-root program (dwdf_program_create_neural)");
    if (B == 0 || T == 0)
        return DWDF_OK;
    if (y == nullptr || weights == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    if ((prog->desc.r_node >= 0) != (r != nullptr))
        return fail (DWDF_ERR_INVALID, "the per-sample resistance channel must be given exactly when the program has an r_node");
    // few sequences, long ones: time-parallel (one lane per pair and 256-sample chunk; clipper_kernels.cu explains the scheme)
    int K = 1;
    AsyncScratch scratch;
    if (! (g_clip_opts & kOptNoChunks) && (B <= kNnChunkedMaxB || (g_clip_opts & kOptForceChunks)) && T >= 2 * kTimeChunk)
    {
        K = nn_time_chunks (T);
        DWDF_CUDA (scratch.alloc ((size_t) 4 * ((B + 1) / 2) * K * sizeof (float), (cudaStream_t) stream));
    }
    DWDF_CUDA (launch_nn_forward (prog->mlp.hidden, prog->mlp.n_hidden, prog->desc.ordering == DWDF_ORDER_PYTHON, x, r, y, params, prog->nodes[0].param, prog->nodes[1].param, prog->desc.fs, weights,
                                  (int) dwdf_mlp_weight_count (&prog->mlp), state, z_ckpt, B, T, K, scratch.p, K > 1 ? redone_counter () : nullptr, (cudaStream_t) stream));
    g_launches.fetch_add (K > 1 ? 2 : 1);
    return DWDF_OK;
}

static int backward_neural_impl (bool raw_only, const dwdf_program* prog, const float* params, const float* weights, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode,
                                 int32_t loss_kind, int64_t skip, float* gx, double* grad_w, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream_)
{
    cudaStream_t stream = (cudaStream_t) stream_;
    if (int rc = check_batch (prog, params, x, B, T))
        return rc;
    if (! prog->is_neural)
        return fail (DWDF_ERR_INVALID, "not a neural-root program (dwdf_program_create_neural)");
    if (weights == nullptr || y == nullptr || z_ckpt == nullptr || gy_or_target == nullptr || grad_w == nullptr || out == nullptr || workspace == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    if (grad_mode != DWDF_GRAD_UPSTREAM && grad_mode != DWDF_GRAD_TARGET)
        return fail (DWDF_ERR_INVALID, "unknown grad mode %d", grad_mode);
    if (! loss_kind_ok (loss_kind))
        return fail (DWDF_ERR_INVALID, "unknown loss kind %d", loss_kind);
    if (B == 0 || T == 0)
        return fail (DWDF_ERR_INVALID, "empty batch has no gradient");
    if (loss_kind == DWDF_LOSS_MSE_ESR_AS_CALLED)
    { // the reference's loss as its loop calls it: batch sums -> dL/dy -> the same sweep in upstream mode
        if (grad_mode != DWDF_GRAD_TARGET || raw_only)
            return fail (DWDF_ERR_INVALID, "DWDF_LOSS_MSE_ESR_AS_CALLED is a fused loss of the whole batch: it needs DWDF_GRAD_TARGET (and has no raw-sum form)");
        AsCalledLoss a;
        if (int rc = as_called_prepare (a, y, gy_or_target, B, T, skip < 0 ? 0 : (skip > T ? T : skip), nullptr, stream))
            return rc;
        if (int rc = backward_neural_impl (false, prog, params, weights, x, r, y, z_ckpt, a.ybar, DWDF_GRAD_UPSTREAM, DWDF_LOSS_MSE, 0, gx, grad_w, out, workspace, workspace_bytes, B, T, stream_))
            return rc;
        DWDF_CUDA (launch_loss_write (a.sums, out, stream));
        g_launches.fetch_add (1);
        return DWDF_OK;
    }
    if (workspace_bytes < dwdf_neural_workspace_bytes (prog, B, T))
        return fail (DWDF_ERR_WORKSPACE, "workspace of %zu bytes, need %zu", workspace_bytes, dwdf_neural_workspace_bytes (prog, B, T));
    if ((prog->desc.r_node >= 0) != (r != nullptr))
        return fail (DWDF_ERR_INVALID, "the per-sample resistance channel must be given exactly when the program has an r_node");
    const bool shape_ok = (prog->mlp.n_hidden == 2 && (prog->mlp.hidden == 4 || prog->mlp.hidden == 8 || prog->mlp.hidden == 16)) || (prog->mlp.n_hidden == 4 && (prog->mlp.hidden == 4 || prog->mlp.hidden == 8));
    if (! shape_ok)
        return fail (DWDF_ERR_UNSUPPORTED, "the adjoint kernel is built for the reference's shapes 2x4, 2x8, 2x16, 4x4, 4x8 (got %dx%d)", prog->mlp.n_hidden, prog->mlp.hidden);
    const bool target = grad_mode == DWDF_GRAD_TARGET;
    const int64_t sk = skip < 0 ? 0 : (skip > T ? T : skip);
    const int nw = (int) dwdf_mlp_weight_count (&prog->mlp);
    const int K = nn_adjoint_chunks (B, T);
    const size_t partials_bytes = ((size_t) nn_adjoint_ctas (B, K) * (nw + 8) * sizeof (double) + 255) / 256 * 256;
    float* scratch = K > 1 ? (float*) ((char*) workspace + partials_bytes) : nullptr;
    DWDF_CUDA (launch_nn_adjoint (prog->mlp.hidden, prog->mlp.n_hidden, prog->desc.ordering == DWDF_ORDER_PYTHON, target, x, r, y, gy_or_target, z_ckpt, params, prog->nodes[0].param, prog->nodes[1].param,
                                  prog->desc.fs, weights, nw, (double*) workspace, (int) sk, B, T, K, scratch, gx, stream));
    DWDF_CUDA (launch_nn_reduce ((const double*) workspace, nn_adjoint_ctas (B, K), nw, (double) B * (double) (T - sk), grad_w, out, stream));
    g_launches.fetch_add (K > 1 ? 4 : 2);
    if (! raw_only)
    {
        DWDF_CUDA (launch_nn_scale (nw, target, loss_kind, grad_w, out, stream));
        g_launches.fetch_add (1);
    }
    return DWDF_OK;
}

int dwdf_backward_neural (const dwdf_program* prog, const float* params, const float* weights, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode,
                          int32_t loss_kind, int64_t skip, float* gx, double* grad_w, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream)
{
    return backward_neural_impl (false, prog, params, weights, x, r, y, z_ckpt, gy_or_target, grad_mode, loss_kind, skip, gx, grad_w, out, workspace, workspace_bytes, B, T, stream);
}

int dwdf_backward_neural_raw (const dwdf_program* prog, const float* params, const float* weights, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode,
                              int64_t skip, float* gx, double* grad_w_raw, double* raw, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream)
{
    return backward_neural_impl (true, prog, params, weights, x, r, y, z_ckpt, gy_or_target, grad_mode, DWDF_LOSS_MSE, skip, gx, grad_w_raw, raw, workspace, workspace_bytes, B, T, stream);
}

int dwdf_finalize_neural (const dwdf_program* prog, int32_t grad_mode, int32_t loss_kind, double* grad_w_inout, double* raw_inout, void* stream)
{
    if (prog == nullptr || ! prog->is_neural || grad_w_inout == nullptr || raw_inout == nullptr)
        return fail (DWDF_ERR_INVALID, "dwdf_finalize_neural needs a neural-root program, the summed weight gradients and the summed raw block");
    if (loss_kind != DWDF_LOSS_MSE && loss_kind != DWDF_LOSS_MSE_ESR)
        return fail (DWDF_ERR_INVALID, "loss kind %d has no raw-sum form", loss_kind);
    DWDF_CUDA (launch_nn_scale ((int) dwdf_mlp_weight_count (&prog->mlp), grad_mode == DWDF_GRAD_TARGET, loss_kind, grad_w_inout, raw_inout, (cudaStream_t) stream));
    g_launches.fetch_add (1);
    return DWDF_OK;
}

int dwdf_adam_step_vec (float* w, const double* grad, float* m, float* v, int32_t* step, int64_t n, float lr, float beta1, float beta2, float eps, double grad_scale, void* stream)
{
    if (w == nullptr || grad == nullptr || m == nullptr || v == nullptr || step == nullptr || n < 1)
        return fail (DWDF_ERR_INVALID, "null argument");
    DWDF_CUDA (launch_adam_vec (w, grad, m, v, step, n, lr, beta1, beta2, eps, grad_scale, (cudaStream_t) stream));
    g_launches.fetch_add (1);
    return DWDF_OK;
}

int dwdf_program_destroy (dwdf_program* prog)
{
    delete prog;
    return DWDF_OK;
}

int dwdf_program_is_clipper (const dwdf_program* prog) { return prog != nullptr && prog->is_clipper ? 1 : 0; }
int dwdf_program_n_states (const dwdf_program* prog) { return prog == nullptr ? 0 : ((prog->is_clipper || prog->is_neural) ? 1 : prog->n_states + 1); }

// ---- run-time specialisation of tree programs (tree_jit.cu) ------------------------------------------------------------
int dwdf_program_specialize (dwdf_program* prog)
{
    if (prog == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    if (prog->is_clipper || prog->is_neural)
        return fail (DWDF_ERR_UNSUPPORTED, "the program already runs on kernels written for its circuit (diode clipper / neural root)");
    std::lock_guard<std::mutex> lock (prog->jit_mu);
    if (prog->jit.load () != nullptr)
        return DWDF_OK;
    std::string err;
    TreeJit* j = tree_jit_create (prog->tree, err);
    if (j == nullptr)
        return fail (DWDF_ERR_UNSUPPORTED, "%s", err.c_str ());
    if (! tree_jit_load (j, err))
    {
        tree_jit_destroy (j);
        return fail (DWDF_ERR_CUDA, "%s", err.c_str ());
    }
    prog->jit.store (j, std::memory_order_release);
    return DWDF_OK;
}

int dwdf_program_is_specialized (const dwdf_program* prog) { return prog != nullptr && prog->jit.load () != nullptr ? 1 : 0; }

size_t dwdf_program_specialized_source (const dwdf_program* prog, int32_t part, char* buf, size_t capacity)
{
    if (prog == nullptr || prog->is_clipper || prog->is_neural || part < 0 || part > 2)
        return 0;
    std::string src;
    if (part == 0)
    {
        if (! tree_jit_supported (prog->tree))
            return 0;
        src = tree_jit_full_source (prog->tree);
    }
    else
        src = tree_jit_header (part - 1);
    if (buf != nullptr && capacity > 0)
    {
        const size_t n = src.size () < capacity - 1 ? src.size () : capacity - 1;
        std::memcpy (buf, src.data (), n);
        buf[n] = '\0';
    }
    return src.size () + 1;
}

size_t dwdf_ckpt_bytes (const dwdf_program* prog, int64_t B, int64_t T)
{
    if (prog == nullptr || B <= 0 || T <= 0)
        return 0;
    if (prog->is_clipper)
        return (size_t) (n_segments (T) * B) * sizeof (float);
    if (prog->jit.load () != nullptr) // specialised tree: every state (and the probe's incident wave) at each segment start
        return (size_t) (n_segments (T) * B) * (size_t) (prog->n_states + 1) * sizeof (float);
    return 16; // the interpreter's adjoint keeps its own tape in the workspace
}

size_t dwdf_workspace_bytes (const dwdf_program* prog, int64_t B, int64_t T)
{
    if (prog == nullptr || B <= 0 || T <= 0)
        return 0;
    size_t bytes = partials_bytes (B);
    if (prog->is_clipper)
        bytes += adj_maps_bytes (B, T); // affine maps of the adjoint's time chunks (small batches)
    if (! prog->is_clipper && prog->jit.load () == nullptr) // per-sample tape of the interpreter adjoint: (n_states + 1) floats per sample (the specialised kernels keep none)
        bytes += (size_t) B * (size_t) T * (size_t) (prog->n_states + 1) * sizeof (float);
    return bytes;
}

static int check_batch (const dwdf_program* prog, const void* params, const void* x, int64_t B, int64_t T)
{
    if (B < 0 || T < 0 || T > (int64_t) 1 << 30 || B > (int64_t) 1 << 36)
        return fail (DWDF_ERR_INVALID, "bad batch shape (%lld, %lld)", (long long) B, (long long) T);
    if (prog == nullptr || params == nullptr || (x == nullptr && B * T > 0))
        return fail (DWDF_ERR_INVALID, "null argument");
    return DWDF_OK;
}

static int forward_impl (const dwdf_program* prog, const float* params, const float* x, const float* r, float* y, float* z_ckpt, float* state, int64_t B, int64_t T, cudaStream_t stream)
{
    if (int rc = check_batch (prog, params, x, B, T))
        return rc;
    if (B == 0 || T == 0)
        return DWDF_OK; // empty batch: nothing to do (and the pointers may be null)
    if (y == nullptr)
        return fail (DWDF_ERR_INVALID, "null output");
    if (prog->is_neural)
        return fail (DWDF_ERR_INVALID, "neural-root programs run through dwdf_forward_neural (they carry a weight vector)");
    if ((prog->desc.r_node >= 0) != (r != nullptr))
        return fail (DWDF_ERR_INVALID, "the per-sample resistance channel must be given exactly when the program has an r_node");
    if (prog->clip_r)
    {
        ClipTmaMaps maps;
        const bool tma = tma_usable (x, y, r, B, T) && make_map (&maps.x, x, B, T, kSeg) && make_map (&maps.y, y, B, T, kSeg) && make_map (&maps.r, r, B, T, kSeg);
        DWDF_CUDA (launch_clipper_forward_r (prog->variant, tma, &maps, prog->clip, params, x, r, y, z_ckpt, state, B, T, stream));
    }
    else if (prog->is_clipper)
    {
        ClipTmaMaps maps;
        const bool tma = tma_usable (x, y, nullptr, B, T) && make_map (&maps.x, x, B, T, kFwdTileSamples) && make_map (&maps.y, y, B, T, kFwdTileSamples);
        // symmetric pair, more than one warp's worth of sequences: two sequences per lane (packed fp32x2)
        maps.pair = tma && (prog->variant.mode == kModeApprox || prog->variant.mode == kModeExact) && ! prog->variant.general && B > 32 && ! (g_clip_opts & kOptNoPair) && make_map (&maps.x2, x, B, T, kFwdTileSamples, 64) && make_map (&maps.y2, y, B, T, kFwdTileSamples, 64);
        // fewer sequences than the SMs hold warps: time chunks. Scratch: assumed / end state per (chunk, sequence) and, when
        // the caller keeps no checkpoints, a checkpoint buffer for the verification pass
        AsyncScratch scratch;
        const int cap = tma ? fwd_chunk_cap (B, T) : 1;
        if (cap > 1)
        {
            const size_t zfloats = (size_t) cap * (size_t) B, ckfloats = z_ckpt == nullptr ? (size_t) n_segments (T) * (size_t) B : 0;
            DWDF_CUDA (scratch.alloc ((2 * zfloats + ckfloats) * sizeof (float), stream));
            maps.kcap_fwd = cap;
            maps.zs = scratch.p;
            maps.ze = scratch.p + zfloats;
            maps.redone = redone_counter ();
            if (z_ckpt == nullptr)
                z_ckpt = scratch.p + 2 * zfloats;
        }
        DWDF_CUDA (launch_clipper_forward (prog->variant, tma, &maps, prog->clip, params, x, y, z_ckpt, state, B, T, stream));
    }
    else if (TreeJit* jit = prog->jit.load (std::memory_order_acquire))
    { // run-time specialised kernels (tree_jit.cu): TMA tiles when the rows allow it, else a lane walks its own row
        CUtensorMap mx, my;
        const bool tma = tma_usable (x, y, nullptr, B, T) && make_map (&mx, x, B, T, 32) && make_map (&my, y, B, T, 32);
        std::string err;
        if (! tree_jit_forward (jit, tma ? &mx : nullptr, tma ? &my : nullptr, params, x, y, z_ckpt, state, B, T, stream, err))
            return fail (DWDF_ERR_CUDA, "specialised tree kernel: %s", err.c_str ());
    }
    else
    {
        DWDF_CUDA (launch_tree_forward (prog->tree, params, x, r, y, state, B, T, stream));
    }
    g_launches.fetch_add (1);
    return DWDF_OK;
}

int dwdf_forward (const dwdf_program* prog, const float* params, const float* x, const float* r, float* y, float* z_ckpt, int64_t B, int64_t T, void* stream)
{
    return forward_impl (prog, params, x, r, y, z_ckpt, nullptr, B, T, (cudaStream_t) stream);
}

int dwdf_process_block (const dwdf_program* prog, const float* params, const float* x, const float* r, float* y, float* state, int64_t B, int64_t T, void* stream)
{
    if (state == nullptr)
        return fail (DWDF_ERR_INVALID, "null state");
    return forward_impl (prog, params, x, r, y, nullptr, state, B, T, (cudaStream_t) stream);
}

static int backward_impl (int raw_only, const dwdf_program* prog, const float* params, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode, int32_t loss_kind, int64_t skip, float* gx, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream_)
{
    cudaStream_t stream = (cudaStream_t) stream_;
    if (int rc = check_batch (prog, params, x, B, T))
        return rc;
    if (gy_or_target == nullptr || out == nullptr || workspace == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    if (grad_mode != DWDF_GRAD_UPSTREAM && grad_mode != DWDF_GRAD_TARGET)
        return fail (DWDF_ERR_INVALID, "unknown grad mode %d", grad_mode);
    if (! loss_kind_ok (loss_kind))
        return fail (DWDF_ERR_INVALID, "unknown loss kind %d", loss_kind);
    if (workspace_bytes < dwdf_workspace_bytes (prog, B, T))
        return fail (DWDF_ERR_WORKSPACE, "workspace of %zu bytes, need %zu", workspace_bytes, dwdf_workspace_bytes (prog, B, T));
    if (B == 0 || T == 0)
        return fail (DWDF_ERR_INVALID, "empty batch has no gradient");
    if (prog->is_neural)
        return fail (DWDF_ERR_INVALID, "neural-root programs differentiate through dwdf_backward_neural (the gradient is a weight vector)");
    if (loss_kind == DWDF_LOSS_MSE_ESR_AS_CALLED)
    { // the reference's loss as its loop calls it: batch sums -> dL/dy -> the same sweep in upstream mode
        if (grad_mode != DWDF_GRAD_TARGET || raw_only != 0)
            return fail (DWDF_ERR_INVALID, "DWDF_LOSS_MSE_ESR_AS_CALLED is a fused loss of the whole batch: it needs DWDF_GRAD_TARGET (and has no raw-sum form)");
        if (y == nullptr)
            return fail (DWDF_ERR_INVALID, "DWDF_LOSS_MSE_ESR_AS_CALLED reads the forward output y: it is null");
        AsCalledLoss a;
        if (int rc = as_called_prepare (a, y, gy_or_target, B, T, skip < 0 ? 0 : (skip > T ? T : skip), nullptr, stream))
            return rc;
        if (int rc = backward_impl (0, prog, params, x, r, y, z_ckpt, a.ybar, DWDF_GRAD_UPSTREAM, DWDF_LOSS_MSE, 0, gx, out, workspace, workspace_bytes, B, T, stream_))
            return rc;
        DWDF_CUDA (launch_loss_write (a.sums, out, stream));
        g_launches.fetch_add (1);
        return DWDF_OK;
    }
    if (prog->desc.root_kind == DWDF_ROOT_DIODE_PAIR && prog->desc.root_mode == DWDF_MODE_APPROX_GOOD)
        return fail (DWDF_ERR_UNSUPPORTED, "the 'Good' diode law is forward only");
    if (! prog->differentiable)
        return fail (DWDF_ERR_UNSUPPORTED, "reverse mode covers the wdf_py element set and the inductor, closed by an ideal voltage source, a diode pair or a neural root; this circuit (alpha-transform / Y-parameter / current-source / diode / switch elements, or a current probe) is forward only");
    const bool target = grad_mode == DWDF_GRAD_TARGET;
    const int64_t sk = skip < 0 ? 0 : (skip > T ? T : skip);
    const double count = (double) B * (double) (T - sk);
    double* partials = (double*) workspace;
    if (prog->clip_r)
    {
        if (z_ckpt == nullptr || y == nullptr || r == nullptr)
            return fail (DWDF_ERR_INVALID, "the clipper adjoint reads the output y, the checkpoints z_ckpt that dwdf_forward wrote and the resistance channel r: one of them is null");
        if (gx != nullptr)
            return fail (DWDF_ERR_UNSUPPORTED, "dL/dx with a per-sample resistance channel is not implemented");
        ClipTmaMaps maps;
        const bool tma = tma_usable (x, gy_or_target, y, B, T) && ((uintptr_t) r & 15u) == 0 && make_map (&maps.x, x, B, T, kSeg) && make_map (&maps.y, y, B, T, kSeg) && make_map (&maps.g, gy_or_target, B, T, kSeg) && make_map (&maps.r, r, B, T, kSeg);
        if (tma && adj_chunk_cap (B, T) > 1)
        {
            maps.kcap_adj = adj_chunk_cap (B, T);
            maps.cmaps = (float*) ((char*) workspace + partials_bytes (B));
        }
        DWDF_CUDA (launch_clipper_adjoint_r (prog->variant, tma, &maps, prog->clip, params, x, r, y, z_ckpt, gy_or_target, target, sk, partials, B, T, stream));
        if (raw_only == 2)
            return fail (DWDF_ERR_UNSUPPORTED, "internal: the fused tail does not cover the resistance channel");
        DWDF_CUDA (launch_clipper_finalize_r (prog->clip, params, partials, n_groups (B), nullptr, raw_only != 0, target, loss_kind, count, out, stream));
    }
    else if (prog->is_clipper)
    {
        if (z_ckpt == nullptr || y == nullptr)
            return fail (DWDF_ERR_INVALID, "the clipper adjoint reads the output y and the checkpoints z_ckpt that dwdf_forward wrote: %s is null", y == nullptr ? "y" : "z_ckpt");
        ClipTmaMaps maps;
        const bool tma = gx == nullptr && tma_usable (x, gy_or_target, y, B, T) && make_map (&maps.x, x, B, T, kSeg) && make_map (&maps.y, y, B, T, kSeg) && make_map (&maps.g, gy_or_target, B, T, kSeg);
        if (tma && adj_chunk_cap (B, T) > 1)
        {
            maps.kcap_adj = adj_chunk_cap (B, T);
            maps.cmaps = (float*) ((char*) workspace + partials_bytes (B));
        }
        DWDF_CUDA (launch_clipper_adjoint (prog->variant, tma, &maps, prog->clip, params, x, y, z_ckpt, gy_or_target, target, sk, gx, partials, B, T, stream));
        if (raw_only == 2)
        { // the caller reduces the partials itself (dwdf_train_step_dp: reduction + exchange + chain rule + Adam in one kernel)
            g_launches.fetch_add (1);
            return DWDF_OK;
        }
        DWDF_CUDA (launch_clipper_finalize (prog->clip, params, partials, n_groups (B), nullptr, raw_only != 0, target, loss_kind, count, out, stream));
    }
    else
    {
        if (gx != nullptr)
            return fail (DWDF_ERR_UNSUPPORTED, "dL/dx is available for the diode-clipper program only");
        if ((prog->desc.r_node >= 0) != (r != nullptr))
            return fail (DWDF_ERR_INVALID, "the per-sample resistance channel must be given exactly when the program has an r_node");
        if (TreeJit* jit = prog->jit.load (std::memory_order_acquire))
        { // specialised reverse sweep: no tape; replays 16-sample segments from the checkpoints the specialised forward wrote
            if (z_ckpt == nullptr)
                return fail (DWDF_ERR_INVALID, "a specialised tree program differentiates from the checkpoints of dwdf_forward (z_ckpt, dwdf_ckpt_bytes): z_ckpt is null");
            CUtensorMap mx, mg;
            const bool tma = tma_usable (x, gy_or_target, nullptr, B, T) && make_map (&mx, x, B, T, kSeg) && make_map (&mg, gy_or_target, B, T, kSeg);
            std::string err;
            if (! tree_jit_adjoint (jit, tma ? &mx : nullptr, tma ? &mg : nullptr, params, x, gy_or_target, z_ckpt, target, (int) sk, partials, B, T, stream, err))
                return fail (DWDF_ERR_CUDA, "specialised tree kernel: %s", err.c_str ());
        }
        else
        {
            float* tape = (float*) ((char*) workspace + (((size_t) n_groups (B) * kTreePartialStride * sizeof (double) + 255) / 256) * 256);
            DWDF_CUDA (launch_tree_adjoint (prog->tree, params, x, r, gy_or_target, target, sk, partials, tape, B, T, stream));
        }
        DWDF_CUDA (launch_tree_finalize (prog->tree, params, partials, n_groups (B), nullptr, raw_only != 0, target, loss_kind, count, out, stream));
    }
    g_launches.fetch_add (2);
    return DWDF_OK;
}

int dwdf_backward (const dwdf_program* prog, const float* params, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode, int32_t loss_kind, int64_t skip, float* gx, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream)
{
    return backward_impl (0, prog, params, x, r, y, z_ckpt, gy_or_target, grad_mode, loss_kind, skip, gx, out, workspace, workspace_bytes, B, T, stream);
}

int dwdf_backward_raw (const dwdf_program* prog, const float* params, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode, int64_t skip, float* gx, double* raw, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream)
{
    return backward_impl (1, prog, params, x, r, y, z_ckpt, gy_or_target, grad_mode, DWDF_LOSS_MSE, skip, gx, raw, workspace, workspace_bytes, B, T, stream);
}

int dwdf_finalize (const dwdf_program* prog, const float* params, int32_t grad_mode, int32_t loss_kind, double* raw_inout, void* stream)
{
    if (prog == nullptr || params == nullptr || raw_inout == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    const bool target = grad_mode == DWDF_GRAD_TARGET;
    if (prog->clip_r)
        DWDF_CUDA (launch_clipper_finalize_r (prog->clip, params, nullptr, 0, raw_inout, false, target, loss_kind, 0.0, raw_inout, (cudaStream_t) stream));
    else if (prog->is_clipper)
        DWDF_CUDA (launch_clipper_finalize (prog->clip, params, nullptr, 0, raw_inout, false, target, loss_kind, 0.0, raw_inout, (cudaStream_t) stream));
    else
        DWDF_CUDA (launch_tree_finalize (prog->tree, params, nullptr, 0, raw_inout, false, target, loss_kind, 0.0, raw_inout, (cudaStream_t) stream));
    g_launches.fetch_add (1);
    return DWDF_OK;
}

static int train_impl (bool raw_only, const dwdf_program* prog, const float* params, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream_)
{
    cudaStream_t stream = (cudaStream_t) stream_;
    if (int rc = check_batch (prog, params, x, B, T))
        return rc;
    if (target == nullptr || out == nullptr || workspace == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    if (! prog->is_clipper || prog->clip_r || r != nullptr)
        return fail (DWDF_ERR_UNSUPPORTED, "the fused training pass exists for the diode-clipper program without a resistance channel only");
    if (loss_kind == DWDF_LOSS_MSE_ESR_AS_CALLED)
        return fail (DWDF_ERR_UNSUPPORTED, "DWDF_LOSS_MSE_ESR_AS_CALLED needs the batch's output energy before the gradient sweep: use dwdf_forward + dwdf_backward (or dwdf_train_step), not the one-sweep pass");
    if (loss_kind != DWDF_LOSS_MSE && loss_kind != DWDF_LOSS_MSE_ESR)
        return fail (DWDF_ERR_INVALID, "unknown loss kind %d", loss_kind);
    if (workspace_bytes < dwdf_workspace_bytes (prog, B, T))
        return fail (DWDF_ERR_WORKSPACE, "workspace of %zu bytes, need %zu", workspace_bytes, dwdf_workspace_bytes (prog, B, T));
    if (B == 0 || T == 0)
        return fail (DWDF_ERR_INVALID, "empty batch has no gradient");
    const int64_t sk = skip < 0 ? 0 : (skip > T ? T : skip);
    double* partials = (double*) workspace;
    ClipTmaMaps maps[2];
    bool tma = tma_usable (x, target, y, B, T) && make_map (&maps[0].x, x, B, T, 32) && make_map (&maps[0].y, target, B, T, 32);
    if (tma && y != nullptr)
        tma = make_map (&maps[1].y, y, B, T, 32);
    else if (tma)
        maps[1].y = maps[0].y;
    // symmetric pair (approx or exact root), more than a warp's worth of sequences: two sequences per lane, [64 x 16] tiles of x, target (and y)
    if (tma && (prog->variant.mode == kModeApprox || prog->variant.mode == kModeExact) && ! prog->variant.general && B > 32 && ! (g_clip_opts & kOptNoPair))
    {
        maps[0].pair = make_map (&maps[0].x2, x, B, T, kSeg, 64) && make_map (&maps[0].y2, target, B, T, kSeg, 64) && (y == nullptr || make_map (&maps[1].y2, y, B, T, kSeg, 64));
        if (maps[0].pair && y == nullptr)
            maps[1].y2 = maps[0].y2;
    }
    DWDF_CUDA (launch_clipper_train (prog->variant, tma, maps, prog->clip, params, x, target, sk, y, partials, B, T, stream));
    DWDF_CUDA (launch_clipper_finalize (prog->clip, params, partials, n_groups (B), nullptr, raw_only, true, loss_kind, (double) B * (double) (T - sk), out, stream));
    g_launches.fetch_add (2);
    return DWDF_OK;
}

int dwdf_train_pass (const dwdf_program* prog, const float* params, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream)
{
    return train_impl (false, prog, params, x, r, target, loss_kind, skip, y, out, workspace, workspace_bytes, B, T, stream);
}

int dwdf_train_pass_raw (const dwdf_program* prog, const float* params, const float* x, const float* r, const float* target, int64_t skip, float* y, double* raw, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream)
{
    return train_impl (true, prog, params, x, r, target, DWDF_LOSS_MSE, skip, y, raw, workspace, workspace_bytes, B, T, stream);
}

int dwdf_adam_step (float* params, const double* out, float* m, float* v, int32_t* step, int32_t n_params, float lr, const float* lr_per_slot, float beta1, float beta2, float eps, double grad_scale, const float* lo, const float* hi, void* stream)
{
    if (params == nullptr || out == nullptr || m == nullptr || v == nullptr || step == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    if (n_params < 1 || n_params > DWDF_MAX_PARAMS)
        return fail (DWDF_ERR_INVALID, "n_params out of range");
    DWDF_CUDA (launch_adam (params, out, m, v, step, n_params, lr, lr_per_slot, beta1, beta2, eps, grad_scale, lo, hi, (cudaStream_t) stream));
    g_launches.fetch_add (1);
    return DWDF_OK;
}

// ---- multi-GPU: one process per GPU, the step's single exchange over peer memory --------------------------------
struct dwdf_comm
{
    int rank = 0, world = 1, device = 0;
    char* mailbox = nullptr; // this rank's mailbox (cudaMalloc: exportable with CUDA IPC)
    char* peers[kDpMaxWorld] = {};
    bool opened[kDpMaxWorld] = {};
    bool connected = false;
    double timeout_s = 20.0;
};

static bool comm_peers (const dwdf_comm* c, DpPeers& dp)
{
    if (c == nullptr || ! c->connected)
        return false;
    dp.rank = c->rank;
    dp.world = c->world;
    dp.timeout_ns = (unsigned long long) (c->timeout_s * 1e9);
    for (int p = 0; p < kDpMaxWorld; ++p)
        dp.mailbox[p] = p < c->world ? c->peers[p] : nullptr;
    return true;
}

int dwdf_comm_create (int32_t rank, int32_t world, dwdf_comm** out)
{
    if (out == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    *out = nullptr;
    if (world < 1 || world > kDpMaxWorld || rank < 0 || rank >= world)
        return fail (DWDF_ERR_INVALID, "rank %d / world %d (at most %d ranks of one node)", rank, world, kDpMaxWorld);
    dwdf_comm* c = new (std::nothrow) dwdf_comm;
    if (c == nullptr)
        return fail (DWDF_ERR_INVALID, "out of memory");
    c->rank = rank;
    c->world = world;
    cudaError_t e = cudaGetDevice (&c->device);
    if (e == cudaSuccess)
        e = cudaMalloc ((void**) &c->mailbox, dp_mailbox_bytes (world));
    if (e == cudaSuccess)
        e = cudaMemset (c->mailbox, 0, dp_mailbox_bytes (world));
    if (e == cudaSuccess)
        e = cudaDeviceSynchronize ();
    if (e != cudaSuccess)
    {
        delete c;
        return cuda_fail (e, "dwdf_comm_create");
    }
    c->peers[rank] = c->mailbox;
    c->connected = world == 1;
    *out = c;
    return DWDF_OK;
}

size_t dwdf_comm_handle_bytes (void) { return sizeof (cudaIpcMemHandle_t); }

int dwdf_comm_get_handle (const dwdf_comm* comm, void* handle_out)
{
    if (comm == nullptr || handle_out == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    DWDF_CUDA (cudaIpcGetMemHandle (&h, comm->mailbox));
    std::memcpy (handle_out, &h, sizeof (h));
    return DWDF_OK;
}

int dwdf_comm_connect (dwdf_comm* comm, const void* handles)
{
    if (comm == nullptr || handles == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    for (int p = 0; p < comm->world; ++p)
    {
        if (p == comm->rank || comm->opened[p])
            continue;
        cudaIpcMemHandle_t h;
        std::memcpy (&h, (const char*) handles + (size_t) p * sizeof (h), sizeof (h));
        void* ptr = nullptr;
        DWDF_CUDA (cudaIpcOpenMemHandle (&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        comm->peers[p] = (char*) ptr;
        comm->opened[p] = true;
    }
    comm->connected = true;
    return DWDF_OK;
}

int dwdf_comm_set_timeout (dwdf_comm* comm, double seconds)
{
    if (comm == nullptr || ! (seconds > 0.0))
        return fail (DWDF_ERR_INVALID, "bad argument");
    comm->timeout_s = seconds;
    return DWDF_OK;
}

int dwdf_comm_destroy (dwdf_comm* comm)
{
    if (comm == nullptr)
        return DWDF_OK;
    for (int p = 0; p < comm->world; ++p)
        if (comm->opened[p])
            cudaIpcCloseMemHandle (comm->peers[p]);
    if (comm->mailbox != nullptr)
        cudaFree (comm->mailbox);
    delete comm;
    return DWDF_OK;
}

int dwdf_allreduce_sum (const dwdf_comm* comm, double* inout, int64_t n, void* stream)
{
    DpPeers dp;
    if (inout == nullptr || ! comm_peers (comm, dp))
        return fail (DWDF_ERR_INVALID, "dwdf_allreduce_sum needs a connected communicator (dwdf_comm_create / _connect)");
    if (n < 1 || n > kDpSlotDoubles - 1)
        return fail (DWDF_ERR_INVALID, "%lld doubles: one exchange carries at most %d", (long long) n, kDpSlotDoubles - 1);
    if (dp.world == 1)
        return DWDF_OK;
    DWDF_CUDA (launch_peer_allreduce (inout, (int) n, dp, (cudaStream_t) stream));
    g_launches.fetch_add (1);
    return DWDF_OK;
}

static int train_step_impl (const dwdf_program* prog, const DpPeers& dp, float* params, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, float* z_ckpt, double* out, void* workspace,
                            size_t workspace_bytes, float* m, float* v, int32_t* step, float lr, const float* lr_per_slot, float beta1, float beta2, float eps, const float* lo, const float* hi, int64_t B, int64_t T, cudaStream_t stream)
{
    if (B < 1 || T < 1)
        return fail (DWDF_ERR_INVALID, "a training step needs at least one sequence (B = %lld, T = %lld)", (long long) B, (long long) T);
    if (! loss_kind_ok (loss_kind))
        return fail (DWDF_ERR_INVALID, "unknown loss kind %d", loss_kind);
    g_prof.mark (0, stream);
    if (int rc = dwdf_forward (prog, params, x, r, y, z_ckpt, B, T, stream))
        return rc;
    g_prof.mark (1, stream);
    const int64_t sk = skip < 0 ? 0 : (skip > T ? T : skip);
    if (loss_kind == DWDF_LOSS_MSE_ESR_AS_CALLED)
    { // batch sums (summed over ranks) -> dL/dy -> reverse sweep in upstream mode -> [exchange of the raw sums] -> chain rule -> Adam
        AsCalledLoss a;
        if (int rc = as_called_prepare (a, y, target, B, T, sk, &dp, stream))
            return rc;
        if (int rc = backward_impl (dp.world > 1 ? 1 : 0, prog, params, x, r, y, z_ckpt, a.ybar, DWDF_GRAD_UPSTREAM, DWDF_LOSS_MSE, 0, nullptr, out, workspace, workspace_bytes, B, T, stream))
            return rc;
        g_prof.mark (2, stream);
        if (dp.world > 1)
        {
            DWDF_CUDA (launch_peer_allreduce (out, DWDF_OUT_LEN, dp, stream));
            g_launches.fetch_add (1);
            if (int rc = dwdf_finalize (prog, params, DWDF_GRAD_UPSTREAM, DWDF_LOSS_MSE, out, stream))
                return rc;
        }
        DWDF_CUDA (launch_loss_write (a.sums, out, stream));
        g_launches.fetch_add (1);
        int rc = m == nullptr ? DWDF_OK : dwdf_adam_step (params, out, m, v, step, prog->desc.n_params, lr, lr_per_slot, beta1, beta2, eps, 1.0, lo, hi, stream);
        g_prof.mark (3, stream);
        return rc;
    }
    if (prog->is_clipper && ! prog->clip_r)
    { // adjoint -> ONE kernel: reduction of the partials + exchange over peer memory + chain rule + loss + Adam
        if (int rc = backward_impl (2, prog, params, x, r, y, z_ckpt, target, DWDF_GRAD_TARGET, loss_kind, skip, nullptr, out, workspace, workspace_bytes, B, T, stream))
            return rc;
        g_prof.mark (2, stream);
        DWDF_CUDA (launch_clipper_finalize_dp (prog->clip, params, (const double*) workspace, n_groups (B), true, loss_kind, (double) B * (double) (T - sk), out, dp, m, v, step, prog->desc.n_params, lr, lr_per_slot, beta1, beta2, eps, lo, hi,
                                               stream));
        g_launches.fetch_add (1);
        g_prof.mark (3, stream);
        return DWDF_OK;
    }
    // other trees: raw sums -> exchange -> finalize -> Adam
    if (int rc = backward_impl (dp.world > 1 ? 1 : 0, prog, params, x, r, y, z_ckpt, target, DWDF_GRAD_TARGET, dp.world > 1 ? DWDF_LOSS_MSE : loss_kind, skip, nullptr, out, workspace, workspace_bytes, B, T, stream))
        return rc;
    g_prof.mark (2, stream);
    if (dp.world > 1)
    {
        DWDF_CUDA (launch_peer_allreduce (out, DWDF_OUT_LEN, dp, stream));
        g_launches.fetch_add (1);
        if (int rc = dwdf_finalize (prog, params, DWDF_GRAD_TARGET, loss_kind, out, stream))
            return rc;
    }
    int rc = m == nullptr ? DWDF_OK : dwdf_adam_step (params, out, m, v, step, prog->desc.n_params, lr, lr_per_slot, beta1, beta2, eps, 1.0, lo, hi, stream);
    g_prof.mark (3, stream);
    return rc;
}

int dwdf_train_step (const dwdf_program* prog, float* params, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, float* z_ckpt, double* out, void* workspace,
                     size_t workspace_bytes, float* m, float* v, int32_t* step, float lr, const float* lr_per_slot, float beta1, float beta2, float eps, const float* lo, const float* hi, int64_t B, int64_t T, void* stream)
{
    if (prog == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    DpPeers solo {};
    solo.world = 1;
    return train_step_impl (prog, solo, params, x, r, target, loss_kind, skip, y, z_ckpt, out, workspace, workspace_bytes, m, v, step, lr, lr_per_slot, beta1, beta2, eps, lo, hi, B, T, (cudaStream_t) stream);
}

int dwdf_train_step_dp (const dwdf_program* prog, const dwdf_comm* comm, float* params, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, float* z_ckpt, double* out, void* workspace,
                        size_t workspace_bytes, float* m, float* v, int32_t* step, float lr, const float* lr_per_slot, float beta1, float beta2, float eps, const float* lo, const float* hi, int64_t B, int64_t T, void* stream)
{
    DpPeers dp;
    if (prog == nullptr || ! comm_peers (comm, dp))
        return fail (DWDF_ERR_INVALID, "dwdf_train_step_dp needs a program and a connected communicator");
    return train_step_impl (prog, dp, params, x, r, target, loss_kind, skip, y, z_ckpt, out, workspace, workspace_bytes, m, v, step, lr, lr_per_slot, beta1, beta2, eps, lo, hi, B, T, (cudaStream_t) stream);
}

// Neural root: one whole training step of clipper_pot.py:246-269 (the trainable variables are the network's kernels and
// biases) in one call, on this rank's shard when a communicator is given: forward + reverse sweep + fixed-order reduction
// [+ exchange of the raw weight-gradient sums and the loss sums over peer memory] + loss scale + Adam on the weights.
int dwdf_train_step_neural (const dwdf_program* prog, const dwdf_comm* comm, const float* params, float* weights, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, float* z_ckpt,
                            double* grad_w, double* out, void* workspace, size_t workspace_bytes, float* m, float* v, int32_t* step, float lr, float beta1, float beta2, float eps, int64_t B, int64_t T, void* stream_)
{
    cudaStream_t stream = (cudaStream_t) stream_;
    if (prog == nullptr || ! prog->is_neural)
        return fail (DWDF_ERR_INVALID, "dwdf_train_step_neural needs a neural-root program");
    if (B < 1 || T < 1)
        return fail (DWDF_ERR_INVALID, "a training step needs at least one sequence (B = %lld, T = %lld)", (long long) B, (long long) T);
    if (! loss_kind_ok (loss_kind))
        return fail (DWDF_ERR_INVALID, "unknown loss kind %d", loss_kind);
    DpPeers dp {};
    dp.world = 1;
    if (comm != nullptr && ! comm_peers (comm, dp))
        return fail (DWDF_ERR_INVALID, "the communicator is not connected (dwdf_comm_connect)");
    if (int rc = dwdf_forward_neural (prog, params, weights, x, r, y, nullptr, z_ckpt, B, T, stream_))
        return rc;
    const int nw = (int) dwdf_mlp_weight_count (&prog->mlp);
    const int64_t sk = skip < 0 ? 0 : (skip > T ? T : skip);
    const bool as_called = loss_kind == DWDF_LOSS_MSE_ESR_AS_CALLED, many = dp.world > 1;
    AsCalledLoss a;
    if (as_called)
        if (int rc = as_called_prepare (a, y, target, B, T, sk, &dp, stream))
            return rc;
    if (int rc = backward_neural_impl (many, prog, params, weights, x, r, y, z_ckpt, as_called ? a.ybar : target, as_called ? DWDF_GRAD_UPSTREAM : DWDF_GRAD_TARGET, as_called ? DWDF_LOSS_MSE : loss_kind, as_called ? 0 : skip,
                                       nullptr, grad_w, out, workspace, workspace_bytes, B, T, stream_))
        return rc;
    if (many)
    {
        for (int w0 = 0; w0 < nw; w0 += kDpSlotDoubles - 1)
        {
            DWDF_CUDA (launch_peer_allreduce (grad_w + w0, nw - w0 < kDpSlotDoubles - 1 ? nw - w0 : kDpSlotDoubles - 1, dp, stream));
            g_launches.fetch_add (1);
        }
        DWDF_CUDA (launch_peer_allreduce (out, DWDF_OUT_LEN, dp, stream));
        DWDF_CUDA (launch_nn_scale (nw, ! as_called, as_called ? DWDF_LOSS_MSE : loss_kind, grad_w, out, stream));
        g_launches.fetch_add (2);
    }
    if (as_called)
    {
        DWDF_CUDA (launch_loss_write (a.sums, out, stream));
        g_launches.fetch_add (1);
    }
    if (m != nullptr)
        return dwdf_adam_step_vec (weights, grad_w, m, v, step, nw, lr, beta1, beta2, eps, 1.0, stream_);
    return DWDF_OK;
}

int dwdf_profile_begin (int32_t max_steps)
{
    if (max_steps < 1 || max_steps > 65536)
        return fail (DWDF_ERR_INVALID, "max_steps outside [1, 65536]");
    std::lock_guard<std::mutex> lock (g_prof.mu);
    for (cudaEvent_t e : g_prof.ev)
        cudaEventDestroy (e);
    g_prof.ev.assign ((size_t) 4 * max_steps, nullptr);
    for (cudaEvent_t& e : g_prof.ev)
        DWDF_CUDA (cudaEventCreate (&e));
    g_prof.cap = max_steps;
    g_prof.n = 0;
    g_prof.on = true;
    return DWDF_OK;
}

int dwdf_profile_end (double* ms_forward, double* ms_adjoint, double* ms_tail, int32_t* steps)
{
    std::lock_guard<std::mutex> lock (g_prof.mu);
    g_prof.on = false;
    double sum[3] = { 0.0, 0.0, 0.0 };
    const int n = g_prof.n;
    for (int i = 0; i < n; ++i)
    {
        DWDF_CUDA (cudaEventSynchronize (g_prof.ev[(size_t) 4 * i + 3]));
        for (int k = 0; k < 3; ++k)
        {
            float ms = 0.0f;
            DWDF_CUDA (cudaEventElapsedTime (&ms, g_prof.ev[(size_t) 4 * i + k], g_prof.ev[(size_t) 4 * i + k + 1]));
            sum[k] += ms;
        }
    }
    if (ms_forward != nullptr)
        *ms_forward = n > 0 ? sum[0] / n : 0.0;
    if (ms_adjoint != nullptr)
        *ms_adjoint = n > 0 ? sum[1] / n : 0.0;
    if (ms_tail != nullptr)
        *ms_tail = n > 0 ? sum[2] / n : 0.0;
    if (steps != nullptr)
        *steps = n;
    for (cudaEvent_t e : g_prof.ev)
        cudaEventDestroy (e);
    g_prof.ev.clear ();
    g_prof.cap = g_prof.n = 0;
    return DWDF_OK;
}

// ---- end-to-end calls with host buffers ----------------------------------------------------------
// Library-owned device arena (grow-only, one per process) and a small set of streams: the batch is
// cut into row chunks so that the host->device copy of chunk k+1, the kernels of chunk k and the
// device->host copy of chunk k-1 overlap (PCIe is full duplex; the copy engines run beside the SMs).
namespace
{
struct Arena
{
    std::mutex mu;
    char* dev = nullptr;
    size_t cap = 0;
    cudaStream_t streams[3] = { nullptr, nullptr, nullptr };
    cudaEvent_t done[3] = { nullptr, nullptr, nullptr };
    cudaEvent_t params_ready = nullptr;
    int ensure (size_t bytes)
    {
        if (streams[0] == nullptr)
        {
            DWDF_CUDA (cudaEventCreateWithFlags (&params_ready, cudaEventDisableTiming));
            for (int i = 0; i < 3; ++i)
            {
                DWDF_CUDA (cudaStreamCreateWithFlags (&streams[i], cudaStreamNonBlocking));
                DWDF_CUDA (cudaEventCreateWithFlags (&done[i], cudaEventDisableTiming));
            }
        }
        if (bytes <= cap)
            return DWDF_OK;
        if (dev != nullptr)
            DWDF_CUDA (cudaFree (dev));
        dev = nullptr;
        cap = 0;
        DWDF_CUDA (cudaMalloc ((void**) &dev, bytes));
        cap = bytes;
        return DWDF_OK;
    }
};
Arena g_arena;
size_t align256 (size_t n) { return (n + 255) / 256 * 256; }
constexpr int64_t kChunkRows = 4096; // rows per pipelined chunk (64 MiB of x at T = 4096)
} // namespace

int dwdf_forward_host (const dwdf_program* prog, const float* params_host, const float* x_host, const float* r_host, float* y_host, int64_t B, int64_t T)
{
    if (int rc = check_batch (prog, params_host, x_host, B, T))
        return rc;
    if (y_host == nullptr)
        return fail (DWDF_ERR_INVALID, "null output");
    if (B == 0 || T == 0)
        return DWDF_OK;
    std::lock_guard<std::mutex> lock (g_arena.mu);
    const size_t row = (size_t) T * sizeof (float), bt = (size_t) B * row;
    const size_t off_x = 256, off_y = off_x + align256 (bt), off_r = off_y + align256 (bt);
    if (int rc = g_arena.ensure (off_r + (r_host != nullptr ? align256 (bt) : 0)))
        return rc;
    float* d_params = (float*) g_arena.dev;
    float *d_x = (float*) (g_arena.dev + off_x), *d_y = (float*) (g_arena.dev + off_y), *d_r = r_host != nullptr ? (float*) (g_arena.dev + off_r) : nullptr;
    DWDF_CUDA (cudaMemcpyAsync (d_params, params_host, sizeof (float) * prog->desc.n_params, cudaMemcpyHostToDevice, g_arena.streams[0]));
    DWDF_CUDA (cudaEventRecord (g_arena.params_ready, g_arena.streams[0]));
    int k = 0;
    for (int64_t b0 = 0; b0 < B; b0 += kChunkRows, ++k)
    {
        const int64_t nb = B - b0 < kChunkRows ? B - b0 : kChunkRows;
        cudaStream_t s = g_arena.streams[k % 3];
        if (k < 3 && k > 0)
            DWDF_CUDA (cudaStreamWaitEvent (s, g_arena.params_ready, 0)); // parameters uploaded
        DWDF_CUDA (cudaMemcpyAsync (d_x + b0 * T, x_host + b0 * T, (size_t) nb * row, cudaMemcpyHostToDevice, s));
        if (d_r != nullptr)
            DWDF_CUDA (cudaMemcpyAsync (d_r + b0 * T, r_host + b0 * T, (size_t) nb * row, cudaMemcpyHostToDevice, s));
        if (int rc = forward_impl (prog, d_params, d_x + b0 * T, d_r != nullptr ? d_r + b0 * T : nullptr, d_y + b0 * T, nullptr, nullptr, nb, T, s))
            return rc;
        DWDF_CUDA (cudaMemcpyAsync (y_host + b0 * T, d_y + b0 * T, (size_t) nb * row, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < 3; ++i)
        DWDF_CUDA (cudaStreamSynchronize (g_arena.streams[i]));
    return DWDF_OK;
}

int dwdf_grad_host (const dwdf_program* prog, const float* params_host, const float* x_host, const float* r_host, const float* g_host, int32_t grad_mode, int32_t loss_kind, int64_t skip, float* y_host, double* out_host, int64_t B, int64_t T)
{
    if (int rc = check_batch (prog, params_host, x_host, B, T))
        return rc;
    if (g_host == nullptr || out_host == nullptr)
        return fail (DWDF_ERR_INVALID, "null argument");
    if (! prog->is_clipper || prog->clip_r || r_host != nullptr)
        return fail (DWDF_ERR_UNSUPPORTED, "dwdf_grad_host covers the diode-clipper program without a resistance channel; use the device API otherwise");
    if (B == 0 || T == 0)
        return fail (DWDF_ERR_INVALID, "empty batch has no gradient");
    if (loss_kind != DWDF_LOSS_MSE && loss_kind != DWDF_LOSS_MSE_ESR)
        return fail (loss_kind == DWDF_LOSS_MSE_ESR_AS_CALLED ? DWDF_ERR_UNSUPPORTED : DWDF_ERR_INVALID, "dwdf_grad_host pipelines the batch in chunks: loss kind %d is not available here (DWDF_LOSS_MSE_ESR_AS_CALLED needs the whole batch's output energy before the sweep; use the device API)", loss_kind);
    std::lock_guard<std::mutex> lock (g_arena.mu);
    const bool target = grad_mode == DWDF_GRAD_TARGET;
    const int64_t sk = skip < 0 ? 0 : (skip > T ? T : skip);
    const size_t row = (size_t) T * sizeof (float), bt = (size_t) B * row;
    const size_t ck = align256 ((size_t) (n_segments (T) * B) * sizeof (float));
    const size_t pb = align256 ((size_t) (n_groups (B) + B / kChunkRows + 2) * kPartialStride * sizeof (double));
    const size_t off_out = 256, off_x = 1024, off_g = off_x + align256 (bt), off_y = off_g + align256 (bt), off_ck = off_y + align256 (bt), off_p = off_ck + ck;
    if (int rc = g_arena.ensure (off_p + pb))
        return rc;
    float* d_params = (float*) g_arena.dev;
    double* d_out = (double*) (g_arena.dev + off_out);
    float *d_x = (float*) (g_arena.dev + off_x), *d_g = (float*) (g_arena.dev + off_g), *d_y = (float*) (g_arena.dev + off_y), *d_ck = (float*) (g_arena.dev + off_ck);
    double* d_part = (double*) (g_arena.dev + off_p);
    DWDF_CUDA (cudaMemcpyAsync (d_params, params_host, sizeof (float) * prog->desc.n_params, cudaMemcpyHostToDevice, g_arena.streams[0]));
    DWDF_CUDA (cudaEventRecord (g_arena.params_ready, g_arena.streams[0]));
    int k = 0;
    int64_t group0 = 0;
    for (int64_t b0 = 0; b0 < B; b0 += kChunkRows, ++k)
    {
        const int64_t nb = B - b0 < kChunkRows ? B - b0 : kChunkRows;
        cudaStream_t s = g_arena.streams[k % 3];
        if (k < 3 && k > 0)
            DWDF_CUDA (cudaStreamWaitEvent (s, g_arena.params_ready, 0));
        DWDF_CUDA (cudaMemcpyAsync (d_x + b0 * T, x_host + b0 * T, (size_t) nb * row, cudaMemcpyHostToDevice, s));
        DWDF_CUDA (cudaMemcpyAsync (d_g + b0 * T, g_host + b0 * T, (size_t) nb * row, cudaMemcpyHostToDevice, s));
        float* ckc = d_ck + (size_t) n_segments (T) * b0; // each chunk owns a (segments, nb) checkpoint block
        if (int rc = forward_impl (prog, d_params, d_x + b0 * T, nullptr, d_y + b0 * T, ckc, nullptr, nb, T, s))
            return rc;
        if (y_host != nullptr)
            DWDF_CUDA (cudaMemcpyAsync (y_host + b0 * T, d_y + b0 * T, (size_t) nb * row, cudaMemcpyDeviceToHost, s));
        ClipTmaMaps maps;
        const bool tma = tma_usable (d_x + b0 * T, d_g + b0 * T, d_y + b0 * T, nb, T) && make_map (&maps.x, d_x + b0 * T, nb, T, kSeg) && make_map (&maps.y, d_y + b0 * T, nb, T, kSeg) && make_map (&maps.g, d_g + b0 * T, nb, T, kSeg);
        DWDF_CUDA (launch_clipper_adjoint (prog->variant, tma, &maps, prog->clip, d_params, d_x + b0 * T, d_y + b0 * T, ckc, d_g + b0 * T, target, sk, nullptr, d_part + group0 * kPartialStride, nb, T, s));
        g_launches.fetch_add (1);
        group0 += n_groups (nb);
        DWDF_CUDA (cudaEventRecord (g_arena.done[k % 3], s));
    }
    // join on stream 0, reduce, read back
    for (int i = 1; i < 3; ++i)
        DWDF_CUDA (cudaStreamWaitEvent (g_arena.streams[0], g_arena.done[i], 0));
    DWDF_CUDA (launch_clipper_finalize (prog->clip, d_params, d_part, group0, nullptr, false, target, loss_kind, (double) B * (double) (T - sk), d_out, g_arena.streams[0]));
    g_launches.fetch_add (1);
    DWDF_CUDA (cudaMemcpyAsync (out_host, d_out, DWDF_OUT_LEN * sizeof (double), cudaMemcpyDeviceToHost, g_arena.streams[0]));
    for (int i = 0; i < 3; ++i)
        DWDF_CUDA (cudaStreamSynchronize (g_arena.streams[i]));
    return DWDF_OK;
}

} // extern "C"
