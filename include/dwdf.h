/* dwdf.h — C ABI of libdwdf.so, the B200-native differentiable wave-digital-filter engine.
 *
 * The reference (jatinchowdhury18/differentiable-wdfs) has no FFI: its Python half calls
 * TensorFlow eager ops sample by sample and its C++ half is header-only templates. The boundary
 * this library sits behind is therefore the reference's *element protocol* and the script-level
 * forward()/GradientTape contract. Each entry point below cites the reference
 * interface it replaces. Plain pointers and sizes only; no torch / C++ types cross this boundary;
 * every function returns a status (0 = ok) and never throws.
 *
 * Buffers are caller-owned. `x`, `y`, `r`, `gy_or_target`, `gx`, `params`, `workspace`, `out` are
 * DEVICE pointers in the dwdf_forward / dwdf_backward / dwdf_train_* calls and HOST pointers in
 * the *_host calls (which stage through library-owned device memory and pinned bounce buffers).
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Launches are
 * asynchronous on that stream.
 *
 * Tensors: x, y, r, gy_or_target, gx are (B, T) fp32, batch-major, contiguous — the layout of the
 * reference's training batches (clipper_pot.py:61-80; channel 0 = x, channel 1 = r).
 */
#ifndef DWDF_H
#define DWDF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DWDF_VERSION 4
#if defined(__GNUC__)
#define DWDF_API __attribute__ ((visibility ("default")))
#else
#define DWDF_API
#endif

/* ---- status codes -------------------------------------------------------------------------- */
enum dwdf_status
{
    DWDF_OK = 0,
    DWDF_ERR_INVALID = 1, /* bad argument / malformed tree                     */
    DWDF_ERR_UNSUPPORTED = 2, /* valid, but not implemented for this combination   */
    DWDF_ERR_CUDA = 3, /* a CUDA runtime / driver call failed               */
    DWDF_ERR_NO_DEVICE = 4, /* no CUDA device: there is NO CPU fallback          */
    DWDF_ERR_WORKSPACE = 5 /* workspace too small                               */
};

/* ---- circuit description ------------------------------------------------------------------- */
/* One-port leaves and adaptors. Same set as wdf_py/lib/tf_wdf.py:31-214 and, on the C++ side,
 * wdf_t.h ResistorT:69, CapacitorT:118, ResistiveVoltageSourceT:693, WDFSeriesT:508,
 * WDFParallelT:448, PolarityInverterT:558. */
enum dwdf_node_kind
{
    DWDF_RESISTOR = 0, /* tf_wdf.py:62-88   value = R [ohm]                                      */
    DWDF_CAPACITOR = 1, /* tf_wdf.py:91-126  value = C [farad]; port resistance 1/(2 C fs)        */
    DWDF_RESISTIVE_VS = 2, /* tf_wdf.py:31-59   value = R [ohm]; reflected wave = source voltage     */
    DWDF_SERIES = 3, /* tf_wdf.py:129-155 child1 = P1, child2 = P2                             */
    DWDF_PARALLEL = 4, /* tf_wdf.py:158-192 child1 = P1, child2 = P2                             */
    DWDF_INVERTER = 5, /* tf_wdf.py:195-214 child1 = P1                                          */
    /* the remaining chowdsp_wdf one-ports and the two-port (wdf_t.h); forward and streaming everywhere, reverse mode for the inductor */
    DWDF_INDUCTOR = 6, /* wdf_t.h:280-348   value = L [henry]; port resistance 2 L fs; reflects -z            */
    DWDF_CAPACITOR_ALPHA = 7, /* wdf_t.h:190-276 value = C, params[param + 1] = alpha (0 backward Euler .. 1 bilinear) */
    DWDF_INDUCTOR_ALPHA = 8, /* wdf_t.h:352-444  value = L, params[param + 1] = alpha                            */
    DWDF_RESISTIVE_CS = 9, /* wdf_t.h:786-846   value = R [ohm]; reflected wave = R * source current          */
    DWDF_Y_PARAMETER = 10 /* wdf_t.h:597-654   child1 = port1; params[param .. param + 3] = y11, y12, y21, y22 */
};

/* Non-adaptable root closing the tree. */
enum dwdf_root_kind
{
    DWDF_ROOT_IDEAL_VS = 0, /* tf_wdf.py:13-28 / wdf_t.h:658-689: b = -a + 2 Vs, Vs = x[n]            */
    DWDF_ROOT_DIODE_PAIR = 1, /* analytic antiparallel diode pair (see dwdf_root_mode)                 */
    DWDF_ROOT_NEURAL = 2, /* b = -MLP(a, ln Rp): layers.py:42-82 (DenseRootModel), DiodePairNeuralModel.h:62-75;
                            programs of this kind are created by dwdf_program_create_neural                */
    DWDF_ROOT_IDEAL_CS = 3, /* wdf_t.h:746-784: b = 2 R Is + a, Is = x[n]                                       */
    DWDF_ROOT_DIODE = 4, /* wdf_t.h:987-1072: single diode, eq. 10 with omega4; param_Is / param_nabla as the pair */
    DWDF_ROOT_SWITCH = 5 /* wdf_t.h:1076-1106: b = -a closed (root_mode 1), b = a open (root_mode 0)             */
};

/* How the diode-pair root evaluates the Wright-omega function. */
enum dwdf_root_mode
{
    DWDF_MODE_APPROX = 0, /* wdft::DiodePairT<Best>, wdf_t.h:917-924 with omega4 of omega.h:172-177  */
    DWDF_MODE_EXACT = 1, /* Toms917DiodePair.h:51-67 / diode_pretraining.py:39-60: omega to fp32
                            round-off (series start + Fritsch-Shafer-Crowley/Newton iterations)   */
    DWDF_MODE_APPROX_GOOD = 2 /* wdft::DiodePairT<Good>, wdf_t.h:907-913 (forward only)               */
};

/* Where the voltage probe sits inside one sample. */
enum dwdf_ordering
{
    DWDF_ORDER_PLUGIN = 0, /* between root.incident and tree.incident (DiodeClipperWDF.cpp:26-28)    */
    DWDF_ORDER_PYTHON = 1 /* after tree.incident (lpf.py:42-45, clipper_pot.py:121-123)             */
};

/* One node of the element tree, listed in post-order (children before parents, top adaptor last).
 * `param` is the index into the parameter vector of this leaf's value (R or C), or -1 for
 * adaptors. */
typedef struct dwdf_node
{
    int32_t kind; /* dwdf_node_kind */
    int32_t child1; /* node index or -1 */
    int32_t child2; /* node index or -1 */
    int32_t param; /* parameter slot or -1 */
} dwdf_node;

typedef struct dwdf_circuit_desc
{
    int32_t root_kind; /* dwdf_root_kind */
    int32_t root_mode; /* dwdf_root_mode (diode pair only) */
    int32_t ordering; /* dwdf_ordering */
    int32_t probe; /* node whose voltage (a+b)/2 is the output (tf_wdf.py:8-10) */
    int32_t source; /* leaf driven by x[n] (a DWDF_RESISTIVE_VS or DWDF_RESISTIVE_CS); -1 when x[n] drives the root (DWDF_ROOT_IDEAL_VS / _CS) */
    int32_t r_node; /* leaf whose resistance is the per-sample channel `r` (clipper_pot.py:116), or -1 */
    int32_t param_Is; /* parameter slot of the diode saturation current, diode pair only */
    int32_t param_nabla; /* parameter slot of the ideality factor / nDiodes (wdf_t.h:875-882) */
    int32_t n_params; /* length of the parameter vector */
    int32_t newton_max_iter; /* exact mode: Fritsch-Shafer-Crowley refinement iterations of omega (toms917.cpp:345-364); <=0: 1, which already reaches fp32 round-off */
    float fs; /* sample rate [Hz] */
    float Vt; /* thermal voltage, 25.85e-3 in the reference */
    float n_up; /* diodes in series, "up" branch (diode_config.py:5-9); 1 = symmetric */
    float n_down;
    float newton_tol; /* exact mode: stop refining when |residual| <= tol (0: run all iterations) */
    int32_t probe_current; /* 0: the output is the probe's voltage (a + b) / 2; 1: its current (a - b) / (2 R) (wdf_t.h:1119-1123) */
} dwdf_circuit_desc;

typedef struct dwdf_program dwdf_program; /* opaque, immutable after creation, thread-shareable */

/* The reference's network shapes (DiodePairNeuralModel.h:5-41, diode_pretraining.py:114-126): dense 2 -> H
 * (tanh), n_hidden x dense H -> H (tanh), dense H -> 1. "2x16" is n_hidden = 2, hidden = 16.
 * Weight vector (device floats): per layer the kernel, (in x out) row-major as in the RTNeural JSON files
 * (model_utils.py:17-85), then the bias; layers in order. */
typedef struct dwdf_mlp_desc
{
    int32_t n_hidden; /* 1 .. 6 */
    int32_t hidden; /* H: 4, 8 or 16 */
} dwdf_mlp_desc;

/* ---- loss / gradient ------------------------------------------------------------------------ */
enum dwdf_grad_mode
{
    DWDF_GRAD_UPSTREAM = 0, /* gy_or_target holds dL/dy (what tape.gradient feeds back)               */
    DWDF_GRAD_TARGET = 1 /* gy_or_target holds the target; loss fused (clipper_pot.py:141-177)     */
};
enum dwdf_loss_kind
{
    DWDF_LOSS_MSE = 0, /* tf.keras.losses.MeanSquaredError, clipper_pot.py:176                   */
    DWDF_LOSS_MSE_ESR = 1, /* MSE + error-to-signal ratio as esr_loss's signature reads (clipper_pot.py:148-156,177):
                            * sqrt(sum (y - target)^2 / (sum target^2 + eps) / N); fused in the adjoint kernels          */
    DWDF_LOSS_MSE_ESR_AS_CALLED = 2 /* the same loss AS THE TRAINING LOOP CALLS IT: loss_func(outs, train_Y)
                            * (clipper_pot.py:248) puts the model output into esr_loss's `target_y` slot, so the energy
                            * is the PREDICTION's, sum y^2, and contributes its own gradient -esr y / E. Needs the batch
                            * sums before the sweep: one reduction over (y, target) + one pass writing dL/dy + the reverse
                            * sweep in upstream mode (DWDF_GRAD_TARGET only; not in dwdf_train_pass / dwdf_grad_host)     */
};

/* Result block written by dwdf_backward / dwdf_train_step: doubles, device (or host in *_host).
 *   [0 .. n_params)   dL/dparam, same slot order as the parameter vector
 *   [DWDF_OUT_LOSS]   loss     [DWDF_OUT_MSE] mse     [DWDF_OUT_ESR] esr   (target mode; else 0)
 * n_params <= DWDF_MAX_PARAMS. */
#define DWDF_MAX_PARAMS 16
#define DWDF_MAX_NODES 16
#define DWDF_OUT_LOSS 16
#define DWDF_OUT_MSE 17
#define DWDF_OUT_ESR 18
#define DWDF_OUT_LEN 24

/* ---- API ------------------------------------------------------------------------------------ */

/* Replaces: building the element object graph (lpf.py:23-28, clipper_pot.py:97-101,
 * DiodeClipperWDF.h:18-25). Validates the tree, recognises topologies that have a specialised
 * kernel (Parallel(ResistiveVs, Capacitor) + DiodePair = the diode clipper) and stores the flat
 * post-order program every kernel consumes. */
DWDF_API int dwdf_program_create (const dwdf_node* nodes, int32_t n_nodes, const dwdf_circuit_desc* desc, dwdf_program** out);
DWDF_API int dwdf_program_destroy (dwdf_program* prog);
/* 1 if the program runs on the specialised diode-clipper kernels, 0 if on the tree interpreter. */
DWDF_API int dwdf_program_is_clipper (const dwdf_program* prog);

/* Run-time specialisation of a tree program (any tree that is not the diode clipper). The interpreter sweeps a node list
 * per sample; dwdf_program_specialize instead generates straight-line CUDA source for THIS circuit (one scalar per wave,
 * the reverse-mode step derived node by node), compiles it with NVRTC for sm_100a and loads it: the same entry points
 * then run TMA-tiled kernels with every state in registers, and reverse mode keeps no tape (16-sample segments replayed
 * from state checkpoints that dwdf_forward writes to z_ckpt; dwdf_ckpt_bytes / dwdf_workspace_bytes change accordingly, so
 * query them after specialising, and run dwdf_forward again before dwdf_backward). Covers every element, root and probe of the
 * interpreter (reverse mode where the interpreter has it) except a per-sample resistance channel: DWDF_ERR_UNSUPPORTED then (or
 * without libnvrtc) and the program stays on the interpreter.
 * Compiles for about a second (serialised per program; launches on other threads keep using the interpreter until the
 * specialised kernels are published); not inside a stream capture.
 * dwdf_program_specialized_source: the text handed to NVRTC (part 0: generated circuit code + kernel skeleton; 1, 2: the
 * two embedded headers it includes, dwdf_math.cuh and dwdf_tma.cuh); returns the size needed including the terminator
 * (0: not a specialisable tree) and copies at most `capacity` bytes. Needs no device: the CPU test-suite compiles it. */
DWDF_API int dwdf_program_specialize (dwdf_program* prog);
DWDF_API int dwdf_program_is_specialized (const dwdf_program* prog);
DWDF_API size_t dwdf_program_specialized_source (const dwdf_program* prog, int32_t part, char* buf, size_t capacity);

/* Floats of streaming state per sequence for dwdf_process_block: 1 for the diode-clipper program
 * (its capacitor), number of capacitors + 1 (the probe's last incident wave) for other trees. */
DWDF_API int dwdf_program_n_states (const dwdf_program* prog);

/* Bytes of device scratch dwdf_forward(z_ckpt)/dwdf_backward need for a (B, T) batch. */
DWDF_API size_t dwdf_ckpt_bytes (const dwdf_program* prog, int64_t B, int64_t T);
DWDF_API size_t dwdf_workspace_bytes (const dwdf_program* prog, int64_t B, int64_t T);

/* Replaces: Model.forward / ClipperModel.forward (lpf.py:30-49, clipper_pot.py:103-127) and
 * processDiodeClipper (DiodeClipperWDF.cpp:18-30): T samples of
 *     root.incident(tree.reflected()); tree.incident(root.reflected()); y[n] = voltage(probe)
 * for B independent sequences, each from reset state. `params` (device, n_params floats) holds the
 * trainable values; `r` is NULL or the per-sample resistance channel; `z_ckpt` is NULL or receives
 * the capacitor-state checkpoints the adjoint replays from (dwdf_ckpt_bytes). */
DWDF_API int dwdf_forward (const dwdf_program* prog, const float* params, const float* x, const float* r, float* y, float* z_ckpt, int64_t B, int64_t T, void* stream);

/* Replaces: tape.gradient(loss, trainable_variables) (lpf.py:87-90, clipper_pot.py:246-269).
 * Hand-written adjoint, no tape, no autodiff framework: one reverse sweep over (x, y, g) that
 * re-derives every step of the recurrence from its two end states — recovered from the forward
 * output `y` (the (B, T) buffer dwdf_forward wrote, unchanged) and re-anchored at the checkpoints
 * `z_ckpt` — and carries the adjoint of the capacitor state backwards. `skip` leading samples are
 * excluded from the fused loss (clipper_pot.py:232,248). `gx` is NULL or receives dL/dx. `out`
 * receives the DWDF_OUT_LEN doubles described above, already reduced over the batch in a fixed
 * order (bit-reproducible). Programs on the tree interpreter ignore `y` and `z_ckpt` (may be NULL).
 * Exact (TOMS-917) root with a symmetric pair: the sweep takes each step's linearisation from `y` alone and does
 * not read `x` (DESIGN.md 4a'); the argument is still validated like the others. */
DWDF_API int dwdf_backward (const dwdf_program* prog, const float* params, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode, int32_t loss_kind, int64_t skip, float* gx, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream);

/* Fused training pass (forward + loss + parameter gradients in ONE sweep, by propagating the
 * parameter sensitivities with the recurrence): same `out` as dwdf_forward followed by
 * dwdf_backward in DWDF_GRAD_TARGET mode; y may be NULL. Diode clipper only. */
DWDF_API int dwdf_train_pass (const dwdf_program* prog, const float* params, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream);

/* Multi-GPU split of the two calls above. The *_raw variants stop after the fixed-order batch
 * reduction and write the RAW sums (before the chain rule and the loss normalisation; raw[23] = number
 * of samples in the loss) so that ranks can all-reduce them (one ncclAllReduce of DWDF_OUT_LEN doubles);
 * dwdf_finalize then turns the summed block, in place, into the `out` block described above. The
 * result is independent of how the batch was sharded. */
DWDF_API int dwdf_backward_raw (const dwdf_program* prog, const float* params, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode, int64_t skip, float* gx, double* raw, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream);
DWDF_API int dwdf_train_pass_raw (const dwdf_program* prog, const float* params, const float* x, const float* r, const float* target, int64_t skip, float* y, double* raw, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream);
DWDF_API int dwdf_finalize (const dwdf_program* prog, const float* params, int32_t grad_mode, int32_t loss_kind, double* raw_inout, void* stream);

/* Replaces: tf.keras.optimizers.Adam.apply_gradients + the Keras clip constraints
 * (clipper_pot.py:180,269; tf_wdf.py:74,104). One tiny kernel; state m, v (n_params floats each)
 * and the step counter live on the device so a whole training step is capturable in a CUDA
 * graph. `grad_scale` multiplies the gradients first (1/world_size after an all-reduce).
 * lr_per_slot: NULL, or n_params device floats overriding `lr` per slot (0 freezes a slot; lpf.py:79-80
 * trains R and C with different rates). lo/hi: per-slot clip bounds (device, n_params floats each) or NULL. */
DWDF_API int dwdf_adam_step (float* params, const double* out, float* m, float* v, int32_t* step, int32_t n_params, float lr, const float* lr_per_slot, float beta1, float beta2, float eps, double grad_scale, const float* lo, const float* hi, void* stream);

/* One whole training step of clipper_pot.py:246-269 in ONE call: dwdf_forward + dwdf_backward (fused loss
 * against `target`) + dwdf_adam_step. Everything is enqueued on `stream` and nothing synchronises, so a
 * step — or a loop of steps — can be captured in a CUDA graph and replayed (small batches are launch-bound).
 * y, z_ckpt, out, workspace as in dwdf_forward / dwdf_backward; m, v, step, lr_per_slot, lo, hi as in dwdf_adam_step. */
DWDF_API int dwdf_train_step (const dwdf_program* prog, float* params, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, float* z_ckpt, double* out, void* workspace, size_t workspace_bytes, float* m, float* v, int32_t* step, float lr, const float* lr_per_slot, float beta1, float beta2, float eps, const float* lo, const float* hi, int64_t B, int64_t T, void* stream);

/* ---- multi-GPU: one process per GPU --------------------------------------------------------------------
 * Replaces: nothing in the reference (it trains on one CPU); realises BASELINE config 5 / SURVEY.md §8e: sequences shard
 * over GPUs with no data-path collective, and a training step exchanges ONE small vector (the DWDF_OUT_LEN raw sums; the
 * weight-gradient vector for the neural root). That message is pure latency, so the exchange runs INSIDE the step's
 * reduction kernel over NVLink peer memory instead of as a library collective between two tiny kernels: every rank owns a
 * mailbox (device memory, mapped by its peers with CUDA IPC); a step writes its vector into its slot of every mailbox as
 * self-validating 8-byte words (epoch tag + half a double: no fence, no separate flag), polls the slots of its own until
 * every word carries the epoch, and sums them in rank order (bit-identical results on every rank, no broadcast).
 *   dwdf_comm_create      allocates this rank's mailbox on the CURRENT device
 *   dwdf_comm_get_handle  the mailbox's CUDA IPC handle (dwdf_comm_handle_bytes() bytes), to be gathered by the launcher
 *                         (torch.distributed / MPI / a file — any out-of-band channel)
 *   dwdf_comm_connect     maps the peers' mailboxes: `handles` = world handles in rank order; every rank must have
 *                         connected (barrier on the launcher's side) before the first exchange
 *   dwdf_allreduce_sum    inout[0 .. n) <- sum over ranks, n <= 2047 doubles, one single-block kernel on `stream`
 *   dwdf_train_step_dp    dwdf_train_step on this rank's shard of B sequences: forward + adjoint + ONE kernel doing the
 *                         fixed-order reduction, the exchange, the chain rule / loss over the GLOBAL batch and Adam.
 *                         Nothing synchronises with the host: capturable in a CUDA graph. m == NULL skips Adam.
 * A peer that does not arrive within the timeout (default 20 s) makes the result block NaN instead of hanging the GPU. */
typedef struct dwdf_comm dwdf_comm;
DWDF_API int dwdf_comm_create (int32_t rank, int32_t world, dwdf_comm** out);
DWDF_API size_t dwdf_comm_handle_bytes (void);
DWDF_API int dwdf_comm_get_handle (const dwdf_comm* comm, void* handle_out);
DWDF_API int dwdf_comm_connect (dwdf_comm* comm, const void* handles);
DWDF_API int dwdf_comm_set_timeout (dwdf_comm* comm, double seconds);
DWDF_API int dwdf_comm_destroy (dwdf_comm* comm);
DWDF_API int dwdf_allreduce_sum (const dwdf_comm* comm, double* inout, int64_t n, void* stream);
DWDF_API int dwdf_train_step_dp (const dwdf_program* prog, const dwdf_comm* comm, float* params, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, float* z_ckpt, double* out, void* workspace, size_t workspace_bytes, float* m, float* v, int32_t* step, float lr, const float* lr_per_slot, float beta1, float beta2, float eps, const float* lo, const float* hi, int64_t B, int64_t T, void* stream);

/* Per-phase device timing of dwdf_train_step / dwdf_train_step_dp: between dwdf_profile_begin and dwdf_profile_end the
 * library records CUDA events on the caller's stream at the step's kernel boundaries (at most max_steps steps) and
 * dwdf_profile_end returns the mean milliseconds of the forward pass, the adjoint pass and the tail (reduction, exchange,
 * chain rule, Adam) over the `steps` recorded — the roofline figures of bench.py, taken inside its timed region. */
DWDF_API int dwdf_profile_begin (int32_t max_steps);
DWDF_API int dwdf_profile_end (double* ms_forward, double* ms_adjoint, double* ms_tail, int32_t* steps);

/* End-to-end variants with HOST buffers (what a training script holding numpy batches calls):
 * host->device copies of x / r / target, the kernels, device->host copy of y / out, all on one
 * internal stream, synchronous on return. */
DWDF_API int dwdf_forward_host (const dwdf_program* prog, const float* params_host, const float* x_host, const float* r_host, float* y_host, int64_t B, int64_t T);
DWDF_API int dwdf_grad_host (const dwdf_program* prog, const float* params_host, const float* x_host, const float* r_host, const float* gy_or_target_host, int32_t grad_mode, int32_t loss_kind, int64_t skip, float* y_host, double* out_host, int64_t B, int64_t T);

/* Neural diode-pair root. dwdf_program_create_neural: like dwdf_program_create with desc->root_kind =
 * DWDF_ROOT_NEURAL; the tree must be the clipper's Parallel(ResistiveVoltageSource, Capacitor) with the
 * probe on the capacitor; desc->r_node may name the source (per-sample resistance, clipper_pot.py:114-117).
 * dwdf_forward_neural replaces ClipperModel.forward (clipper_pot.py:103-127) and DiodeClipperWDF::process
 * for plugin models 2-11 (DiodeClipperWDF.cpp:32-166): `weights` = dwdf_mlp_weight_count() device floats,
 * `state` NULL (reset state) or B capacitor states (streaming), `z_ckpt` NULL or dwdf_neural_ckpt_bytes()
 * of device scratch for dwdf_backward_neural.
 * dwdf_backward_neural replaces tape.gradient(loss, model.trainable_variables) (clipper_pot.py:246-269;
 * the trainable variables are the network's kernels and biases): one reverse sweep over (x, y, g), the
 * network differentiated by hand per sample; grad_w receives dL/d(weights) (n_weights doubles, same layout
 * as `weights`), out[DWDF_OUT_LOSS / _MSE / _ESR] the fused loss, gx (NULL or (B, T) floats) dL/dx. Implemented for the reference's shapes
 * 2xH (H = 4, 8, 16) and 4xH (H = 4, 8). dwdf_adam_step_vec: Adam on a vector of any length. */
DWDF_API size_t dwdf_mlp_weight_count (const dwdf_mlp_desc* mlp);
DWDF_API int dwdf_program_create_neural (const dwdf_node* nodes, int32_t n_nodes, const dwdf_circuit_desc* desc, const dwdf_mlp_desc* mlp, dwdf_program** out);
DWDF_API size_t dwdf_neural_ckpt_bytes (const dwdf_program* prog, int64_t B, int64_t T);
DWDF_API size_t dwdf_neural_workspace_bytes (const dwdf_program* prog, int64_t B, int64_t T);
DWDF_API int dwdf_forward_neural (const dwdf_program* prog, const float* params, const float* weights, const float* x, const float* r, float* y, float* state, float* z_ckpt, int64_t B, int64_t T, void* stream);
DWDF_API int dwdf_backward_neural (const dwdf_program* prog, const float* params, const float* weights, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode, int32_t loss_kind, int64_t skip, float* gx, double* grad_w, double* out, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream);
/* Multi-GPU split, as dwdf_backward_raw / dwdf_finalize for the analytic root: the *_raw variant stops after the fixed-order
 * batch reduction (grad_w_raw = unscaled sums over this rank's sequences, raw = the loss sums with raw[23] = samples in the
 * loss); ranks sum both vectors (dwdf_allreduce_sum) and dwdf_finalize_neural applies the loss's scale in place.
 * dwdf_train_step_neural is the whole step of clipper_pot.py:246-269 on the network's weights in one call: forward + reverse
 * sweep + reduction [+ the exchange when `comm` is not NULL: x / target are then this rank's shard] + loss + Adam (m == NULL
 * skips the update). Nothing synchronises with the host. */
DWDF_API int dwdf_backward_neural_raw (const dwdf_program* prog, const float* params, const float* weights, const float* x, const float* r, const float* y, const float* z_ckpt, const float* gy_or_target, int32_t grad_mode, int64_t skip, float* gx, double* grad_w_raw, double* raw, void* workspace, size_t workspace_bytes, int64_t B, int64_t T, void* stream);
DWDF_API int dwdf_finalize_neural (const dwdf_program* prog, int32_t grad_mode, int32_t loss_kind, double* grad_w_inout, double* raw_inout, void* stream);
DWDF_API int dwdf_train_step_neural (const dwdf_program* prog, const dwdf_comm* comm, const float* params, float* weights, const float* x, const float* r, const float* target, int32_t loss_kind, int64_t skip, float* y, float* z_ckpt, double* grad_w, double* out, void* workspace, size_t workspace_bytes, float* m, float* v, int32_t* step, float lr, float beta1, float beta2, float eps, int64_t B, int64_t T, void* stream);
DWDF_API int dwdf_adam_step_vec (float* w, const double* grad, float* m, float* v, int32_t* step, int64_t n, float lr, float beta1, float beta2, float eps, double grad_scale, void* stream);

/* Streaming twin of DiodeClipperWDF::{prepare, process} (DiodeClipperWDF.cpp:3-30): like
 * dwdf_forward, but the capacitor states are read from and written back to `state`
 * (dwdf_program_n_states() * B floats, device, state-major) so consecutive blocks continue the same signal. */
DWDF_API int dwdf_process_block (const dwdf_program* prog, const float* params, const float* x, const float* r, float* y, float* state, int64_t B, int64_t T, void* stream);

/* Last error message of the calling thread ("" if none). */
DWDF_API const char* dwdf_last_error (void);
/* Compile-time facts for the drop-in check: "sm_100a", version, feature flags. */
DWDF_API const char* dwdf_build_info (void);
/* Number of kernels this library has launched so far in this process (bench.py's gpu_launches). */
DWDF_API int64_t dwdf_launch_count (void);
/* Selects the data-movement path of the clipper kernels: 1 = TMA tiles (default when usable),
 * 0 = direct global loads. Returns the previous value. For tests and A/B timing. */
DWDF_API int dwdf_set_tma (int enable);
/* Kernel-variant switches for A/B timing (bit values). 1: forward approx root evaluated sample by sample
 * (no packed fast step); 2: one sequence per lane instead of the packed-fp32x2 pair kernels; 4: TMA L2 prefetch
 * run-ahead; 8: never use time chunks (small batches run one lane per sequence over the whole T); 16: use time
 * chunks whatever the batch size; 64 / 128: time-chunk warm-up until the off-state decay is 1e-13 / 1e-8 instead
 * of 1e-10 (same bits out: the verification pass repairs what the speculation missed); 256: no programmatic dependent launches.
 * 0 = shipped behaviour. Returns the previous bits. */
DWDF_API int dwdf_set_option (int bits);
/* Diagnostics: chunks the time-parallel forward kernels (small batches) had to recompute so far because
 * the speculated start state did not match bit for bit (synchronises the device). */
DWDF_API int64_t dwdf_time_parallel_redone (void);

#ifdef __cplusplus
}
#endif
#endif /* DWDF_H */
