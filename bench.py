#!/usr/bin/env python
"""bench.py — diode-clipper audio samples/sec, forward + backward, on N B200s (one rank per GPU).

Workload (BASELINE.json configs[4] = the configuration the metric is quoted on; it fits one GPU):
1N4148 diode clipper (R 47 kΩ, C 2.2 nF, fs 48 kHz; DiodeClipperWDF.h:18-25), a GLOBAL batch of
B = 65536 sequences × T = 4096 samples, generated once (seed 1237) and sharded by sequence over the N ranks
(strong scaling, SURVEY.md §8d config 5: rank r owns rows [r B/N, (r+1) B/N)), root = `approx`
(wdft::DiodePairT / omega4, the plugin's chowdsp_wdf path), probe ordering of the training script.
One step = the full training step of clipper_pot.py:246-269 on an analytic root, ONE library call
(dwdf_train_step_dp): forward kernel (x -> y + state checkpoints) + adjoint kernel (x, y, target -> raw
gradient sums) + one tail kernel (fixed-order reduction, exchange of the 24 raw sums with the other ranks over
NVLink peer memory, chain rule to Is/nF/R/C, MSE loss, Adam update).

Prints ONE JSON line (contract in the task statement): value = samples/s over all ranks with the
buffers resident in HBM; e2e = the same step with this rank's x and target copied from pinned host memory
every step and the loss read back; roofline for the dominant kernel from CUDA events taken at the
kernel boundaries inside the timed region (dwdf_profile_begin / _end); cpu_baseline timed on this box's
host cores on bounded samples.

`--impl reference` times the CPU implementation on the host cores instead (see reference_arm()).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000.0
T = 4096
B_GLOBAL = 65536
METRIC = "diode-clipper samples/sec fwd+bwd"
UNIT = "samples/s"
# algorithmic bytes per sample: forward reads x, writes y; adjoint re-reads x, reads target
BYTES_FWD, BYTES_ADJ = 8, 8
# measured DRAM traffic per sample (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture at
# B = 65536, T = 4096; the files named in TRAFFIC_SOURCE). Approx root: the adjoint also reads the forward output y
# (4 B/sample) — that read replaces the replay of the forward recurrence. Exact root: the adjoint reads y and
# the target only (clip_step_recover_yv: states and linearisation from the output alone), traffic = algorithmic.
TRAFFIC = {"approx": ((1.073804e9 + 1.093476e9) / (65536 * 4096), (3.289196e9 + 0.003737e9) / (65536 * 4096)),
           "exact": ((1.073811e9 + 1.099483e9) / (65536 * 4096), (2.214953e9 + 0.003812e9) / (65536 * 4096))}
TRAFFIC_SOURCE = ("ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum per launch at 65536 x 4096): approx root profiles/r02_c_ncu_forward_summary.txt, "
                  "profiles/r02_d_ncu_adjoint_summary.txt; exact root profiles/r01_e_ncu_exact_packed_summary.txt (forward), profiles/r02_s_ncu_exact_adjoint_summary.txt")
# fp32 side of the roofline (op-counted algorithmic flops per sample; the roof is the MEASURED FFMA issue rate,
# tools/micro/ffma2_bench.cu on this pool's B200: 120.8 fma lanes/clk/SM x 2 x 148 SMs x 1.965 GHz)
# (forward, adjoint). Exact root, end of round 2: the forward's x <= -2 region lost its Newton step (190 -> 155) and the
# adjoint no longer evaluates an omega at all (260 -> 95), counted op by op from the shipped step functions like the others
FLOPS = {"approx": (70, 130), "exact": (155, 95)}
FP32_PEAK_TFLOPS = 70.3
FP32_PEAK_SOURCE = "measured FFMA / FFMA2 issue rate, profiles/r02_d_ffma2_microbench.txt (nominal 128 lanes/clk/SM: 74.4)"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def synth_inputs(torch, B, seed, device):
    """SURVEY.md §8(d) input law: per-sequence sine burst + noise, generated on the device."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    A = torch.empty(B, 1, device=device).uniform_(0.1, 2.0, generator=g)
    f = torch.exp(torch.empty(B, 1, device=device).uniform_(float(np.log(50.0)), float(np.log(5000.0)), generator=g))
    n = torch.arange(T, device=device, dtype=torch.float32)[None, :]
    x = A * torch.sin((2 * np.pi / FS) * f * n)
    x.add_(torch.randn(B, T, device=device, generator=g), alpha=0.05)
    return x.contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = str(gpu_index)
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9 or f[0] != self.idx:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # the sampler also sees idle gaps; report the loaded half
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU legs (the only places outside tests/ and smoke() that execute oracle/) -------------------------------------
def cpu_inputs(rows, seed=0):
    rng = np.random.default_rng(seed)
    n = np.arange(T)
    x = (rng.uniform(0.1, 2.0, (rows, 1)) * np.sin(2 * np.pi * np.exp(rng.uniform(np.log(50), np.log(5000), (rows, 1))) * n / FS)).astype(np.float32)
    return x, (np.roll(x, 1, 0) * 0.3).astype(np.float32)


def cpu_baseline(sample_rows, threads, budget_s=10.0, passes=None, warmup=1):
    """The CPU implementation of the same step (forward + reverse sweep + loss) timed on the host.
    kind "port": oracle/wdf_oracle.c (the reference's arithmetic restated in C; the reference's C++
    half has no backward pass and its TensorFlow half cannot be installed here). Repeats passes over
    one bounded sample until ~budget_s of CPU work is done; returns (median samples/s, seconds, passes)."""
    from oracle.cpu import ClipperParams, Oracle

    x, target = cpu_inputs(sample_rows)
    orc = Oracle()
    p = ClipperParams()
    for _ in range(warmup):
        orc.clipper_grad(x, target, p, exact=False, mode="target", dtype=np.float32, threads=threads)
    times = []
    t_all = time.perf_counter()
    while (len(times) < passes) if passes is not None else (time.perf_counter() - t_all < budget_s and len(times) < 200):
        t0 = time.perf_counter()
        orc.clipper_grad(x, target, p, exact=False, mode="target", dtype=np.float32, threads=threads)
        times.append(time.perf_counter() - t0)
    return x.size / float(np.median(times)), float(np.sum(times)), len(times)


def reference_forward(sample_rows, threads, root=0, t_len=T, block=None, fs=FS):
    """The UNMODIFIED reference C++ (chowdsp_wdf tree + DiodePairT / Toms917DiodePairT, compiled in place into
    oracle/_ref), forward only: root 0 = omega4 (plugin model 1), root 1 = TOMS-917 (plugin model 0)."""
    from oracle.cpu import ClipperParams, Ref

    try:
        ref = Ref(fast=True)
    except Exception:
        return None
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (sample_rows, t_len)).astype(np.float32)
    y = np.empty_like(x)
    p = ClipperParams(fs=fs)
    ref.clipper(x[: min(64, sample_rows)], p, root=root, threads=threads)
    t0 = time.perf_counter()
    ref.clipper(x, p, root=root, ordering=1, threads=threads, out=y)
    return x.size / (time.perf_counter() - t0)


def torch_eager_stand_in(B=1024, t_len=2048, threads=None):
    """the wdf_py path. TensorFlow 2.5 cannot be installed here (Python 3.12, no network), so this is a
    STAND-IN, labelled as such: the PyTorch-eager restatement of tf_wdf.py driven by the script's own per-sample Python
    loop (clipper_pot.py:103-127 shape, B = 1024 windows x T = 2048) with the analytic DiodePair root, forward +
    autograd backward, all host cores."""
    import torch

    from oracle import torch_wdf

    if threads:
        torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (B, t_len)).astype(np.float32)
    t0 = time.perf_counter()
    y, leaves = torch_wdf.clipper_forward(x, mode="approx", dtype=torch.float32)
    loss = torch.mean((y - 0.3 * torch.as_tensor(np.roll(x, 1, 0))) ** 2)
    loss.backward()
    return x.size / (time.perf_counter() - t0)


def cpu_legs(threads, port_rows=8192):
    """Every CPU figure the north-star names, on this box's host cores (count stated): the C port fwd+bwd (all cores and
    one), the unmodified chowdsp_wdf forward with both roots (all cores and one), the reference bench's own shape, and the
    torch-eager stand-in for the wdf_py TensorFlow path."""
    out = {}
    v, dt, passes = cpu_baseline(port_rows, threads)
    out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "seconds": dt,
                           "sample": f"{port_rows} sequences x {T} samples x {passes} passes (median), forward + reverse sweep + MSE (oracle/wdf_oracle.c, fp32), {threads} threads"}
    v1, dt1, p1 = cpu_baseline(256, 1, budget_s=4.0)
    out["cpu_baseline_1core"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"256 sequences x {T} samples x {p1} passes, fwd + reverse sweep + MSE, 1 thread"}
    legs = {
        "cpu_reference_forward": (min(port_rows, 8192), threads, 0, "unmodified chowdsp_wdf DiodePairT (omega4, plugin model 1) forward only (oracle/_ref)"),
        "cpu_reference_forward_1core": (256, 1, 0, "same, 1 thread"),
        "cpu_reference_forward_toms917": (1024, threads, 1, "unmodified Toms917DiodePairT (exact Wright omega, plugin model 0) forward only"),
        "cpu_reference_forward_toms917_1core": (64, 1, 1, "same, 1 thread"),
    }
    for key, (rows, th, root, what) in legs.items():
        out[key] = {"value": reference_forward(rows, th, root), "unit": UNIT, "cores": th, "kind": "reference", "what": what, "sample": f"{rows} sequences x {T} samples"}
    # plugin/bench/diode_clipper_bench.cpp:7-31: one mono stream, fs 96 kHz, 0.1 s of audio (9600 samples) in 2048-sample blocks
    out["cpu_reference_bench_shape"] = {"value": reference_forward(1, 1, 0, t_len=5 * 2048, fs=96000.0), "value_toms917": reference_forward(1, 1, 1, t_len=5 * 2048, fs=96000.0), "unit": UNIT, "cores": 1, "kind": "reference",
                                        "what": "the reference's own benchmark shape (diode_clipper_bench.cpp: 1 stream, 96 kHz, 0.1 s in 2048-sample blocks), models 1 (omega4) and 0 (TOMS-917)"}
    try:
        out["cpu_torch_eager_stand_in"] = {"value": torch_eager_stand_in(1024, 2048, threads), "unit": UNIT, "cores": threads, "kind": "stand-in",
                                           "what": "STAND-IN for the wdf_py TensorFlow 2.5 eager path (not installable here): torch-eager restatement of tf_wdf.py, per-sample Python loop of clipper_pot.py:103-127, "
                                                   "B=1024 x T=2048, analytic root, forward + autograd backward"}
    except Exception as e:
        out["cpu_torch_eager_stand_in"] = {"error": repr(e)}
    return out


def reference_arm(args):
    """--impl reference: the CPU implementation of the path on this box's host cores, all threads.
    fwd+bwd like the metric: the C port of the reference arithmetic (the reference itself has no
    runnable backward: C++ half is forward-only, TensorFlow half is not installable); the unmodified
    reference C++ forward is timed beside it and reported as reference_forward."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = 4096  # bounded sample: 4096 x 4096 samples per step
    steps = min(args.steps, 40)  # a step is ~0.05 s of all-core work; keeps the arm within a minute whatever K is asked for
    value, secs, passes = cpu_baseline(rows, threads, passes=steps, warmup=args.warmup)
    ms = rows * T / value * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": passes, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1N4148 diode clipper fwd+bwd (MSE, grads Is/nF/R/C), approx root, T=4096, fs=48k; CPU sample of 4096 sequences per step", "sample_rows": rows, "T": T,
                   "same_config": "same circuit, root, T, loss and per-sample work as the GPU arm; a step covers a bounded sample of 4096 of its 65536 sequences (sequences are independent: samples/s does not depend on the count)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{rows} sequences x {T} samples per step, forward + reverse sweep + MSE, fp32, {threads} threads"},
        "reference_forward": {"value": reference_forward(rows, threads, 0), "unit": UNIT, "cores": threads, "kind": "reference", "what": "unmodified chowdsp_wdf DiodePairT (omega4) forward only, oracle/_ref/libdwdf_ref_fast.so"},
        "reference_forward_toms917": {"value": reference_forward(512, threads, 1), "unit": UNIT, "cores": threads, "kind": "reference", "what": "unmodified Toms917DiodePairT forward only"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def other_paths(torch, dwdf, device, x, target):
    """samples/s (CUDA events, a few repetitions after 2 warm-ups) of the rows of the hot path that `value` does not
    cover: the exact (TOMS-917) root, BASELINE configs 2-4, the per-sample resistance channel, the fused one-sweep
    step, the neural root (forward and training) on the reference's own 2x8 / 2x16 weights, the tree interpreter."""
    out = {}

    def timed(fn, reps=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    def clipper(mode, **kw):
        Vs = dwdf.ResistiveVoltageSource(47000.0, True)
        Cc = dwdf.Capacitor(2.2e-9, FS, True)
        dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), kw.pop("Is", 4.352e-9), 25.85e-3, kw.pop("nabla", 1.906), trainable=True, mode=mode, **kw)
        return dwdf.compile_circuit(dp, probe=Cc, ordering="python", device=device)

    def fwd_bwd(circ, xs, ts):
        circ.forward(xs)
        circ.backward(target=ts, loss="mse")

    def adam(c):
        return dwdf.Adam(c, lr={s: 1e-4 * float(c.params[s]) for s in range(c.n_params)}, beta_1=0.5)

    try:
        ce = clipper("exact")
        out["exact_root_fwd_bwd"] = {"value": x.numel() / timed(lambda: fwd_bwd(ce, x, target)), "unit": UNIT, "B": x.shape[0], "T": T}
        ca = clipper("approx")
        # the whole training step in ONE sweep (forward + loss + gradients by forward-mode tangents, two sequences per lane) + Adam;
        # not the headline path (north_star asks for the reverse-mode adjoint kernel), reported beside it
        oa = adam(ca)
        yb = torch.empty_like(x)

        def fused_step():
            ca.train_pass(x, target, loss="mse", y=yb)
            oa.apply()
        out["fused_tangent_training_step"] = {"value": x.numel() / timed(fused_step, reps=5), "unit": UNIT, "B": x.shape[0], "T": T, "kernels": "clipper_train_pair_tma + finalize + adam"}
        oe = adam(ce)
        out["fused_tangent_training_step_exact_root"] = {"value": x.numel() / timed(lambda: ce.train_step(x, target, oe, loss="mse", out=yb, engine="tangent"), reps=5), "unit": UNIT, "B": x.shape[0], "T": T,
                                                         "kernels": "train_step(engine='tangent'): clipper_train_pair_tma<exact> + finalize + adam"}
        del yb
        # BASELINE configs 2-3 (few long sequences: time chunks), as stated: config 2 forward only, config 3 forward + backward
        xs2 = x[:256].contiguous()
        for mode, c in (("approx", ca), ("exact", ce)):
            out[f"config2_B256_forward_{mode}"] = {"value": xs2.numel() / timed(lambda: c.forward(xs2, keep_for_backward=False), reps=20), "unit": UNIT, "B": 256, "T": T, "what": "forward only (BASELINE configs[1])"}
        xs3, ts3 = x[:1024].contiguous(), target[:1024].contiguous()
        out["config3_B1024_fwd_bwd"] = {"value": xs3.numel() / timed(lambda: fwd_bwd(ca, xs3, ts3), reps=20), "unit": UNIT, "B": 1024, "T": T, "what": "forward + backward, grads Is/nF/R/C (BASELINE configs[2])"}
        # BASELINE config 4: asymmetric pair (N_up = 1, N_down = 2; synthetic germanium-like constants), exact root with
        # Newton tolerance 1e-9, B = 4096, FULL training step (forward + MSE+ESR loss + backward + Adam) in one library call
        c4 = clipper("exact", Is=1.0e-6, nabla=1.3, N_up=1, N_down=2, newton_max_iter=4, newton_tol=1e-9)
        o4 = adam(c4)
        xs4, ts4 = x[:4096].contiguous(), target[:4096].contiguous()
        y4 = torch.empty_like(xs4)
        out["config4_B4096_asymmetric_exact_full_step"] = {"value": xs4.numel() / timed(lambda: c4.train_step(xs4, ts4, o4, loss="mse+esr", out=y4), reps=5), "unit": UNIT, "B": 4096, "T": T,
                                                           "what": "N_up=1, N_down=2, exact root (Newton tol 1e-9, <= 4 iterations), forward + MSE+ESR + backward + Adam (BASELINE configs[3])"}
        # the reference's own training layout: per-sample resistance channel (clipper_pot.py:67-69,114-117)
        try:
            Vs = dwdf.ResistiveVoltageSource(47000.0)
            Cc = dwdf.Capacitor(2.2e-9, FS, True)
            dpr = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode="approx")
            cr = dwdf.compile_circuit(dpr, probe=Cc, ordering="python", r_element=Vs, device=device)
            r = torch.full_like(x, 47000.0)
            r[::2] = 10000.0

            def fb_r():
                cr.forward(x, r=r)
                cr.backward(target=target, loss="mse")
            out["resistance_channel_fwd_bwd"] = {"value": x.numel() / timed(fb_r, reps=2), "unit": UNIT, "B": x.shape[0], "T": T, "specialised_kernels": bool(cr.is_clipper),
                                                 "what": "input (B, T, 2) = (x, R) as clipper_pot.py trains: calc_impedance every sample"}
            del r
        except Exception as e:
            out["resistance_channel_fwd_bwd"] = {"error": repr(e)}
        nn_path = os.path.join(ROOT, "tests", "golden", "nn_vectors.npz")
        if os.path.exists(nn_path):
            nnv = np.load(nn_path)
            for name, b in (("2x8", 8192), ("2x16", 4096)):
                mj = dwdf.model_io.json_from_weights(nnv[f"{name}_weights"], [int(v) for v in nnv[f"{name}_sizes"]])
                Vs, Cc = dwdf.ResistiveVoltageSource(47000.0), dwdf.Capacitor(2.2e-9, FS)
                cn = dwdf.compile_circuit(dwdf.DenseRootModel(mj), tree=dwdf.Parallel(Vs, Cc), probe=Cc, ordering="python", device=device)
                xs, ts = x[:b].contiguous(), target[:b].contiguous()
                out[f"neural_root_{name}_forward"] = {"value": xs.numel() / timed(lambda: cn.forward(xs, keep_for_backward=False)), "unit": UNIT, "B": b, "T": T}
                out[f"neural_root_{name}_fwd_bwd"] = {"value": xs.numel() / timed(lambda: fwd_bwd(cn, xs, ts), reps=2), "unit": UNIT, "B": b, "T": T}
                if name == "2x16":  # the reference's largest network at the headline batch (one lane per pair of sequences)
                    out["neural_root_2x16_fwd_bwd_B65536"] = {"value": x.numel() / timed(lambda: fwd_bwd(cn, x, target), reps=2), "unit": UNIT, "B": x.shape[0], "T": T}
        # tree programs. lpf.py's circuit (IdealVoltageSource root, Inverter(Series(R, C)), probe C) and the plugin's HPF clipper
        # (HPFDiodeClipper.h:25-37: Parallel(R, Series(Vs, C)) + DiodePair): run-time specialised kernels (dwdf_program_specialize:
        # generated straight-line source, NVRTC, TMA tiles, no tape) and, on a slice, the node-list interpreter beside them
        try:
            def lpf():
                R1, C1 = dwdf.Resistor(1000.0, True), dwdf.Capacitor(1.0e-6, FS, True)
                return dwdf.compile_circuit(dwdf.IdealVoltageSource(), tree=dwdf.Inverter(dwdf.Series(R1, C1)), probe=C1, device=device)

            def hpf(mode):
                Vs, Ch, Rh = dwdf.ResistiveVoltageSource(4700.0, True), dwdf.Capacitor(2.2e-9, FS, True), dwdf.Resistor(47000.0, True)
                dph = dwdf.DiodePair(dwdf.Parallel(Rh, dwdf.Series(Vs, Ch)), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=mode)
                return dwdf.compile_circuit(dph, probe=Ch, ordering="plugin", device=device)

            xi, ti = x[:4096].contiguous(), target[:4096].contiguous()
            for key, make in (("rc_lowpass", lpf), ("hpf_clipper_approx", lambda: hpf("approx")), ("hpf_clipper_exact", lambda: hpf("exact"))):
                ct = make()
                t0 = time.perf_counter()
                spec = ct.specialize(quiet=True)
                t_spec = time.perf_counter() - t0
                entry = {"unit": UNIT, "B": x.shape[0], "T": T, "specialised": bool(spec), "specialise_s": t_spec}
                if spec:
                    entry["forward"] = x.numel() / timed(lambda: ct.forward(x, keep_for_backward=False), reps=2)
                    entry["value"] = x.numel() / timed(lambda: fwd_bwd(ct, x, target), reps=2)
                    entry["what"] = "forward + reverse mode (fused MSE), specialised kernels; 'forward' = inference alone"
                ci = make()
                ci._specialize_error = RuntimeError("interpreter wanted")  # keep this copy on the node-list interpreter
                entry["interpreter_value"] = xi.numel() / timed(lambda: fwd_bwd(ci, xi, ti), reps=1)
                entry["interpreter_B"] = int(xi.shape[0])
                out[f"tree_{key}_fwd_bwd"] = entry
        except Exception as e:
            out["tree_error"] = repr(e)
    except Exception as e:  # the headline line must not depend on these
        out["error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200, help="timed steps (200 x ~1.1 ms: a timed region of ~0.2 s at N = 1)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_GLOBAL, help="sequences: the global batch (strong scaling) or per GPU (--scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="peer: the step's tail kernel exchanges the raw sums over NVLink peer memory (dwdf_train_step_dp); nccl: torch.distributed.all_reduce between two kernels")
    ap.add_argument("--mode", default="approx", choices=["approx", "exact"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the brief timings of the other paths (exact root, small batches, neural root)")
    ap.add_argument("--opts", type=int, default=0, help="development: dwdf_set_option bits for A/B timing (0 = shipped behaviour; recorded in config when set)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)
    # the contract is ONE JSON line on stdout: libraries that write to fd 1 on their own (NCCL's version banner at communicator
    # creation) go to stderr while the benchmark runs; the line itself is printed to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    dwdf = importlib.import_module("differentiable-wdfs_b200")
    if args.opts:
        dwdf.set_option(args.opts)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def make_circuit():
        # circuit: clipper_pot.py:97-101 with the analytic DiodePair root; plugin constants DiodeClipperWDF.h:18-25
        Vs = dwdf.ResistiveVoltageSource(47000.0, True)
        Cc = dwdf.Capacitor(2.2e-9, FS, True)
        dpair = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=args.mode)
        c = dwdf.compile_circuit(dpair, probe=Cc, ordering="python", device=device)
        scale = torch.tensor([1.1 if s == c.slot(Vs, "R") else 0.9 if s == c.slot(Cc, "C") else 2.0 if s == c.slot(dpair, "Is") else 1.05 for s in range(c.n_params)], device=device)
        # Adam(beta_1=0.5) of clipper_pot.py:180; one rate per slot, 1e-4 of the value (R, C, Is, nF span 13 decades)
        return c, dwdf.Adam(c, lr={s: 1e-4 * float(c.params[s]) for s in range(c.n_params)}, beta_1=0.5), scale

    circ, opt, perturb = make_circuit()

    def make_target(c, xs):
        """target = output of a perturbed parameter set, synthesised with the same engine"""
        base = c.params.clone()
        c.params.mul_(perturb)
        t = c.forward(xs, keep_for_backward=False).clone()
        c.params.copy_(base)
        return t

    # ---- inputs: generated globally (rank-independent), then sliced -------------------------------------------------------
    strong = args.scaling == "strong"
    if strong:
        B_total = args.batch
        lo, hi = dwdf.shard_rows(B_total, world, rank)
        x_full = synth_inputs(torch, B_total, 1237, device)
        x = x_full[lo:hi].clone() if world > 1 else x_full
    else:
        B_total = args.batch * world
        x_full = None
        x = synth_inputs(torch, args.batch, 1237 + rank, device)
    B = x.shape[0]
    target = make_target(circ, x)
    y = torch.empty_like(x)
    torch.cuda.synchronize()

    # ---- the exchange ------------------------------------------------------------------------------------------------------------
    comm, exchange = None, "none (1 GPU)"
    if world > 1:
        exchange = args.exchange
        if exchange == "peer":
            try:
                comm = dwdf.PeerComm(device)
            except Exception as e:  # CUDA IPC unavailable on this box: the library collective between two kernels instead
                exchange = f"nccl (peer memory unavailable: {e!r})"
        ok = torch.tensor([1 if comm is not None else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0 and comm is not None:
            comm.close()
            comm, exchange = None, "nccl (peer memory unavailable on some rank)"

    def step(c=circ, o=opt, xs=x, ts=target, ys=y):
        if world == 1 or comm is not None:
            c.train_step(xs, ts, o, loss="mse", out=ys, comm=comm)  # ONE library call: forward + adjoint + [reduce + exchange + chain rule + Adam]
        else:
            c.forward(xs, out=ys)
            c.backward(target=ts, loss="mse", raw=True)
            dist.all_reduce(c.out, op=dist.ReduceOp.SUM)
            c.finalize(target=True, loss="mse")
            o.apply()

    # ---- N ranks == 1 rank: the sharded step's loss and gradients against the unsharded step on rank 0 ------------------------------
    dp_check = None
    if world > 1 and strong:
        dp_check = {}
        for label, opts in (("rel_err_one_chunk", 8), ("rel_err", 0)):  # 8 = kOptNoChunks: same summation order on every rank count -> fp64 round-off only
            prev = dwdf.set_option(opts)
            c2, o2, _ = make_circuit()
            o2.lr.zero_()  # a step that leaves the parameters alone
            step(c2, o2)
            got = c2.out.clone()
            if rank == 0:
                c1, o1, _ = make_circuit()
                o1.lr.zero_()
                t_full = make_target(c1, x_full)
                c1.train_step(x_full, t_full, o1, loss="mse")
                idx = list(range(c1.n_params)) + [dwdf._lib.OUT_LOSS]
                dp_check[label] = float(torch.max(torch.abs(got[idx] / c1.out[idx] - 1.0)))
                del t_full, c1
            dwdf.set_option(prev)
            del c2
        if rank == 0:
            assert dp_check["rel_err_one_chunk"] < 1e-10 and dp_check["rel_err"] < 1e-4, dp_check
    del x_full
    torch.cuda.empty_cache()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = dwdf.launch_count()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()  # every rank enters the timed region together (rank 0 has just started the clock sampler)
        torch.cuda.synchronize()
    profiled = world == 1 or comm is not None
    # per-phase CUDA events (for the roofline block) on the FIRST tenth of the timed steps only: an event record between two kernels
    # costs nothing at 65536 sequences per GPU but ~4 us per boundary on a 200 us step (it also defeats the programmatic dependent
    # launches), i.e. 8 % of an 8-GPU shard's step — measured with tools/graph_step_bench.py
    n_prof_steps = max(1, min(args.steps, max(10, args.steps // 10)))
    if profiled:
        dwdf._lib.profile_begin(n_prof_steps)
    t_beg.record()
    for i in range(args.steps):
        step()
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = dwdf.launch_count() - launches0
    elapsed_ms = t_beg.elapsed_time(t_end)
    fwd_ms, adj_ms, tail_ms, n_prof = dwdf._lib.profile_end() if profiled else (float("nan"), float("nan"), float("nan"), 0)
    loss = float(circ.out[dwdf._lib.OUT_LOSS])
    tmax = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tmax)
    value = B_total * T * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: the same training step with this rank's x and target copied from pinned host memory every step, loss read back ----------
    e2e = None
    if not args.no_e2e:
        xh = x.cpu().pin_memory()
        th = target.cpu().pin_memory()
        n_e2e = max(3, min(args.steps, 10))

        def e2e_step():
            x.copy_(xh, non_blocking=True)
            target.copy_(th, non_blocking=True)
            step()
            return float(circ.out[dwdf._lib.OUT_LOSS])  # device -> host read of the step's result (synchronises)
        for _ in range(2):
            e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_loss = e2e_step()
        dt = time.perf_counter() - t0
        tm = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e = {"value": B_total * T * n_e2e / float(tm), "unit": UNIT, "h2d_bytes_per_step": int(2 * B * T * 4), "d2h_bytes_per_step": 8, "steps": n_e2e,
               "api": "CompiledCircuit.train_step (dwdf_train_step_dp) after copying this rank's x and target from pinned host memory; loss read back every step; bytes are per rank", "loss": e2e_loss}
        del xh, th
    clocks = sampler.stop() if rank == 0 else None

    # ---- weak scaling beside the strong headline: 65536 sequences PER GPU ------------------------------------------------------------
    weak = None
    if world > 1 and strong:
        xw = synth_inputs(torch, args.batch, 1237 + rank, device)
        cw, ow, _ = make_circuit()
        tw = make_target(cw, xw)
        yw = torch.empty_like(xw)
        n_w = max(5, min(args.steps, 20))
        for _ in range(3):
            step(cw, ow, xw, tw, yw)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_w):
            step(cw, ow, xw, tw, yw)
        e1.record()
        torch.cuda.synchronize()
        tw_ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        dist.all_reduce(tw_ms, op=dist.ReduceOp.MAX)
        weak = {"value": world * args.batch * T * n_w / (float(tw_ms) * 1e-3), "unit": UNIT, "ms_per_step": float(tw_ms) / n_w, "batch_per_gpu": args.batch, "steps": n_w, "scaling": "weak"}
        del xw, tw, yw, cw

    # ---- the other rows of the hot path, timed briefly on rank 0 (not part of `value`) ----------------------
    other = None
    if rank == 0 and not args.no_extra:
        if x.shape[0] < 4096:
            x = synth_inputs(torch, 8192, 1237, device)
            target = make_target(circ, x)
        other = other_paths(torch, dwdf, device, x, target)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        pair = args.mode in ("approx", "exact")
        dom = "clipper_adjoint_tma" if adj_ms >= fwd_ms else ("clipper_forward_pair_tma" if pair else "clipper_forward_tma")
        traffic_fwd, traffic_adj = TRAFFIC[args.mode]
        dom_ms, dom_bytes, dom_traffic = (adj_ms, BYTES_ADJ, traffic_adj) if adj_ms >= fwd_ms else (fwd_ms, BYTES_FWD, traffic_fwd)
        achieved = B * T * dom_bytes / (dom_ms * 1e-3) / 1e9 if n_prof else None
        step_bytes = B * T * (BYTES_FWD + BYTES_ADJ)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[4]: 1N4148 diode clipper fwd+bwd (grads Is,nF,R,C; MSE; Adam), {args.mode} root, "
                                   + (f"GLOBAL batch {B_total} seqs x T={T} @48kHz generated once (seed 1237) and sharded by sequence over {world} GPU(s): {B} per GPU" if strong else f"{args.batch} seqs/GPU x T={T} @48kHz (weak scaling)"),
                       "global_batch": B_total, "batch_per_gpu": B, "T": T, "root_mode": args.mode,
                       "l2": f"inputs larger than L2 ({3 * B * T * 4 / 2**20:.0f} MiB working set per GPU against 126 MB, no flush needed)",
                       "parallelism": f"dp{world}: sequences sharded, one exchange of 24 doubles per step ({exchange})"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                         "traffic": dom_traffic * B * T if B == 65536 else None, "traffic_source": TRAFFIC_SOURCE, "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": dom_bytes, "kernel_ms": dom_ms, "timing": f"CUDA events at the kernel boundaries inside the timed region, mean of its first {n_prof} steps (dwdf_profile_begin/_end)",
                         "fp32": {"what": "second roof (SURVEY.md §8d): algorithmic flops per sample / kernel time against the measured fp32 FMA issue rate; pipe utilisations from ncu are in profiles/r02_*_ncu_*_summary.txt",
                                  "algorithmic_flops_per_sample": {"forward": FLOPS[args.mode][0], "adjoint": FLOPS[args.mode][1]}, "peak_TFLOPs": FP32_PEAK_TFLOPS, "peak_source": FP32_PEAK_SOURCE,
                                  "kernel_TFLOPs": (FLOPS[args.mode][1] if dom == "clipper_adjoint_tma" else FLOPS[args.mode][0]) * B * T / (dom_ms * 1e-3) / 1e12 if dom_ms else None,
                                  "kernel_frac": (FLOPS[args.mode][1] if dom == "clipper_adjoint_tma" else FLOPS[args.mode][0]) * B * T / (dom_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS if dom_ms else None,
                                  "step_TFLOPs": sum(FLOPS[args.mode]) * B * T / ((fwd_ms + adj_ms + tail_ms) * 1e-3) / 1e12 if n_prof else None,
                                  "step_frac": sum(FLOPS[args.mode]) * B * T / ((fwd_ms + adj_ms + tail_ms) * 1e-3) / 1e12 / FP32_PEAK_TFLOPS if n_prof else None},
                         "step": {"forward_ms": fwd_ms, "adjoint_ms": adj_ms, "tail_ms": tail_ms, "bytes_per_sample": BYTES_FWD + BYTES_ADJ,
                                  "achieved_GBs": step_bytes / ((fwd_ms + adj_ms + tail_ms) * 1e-3) / 1e9 if n_prof else None,
                                  "frac": step_bytes / ((fwd_ms + adj_ms + tail_ms) * 1e-3) / 1e9 / peak if n_prof else None}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "loss": loss, "dp_check": dp_check, "weak_scaling": weak, "other_paths": other,
        }
        if world == 1 and not args.no_cpu:
            line.update(cpu_legs(os.cpu_count() or 1))
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
