#!/usr/bin/env python
"""bench.py — diode-clipper audio samples/sec, forward + backward, on N B200s (one rank per GPU).

Workload (BASELINE.json configs[4] = the configuration the metric is quoted on; it fits one GPU):
1N4148 diode clipper (R 47 kΩ, C 2.2 nF, fs 48 kHz; DiodeClipperWDF.h:18-25), B = 65536 sequences ×
T = 4096 samples PER GPU (weak scaling: sequences are independent, each rank owns its shard), root
= `approx` (wdft::DiodePairT / omega4, the plugin's chowdsp_wdf path), probe ordering of the training
script. One step = forward kernel (x -> y + state checkpoints) + adjoint kernel (x, target ->
gradients w.r.t. Is, nF, R, C and the MSE loss) + the single all-reduce of the raw sums (N > 1) +
finalize + Adam update — the full training step of clipper_pot.py:246-269 on an analytic root.

Prints ONE JSON line (contract in the task statement): value = samples/s over all ranks with the
buffers resident in HBM; e2e = the same through the host-buffer entry point (dwdf_grad_host: pinned
host x and target copied host->device every step, gradients/loss read back); roofline for the
dominant kernel from CUDA events taken inside the timed region; cpu_baseline timed on this box's
host cores on a bounded sample.

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
B_PER_GPU = 65536
METRIC = "diode-clipper samples/sec fwd+bwd"
UNIT = "samples/s"
# algorithmic bytes per sample: forward reads x, writes y; adjoint re-reads x, reads target
BYTES_FWD, BYTES_ADJ = 8, 8
# measured DRAM traffic per sample (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture at
# B = 65536, T = 4096: profiles/r01_b_ncu_full_summary.txt). The adjoint also reads the forward output y (4 B/sample):
# that read replaces the replay of the forward recurrence.
TRAFFIC_FWD, TRAFFIC_ADJ = 2.166901e9 / (65536 * 4096), 3.293523e9 / (65536 * 4096)


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


def cpu_baseline(sample_rows, threads, budget_s=12.0, passes=None, warmup=1):
    """The CPU implementation of the same step (forward + reverse sweep + loss) timed on the host.
    kind "port": oracle/wdf_oracle.c (the reference's arithmetic restated in C; the reference's C++
    half has no backward pass and its TensorFlow half cannot be installed here). Repeats passes over
    one bounded sample until ~budget_s of CPU work is done; returns (median samples/s, seconds, passes)."""
    from oracle.cpu import ClipperParams, Oracle

    rng = np.random.default_rng(0)
    n = np.arange(T)
    x = (rng.uniform(0.1, 2.0, (sample_rows, 1)) * np.sin(2 * np.pi * np.exp(rng.uniform(np.log(50), np.log(5000), (sample_rows, 1))) * n / FS)).astype(np.float32)
    target = np.roll(x, 1, 0) * 0.3
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


def reference_forward(sample_rows, threads):
    """The UNMODIFIED reference C++ (chowdsp_wdf DiodePairT, compiled in place into oracle/_ref): forward only."""
    from oracle.cpu import ClipperParams, Ref

    try:
        ref = Ref(fast=True)
    except Exception:
        return None
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (sample_rows, T)).astype(np.float32)
    y = np.empty_like(x)
    ref.clipper(x[:64], ClipperParams(), root=0, threads=threads)
    t0 = time.perf_counter()
    ref.clipper(x, ClipperParams(), root=0, ordering=1, threads=threads, out=y)
    return x.size / (time.perf_counter() - t0)


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
    value, secs, passes = cpu_baseline(rows, threads, passes=args.steps, warmup=args.warmup)  # exactly W untimed + K timed steps
    ms = rows * T / value * 1e3
    vals = [None] * passes
    fwd = reference_forward(rows, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1N4148 diode clipper fwd+bwd (MSE, grads Is/nF/R/C), approx root, T=4096, fs=48k; CPU sample of 4096 sequences per step", "sample_rows": rows, "T": T},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{rows} sequences x {T} samples per step, forward + reverse sweep + MSE, fp32, {threads} threads"},
        "reference_forward": {"value": fwd, "unit": UNIT, "cores": threads, "kind": "reference", "what": "unmodified chowdsp_wdf DiodePairT (omega4) forward only, oracle/_ref/libdwdf_ref_fast.so"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def other_paths(torch, dwdf, device, x, target):
    """samples/s (forward + adjoint kernels, CUDA events, 3 repetitions after 2 warm-ups) of the rows of the hot
    path that `value` does not cover: the exact (TOMS-917) root, BASELINE configs 2-3 (small batches: the
    time-parallel kernels) and the neural root (forward and training) on the reference's own 2x8 / 2x16 weights."""
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

    def clipper(mode):
        Vs = dwdf.ResistiveVoltageSource(47000.0, True)
        Cc = dwdf.Capacitor(2.2e-9, FS, True)
        dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=mode)
        return dwdf.compile_circuit(dp, probe=Cc, ordering="python", device=device)

    def fwd_bwd(circ, xs, ts):
        circ.forward(xs)
        circ.backward(target=ts, loss="mse")

    try:
        ce = clipper("exact")
        out["exact_root_fwd_bwd"] = {"value": x.numel() / timed(lambda: fwd_bwd(ce, x, target)), "unit": UNIT, "B": x.shape[0], "T": T}
        ca = clipper("approx")
        # the whole training step in ONE sweep (forward + loss + gradients by forward-mode tangents, two sequences per lane) + Adam;
        # not the headline path (north_star asks for the reverse-mode adjoint kernel), reported beside it
        oa = dwdf.Adam(ca, lr={s: 1e-4 * float(ca.params[s]) for s in range(ca.n_params)}, beta_1=0.5)
        yb = torch.empty_like(x)

        def fused_step():
            ca.train_pass(x, target, loss="mse", y=yb)
            oa.apply()
        out["fused_tangent_training_step"] = {"value": x.numel() / timed(fused_step, reps=5), "unit": UNIT, "B": x.shape[0], "T": T, "kernels": "clipper_train_pair_tma + finalize + adam"}
        del yb
        for name, b in (("config2_B256_fwd_bwd", 256), ("config3_B1024_fwd_bwd", 1024)):
            xs, ts = x[:b].contiguous(), target[:b].contiguous()
            out[name] = {"value": xs.numel() / timed(lambda: fwd_bwd(ca, xs, ts), reps=20), "unit": UNIT, "B": b, "T": T, "kernels": "time-parallel (256-sample chunks)"}
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
        # the generic tree interpreter on lpf.py's circuit (IdealVoltageSource root, Inverter(Series(R, C)), probe C)
        R1, C1 = dwdf.Resistor(1000.0, True), dwdf.Capacitor(1.0e-6, FS, True)
        ct = dwdf.compile_circuit(dwdf.IdealVoltageSource(), tree=dwdf.Inverter(dwdf.Series(R1, C1)), probe=C1, device=device)
        out["tree_interpreter_rc_lowpass_fwd_bwd"] = {"value": x.numel() / timed(lambda: fwd_bwd(ct, x, target), reps=2), "unit": UNIT, "B": x.shape[0], "T": T}
    except Exception as e:  # the headline line must not depend on these
        out["error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="sequences per GPU")
    ap.add_argument("--mode", default="approx", choices=["approx", "exact"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the brief timings of the other paths (exact root, small batches, neural root)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    dwdf = importlib.import_module("differentiable-wdfs_b200")
    dp_mod = importlib.import_module("differentiable-wdfs_b200.data_parallel")
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
    B = args.batch

    # circuit: clipper_pot.py:97-101 with the analytic DiodePair root; plugin constants DiodeClipperWDF.h:18-25
    Vs = dwdf.ResistiveVoltageSource(47000.0, True)
    Cc = dwdf.Capacitor(2.2e-9, FS, True)
    P1 = dwdf.Parallel(Vs, Cc)
    dpair = dwdf.DiodePair(P1, 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=args.mode)
    circ = dwdf.compile_circuit(dpair, probe=Cc, ordering="python", device=device)
    # Adam(beta_1=0.5) of clipper_pot.py:180; one rate per slot, 1e-4 of the value (R, C, Is, nF span 13 decades)
    opt = dwdf.Adam(circ, lr={s: 1e-4 * float(circ.params[s]) for s in range(circ.n_params)}, beta_1=0.5)

    x = synth_inputs(torch, B, 1237 + rank, device)
    y = torch.empty_like(x)
    # target = output of a perturbed parameter set, synthesised with the same engine
    base = circ.params.clone()
    circ.params.mul_(torch.tensor([1.1 if s == circ.slot(Vs, "R") else 0.9 if s == circ.slot(Cc, "C") else 2.0 if s == circ.slot(dpair, "Is") else 1.05 for s in range(circ.n_params)], device=device))
    target = circ.forward(x, keep_for_backward=False).clone()
    circ.params.copy_(base)
    torch.cuda.synchronize()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]

    def step(i=None):
        if i is not None:
            ev[i][0].record()
        circ.forward(x, out=y)
        if i is not None:
            ev[i][1].record()
        circ.backward(target=target, loss="mse", raw=True)
        if i is not None:
            ev[i][2].record()
        if world > 1:
            dist.all_reduce(circ.out, op=dist.ReduceOp.SUM)
        circ.finalize(target=True, loss="mse")
        opt.apply()

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
    t_beg.record()
    for i in range(args.steps):
        step(i)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = dwdf.launch_count() - launches0
    elapsed_ms = t_beg.elapsed_time(t_end)
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    adj_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    loss = float(circ.out[dwdf._lib.OUT_LOSS])
    tmax = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tmax)
    value = world * B * T * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: the host-buffer entry point, pinned x and target copied in every step ------------------
    e2e = None
    if not args.no_e2e:
        xh = x.cpu().pin_memory()
        th = target.cpu().pin_memory()
        outh = torch.zeros(24, dtype=torch.float64).pin_memory()
        ph = circ.params.cpu()
        n_e2e = max(3, min(args.steps, 10))
        for _ in range(2):
            circ.grad_host(xh, th, outh, params_host=ph, loss="mse")
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            circ.grad_host(xh, th, outh, params_host=ph, loss="mse")  # synchronous: returns with the result on the host
        dt = time.perf_counter() - t0
        tm = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * T * n_e2e / float(tm), "unit": UNIT, "h2d_bytes_per_step": int(2 * B * T * 4 + 4 * circ.n_params), "d2h_bytes_per_step": 24 * 8, "steps": n_e2e,
               "api": "dwdf_grad_host (forward + adjoint + finalize, chunk-pipelined copies)", "loss": float(outh[dwdf._lib.OUT_LOSS])}
    clocks = sampler.stop() if rank == 0 else None

    # ---- the other rows of the hot path, timed briefly on rank 0 (not part of `value`) ----------------------
    other = None
    if rank == 0 and not args.no_extra:
        other = other_paths(torch, dwdf, device, x, target)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # dominant kernel = the adjoint (replay + reverse sweep); its algorithmic traffic is x + target
        dom = "clipper_adjoint_tma" if adj_ms >= fwd_ms else ("clipper_forward_pair_tma" if args.mode == "approx" else "clipper_forward_tma")
        dom_ms, dom_bytes, dom_traffic = (adj_ms, BYTES_ADJ, TRAFFIC_ADJ) if adj_ms >= fwd_ms else (fwd_ms, BYTES_FWD, TRAFFIC_FWD)
        achieved = B * T * dom_bytes / (dom_ms * 1e-3) / 1e9
        step_bytes = B * T * (BYTES_FWD + BYTES_ADJ)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[4]: 1N4148 diode clipper fwd+bwd (grads Is,nF,R,C; MSE; Adam), {args.mode} root, B={B} seqs/GPU x T={T} @48kHz, sharded by sequence", "batch_per_gpu": B, "T": T,
                       "root_mode": args.mode, "l2": "inputs larger than L2 (3 GiB working set per GPU, no flush needed)", "parallelism": f"dp{world} (one all-reduce of 24 doubles per step)"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": dom_traffic * B * T if args.mode == "approx" else None,
                         "traffic_source": "ncu --set full, profiles/r01_b_ncu_full_summary.txt (bytes per sample x samples per launch)", "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": dom_bytes, "kernel_ms": dom_ms,
                         "step": {"forward_ms": fwd_ms, "adjoint_ms": adj_ms, "bytes_per_sample": BYTES_FWD + BYTES_ADJ, "achieved_GBs": step_bytes / ((fwd_ms + adj_ms) * 1e-3) / 1e9,
                                  "frac": step_bytes / ((fwd_ms + adj_ms) * 1e-3) / 1e9 / peak}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "loss": loss, "other_paths": other,
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            rows = 8192
            v, dt, passes = cpu_baseline(rows, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "seconds": dt,
                                    "sample": f"{rows} sequences x {T} samples x {passes} passes (median), forward + reverse sweep + MSE (oracle/wdf_oracle.c, fp32), {threads} threads"}
            line["cpu_reference_forward"] = {"value": reference_forward(min(rows, 8192), threads), "unit": UNIT, "cores": threads, "kind": "reference",
                                             "what": "unmodified chowdsp_wdf DiodePairT forward only (oracle/_ref)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
