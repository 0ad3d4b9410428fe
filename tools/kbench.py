#!/usr/bin/env python
"""Kernel micro-benchmark (development tool): CUDA-event timings of the clipper kernels at a given shape.
python tools/kbench.py [--B 65536] [--T 4096] [--mode approx] [--opts 0,1] [--iters 10]"""
import argparse, importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_inputs, FS

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--T", type=int, default=4096)
ap.add_argument("--mode", default="approx")
ap.add_argument("--opts", default="0")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--amp", type=float, default=1.0, help="input amplitude scale")
ap.add_argument("--n-up", type=float, default=1)
ap.add_argument("--n-down", type=float, default=1)
ap.add_argument("--r", action="store_true", help="per-sample resistance channel (clipper_pot.py's (B, T, 2) input)")
a = ap.parse_args()
dwdf = importlib.import_module("differentiable-wdfs_b200")
dev = torch.device("cuda", 0)
import bench
bench.T = a.T
Vs = dwdf.ResistiveVoltageSource(47000.0, True); Cc = dwdf.Capacitor(2.2e-9, FS, True); P1 = dwdf.Parallel(Vs, Cc)
dp = dwdf.DiodePair(P1, 4.352e-9, 25.85e-3, 1.906, a.n_up, a.n_down, trainable=True, mode=a.mode)
circ = dwdf.compile_circuit(dp, probe=Cc, ordering="python", device=dev, r_element=Vs if a.r else None)
x = synth_inputs(torch, a.B, 1, dev) * a.amp
y = torch.empty_like(x)
r = None
if a.r:
    r = torch.full_like(x, 47000.0)
    r[::2] = 10000.0
target = (0.9 * circ.forward(x, r=r, keep_for_backward=False)).clone()
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for e0, e1 in ev:
        e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return ts[len(ts) // 2]
n = a.B * a.T
for o in [int(v) for v in a.opts.split(",")]:
    dwdf.set_option(o)
    f = timed(lambda: circ.forward(x, r=r, out=y))
    circ.forward(x, r=r, out=y)
    b = timed(lambda: circ.backward(target=target, loss="mse", raw=True))
    t = timed(lambda: circ.train_pass(x, target, raw=True)) if not a.r else float("nan")
    print(f"[redone chunks so far: {dwdf.time_parallel_redone()}] ", end="")
    print(f"B={a.B} T={a.T} {a.mode} amp={a.amp} opts={o}: forward {f:.4f} ms ({n*8/f/1e6:.0f} GB/s)  adjoint {b:.4f} ms ({n*8/b/1e6:.0f} GB/s alg)  train_pass {t:.4f} ms  fwd+adj {n/(f+b)/1e6:.1f} Gsamples/s", flush=True)
