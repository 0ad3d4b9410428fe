#!/usr/bin/env python
"""Development tool: the training step (dwdf_train_step) eagerly vs captured in a CUDA graph and replayed, at a given batch.
python tools/graph_step_bench.py [--B 8192] [--T 4096] [--iters 400]"""
import argparse, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8192)
ap.add_argument("--iters", type=int, default=400)
ap.add_argument("--opts", type=int, default=0)
a = ap.parse_args()
dwdf = importlib.import_module("differentiable-wdfs_b200")
dev = torch.device("cuda", 0)
dwdf.set_option(a.opts)
Vs = dwdf.ResistiveVoltageSource(47000.0, True); Cc = dwdf.Capacitor(2.2e-9, bench.FS, True)
dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode="approx")
circ = dwdf.compile_circuit(dp, probe=Cc, ordering="python", device=dev)
opt = dwdf.Adam(circ, lr={s: 1e-4 * float(circ.params[s]) for s in range(circ.n_params)}, beta_1=0.5)
x = bench.synth_inputs(torch, a.B, 1237, dev)
target = (0.9 * circ.forward(x, keep_for_backward=False)).clone()
y = torch.empty_like(x)
def step():
    circ.train_step(x, target, opt, loss="mse", out=y)
def timed(fn, n):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_eager = timed(step, a.iters)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    step()
t_graph = timed(g.replay, a.iters)
n = a.B * bench.T
print(f"B={a.B} opts={a.opts}: eager {t_eager*1e3:.1f} us/step ({n/t_eager/1e6:.1f} Gsamples/s)   graph replay {t_graph*1e3:.1f} us/step ({n/t_graph/1e6:.1f} Gsamples/s)")
