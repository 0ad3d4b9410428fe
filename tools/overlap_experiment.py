#!/usr/bin/env python
"""Development experiment: forward(second half) concurrently with adjoint(first half) on two streams, against the plain step.
python tools/overlap_experiment.py [--B 65536] [--parts 2]"""
import argparse, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--parts", type=int, default=2)
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
dwdf = importlib.import_module("differentiable-wdfs_b200")
dev = torch.device("cuda", 0)
def make():
    Vs = dwdf.ResistiveVoltageSource(47000.0, True); Cc = dwdf.Capacitor(2.2e-9, bench.FS, True)
    dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode="approx")
    return dwdf.compile_circuit(dp, probe=Cc, ordering="python", device=dev)
x = bench.synth_inputs(torch, a.B, 1237, dev)
c0 = make()
target = (0.9 * c0.forward(x, keep_for_backward=False)).clone()
y = torch.empty_like(x)
P = a.parts
h = a.B // P
circs = [make() for _ in range(P)]
xs = [x[i * h:(i + 1) * h] for i in range(P)]
ts = [target[i * h:(i + 1) * h] for i in range(P)]
ys = [y[i * h:(i + 1) * h] for i in range(P)]
s_main = torch.cuda.current_stream()
s_adj = torch.cuda.Stream()
def plain():
    c0.forward(x, out=y)
    c0.backward(target=target, loss="mse", raw=True)
def overlapped():
    evs = []
    for i in range(P):
        circs[i].forward(xs[i], out=ys[i])           # main stream: forwards back to back
        e = torch.cuda.Event(); e.record(s_main); evs.append(e)
        s_adj.wait_event(e)
        with torch.cuda.stream(s_adj):                   # adjoint of part i runs beside the forward of part i + 1
            circs[i].backward(target=ts[i], loss="mse", raw=True)
    e = torch.cuda.Event(); e.record(s_adj); s_main.wait_event(e)
    tot = circs[0].out.clone()
    for c in circs[1:]:
        tot += c.out
    return tot
def timed(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
plain(); ref = c0.out.clone()
got = overlapped()
torch.cuda.synchronize()
print("raw sums agree:", float(torch.max(torch.abs(got[:5] / ref[:5] - 1))))
tp, to = timed(plain, a.iters), timed(overlapped, a.iters)
n = a.B * bench.T
print(f"B={a.B} parts={P}: plain forward+adjoint {tp:.3f} ms ({n/tp/1e6:.1f} Gsamples/s)   overlapped {to:.3f} ms ({n/to/1e6:.1f} Gsamples/s)")
