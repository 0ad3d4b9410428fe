#!/usr/bin/env python
"""Development probe: host-side cost of one forward / backward call (tiny problem: the kernels are a few microseconds)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dwdf = importlib.import_module("differentiable-wdfs_b200")
Vs = dwdf.ResistiveVoltageSource(47000.0, True); Cc = dwdf.Capacitor(2.2e-9, 48000.0, True)
dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode="approx")
circ = dwdf.compile_circuit(dp, probe=Cc)
for B, T in ((64, 128), (64, 4096), (256, 4096), (1024, 4096)):
    x = torch.randn(B, T, device="cuda") * 0.5
    y = torch.empty_like(x)
    t = (0.9 * circ.forward(x)).clone()
    for name, fn in (("forward", lambda: circ.forward(x, out=y)), ("forward_nokeep", lambda: circ.forward(x, out=y, keep_for_backward=False)), ("backward", lambda: circ.backward(target=t, loss="mse", raw=True))):
        for tma in (1, 0):
            prev = dwdf.set_tma(tma)
            circ.forward(x, out=y)
            for _ in range(20): fn()
            torch.cuda.synchronize()
            n = 300
            t0 = time.perf_counter()
            for _ in range(n): fn()
            t_issue = (time.perf_counter() - t0) / n * 1e6
            torch.cuda.synchronize()
            t_total = (time.perf_counter() - t0) / n * 1e6
            dwdf.set_tma(prev)
            print(f"B={B} T={T} {name:15s} tma={tma}: host issue {t_issue:6.1f} us/call, with sync {t_total:6.1f} us/call", flush=True)
