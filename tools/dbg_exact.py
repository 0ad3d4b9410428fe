import sys, os, importlib
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from conftest import make_inputs
dwdf = importlib.import_module("differentiable-wdfs_b200")
from oracle.cpu import ClipperParams
p = ClipperParams()
def mk(ordering="plugin"):
    Vs = dwdf.ResistiveVoltageSource(p.R, True); C = dwdf.Capacitor(p.C, p.fs, True); P1 = dwdf.Parallel(Vs, C)
    dp = dwdf.DiodePair(P1, p.Is, p.Vt, p.nabla, 1, 1, trainable=True, mode="exact")
    return dwdf.compile_circuit(dp, probe=C, ordering=ordering)
x = make_inputs(64, 512, seed=31, amp=(0.1, 2.0))
xd = torch.from_numpy(x).cuda()
dwdf.set_option(8)
c = mk()
y_pair = c.forward(xd).cpu().numpy()
dwdf.set_tma(False)
y_dir = c.forward(xd).cpu().numpy()
dwdf.set_tma(True)
y_tma32 = c.forward(xd[:32].contiguous()).cpu().numpy()
print("pair vs direct equal:", np.array_equal(y_pair, y_dir), " tma(f1) vs direct equal:", np.array_equal(y_tma32, y_dir[:32]))
bad = np.argwhere(y_pair != y_dir)
print("mismatches", len(bad), "of", y_pair.size)
for b, n in bad[:8]:
    print(b, n, x[b, n], y_pair[b, n], y_dir[b, n], "prev equal:", y_pair[b, n-1] == y_dir[b, n-1])
rows = np.unique(bad[:, 0]) if len(bad) else []
print("rows with mismatches:", rows[:40])
