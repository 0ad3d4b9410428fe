import importlib, os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from bench import synth_inputs, FS
dwdf = importlib.import_module("differentiable-wdfs_b200")
nnv = np.load("/root/repo/tests/golden/nn_vectors.npz")
dev = torch.device("cuda", 0)
x = synth_inputs(torch, 512, 1, dev)
for name in ("2x8", "2x16"):
    mj = dwdf.model_io.json_from_weights(nnv[f"{name}_weights"], [int(v) for v in nnv[f"{name}_sizes"]])
    Vs = dwdf.ResistiveVoltageSource(47000.0); Cc = dwdf.Capacitor(2.2e-9, FS)
    circ = dwdf.compile_circuit(dwdf.DenseRootModel(mj), tree=dwdf.Parallel(Vs, Cc), probe=Cc, ordering="plugin", device=dev)
    r0 = dwdf.time_parallel_redone()
    y = circ.forward(x, keep_for_backward=False)
    torch.cuda.synchronize()
    print(name, "redone", dwdf.time_parallel_redone() - r0, "of", 256 * 15)
    dwdf.set_option(8); y2 = circ.forward(x, keep_for_backward=False); dwdf.set_option(0)
    print("max diff", float((y - y2).abs().max()), float(y2.abs().max()))
