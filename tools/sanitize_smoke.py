"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dwdf = importlib.import_module("differentiable-wdfs_b200")
rng = np.random.default_rng(0)
for mode in ("approx", "exact"):
    for B, T in ((70, 96), (33, 2052), (5, 37)):
        x = torch.from_numpy((rng.standard_normal((B, T)) * 0.5).astype(np.float32)).cuda()
        Vs = dwdf.ResistiveVoltageSource(47000.0, True); Cc = dwdf.Capacitor(2.2e-9, 48000.0, True)
        dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=mode)
        circ = dwdf.compile_circuit(dp, probe=Cc)
        y = circ.forward(x)
        res = circ.backward(target=(0.5 * y).contiguous(), loss="mse+esr", skip=8)
        circ.train_pass(x, (0.5 * y).contiguous(), y=torch.empty_like(x))
        circ.backward(gy=torch.ones_like(x), want_gx=True)
nnv = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "nn_vectors.npz"))
mj = dwdf.model_io.json_from_weights(nnv["2x8_weights"], [int(v) for v in nnv["2x8_sizes"]])
Vs = dwdf.ResistiveVoltageSource(47000.0); Cc = dwdf.Capacitor(2.2e-9, 48000.0)
cn = dwdf.compile_circuit(dwdf.DenseRootModel(mj), tree=dwdf.Parallel(Vs, Cc), probe=Cc)
x = torch.from_numpy((rng.standard_normal((9, 130)) * 0.5).astype(np.float32)).cuda()
y = cn.forward(x); cn.backward(target=(0.5 * y).contiguous())
# the staged outer-product adjoint at every tile shape (H = 4, 8, 16; 2 and 4 hidden layers), ragged pairs, several warps,
# and its time-parallel phases (option 16 forces them)
for name, B, T in (("2x16", 70, 140), ("4x8", 67, 96), ("2x4", 3, 70), ("4x4", 40, 65)):
    mjn = dwdf.model_io.json_from_weights(nnv[f"{name}_weights"], [int(v) for v in nnv[f"{name}_sizes"]])
    Vn = dwdf.ResistiveVoltageSource(47000.0); Cn = dwdf.Capacitor(2.2e-9, 48000.0)
    cnn = dwdf.compile_circuit(dwdf.DenseRootModel(mjn), tree=dwdf.Parallel(Vn, Cn), probe=Cn)
    xn = torch.from_numpy((rng.standard_normal((B, T)) * 0.5).astype(np.float32)).cuda()
    yn = cnn.forward(xn); cnn.backward(target=(0.5 * yn).contiguous(), loss="mse+esr", skip=5)
prev = dwdf.set_option(16)
xn = torch.from_numpy((rng.standard_normal((5, 600)) * 0.5).astype(np.float32)).cuda()
yn = cn.forward(xn); cn.backward(target=(0.5 * yn).contiguous())
dwdf.set_option(prev)
R1 = dwdf.Resistor(1000.0, True); C1 = dwdf.Capacitor(1.0e-6, 48000.0, True)
ct = dwdf.compile_circuit(dwdf.IdealVoltageSource(), tree=dwdf.Inverter(dwdf.Series(R1, C1)), probe=C1)
y = ct.forward(x); ct.backward(target=(0.5 * y).contiguous())
# ---- round 2 ----------------------------------------------------------------------------------------------------------
# time chunks as a grid dimension (forced), the as-called loss composition, the resistance-channel kernels, train_step
prev = dwdf.set_option(16)
for mode in ("approx", "exact"):
    Vs = dwdf.ResistiveVoltageSource(47000.0, True); Cc = dwdf.Capacitor(2.2e-9, 48000.0, True)
    dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=mode)
    circ = dwdf.compile_circuit(dp, probe=Cc)
    xc = torch.from_numpy((rng.standard_normal((70, 2048)) * 0.5).astype(np.float32)).cuda()
    yc = circ.forward(xc); circ.backward(target=(0.5 * yc).contiguous(), loss="mse+esr_as_called", skip=8)
    opt = dwdf.Adam(circ, lr=1e-6)
    circ.train_step(xc, (0.5 * yc).contiguous(), opt, loss="mse+esr", skip=8)
    circ.train_step(xc, (0.5 * yc).contiguous(), opt, loss="mse", engine="tangent")
    Vr = dwdf.ResistiveVoltageSource(47000.0); Cr = dwdf.Capacitor(2.2e-9, 48000.0, True)
    dr = dwdf.DiodePair(dwdf.Parallel(Vr, Cr), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=mode)
    cr = dwdf.compile_circuit(dr, probe=Cr, r_element=Vr)
    for Bq, Tq in ((70, 2048), (5, 37)):
        xr = torch.from_numpy((rng.standard_normal((Bq, Tq)) * 0.5).astype(np.float32)).cuda()
        rr = torch.full_like(xr, 47000.0); rr[::2] = 10000.0
        yr = cr.forward(xr, r=rr); cr.backward(target=(0.5 * yr).contiguous(), loss="mse")
dwdf.set_option(prev)
# neural root: dL/dx, raw sums + finalize, the one-call step, the as-called loss
yn = cn.forward(x); cn.backward(gy=torch.ones_like(x), want_gx=True)
cn.forward(x); cn.backward(target=(0.5 * yn).contiguous(), raw=True); cn.finalize(target=True, loss="mse+esr")
cn.train_step(x, (0.5 * yn).contiguous(), dwdf.AdamWeights(cn, lr=1e-4), loss="mse+esr_as_called", skip=5)
# run-time specialised tree programs: TMA kernels (aligned) and the direct twins (ragged), linear and diode roots, both orderings
for ordering in ("python", "plugin"):
    for mode in (None, "approx", "exact"):
        Rt = dwdf.Resistor(47000.0, True); Vt_ = dwdf.ResistiveVoltageSource(4700.0, True); Ct = dwdf.Capacitor(2.2e-9, 48000.0, True)
        top = dwdf.Parallel(Rt, dwdf.Series(Vt_, Ct))
        root = dwdf.IdealVoltageSource() if mode is None else dwdf.DiodePair(top, 4.352e-9, 25.85e-3, 1.906, trainable=True, mode=mode)
        cj = dwdf.compile_circuit(root, tree=top, probe=Ct, ordering=ordering)
        assert cj.specialize()
        for Bq, Tq in ((70, 256), (33, 203)):
            xj = torch.from_numpy((rng.standard_normal((Bq, Tq)) * 0.5).astype(np.float32)).cuda()
            yj = cj.forward(xj); cj.backward(target=(0.5 * yj).contiguous(), loss="mse+esr", skip=3)
            st = cj.new_state(Bq); cj.process_block(xj, st)
# ---- round 2, last session: the exact root's reverse sweep from the output alone (two-tile ring slots, three of them), both
# probe orderings, one chunk and forced time chunks, and the x-reading sweep above the fromy_ok switch (Is = 4e-8)
for ordering in ("python", "plugin"):
    for Is in (4.352e-9, 4.0e-8):
        Vs = dwdf.ResistiveVoltageSource(47000.0, True); Cc = dwdf.Capacitor(2.2e-9, 48000.0, True)
        dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), Is, 25.85e-3, 1.906, trainable=True, mode="exact")
        circ = dwdf.compile_circuit(dp, probe=Cc, ordering=ordering)
        for force in (0, 16):
            prev = dwdf.set_option(force)
            for Bq, Tq in ((70, 1024), (33, 203)):
                xq = torch.from_numpy((rng.standard_normal((Bq, Tq)) * 0.5).astype(np.float32)).cuda()
                yq = circ.forward(xq); circ.backward(target=(0.5 * yq).contiguous(), loss="mse+esr", skip=8)
                circ.forward(xq); circ.backward(gy=torch.ones_like(xq), want_gx=True)
            dwdf.set_option(prev)
# loud inputs (+-8 V on every second row): the warp-voted LOUD step of the approx root's forward kernels (pair, single, direct, forced chunks)
for force in (0, 16, 2):
    prev = dwdf.set_option(force)
    Vs = dwdf.ResistiveVoltageSource(47000.0, True); Cc = dwdf.Capacitor(2.2e-9, 48000.0, True)
    dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, trainable=True, mode="approx")
    circ = dwdf.compile_circuit(dp, probe=Cc)
    for Bq, Tq in ((70, 1024), (33, 203)):
        xq = torch.from_numpy((rng.standard_normal((Bq, Tq)) * 0.3).astype(np.float32)).cuda()
        xq[1::2] *= 25.0
        yq = circ.forward(xq); circ.backward(target=(0.5 * yq).contiguous(), loss="mse", skip=8)
    dwdf.set_option(prev)
torch.cuda.synchronize()
print("sanitize smoke done")
