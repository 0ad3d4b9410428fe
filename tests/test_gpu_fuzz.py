"""Randomised parity sweep (-m gpu): shapes, circuit constants, diode laws, root modes and probe orderings
drawn from seeded distributions wide enough to leave every fast path's precondition (L < -3.34, L > -39,
N_up = N_down, T % 4 = 0, B > 32, B <= 8192, ...) on both sides; forward against the C oracle
(<= 1e-5 of the peak per sequence), gradients against its fp64 reverse sweep."""
import numpy as np
import pytest
import torch

from conftest import make_inputs, seq_rel_err
from oracle.cpu import ORDER_PLUGIN, ORDER_PYTHON, ClipperParams

pytestmark = pytest.mark.gpu


def draw(rng):
    logu = lambda lo, hi: float(np.exp(rng.uniform(np.log(lo), np.log(hi))))
    p = ClipperParams(fs=float(rng.choice([44100.0, 48000.0, 96000.0])), R=logu(1e3, 1e6), C=logu(1e-10, 1e-6), Is=logu(1e-13, 1e-5), nabla=float(rng.uniform(1.0, 2.5)),
                      n_up=int(rng.choice([1, 1, 2, 3])), n_down=int(rng.choice([1, 1, 2, 3])))
    B = int(rng.choice([1, 2, 31, 32, 33, 64, 65, 100, 257]))
    T = int(rng.choice([1, 3, 4, 16, 17, 100, 512, 515, 1024, 1500, 2048]))
    return p, B, T, str(rng.choice(["approx", "exact"])), str(rng.choice(["plugin", "python"])), float(rng.choice([0.3, 1.0, 3.0]))


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("DWDF_FUZZ_N", "48"))))
def test_random_circuit(dwdf, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    p, B, T, mode, ordering, gain = draw(rng)
    oord = ORDER_PLUGIN if ordering == "plugin" else ORDER_PYTHON
    x = (make_inputs(B, T, fs=p.fs, seed=seed) * gain).astype(np.float32)
    Vs = dwdf.ResistiveVoltageSource(p.R, True)
    C = dwdf.Capacitor(p.C, p.fs, True)
    dp = dwdf.DiodePair(dwdf.Parallel(Vs, C), p.Is, p.Vt, p.nabla, p.n_up, p.n_down, trainable=True, mode=mode)
    circ = dwdf.compile_circuit(dp, probe=C, ordering=ordering)
    xd = torch.from_numpy(x).cuda()
    y = circ.forward(xd)
    yn = y.cpu().numpy()
    ref = oracle.clipper_forward(x, p, exact=(mode == "exact"), ordering=oord)
    # How well is this circuit conditioned at all? The reference's own fp32 and fp64 runs part by `cond`: hard-conducting
    # constants (L = ln(Rp Is / V) > 0: |dz'/dz| -> 1) and the seams of omega4 make some draws chaotic at the 1e-5 level.
    cond = seq_rel_err(ref, oracle.clipper_forward(x, p, exact=(mode == "exact"), ordering=oord, dtype=np.float64))
    assert np.all(np.isfinite(yn))
    # per sequence, relative to the output's peak — but not below -60 dB of the input's (a 4-sample sequence through a long RC
    # has an output of a few microvolts, where 1e-11 V of rounding is "1e-5")
    den = np.maximum(np.max(np.abs(ref), axis=1), 1e-3 * np.max(np.abs(x), axis=1) + 1e-30)
    err = float(np.max(np.max(np.abs(yn - ref), axis=1) / den))
    assert err < max(1e-5, 5.0 * cond), (p, B, T, mode, ordering, gain, cond)
    loud = cond > 3e-6
    if T < 8:
        return
    target = (0.6 * ref + 0.01).astype(np.float32)
    res = circ.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr", skip=min(4, T // 2))
    g = res["grads"].cpu().numpy()[[circ.slot(dp, "Is"), circ.slot(dp, "nabla"), circ.slot(Vs, "R"), circ.slot(C, "C")]]
    assert np.all(np.isfinite(g))
    if loud:
        return
    gref = oracle.clipper_grad(x, target, p, exact=(mode == "exact"), ordering=oord, mode="target", loss="mse+esr", skip=min(4, T // 2), dtype=np.float64)
    scale = np.abs(gref["grads"]) + 1e-3 * np.max(np.abs(gref["grads"] * np.array([p.Is, p.nabla, p.R, p.C]))) / np.array([p.Is, p.nabla, p.R, p.C])
    assert np.max(np.abs(g - gref["grads"]) / scale) < 2e-3, (p, B, T, mode, ordering, g, gref["grads"])
    assert abs(float(res["loss"]) / gref["loss"] - 1) < 1e-4


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("DWDF_FUZZ_N2", "24"))))
def test_random_circuit_other_entry_points(dwdf, oracle, seed):
    """Same draws through the other entry points: upstream-gradient mode with dL/dx, the fused one-sweep training pass
    against forward + adjoint, and streaming in ragged blocks against one long block."""
    rng = np.random.default_rng(5000 + seed)
    p, B, T, mode, ordering, gain = draw(rng)
    T = max(T, 24)
    oord = ORDER_PLUGIN if ordering == "plugin" else ORDER_PYTHON
    x = (make_inputs(B, T, fs=p.fs, seed=seed) * gain).astype(np.float32)
    ref = oracle.clipper_forward(x, p, exact=(mode == "exact"), ordering=oord)
    cond = seq_rel_err(ref, oracle.clipper_forward(x, p, exact=(mode == "exact"), ordering=oord, dtype=np.float64))
    Vs = dwdf.ResistiveVoltageSource(p.R, True)
    C = dwdf.Capacitor(p.C, p.fs, True)
    dp = dwdf.DiodePair(dwdf.Parallel(Vs, C), p.Is, p.Vt, p.nabla, p.n_up, p.n_down, trainable=True, mode=mode)
    circ = dwdf.compile_circuit(dp, probe=C, ordering=ordering)
    order = [circ.slot(dp, "Is"), circ.slot(dp, "nabla"), circ.slot(Vs, "R"), circ.slot(C, "C")]
    xd = torch.from_numpy(x).cuda()
    y = circ.forward(xd)
    # streaming: three ragged blocks continue the same signal
    cuts = sorted(set([0, T // 3 + 1, (2 * T) // 3 + 2, T]))
    st = circ.new_state(B)
    parts = [circ.process_block(xd[:, a:b].contiguous(), st) for a, b in zip(cuts[:-1], cuts[1:])]
    assert seq_rel_err(torch.cat(parts, 1).cpu().numpy(), y.cpu().numpy()) < max(2e-6, 5.0 * cond)
    if cond > 3e-6:
        return
    # upstream gradient + dL/dx against the fp64 oracle
    gy = np.random.default_rng(seed).standard_normal(x.shape).astype(np.float32)
    res = circ.backward(gy=torch.from_numpy(gy).cuda(), want_gx=True)
    gref = oracle.clipper_grad(x, gy, p, exact=(mode == "exact"), ordering=oord, mode="upstream", dtype=np.float64, want_gx=True)
    g = res["grads"].cpu().numpy()[order]
    pv = np.array([p.Is, p.nabla, p.R, p.C])
    scale = np.abs(gref["grads"]) + 1e-3 * np.max(np.abs(gref["grads"] * pv)) / pv
    assert np.max(np.abs(g - gref["grads"]) / scale) < 2e-3, (p, B, T, mode, ordering, g, gref["grads"])
    gx = res["gx"].cpu().numpy()
    assert np.max(np.abs(gx - gref["gx"])) < 2e-4 * np.max(np.abs(gref["gx"])), (p, B, T, mode, ordering)
    # fused training pass == forward + adjoint
    target = (0.6 * ref + 0.01).astype(np.float32)
    td = torch.from_numpy(target).cuda()
    circ.forward(xd)
    a = circ.backward(target=td, loss="mse+esr", skip=4)
    ga, la = a["grads"].cpu().numpy()[order].copy(), float(a["loss"])
    b = circ.train_pass(xd, td, loss="mse+esr", skip=4)
    gb, lb = b["grads"].cpu().numpy()[order], float(b["loss"])
    assert abs(la / lb - 1) < 1e-4 and np.max(np.abs(ga - gb) / (np.abs(ga) + 1e-3 * np.max(np.abs(ga * pv)) / pv)) < 2e-3, (p, B, T, mode, ordering, ga, gb)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("DWDF_FUZZ_N3", "16"))))
def test_random_neural_root(dwdf, seed):
    """Neural root with random (orthogonal-ish) weights of every supported shape, random circuit constants, with and
    without the per-sample resistance channel: forward against the numpy oracle, weight gradients against fp64 autograd."""
    from oracle import nn

    rng = np.random.default_rng(9000 + seed)
    n_hidden, H = [(2, 4), (2, 8), (2, 16), (4, 4), (4, 8)][int(rng.integers(5))]
    sizes = [2] + [H] * (n_hidden + 1) + [1]
    w = []
    for i, o in zip(sizes[:-1], sizes[1:]):
        w += [(rng.standard_normal((i, o)) * (0.9 / np.sqrt(i))).astype(np.float32).ravel(), (0.1 * rng.standard_normal(o)).astype(np.float32)]
    w = np.concatenate(w)
    fs = float(rng.choice([44100.0, 48000.0, 50000.0]))
    R, Cv = float(np.exp(rng.uniform(np.log(5e3), np.log(2e5)))), float(np.exp(rng.uniform(np.log(1e-9), np.log(5e-8))))
    B, T = int(rng.choice([1, 2, 3, 65, 130])), int(rng.choice([8, 100, 700, 1100]))
    ordering = str(rng.choice(["plugin", "python"]))
    order = nn.ORDER_PLUGIN if ordering == "plugin" else nn.ORDER_PYTHON
    with_r = bool(rng.integers(2))
    x = make_inputs(B, T, fs=fs, seed=seed)
    r = (np.exp(rng.uniform(np.log(1e4), np.log(1e5), (B, 1))) * np.ones((1, T))).astype(np.float32) if with_r else None
    Vs, Cc = dwdf.ResistiveVoltageSource(R), dwdf.Capacitor(Cv, fs)
    circ = dwdf.compile_circuit(dwdf.DenseRootModel(dwdf.model_io.json_from_weights(w, sizes)), tree=dwdf.Parallel(Vs, Cc), probe=Cc, ordering=ordering, r_element=Vs if with_r else None)
    xd = torch.from_numpy(x).cuda()
    rd = torch.from_numpy(r).cuda() if with_r else None
    y = circ.forward(xd, r=rd).cpu().numpy()
    ref = nn.nn_clipper_forward(x, w, sizes, fs, R, Cv, order, r=r, dtype=np.float64)
    assert np.all(np.isfinite(y))
    # (a random network need not be contractive like a diode: compare where the fp32 and fp64 oracles agree themselves)
    cond = seq_rel_err(nn.nn_clipper_forward(x, w, sizes, fs, R, Cv, order, r=r, dtype=np.float32), ref)
    assert seq_rel_err(y, ref) < max(5e-5, 5.0 * cond), (sizes, B, T, ordering, with_r, cond)
    if cond > 1e-5 or B * T > 20000:
        return
    target = (0.7 * ref + 0.02).astype(np.float32)
    res = circ.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr", skip=min(4, T // 2))
    gref = nn.nn_clipper_grad_torch(x, target, w, sizes, fs, R, Cv, order, r=r, loss="mse+esr", skip=min(4, T // 2))
    g = res["grads"].cpu().numpy()
    assert np.max(np.abs(g - gref["grad_w"])) < 5e-4 * np.max(np.abs(gref["grad_w"])), (sizes, B, T, ordering, with_r)
    assert abs(float(res["loss"]) / gref["loss"] - 1) < 1e-4
