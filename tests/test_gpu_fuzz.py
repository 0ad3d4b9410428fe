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


def _random_tree(rng, dwdf, fs, need_source):
    """A random binary tree of the reference's elements; returns (top element, oracle node list, leaves, source, index map)."""
    from oracle.cpu import CAPACITOR, INVERTER, PARALLEL, RESISTOR, RESVS, SERIES

    nodes, elems = [], []

    def leaf(kind):
        if kind == "R":
            v = float(np.exp(rng.uniform(np.log(500.0), np.log(2e5))))
            e = dwdf.Resistor(v, True)
            nodes.append((RESISTOR, -1, -1, v))
        elif kind == "C":
            v = float(np.exp(rng.uniform(np.log(1e-9), np.log(1e-6))))
            e = dwdf.Capacitor(v, fs, True)
            nodes.append((CAPACITOR, -1, -1, v))
        else:
            v = float(np.exp(rng.uniform(np.log(500.0), np.log(2e5))))
            e = dwdf.ResistiveVoltageSource(v, True)
            nodes.append((RESVS, -1, -1, v))
        elems.append(e)
        return len(nodes) - 1

    def build(kinds):
        if len(kinds) == 1:
            return leaf(kinds[0])
        k = int(rng.integers(1, len(kinds)))
        a, b = build(kinds[:k]), build(kinds[k:])
        if rng.random() < 0.5:
            e = dwdf.Series(elems[a], elems[b])
            nodes.append((SERIES, a, b, 0.0))
        else:
            e = dwdf.Parallel(elems[a], elems[b])
            nodes.append((PARALLEL, a, b, 0.0))
        elems.append(e)
        i = len(nodes) - 1
        if rng.random() < 0.25:
            elems.append(dwdf.Inverter(elems[i]))
            nodes.append((INVERTER, i, -1, 0.0))
            i = len(nodes) - 1
        return i

    n_leaves = int(rng.integers(2, 6))
    kinds = ["C"] + [str(rng.choice(["R", "C"])) for _ in range(n_leaves - 1)]
    if need_source:
        kinds[int(rng.integers(1, n_leaves))] = "V"
    order = rng.permutation(n_leaves)
    kinds = [kinds[i] for i in order]
    top = build(kinds)
    return elems[top], nodes, elems


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("DWDF_FUZZ_N4", "24"))))
def test_random_tree(dwdf, oracle, seed):
    """The generic interpreter on random trees (2-5 leaves of R / C / one resistive source, Series / Parallel /
    Inverter adaptors) closed by an ideal voltage source or a diode pair, probe on a random leaf: forward against
    the oracle's tree executor; gradients are finite and reproducible."""
    from oracle.cpu import ROOT_DIODE_PAIR, ROOT_IDEAL_VS

    rng = np.random.default_rng(13000 + seed)
    fs = float(rng.choice([44100.0, 48000.0, 96000.0]))
    diode = bool(rng.integers(2))
    mode = str(rng.choice(["approx", "exact"]))
    ordering = str(rng.choice(["plugin", "python"]))
    oord = ORDER_PLUGIN if ordering == "plugin" else ORDER_PYTHON
    top, nodes, elems = _random_tree(rng, dwdf, fs, need_source=diode)
    leaves = [i for i, n in enumerate(nodes) if n[1] < 0]
    probe = int(rng.choice(leaves))
    B, T = int(rng.choice([1, 5, 40])), int(rng.choice([3, 50, 300]))
    x = (make_inputs(B, T, fs=fs, seed=seed) * float(rng.choice([0.3, 1.0]))).astype(np.float32)
    p = ClipperParams()
    if diode:
        root = dwdf.DiodePair(top, p.Is, p.Vt, p.nabla, trainable=True, mode=mode)
        source = [i for i, e in enumerate(elems) if isinstance(e, dwdf.ResistiveVoltageSource)][0]
        ref_args = dict(root_kind=ROOT_DIODE_PAIR, source=source, root_par=[float(mode == "exact"), 0, p.Is, p.Vt, p.nabla, 1, 1])
    else:
        root = dwdf.IdealVoltageSource()
        ref_args = dict(root_kind=ROOT_IDEAL_VS, source=-1, root_par=None)
    circ = dwdf.compile_circuit(root, tree=top, probe=elems[probe], ordering=ordering)
    xd = torch.from_numpy(x).cuda()
    y = circ.forward(xd).cpu().numpy()
    ref = oracle.tree_run(nodes, fs, ref_args["root_kind"], x, probe=probe, source=ref_args["source"], root_par=ref_args["root_par"], ordering=oord)
    ref64 = oracle.tree_run(nodes, fs, ref_args["root_kind"], x, probe=probe, source=ref_args["source"], root_par=ref_args["root_par"], ordering=oord, dtype=np.float64)
    den = np.maximum(np.max(np.abs(ref), axis=1), 1e-3 * np.max(np.abs(x), axis=1) + 1e-30)
    cond = float(np.max(np.max(np.abs(ref - ref64), axis=1) / den))
    err = float(np.max(np.max(np.abs(y - ref), axis=1) / den))
    assert np.all(np.isfinite(y)) and err < max(1e-5, 5.0 * cond), (nodes, probe, diode, mode, ordering, err, cond)
    if T >= 8 and not circ.is_clipper:
        g1 = circ.backward(target=torch.from_numpy((0.5 * ref).astype(np.float32)).cuda(), loss="mse")["grads"].clone()
        circ.forward(xd)
        g2 = circ.backward(target=torch.from_numpy((0.5 * ref).astype(np.float32)).cuda(), loss="mse")["grads"]
        assert torch.all(torch.isfinite(g1)) and torch.equal(g1, g2)
    if T >= 50 and not circ.is_clipper and (not diode or mode == "exact") and cond < 1e-6:
        # leaf gradients against central differences of the fp64 oracle (exact root or linear circuit: smooth in the values)
        gy = np.random.default_rng(seed).standard_normal(x.shape).astype(np.float32)
        circ.forward(xd)
        g = circ.backward(gy=torch.from_numpy(gy).cuda())["grads"].cpu().numpy()
        fds, gs = [], []
        for i in leaves:
            h = 1e-5 * nodes[i][3]
            yp, ym = [oracle.tree_run([n if j != i else (n[0], n[1], n[2], n[3] + sgn * h) for j, n in enumerate(nodes)], fs, ref_args["root_kind"], x, probe=probe, source=ref_args["source"],
                                      root_par=ref_args["root_par"], ordering=oord, dtype=np.float64) for sgn in (1, -1)]
            fds.append(float(np.sum(gy.astype(np.float64) * (yp - ym)) / (2 * h)) * nodes[i][3])
            gs.append(float(g[circ.slot(elems[i], "C" if isinstance(elems[i], dwdf.Capacitor) else "R")]) * nodes[i][3])
        fds, gs = np.array(fds), np.array(gs)  # d/d ln(value): comparable across R (ohms) and C (farads)
        floor = 1e-6 * float(np.linalg.norm(gy) * np.linalg.norm(ref64))  # sensitivities that are zero by topology: finite-difference noise
        assert np.max(np.abs(gs - fds)) < 2e-3 * np.max(np.abs(fds)) + floor, (nodes, probe, diode, mode, ordering, gs, fds)
