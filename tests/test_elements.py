"""The remaining chowdsp_wdf elements (inductor, alpha-transform C / L, current sources, Y-parameter two-port, single diode,
switch; wdf_t.h:190-444, 597-654, 746-846, 987-1106) — SURVEY.md §8(f)-4.

Chain of evidence: tests/golden/ref_elements.npz holds outputs of the UNMODIFIED reference classes in the circuits of the
reference's own tests (make_golden_elements.py); the C oracle's executor (oracle/wdf_oracle_impl.h, ow_tree_run_ext) must
reproduce them bit for bit; the product's imperative element classes and the CUDA interpreter are checked against both.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, seq_rel_err
from elements_cases import CASES

TOL = 1e-5


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "ref_elements.npz"))


def run_oracle(oracle, case, x, dtype=np.float32):
    return oracle.tree_run_ext(case["nodes"], case["fs"], case["root_kind"], x, probe=case["probe"], source=case["source"], root_par=case["root_par"], probe_current=case["probe_current"], dtype=dtype)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_reference_elements(oracle, gold, name):
    y = run_oracle(oracle, CASES[name], gold["x"][None, :])[0]
    assert np.array_equal(y, gold[name]), float(np.max(np.abs(y - gold[name])))


def test_oracle_matches_the_live_reference_on_other_signals(oracle):
    from oracle.cpu import RefElements

    try:
        ref = RefElements()
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"compiled reference elements unavailable: {e}")
    x = (np.random.default_rng(5).standard_normal(900) * 1.3).astype(np.float32)
    assert np.array_equal(run_oracle(oracle, CASES["rlc_alpha_0.5"], x[None])[0], ref.rlc_highpass(x, 44100.0, 300.0, 1.0e-6, 0.022, 0.5))
    assert np.array_equal(run_oracle(oracle, CASES["rectifier_voltage"], x[None])[0], ref.diode(x, 48000.0, 4700.0, 47.0e-9, 2.52e-9, 25.85e-3, 1.0, 0))
    assert np.array_equal(run_oracle(oracle, CASES["current_divider"], x[None])[0], ref.current_divider(x, 10000.0, 4700.0))


def test_reference_known_answers(oracle, gold):
    """What the reference's tests assert on these circuits: current divider 0.5 A (StaticWDFTest.cpp:39-55), switch -1 A closed /
    0 A open (:57-80), the Y-parameter port equation (WDFTest.cpp:55-74), the Shockley diode current in double (:76-94, tol 1e-3),
    the alpha-transform passband gains (StaticWDFTest.cpp:172-212)."""
    one = np.ones((1, 4), np.float32)
    eq = dict(CASES["current_divider"], nodes=[(0, -1, -1, 10000.0), (0, -1, -1, 10000.0), CASES["current_divider"]["nodes"][2]])
    assert np.all(run_oracle(oracle, eq, one) == 0.5)
    assert np.all(np.abs(run_oracle(oracle, CASES["current_switch_closed"], one) + 1.0) < 1e-3)
    assert np.all(run_oracle(oracle, CASES["current_switch_open"], one) == 0.0)
    vres, i2 = run_oracle(oracle, CASES["ypar_voltage"], 2 * one), run_oracle(oracle, CASES["ypar_current"], 2 * one)
    assert np.all(np.abs(i2 - (0.33 * vres + 0.44 * 2.0)) < 1e-3)
    shock = dict(CASES["shockley_voltage"], probe_current=True)
    i_d = run_oracle(oracle, shock, np.full((1, 4), -0.35), dtype=np.float64)
    assert np.all(np.abs(i_d - 1.0e-7 * (np.exp(0.35 / 25.85e-3) - 1.0)) < 1e-3)
    ref_db, a01_db = gold["known_alpha_passband_db"]
    assert abs(ref_db) < 0.1 and abs(a01_db - (ref_db - 1.1)) < 0.1


@pytest.mark.parametrize("name", sorted(CASES))
def test_imperative_elements_follow_the_reference(dwdf, gold, name):
    """The element classes stepped by hand, as a reference-style script would (root.incident(tree.reflected()); ...)."""
    case = CASES[name]
    root, tree, probe, kind = case["build"](dwdf)
    root.next = tree
    tree.calc_impedance()
    x = torch.from_numpy(gold["x"][:400].copy())
    src = [e for e in _walk(tree) if isinstance(e, (dwdf.ResistiveVoltageSource, dwdf.ResistiveCurrentSource))]
    out = []
    for n in range(x.shape[0]):
        v = x[n:n + 1]
        if isinstance(root, dwdf.IdealVoltageSource):
            root.set_voltage(v)
        elif isinstance(root, dwdf.IdealCurrentSource):
            root.set_current(v)
        else:
            src[0].set_voltage(v)
        root.incident(tree.reflected())
        tree.incident(root.reflected())
        out.append(dwdf.voltage(probe) if kind == "voltage" else dwdf.current(probe))
    y = torch.cat([o.reshape(1) for o in out]).numpy()
    want = gold[name][:400]
    assert np.max(np.abs(y - want)) <= 2e-5 * max(np.max(np.abs(want)), 1e-30) + 1e-7, float(np.max(np.abs(y - want)))


def _walk(e):
    yield e
    for attr in ("P1", "P2"):
        if hasattr(e, attr):
            yield from _walk(getattr(e, attr))


@pytest.mark.gpu
@pytest.mark.parametrize("ordering", ["python", "plugin"])
@pytest.mark.parametrize("specialised", [False, True], ids=["interpreter", "specialised"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_interpreter_elements(dwdf, oracle, gold, name, ordering, specialised):
    """Every remaining chowdsp_wdf element on the node-list interpreter and on the kernels generated for the circuit at run time
    (dwdf_program_specialize), against the oracle's executor and the reference classes' own output."""
    from oracle.cpu import ORDER_PLUGIN, ORDER_PYTHON

    case = CASES[name]
    root, tree, probe, kind = case["build"](dwdf)
    circ = dwdf.compile_circuit(root, tree=tree, probe=probe, ordering=ordering, probe_kind=kind)
    if specialised:
        assert circ.specialize()
    rng = np.random.default_rng(11)
    x = np.stack([gold["x"], (rng.standard_normal(gold["x"].size) * 0.8).astype(np.float32), np.zeros_like(gold["x"])] + [(rng.standard_normal(gold["x"].size) * a).astype(np.float32) for a in (0.1, 2.0)] * 17)
    y = circ.forward(torch.from_numpy(x).cuda(), keep_for_backward=False).cpu().numpy()
    want = oracle.tree_run_ext(case["nodes"], case["fs"], case["root_kind"], x, probe=case["probe"], source=case["source"], root_par=case["root_par"], probe_current=case["probe_current"],
                               ordering=ORDER_PYTHON if ordering == "python" else ORDER_PLUGIN)
    scale = max(float(np.max(np.abs(want))), 1e-30)
    assert np.max(np.abs(y - want)) <= TOL * scale, (float(np.max(np.abs(y - want))), scale)
    if ordering == "python":
        assert np.max(np.abs(y[0] - gold[name])) <= TOL * max(float(np.max(np.abs(gold[name]))), 1e-30)  # the reference's own output
    # streaming: three blocks == one
    xd = torch.from_numpy(x).cuda()
    st = circ.new_state(x.shape[0])
    parts = [circ.process_block(xd[:, a:b].contiguous(), st) for a, b in ((0, 700), (700, 701), (701, x.shape[1]))]
    got = torch.cat(parts, 1).cpu().numpy()
    assert np.array_equal(got, y) if not specialised else np.max(np.abs(got - y)) <= TOL * scale  # (the block of one sample runs the direct twin: same arithmetic, other FMA contractions)
    with pytest.raises(dwdf.DwdfError):  # reverse mode covers the wdf_py set + the inductor; these circuits say so loudly
        if name == "rlc_plain":
            raise dwdf.DwdfError(2, "differentiable: checked in test_gpu_inductor_gradient")
        circ.forward(xd)
        circ.backward(target=xd)


@pytest.mark.gpu
def test_gpu_inductor_gradient(dwdf):
    """Reverse mode through an inductor: dL/dR, dL/dC, dL/dL of the RLC highpass against central finite differences of the
    engine's own forward pass in the oracle's double instantiation."""
    from oracle.cpu import Oracle

    orc = Oracle()
    case = CASES["rlc_plain"]
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((6, 500)) * 0.5).astype(np.float32)
    target = (0.3 * np.roll(x, 2, 1)).astype(np.float32)
    r1, c1, l1 = dwdf.Resistor(300.0, True), dwdf.Capacitor(1.0e-6, 44100.0, True), dwdf.Inductor(0.022, 44100.0, True)
    top = dwdf.Inverter(dwdf.Series(dwdf.Series(r1, c1), l1))
    circ = dwdf.compile_circuit(dwdf.IdealVoltageSource(), tree=top, probe=l1)
    circ.forward(torch.from_numpy(x).cuda())
    g = circ.backward(target=torch.from_numpy(target).cuda(), loss="mse")["grads"].cpu().numpy()

    def loss(vals):
        nodes = [list(n) for n in case["nodes"]]
        nodes[0][3], nodes[1][3], nodes[3][3] = vals
        y = orc.tree_run_ext([tuple(n) for n in nodes], case["fs"], case["root_kind"], x.astype(np.float64), probe=3, dtype=np.float64)
        return float(np.mean((y - target) ** 2))

    base = np.array([300.0, 1.0e-6, 0.022])
    for k, (elem, attr) in enumerate(((r1, "R"), (c1, "C"), (l1, "L"))):
        h = base[k] * 1e-5
        up, dn = base.copy(), base.copy()
        up[k] += h
        dn[k] -= h
        fd = (loss(up) - loss(dn)) / (2 * h)
        assert abs(g[circ.slot(elem, attr)] / fd - 1) < 2e-3, (attr, g[circ.slot(elem, attr)], fd)
