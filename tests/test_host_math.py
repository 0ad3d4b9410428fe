"""CPU-side checks of the product's device math (csrc/dwdf_math.cuh compiled for the host by
tests/host_math/host_math.cpp): same source the kernels inline, with the MUFU / round-down
intrinsics replaced by IEEE host equivalents. Catches logic errors without a GPU; the GPU parity
tests (-m gpu) are the ones that count."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_inputs, seq_rel_err
from oracle.cpu import ORDER_PLUGIN, ORDER_PYTHON, ClipperParams

SRC = os.path.join(ROOT, "tests", "host_math", "host_math.cpp")
OUT = os.path.join(ROOT, "tests", "host_math", "_build", "libhostmath.so")
INC = os.path.join(ROOT, "differentiable-wdfs_b200", "csrc")


@pytest.fixture(scope="session")
def hm():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-I", INC, SRC, "-o", OUT], check=True)
    return C.CDLL(OUT)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def scalar(hm, kind, x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    hm.hm_scalar(C.c_int(kind), P(x), P(out), C.c_int64(x.size))
    return out


def clip(hm, mode, py, p, x, g=None, recover=False):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    acc = np.zeros(3)
    g = None if g is None else np.ascontiguousarray(g, np.float32)
    general = int(not (p.n_up == 1 and p.n_down == 1))
    (hm.hm_clipper_recover if recover else hm.hm_clipper)(C.c_int(mode), C.c_int(general), C.c_int(py), C.c_float(p.fs), C.c_float(p.R), C.c_float(p.C), C.c_float(p.Is), C.c_float(p.Vt), C.c_float(p.nabla), C.c_float(p.n_up),
                  C.c_float(p.n_down), P(x), None if g is None else P(g), P(y), P(acc), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]))
    return y, acc


def test_omega_against_reference_vectors(hm, golden, known):
    """The throughput-oriented omega3/omega4/exp/log of dwdf_math.cuh against the reference's own
    outputs (golden) and the OmegaTest.cpp table."""
    x = golden["omega_x"]
    assert np.max(np.abs(scalar(hm, 0, x) - golden["omega3_f32"])) < 4e-6
    w4, r4 = scalar(hm, 1, x), golden["omega4_f32"]
    assert np.max(np.abs(w4 - r4) / np.maximum(np.abs(r4), 1e-30)) < 2e-6
    e, re_ = scalar(hm, 2, x), golden["exp_approx_f32"]
    integer = (np.round(x * 1.442695) == x * 1.442695)  # the reference's truncation quirk at negative integers
    assert np.max((np.abs(e - re_) / np.maximum(np.abs(re_), 1e-37))[~integer]) < 2e-6
    assert np.max(np.abs(scalar(hm, 3, golden["log_x"]) - golden["log_approx_f32"])) < 4e-6
    tab = np.array(known["omega_table"])
    assert np.max(np.abs(scalar(hm, 1, tab[:, 0]) - tab[:, 1])) < known["omega_tolerances"]["omega4"]
    for kind in (4, 5):  # exact omega, 2 and 1 refinement iterations
        assert np.max(np.abs(scalar(hm, kind, tab[:, 0]) / tab[:, 1] - 1)) < 5e-7
        assert np.max(np.abs(scalar(hm, kind, x) / golden["toms917_f64"] - 1)) < 5e-7


@pytest.mark.parametrize("mode,name", [(0, "best"), (1, "toms"), (2, "good")])
def test_pair_law_against_reference_vectors(hm, golden, mode, name):
    a = golden["pair_a"]
    p = ClipperParams()
    for Rp in ("4301.51", "100", "1e+06"):
        b = np.empty_like(a)
        hm.hm_pair(C.c_int(mode), C.c_int(0), C.c_int(0), C.c_float(float(Rp)), C.c_float(p.Is), C.c_float(p.Vt), C.c_float(p.nabla), C.c_float(1), C.c_float(1), P(a), P(b), None, C.c_int64(a.size))
        ref = golden[f"pair_{name}_f32_Rp{Rp}"]
        assert np.max(np.abs(b - ref)) / np.max(np.abs(ref)) < 2e-6, (name, Rp)


def test_general_pair_law_eq45(hm, golden, diode_configs):
    """N_up / N_down law against diode_pretraining.py:39-60 executed as is (golden eq45_*)."""
    a = golden["eq45_a"].astype(np.float32)
    for name, cfg in diode_configs.items():
        if f"eq45_{name}" not in golden.files:
            continue
        for i, R in enumerate(golden["eq45_R"]):
            b = np.empty_like(a)
            hm.hm_pair(C.c_int(1), C.c_int(1), C.c_int(0), C.c_float(R), C.c_float(cfg["Is"]), C.c_float(cfg["Vt"]), C.c_float(cfg["nabla"]), C.c_float(cfg["N_up"]), C.c_float(cfg["N_down"]), P(a), P(b),
                       None, C.c_int64(a.size))
            ref = golden[f"eq45_{name}"][i]
            assert np.max(np.abs(b - ref)) / np.max(np.abs(ref)) < 5e-6, (name, R)


@pytest.mark.parametrize("circuit", ["plugin", "training"])
@pytest.mark.parametrize("mode,mname", [(0, "approx"), (1, "exact")])
@pytest.mark.parametrize("py,oname", [(0, "plugin"), (1, "python")])
def test_clipper_recurrence_against_reference_vectors(hm, golden, circuit, mode, mname, py, oname):
    p = ClipperParams() if circuit == "plugin" else ClipperParams(R=45000.0, C=4.7e-9)
    y, _ = clip(hm, mode, py, p, golden["clip_x"])
    assert seq_rel_err(y, golden[f"clip_{circuit}_{mname}_{oname}_f32"]) < 1e-5
    assert seq_rel_err(y, golden[f"clip_{circuit}_{mname}_{oname}_f64"]) < 1e-5


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("py,oord", [(0, ORDER_PLUGIN), (1, ORDER_PYTHON)])
@pytest.mark.parametrize("n_up,n_down", [(1, 1), (1, 2)])
def test_step_tape_adjoint_sums(hm, oracle, mode, py, oord, n_up, n_down):
    """clip_step_tape's (A, cg, cl, cv) swept backwards reproduce the oracle's raw adjoint sums."""
    p = ClipperParams(n_up=n_up, n_down=n_down)
    x = make_inputs(8, 1024, seed=5)
    g = np.random.default_rng(5).standard_normal(x.shape).astype(np.float32)
    _, acc = clip(hm, mode, py, p, x, g)
    ref = oracle.clipper_grad(x, g, p, exact=bool(mode), ordering=oord, mode="upstream", dtype=np.float64)
    assert np.max(np.abs(acc / ref["raw"][:3] - 1)) < 2e-5


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("py,oord", [(0, ORDER_PLUGIN), (1, ORDER_PYTHON)])
@pytest.mark.parametrize("n_up,n_down", [(1, 1), (1, 2), (3, 1)])
@pytest.mark.parametrize("amp", [(0.1, 2.0), (3.0, 10.0)])
def test_step_recover_adjoint_sums(hm, oracle, mode, py, oord, n_up, n_down, amp):
    """The adjoint kernel's scheme — states recovered from the forward OUTPUT, linearisation read off
    (x, z, z') by clip_step_recover without evaluating the forward-biased omega — gives the oracle's
    raw adjoint sums (quiet and loud inputs, symmetric and asymmetric pairs, T not a multiple of 16)."""
    p = ClipperParams(n_up=n_up, n_down=n_down)
    x = make_inputs(8, 1000, seed=6, amp=amp)
    g = np.random.default_rng(6).standard_normal(x.shape).astype(np.float32)
    y, acc = clip(hm, mode, py, p, x, g, recover=True)
    y0, acc0 = clip(hm, mode, py, p, x, g)
    assert np.array_equal(y, y0)
    assert np.max(np.abs(acc / acc0 - 1)) < 5e-5, acc / acc0 - 1  # same sums as the replayed tape
    if amp[1] <= 2.0 or mode == 1:
        # (loud inputs cross the seams of omega4's piecewise approximation, where the fp32 and fp64 trajectories part: no oracle there)
        ref = oracle.clipper_grad(x, g, p, exact=bool(mode), ordering=oord, mode="upstream", dtype=np.float64)
        assert np.max(np.abs(acc / ref["raw"][:3] - 1)) < 5e-5, (acc / ref["raw"][:3] - 1, acc0 / ref["raw"][:3] - 1)


def fast(hm, pairs, py, p, x):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    nfb = C.c_int64(0)
    rc = hm.hm_clipper_fast(C.c_int(pairs), C.c_int(py), C.c_float(p.fs), C.c_float(p.R), C.c_float(p.C), C.c_float(p.Is), C.c_float(p.Vt), C.c_float(p.nabla), P(x), P(y), C.byref(nfb),
                            C.c_int64(x.shape[0]), C.c_int64(x.shape[1]))
    assert rc == 0
    assert nfb.value >= 0, f"{-nfb.value} samples below omega3's log branch differ between the plain and the LOUD step"
    return y, nfb.value


@pytest.mark.parametrize("circuit", ["plugin", "training"])
@pytest.mark.parametrize("py,oname", [(0, "plugin"), (1, "python")])
def test_fast_forward_step_against_reference_vectors(hm, golden, circuit, py, oname):
    """clip_step_fastv — the forward kernels' sample (packed pairs f2 and its one-sequence twin f1) with
    the loud chunks redone by the LOUD step (omega3's log branch, still packed) — against the reference's own
    outputs, quiet and loud (+-10 V: nearly every chunk is loud); f1 and f2 agree bit for bit; an instance below the
    branch gets the same bits from the plain and the LOUD step (checked on every chunk); silence stays exact silence."""
    p = ClipperParams() if circuit == "plugin" else ClipperParams(R=45000.0, C=4.7e-9)
    for xname, refname, expect_fallbacks in (("clip_x", f"clip_{circuit}_approx_{oname}_f32", False), ("clip_loud_x", "clip_loud_approx_python_f32", True)):
        if xname == "clip_loud_x" and (circuit != "plugin" or not py):
            continue
        y1, n1 = fast(hm, 0, py, p, golden[xname])
        y2, n2 = fast(hm, 1, py, p, golden[xname])
        assert np.array_equal(y1, y2) and n1 == n2
        assert seq_rel_err(y2, golden[refname]) < 1e-5
        assert (n2 > 0) == expect_fallbacks, n2
    x = make_inputs(5, 1003, seed=3)  # odd row count, T % 4 != 0
    y1, _ = fast(hm, 0, py, p, x)
    y2, _ = fast(hm, 1, py, p, x)
    assert np.array_equal(y1, y2)
    assert seq_rel_err(y2, clip(hm, 0, py, p, x)[0]) < 2e-6
    ys, _ = fast(hm, 1, py, p, np.zeros((3, 64), np.float32))
    assert not ys.any()


def test_recover_pairs_equal_scalar(hm):
    """The adjoint kernel's packed pair step gives the scalar step's linearisation (same formulas)."""
    p = ClipperParams()
    x = make_inputs(4, 1000, seed=11, amp=(0.1, 6.0))
    y, _ = clip(hm, 0, 0, p, x)  # plugin ordering: y[n] = z[n]
    xs, z, zn = x[:, :-1].ravel(), y[:, :-1].ravel(), y[:, 1:].ravel()
    n = xs.size - xs.size % 2
    out = np.zeros((n, 8), np.float32)
    rc = hm.hm_recover_pairs(C.c_float(p.fs), C.c_float(p.R), C.c_float(p.C), C.c_float(p.Is), C.c_float(p.Vt), C.c_float(p.nabla), P(np.ascontiguousarray(xs[:n])), P(np.ascontiguousarray(z[:n])),
                             P(np.ascontiguousarray(zn[:n])), P(out), C.c_int64(n))
    assert rc == 0
    scale = np.max(np.abs(out[:, 4:]), axis=0)
    assert np.all(np.max(np.abs(out[:, :4] - out[:, 4:]), axis=0) <= 2e-6 * scale)


@pytest.mark.parametrize("circuit", ["plugin", "training"])
@pytest.mark.parametrize("py,oname", [(0, "plugin"), (1, "python")])
def test_exact_forward_step_against_reference_vectors(hm, golden, circuit, py, oname):
    """clip_step_exactv — the exact (TOMS-917) root's forward sample as the kernels run it, packed pairs (f2) and its
    one-sequence twin (f1) — against the reference's own C++ outputs, quiet and loud; f1 and f2 agree bit for bit;
    silence in, silence out."""
    p = ClipperParams() if circuit == "plugin" else ClipperParams(R=45000.0, C=4.7e-9)
    outs = []
    for pairs in (0, 1):
        x = np.ascontiguousarray(golden["clip_x"], np.float32)
        y = np.empty_like(x)
        rc = hm.hm_clipper_exactv(C.c_int(pairs), C.c_int(py), C.c_float(p.fs), C.c_float(p.R), C.c_float(p.C), C.c_float(p.Is), C.c_float(p.Vt), C.c_float(p.nabla), P(x), P(y),
                                  C.c_int64(x.shape[0]), C.c_int64(x.shape[1]))
        assert rc == 0
        assert seq_rel_err(y, golden[f"clip_{circuit}_exact_{oname}_f64"]) < 3e-6
        outs.append(y)
    assert np.array_equal(outs[0], outs[1])
    z = np.zeros((3, 40), np.float32)
    yz = np.ones_like(z)
    assert hm.hm_clipper_exactv(C.c_int(1), C.c_int(py), C.c_float(p.fs), C.c_float(p.R), C.c_float(p.C), C.c_float(p.Is), C.c_float(p.Vt), C.c_float(p.nabla), P(z), P(yz), C.c_int64(3), C.c_int64(40)) == 0
    assert not yz.any()


def test_recover_step_on_pairs_exact_root(hm, oracle):
    """The same for the exact (TOMS-917) root: omega_exact_low over packed pairs against the scalar step."""
    p = ClipperParams()
    x = make_inputs(8, 600, seed=12)
    y = oracle.clipper_forward(x, p, exact=True, ordering=ORDER_PLUGIN)  # plugin ordering: y[n] = z[n]
    xs, z, zn = x[:, :-1].ravel(), y[:, :-1].ravel(), y[:, 1:].ravel()
    n = xs.size - xs.size % 2
    out = np.zeros((n, 8), np.float32)
    rc = hm.hm_recover_pairs_mode(C.c_int(1), C.c_float(p.fs), C.c_float(p.R), C.c_float(p.C), C.c_float(p.Is), C.c_float(p.Vt), C.c_float(p.nabla), P(np.ascontiguousarray(xs[:n])),
                                  P(np.ascontiguousarray(z[:n])), P(np.ascontiguousarray(zn[:n])), P(out), C.c_int64(n))
    assert rc == 0
    scale = np.max(np.abs(out[:, 4:]), axis=0)
    assert np.all(np.max(np.abs(out[:, :4] - out[:, 4:]), axis=0) <= 2e-6 * scale)


def test_fast_step_long_time_constant(hm, oracle):
    """gamma = 1e-3 (R 866 kOhm against 5.5 nF at 96 kHz: an RC memory of ~500 samples): the fast step must keep
    a = z + gamma (x - z) in the adaptor's own form; (1 - gamma) z + gamma x rounds the pole by half an ulp of one
    and the long memory amplifies that 1 / (2 gamma)-fold into the DC gain (1e-5; found by tests/test_gpu_fuzz.py)."""
    p = ClipperParams(fs=96000.0, R=866000.0, C=5.5e-9, Is=1.3e-12, nabla=1.9)
    x = (make_inputs(16, 3000, fs=p.fs, seed=7) * 0.3).astype(np.float32)
    for py, oord in ((0, ORDER_PLUGIN), (1, ORDER_PYTHON)):
        ref = oracle.clipper_forward(x, p, ordering=oord)
        for pairs in (0, 1):
            y, _ = fast(hm, pairs, py, p, x)
            assert seq_rel_err(y, ref) < 2e-6


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n_up,n_down", [(1, 1), (1, 2)])
@pytest.mark.parametrize("recover", [False, True])
def test_adjoint_sums_of_a_quiet_signal(hm, oracle, mode, n_up, n_down, recover):
    """Diodes that never conduct (L = -17, |a| < 0.03 V): the parameter sensitivities are sums of terms far below
    fp32's resolution of the waves themselves. Two formulations keep them exact (both found by tests/test_gpu_fuzz.py):
    dV without the (b - a)/V + 2 M1 cancellation (pair_deriv), and the forward-biased omega evaluated directly while the
    diode is off instead of read off a - b (clip_step_recover)."""
    p = ClipperParams(fs=48000.0, R=9778.8, C=2.8895e-07, Is=3.6248e-11, nabla=1.0321, n_up=n_up, n_down=n_down)
    x = make_inputs(3, 400, fs=p.fs, seed=475)
    gy = np.random.default_rng(475).standard_normal(x.shape).astype(np.float32)
    _, acc = clip(hm, mode, 1, p, x, gy, recover=recover)
    ref = oracle.clipper_grad(x, gy, p, exact=bool(mode), ordering=ORDER_PYTHON, mode="upstream", dtype=np.float64)
    assert np.max(np.abs(acc / ref["raw"][:3] - 1)) < 1e-4, acc / ref["raw"][:3] - 1


def clip_from_y(hm, mode, py, p, x, g):
    x = np.ascontiguousarray(x, np.float32)
    g = np.ascontiguousarray(g, np.float32)
    y = np.empty_like(x)
    acc = np.zeros(3)
    rc = hm.hm_clipper_recover_y(C.c_int(mode), C.c_int(py), C.c_float(p.fs), C.c_float(p.R), C.c_float(p.C), C.c_float(p.Is), C.c_float(p.Vt), C.c_float(p.nabla), P(x), P(g), P(y), P(acc),
                                 C.c_int64(x.shape[0]), C.c_int64(x.shape[1]))
    assert rc == 0
    return y, acc


@pytest.mark.parametrize("py,oord", [(0, ORDER_PLUGIN), (1, ORDER_PYTHON)])
@pytest.mark.parametrize("amp", [(0.001, 0.02), (0.1, 2.0), (3.0, 10.0)])
def test_step_recover_from_output_alone(hm, oracle, py, oord, amp):
    """clip_step_recover_yv — the step of the exact root's reverse sweep: states AND linearisation from the forward
    output, x never read (a + b = 2 v, the omegas explicit in the diode voltage) — gives the replayed tape's sums and the
    fp64 oracle's, from signals that never open the diodes to +-10 V, both probe orderings, T not a multiple of 16."""
    p = ClipperParams()
    x = make_inputs(8, 1000, seed=6, amp=amp)
    g = np.random.default_rng(6).standard_normal(x.shape).astype(np.float32)
    _, acc = clip_from_y(hm, 1, py, p, x, g)
    _, acc0 = clip(hm, 1, py, p, x, g)
    assert np.max(np.abs(acc / acc0 - 1)) < 2e-5, acc / acc0 - 1
    ref = oracle.clipper_grad(x, g, p, exact=True, ordering=oord, mode="upstream", dtype=np.float64)
    assert np.max(np.abs(acc / ref["raw"][:3] - 1)) < 2e-5, (acc / ref["raw"][:3] - 1, acc0 / ref["raw"][:3] - 1)


def test_step_recover_from_output_is_not_for_the_approx_root(hm, oracle):
    """Why the approx root's reverse sweep keeps reading x: omega4's own error (~1e-3 w0) is exponentiated when a is
    recovered from the diode voltage, and the sums leave the gradient bar (5e-4) by an order of magnitude."""
    p = ClipperParams()
    x = make_inputs(8, 1000, seed=6, amp=(0.1, 2.0))
    g = np.random.default_rng(6).standard_normal(x.shape).astype(np.float32)
    _, acc = clip_from_y(hm, 0, 1, p, x, g)
    ref = oracle.clipper_grad(x, g, p, exact=False, ordering=ORDER_PYTHON, mode="upstream", dtype=np.float64)
    assert 2e-3 < np.max(np.abs(acc / ref["raw"][:3] - 1)) < 0.2


@pytest.mark.parametrize("params", [dict(fs=48000.0, R=9778.8, C=2.8895e-07, Is=3.6248e-11, nabla=1.0321),  # diodes never conduct (L = -17)
                                    dict(fs=96000.0, R=866000.0, C=5.5e-9, Is=1.3e-12, nabla=1.9),  # gamma = 1e-3
                                    dict(R=1.0e6, C=1.0e-9, Is=6.0e-9, nabla=1.0),  # k = Rp Is / V = 2.4e-3: the fromy_ok boundary (e^-6)
                                    dict(R=180.0, C=1.0e-6, Is=1.0e-15, nabla=2.0)])
def test_step_recover_from_output_parameter_corners(hm, oracle, params):
    p = ClipperParams(**params)
    x = make_inputs(4, 600, fs=p.fs, seed=475)
    gy = np.random.default_rng(475).standard_normal(x.shape).astype(np.float32)
    _, acc = clip_from_y(hm, 1, 1, p, x, gy)
    ref = oracle.clipper_grad(x, gy, p, exact=True, ordering=ORDER_PYTHON, mode="upstream", dtype=np.float64)
    assert np.max(np.abs(acc / ref["raw"][:3] - 1)) < 1e-4, acc / ref["raw"][:3] - 1
