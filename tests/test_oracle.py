"""Pins the CPU oracle (oracle/wdf_oracle.c, oracle/torch_wdf.py) against the reference:
every known-answer the reference's own tests hold for the path and the golden
vectors produced by the reference itself (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle.cpu import (CAPACITOR, INVERTER, ORDER_PLUGIN, ORDER_PYTHON, PARALLEL, RESISTOR, RESVS, ROOT_APPROX, ROOT_DIODE_PAIR, ROOT_EXACT,
                        ROOT_IDEAL_VS, SERIES, ClipperParams)


def bits(a):
    return np.ascontiguousarray(a).view(np.int32 if a.dtype == np.float32 else np.int64)


# ---- known answers held by the reference's tests -------------------------------------------------

def test_omega_table(oracle, known):
    """OmegaTest.cpp:6-48 table with the tolerances of OmegaTest.cpp:146-164."""
    tab = np.array(known["omega_table"])
    tol = known["omega_tolerances"]
    for dtype in (np.float32, np.float64):
        for k in ("omega1", "omega2", "omega3", "omega4"):
            err = np.max(np.abs(oracle.omega(k, tab[:, 0], dtype) - tab[:, 1]))
            assert err < tol[k], (k, dtype, err)
    # exact omega: TOMS-917 restatement reproduces the table to double round-off
    assert np.max(np.abs(oracle.toms917(tab[:, 0]) - tab[:, 1]) / tab[:, 1]) < 1e-14


def test_log_exp_approx_tolerances(oracle, known):
    """OmegaTest.cpp:104-144: log2/log/pow2/exp approximations vs libm."""
    tol = known["omega_tolerances"]
    x = np.linspace(1.0, 2.0, 200)
    assert np.max(np.abs(oracle.omega("log2_approx", x, np.float32) - np.log2(x))) < tol["log2_approx"]
    x = np.linspace(8.0, 12.0, 200)
    assert np.max(np.abs(oracle.omega("log_approx", x, np.float32) - np.log(x))) < tol["log_approx"]
    x = np.linspace(0.0, 1.0, 200)
    assert np.max(np.abs(oracle.omega("pow2_approx", x, np.float32) - 2.0 ** x)) < tol["pow2_approx"]
    x = np.linspace(-4.0, 2.0, 200)
    assert np.max(np.abs(oracle.omega("exp_approx", x, np.float32) - np.exp(x))) < 0.03  # reference tol is relative to its range


def test_standalone_smoke(oracle, known):
    """wdf_standalone_test.cpp:16-36: Series(R 1k, Inverter(ResVs 1k)) + DiodePair(Is 1e-10), 10 V -> 4.77 +- 0.1."""
    nodes = [(RESISTOR, -1, -1, 1.0e3), (RESVS, -1, -1, 1.0e3), (INVERTER, 1, -1, 0.0), (SERIES, 0, 2, 0.0)]
    y = oracle.tree_run(nodes, 48000.0, ROOT_DIODE_PAIR, np.array([[10.0]]), probe=0, source=1, root_par=[0, 0, 1.0e-10, 25.85e-3, 1, 1, 1], ordering=ORDER_PYTHON)
    assert abs(y[0, 0] - known["standalone_test"]["expected"]) < known["standalone_test"]["tol"]
    assert abs(y[0, 0] - known["standalone_test"]["value"]) < 1e-6


def test_static_wdf_fixture(oracle, known):
    """StaticWDFTest.cpp:216-271: Inverter(Parallel(Series(ResVs, R 4.7k), C 47n)) + DiodePair(2.52e-9), Best and Good."""
    k = known["static_wdf_test"]
    nodes = [(RESVS, -1, -1, 1.0e-9), (RESISTOR, -1, -1, k["R"]), (SERIES, 0, 1, 0.0), (CAPACITOR, -1, -1, k["C"]), (PARALLEL, 2, 3, 0.0)]
    x = np.array([k["inputs"]])
    for good, key in ((0, "best"), (1, "good")):
        y = oracle.tree_run(nodes, k["fs"], ROOT_DIODE_PAIR, x, probe=3, source=0, root_par=[0, good, k["Is"], 25.85e-3, 1, 1, 1], ordering=ORDER_PLUGIN, dtype=np.float64)
        np.testing.assert_allclose(y[0], k[key], rtol=0, atol=1e-15)


def test_rc_lowpass_magnitudes(oracle, known):
    """WDFTest.cpp:96-140: -7 / -3 / -1 dB +- 0.1 at 2fc / fc / fc/2."""
    k = known["rc_lowpass_mag_db"]
    nodes = [(RESISTOR, -1, -1, k["R"]), (CAPACITOR, -1, -1, k["C"]), (SERIES, 0, 1, 0.0), (INVERTER, 2, -1, 0.0)]
    n = np.arange(int(k["fs"]))
    for name, f in (("2fc", 2 * k["fc"]), ("fc", k["fc"]), ("fc/2", k["fc"] / 2)):
        y = oracle.tree_run(nodes, k["fs"], ROOT_IDEAL_VS, np.sin(2 * np.pi * f * n / k["fs"])[None], probe=1, dtype=np.float64)[0]
        mag = 20 * np.log10(np.max(np.abs(y[len(y) // 2:])))
        assert abs(mag - k["expected"][name]) < k["tol"]
        assert abs(mag - k["measured"][name]) < 1e-9


def test_divider(oracle, known):
    """CommonWDFTests.h:6-23: equal resistors halve 10 V."""
    nodes = [(RESISTOR, -1, -1, 10000.0), (RESISTOR, -1, -1, 10000.0), (SERIES, 0, 1, 0.0), (INVERTER, 2, -1, 0.0)]
    y = oracle.tree_run(nodes, 48000.0, ROOT_IDEAL_VS, np.array([[10.0]]), probe=0, dtype=np.float64)
    assert y[0, 0] == known["divider"]["expected"]


def test_plugin_impulse_and_spot_values(oracle, known):
    imp = np.zeros((1, 16), np.float32)
    imp[0, 0] = 1
    p = ClipperParams()
    np.testing.assert_array_equal(oracle.clipper_forward(imp, p, exact=True, ordering=ORDER_PLUGIN)[0], np.float32(known["plugin_impulse_response_toms"]))
    np.testing.assert_array_equal(oracle.clipper_forward(imp, p, exact=False, ordering=ORDER_PLUGIN)[0], np.float32(known["plugin_impulse_response_omega4"]))
    s = known["pair_law_spot"]
    np.testing.assert_array_equal(oracle.diode_pair(s["a"], s["Rp"], p, exact=True), np.float32(s["toms_f32"]))
    np.testing.assert_array_equal(oracle.diode_pair(s["a"], s["Rp"], p, exact=False), np.float32(s["omega4_f32"]))


# ---- golden vectors generated by the reference itself --------------------------------------------

def test_scalar_functions_bit_exact(oracle, golden):
    x = golden["omega_x"]
    for k in ("omega3", "omega4", "exp_approx"):
        np.testing.assert_array_equal(bits(oracle.omega(k, x)), bits(golden[f"{k}_f32"]))
    np.testing.assert_array_equal(bits(oracle.omega("omega4", x.astype(np.float64), np.float64)), bits(golden["omega4_f64"]))
    np.testing.assert_array_equal(bits(oracle.omega("log_approx", golden["log_x"])), bits(golden["log_approx_f32"]))
    np.testing.assert_array_equal(bits(oracle.toms917(x.astype(np.float64))), bits(golden["toms917_f64"]))


def test_pair_laws_bit_exact(oracle, golden):
    a = golden["pair_a"]
    p = ClipperParams()
    for Rp in (4301.5083, 100.0, 1.0e6):
        tag = f"Rp{Rp:g}"
        np.testing.assert_array_equal(bits(oracle.diode_pair(a, Rp, p)), bits(golden[f"pair_best_f32_{tag}"]))
        np.testing.assert_array_equal(bits(oracle.diode_pair(a, Rp, p, good=True)), bits(golden[f"pair_good_f32_{tag}"]))
        np.testing.assert_array_equal(bits(oracle.diode_pair(a, Rp, p, exact=True)), bits(golden[f"pair_toms_f32_{tag}"]))
        np.testing.assert_array_equal(bits(oracle.diode_pair(a.astype(np.float64), Rp, p, exact=True, dtype=np.float64)), bits(golden[f"pair_toms_f64_{tag}"]))


def test_eq45_python_law(oracle, golden, diode_configs):
    """diode_pretraining.py:39-60 executed as is (fixture) vs the oracle's general law in fp64."""
    a, R = golden["eq45_a"], golden["eq45_R"]
    for name, d in diode_configs.items():
        p = ClipperParams(Is=d["Is"], nabla=d["nabla"], Vt=d["Vt"], n_up=d["N_up"], n_down=d["N_down"])
        for i, Rp in enumerate(R):
            b = oracle.diode_pair(a, Rp, p, exact=True, dtype=np.float64)
            want = golden[f"eq45_{name}"][i]
            assert np.max(np.abs(b.astype(np.float32) - want)) <= 4e-7 * max(1.0, np.max(np.abs(want))), (name, Rp)


def test_clipper_trajectories_bit_exact(oracle, golden):
    x = golden["clip_x"]
    cases = {"plugin": ClipperParams(), "training": ClipperParams(R=45.0e3, C=4.7e-9)}
    for cname, p in cases.items():
        for rname, exact in (("approx", False), ("exact", True)):
            for oname, order in (("plugin", ORDER_PLUGIN), ("python", ORDER_PYTHON)):
                np.testing.assert_array_equal(bits(oracle.clipper_forward(x, p, exact=exact, ordering=order)), bits(golden[f"clip_{cname}_{rname}_{oname}_f32"]))
                y64 = oracle.clipper_forward(x.astype(np.float64), p, exact=exact, ordering=order, dtype=np.float64)
                np.testing.assert_array_equal(bits(y64), bits(golden[f"clip_{cname}_{rname}_{oname}_f64"]))
    xl = golden["clip_loud_x"]
    np.testing.assert_array_equal(bits(oracle.clipper_forward(xl, ClipperParams(), exact=True)), bits(golden["clip_loud_exact_python_f32"]))
    np.testing.assert_array_equal(bits(oracle.clipper_forward(xl, ClipperParams(), exact=False)), bits(golden["clip_loud_approx_python_f32"]))


def test_tree_interpreter_equals_clipper(oracle, golden):
    """The generic element-by-element interpreter and the written-out clipper loop agree bit for bit."""
    x = golden["clip_x"]
    p = ClipperParams()
    nodes = [(RESVS, -1, -1, p.R), (CAPACITOR, -1, -1, p.C), (PARALLEL, 0, 1, 0.0)]
    for exact in (0, 1):
        for order in (ORDER_PLUGIN, ORDER_PYTHON):
            a = oracle.clipper_forward(x, p, exact=bool(exact), ordering=order)
            b = oracle.tree_run(nodes, p.fs, ROOT_DIODE_PAIR, x, probe=1, source=0, root_par=[exact, 0, p.Is, p.Vt, p.nabla, 1, 1], ordering=order)
            np.testing.assert_array_equal(bits(a), bits(b))


def test_lpf_config1(oracle, golden):
    """BASELINE config 1: RC low-pass of lpf.py:23-28 forward, 1 x 1024 — vs reference output, the closed
    form and scipy.signal.lfilter of the bilinear RC."""
    from scipy.signal import lfilter

    x = golden["lpf_x"]
    nodes = [(RESISTOR, -1, -1, 1000.0), (CAPACITOR, -1, -1, 1.0e-6), (SERIES, 0, 1, 0.0), (INVERTER, 2, -1, 0.0)]
    y64 = oracle.tree_run(nodes, 48000.0, ROOT_IDEAL_VS, x[None].astype(np.float64), probe=1, dtype=np.float64)[0]
    np.testing.assert_array_equal(bits(y64), bits(golden["lpf_y_f64"]))
    y32 = oracle.tree_run(nodes, 48000.0, ROOT_IDEAL_VS, x[None], probe=1)[0]
    np.testing.assert_array_equal(bits(y32), bits(golden["lpf_y_f32"]))
    vr = oracle.tree_run(nodes, 48000.0, ROOT_IDEAL_VS, x[None].astype(np.float64), probe=0, dtype=np.float64)[0]
    np.testing.assert_array_equal(bits(vr), bits(golden["lpf_vr_f64"]))
    # closed form: z' = z + 2(1-g)(x-z), y = (z'+z)/2, g = R/(R+Rc)
    Rc = 1.0 / (2 * 1.0e-6 * 48000.0)
    g = 1000.0 / (1000.0 + Rc)
    z, yc = 0.0, []
    for v in x.astype(np.float64):
        zn = z + 2 * (1 - g) * (v - z)
        yc.append(0.5 * (zn + z))
        z = zn
    np.testing.assert_allclose(y64, yc, rtol=0, atol=1e-13)
    # bilinear transform of 1/(1+sRC): b = [k, k]/(1+k)..., k = 1/(2 fs R C)
    k = 1.0 / (2 * 48000.0 * 1000.0 * 1.0e-6)
    yb = lfilter([k / (1 + k), k / (1 + k)], [1.0, (k - 1) / (1 + k)], x.astype(np.float64))
    np.testing.assert_allclose(y64, yb, rtol=0, atol=1e-12)


def test_divider_vector(oracle, golden):
    nodes = [(RESISTOR, -1, -1, 2000.0), (RESISTOR, -1, -1, 100.0), (SERIES, 0, 1, 0.0), (INVERTER, 2, -1, 0.0)]
    y = oracle.tree_run(nodes, 48000.0, ROOT_IDEAL_VS, golden["lpf_x"][None].astype(np.float64), probe=0, dtype=np.float64)[0]
    np.testing.assert_array_equal(bits(y), bits(golden["divider_y_f64"]))


# ---- the compiled reference, where present (it travels to the GPU box) ----------------------------

def test_oracle_vs_compiled_reference(oracle, ref):
    from conftest import make_inputs

    x = make_inputs(6, 700, seed=7)
    p = ClipperParams(R=33.0e3, C=3.3e-9, Is=2.52e-9, nabla=1.0)
    for exact, root in ((False, ROOT_APPROX), (True, ROOT_EXACT)):
        for order in (ORDER_PLUGIN, ORDER_PYTHON):
            np.testing.assert_array_equal(bits(oracle.clipper_forward(x, p, exact=exact, ordering=order)), bits(ref.clipper(x, p, root, order)))


# ---- gradient oracle: analytic adjoint (C, fp64) vs torch.autograd vs central differences ---------

@pytest.mark.parametrize("params", [ClipperParams(), ClipperParams(Is=1.0e-6, nabla=1.3, n_up=1, n_down=2)])
@pytest.mark.parametrize("ordering", [ORDER_PYTHON, ORDER_PLUGIN])
def test_gradient_oracles_agree(oracle, params, ordering):
    import torch

    from conftest import make_inputs
    from oracle import torch_wdf as tw

    p = params
    x = make_inputs(3, 80, seed=11).astype(np.float64)
    pt = ClipperParams(**{**p.__dict__, "R": p.R * 1.1, "C": p.C * 0.9, "Is": p.Is * 2, "nabla": p.nabla * 1.05})
    tgt = oracle.clipper_forward(x, pt, exact=True, ordering=ordering, dtype=np.float64)
    skip = 10
    for loss in ("mse", "mse+esr"):
        res = oracle.clipper_grad(x, tgt, p, exact=True, ordering=ordering, mode="target", loss=loss, skip=skip)
        y, leaf = tw.clipper_forward(x, p, "exact", ordering)
        np.testing.assert_allclose(y.detach().numpy(), res["y"], rtol=0, atol=1e-13)
        L = tw.mse_esr_loss(torch.as_tensor(tgt)[:, skip:], y[:, skip:], loss == "mse+esr")
        g = np.array([v.item() for v in torch.autograd.grad(L, [leaf["Is"], leaf["nabla"], leaf["R"], leaf["C"]])])
        assert abs(L.item() - res["loss"]) < 1e-12 * max(1.0, abs(L.item()))
        np.testing.assert_allclose(res["grads"], g, rtol=1e-9)

        def lossf(pp):
            return oracle.clipper_grad(x, tgt, pp, exact=True, ordering=ordering, mode="target", loss=loss, skip=skip)["loss"]

        for i, k in enumerate(("Is", "nabla", "R", "C")):
            h = getattr(p, k) * 1e-6
            fd = (lossf(ClipperParams(**{**p.__dict__, k: getattr(p, k) + h})) - lossf(ClipperParams(**{**p.__dict__, k: getattr(p, k) - h}))) / (2 * h)
            assert abs(fd - res["grads"][i]) <= 1e-6 * abs(fd), (k, fd, res["grads"][i])


def test_upstream_gradient_mode_and_gx(oracle):
    """mode 'gy': arbitrary upstream gradient; dL/dx checked by finite differences."""
    from conftest import make_inputs

    p = ClipperParams()
    x = make_inputs(2, 48, seed=5).astype(np.float64)
    rng = np.random.default_rng(3)
    gy = rng.standard_normal(x.shape)
    res = oracle.clipper_grad(x, gy, p, exact=True, mode="gy", want_gx=True)

    def f(xx):
        return float(np.sum(gy * oracle.clipper_forward(xx, p, exact=True, dtype=np.float64)))

    for (b, n) in ((0, 3), (1, 20), (1, 47)):
        h = 1e-6
        xp, xm = x.copy(), x.copy()
        xp[b, n] += h
        xm[b, n] -= h
        assert abs((f(xp) - f(xm)) / (2 * h) - res["gx"][b, n]) < 1e-7
