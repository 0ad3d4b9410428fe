"""Pinned to the reference's OWN Python code: tests/golden/py_reference_vectors.npz holds outputs, losses and tape.gradient
results of the unmodified wdf_py sources (tf_wdf.py, layers.py, lpf.py's Model, clipper_pot.py's ClipperModel / esr_loss /
loss_func) executed on the TensorFlow look-alike of oracle/shim_tf (tests/golden/make_golden_py_reference.py).

CPU: the oracle's restatements (oracle/torch_wdf.py, oracle/nn.py) reproduce them in fp64. GPU: the kernels — tree programs on the
interpreter and specialised, the neural root with the per-sample resistance channel and the loss as the training loop calls it —
against the same vectors.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, seq_rel_err
from oracle import nn


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLDEN, "py_reference_vectors.npz"))


def test_shim_runs_the_reference_when_present():
    """Where /root/reference exists (this container), the generator itself is the check that the unmodified sources still
    produce the committed vectors; elsewhere (the GPU box) the vectors are all there is."""
    if not os.path.isdir("/root/reference/wdf_py"):
        pytest.skip("the reference tree is not on this machine")
    import subprocess
    import sys
    import tempfile

    here = os.path.dirname(os.path.abspath(__file__))
    src = open(os.path.join(here, "golden", "make_golden_py_reference.py")).read()
    with tempfile.TemporaryDirectory() as d:  # same script, output redirected
        path = os.path.join(d, "gen.py")
        open(path, "w").write(src.replace('HERE = os.path.dirname(os.path.abspath(__file__))', f'HERE = {d!r}').replace('ROOT = os.path.dirname(os.path.dirname(HERE))', f'ROOT = {os.path.dirname(here)!r}'))
        r = subprocess.run([sys.executable, path], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        new = np.load(os.path.join(d, "py_reference_vectors.npz"))
        old = np.load(os.path.join(GOLDEN, "py_reference_vectors.npz"))
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            np.testing.assert_allclose(new[k], old[k], rtol=1e-6, atol=1e-12, err_msg=k)


def test_torch_restatement_of_lpf_matches_the_reference(ref):
    from oracle import torch_wdf as tw

    y, leaves = tw.lpf_forward(ref["lpf_x"], 1000.0, 1.0e-6, 48000.0)
    np.testing.assert_allclose(y[..., 0].t().detach().numpy(), ref["lpf_y_f64"], rtol=1e-12, atol=1e-15)
    loss = torch.mean((y[..., 0].t() - torch.from_numpy(ref["lpf_target"]).double()) ** 2)
    gR, gC = torch.autograd.grad(loss, [leaves["R"], leaves["C"]])
    assert abs(float(loss) / float(ref["lpf_loss_f64"]) - 1) < 1e-12
    np.testing.assert_allclose([float(gR), float(gC)], ref["lpf_grad_RC_f64"], rtol=1e-9)


@pytest.mark.parametrize("name", ["2x8", "2x16"])
def test_numpy_and_torch_restatements_of_the_neural_clipper_match_the_reference(ref, name):
    w, sizes = ref[f"clip_{name}_weights"], [int(v) for v in ref[f"clip_{name}_sizes"]]
    x, r, t = ref[f"clip_{name}_x"], ref[f"clip_{name}_r"], ref[f"clip_{name}_target"]
    fs, Cv, skip = float(ref["clip_fs"]), float(ref["clip_C"]), int(ref["clip_skip"])
    y = nn.nn_clipper_forward(x, w, sizes, fs, 45000.0, Cv, nn.ORDER_PYTHON, r=r, dtype=np.float64)
    np.testing.assert_allclose(y, ref[f"clip_{name}_y_f64"], rtol=1e-9, atol=1e-13)
    g = nn.nn_clipper_grad_torch(x, t, w, sizes, fs, 45000.0, Cv, nn.ORDER_PYTHON, r=r, loss="mse+esr_as_called", skip=skip)
    assert abs(g["loss"] / float(ref[f"clip_{name}_loss_f64"]) - 1) < 1e-10
    gw = ref[f"clip_{name}_grad_w_f64"]
    assert np.max(np.abs(g["grad_w"] - gw)) < 1e-9 * np.max(np.abs(gw))
    # the textbook ESR (energy of the target) is a different loss: the as-called form is what the reference's loop optimises
    g2 = nn.nn_clipper_grad_torch(x, t, w, sizes, fs, 45000.0, Cv, nn.ORDER_PYTHON, r=r, loss="mse+esr", skip=skip)
    assert abs(g2["loss"] / float(ref[f"clip_{name}_loss_f64"]) - 1) > 1e-3


def test_clipper_tree_built_from_the_reference_elements(ref):
    """The diode clipper's tree — ResistiveVoltageSource, Capacitor, Parallel — taken from the reference's tf_wdf.py (unmodified, on
    the shim, fp64) and closed by the oracle's analytic DiodePair (the reference has no differentiable analytic root in Python): output
    and d/d(R, C, Is, nabla) equal the fully restated oracle's, so the gradient oracle of the analytic clipper rests on the reference's
    own adaptor code for everything but the root law (which is pinned separately to diode_pair_func and to the C++ roots)."""
    if not os.path.isdir("/root/reference/wdf_py/lib"):
        pytest.skip("the reference tree is not on this machine")
    import sys

    from conftest import ROOT, make_inputs
    from oracle import torch_wdf as tw
    from oracle.cpu import ClipperParams

    for pth in (os.path.join(ROOT, "oracle", "shim_tf"), "/root/reference/wdf_py/lib"):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import tensorflow as tf  # the shim
    import tf_wdf as rwdf  # the reference

    tf.set_dtype(torch.float64)
    try:
        p = ClipperParams()
        x = make_inputs(3, 200, seed=17)
        target = torch.from_numpy(0.5 * x).double()
        Vs = rwdf.ResistiveVoltageSource(p.R, True)
        C = rwdf.Capacitor(p.C, p.fs, True)
        P1 = rwdf.Parallel(Vs, C)
        dp = tw.DiodePair(P1, p.Is, p.Vt, p.nabla, 1, 1, True, "exact")
        xin = torch.from_numpy(x).double().unsqueeze(-1)
        Vs.reset()
        C.reset()
        P1.calc_impedance()
        outs = []
        for i in range(x.shape[1]):  # clipper_pot.py:110-124 with the analytic root
            Vs.set_voltage(xin[:, i])
            dp.incident(P1.reflected())
            P1.incident(dp.reflected())
            outs.append(rwdf.voltage(C))
        y = torch.stack(outs, dim=1)[..., 0]
        loss = torch.mean((y - target) ** 2)
        g = torch.autograd.grad(loss, [dp.Is, dp.nabla, Vs.R, C.C])
        y2, leaves = tw.clipper_forward(x, p, mode="exact")
        loss2 = torch.mean((y2 - target) ** 2)
        g2 = torch.autograd.grad(loss2, [leaves["Is"], leaves["nabla"], leaves["R"], leaves["C"]])
        np.testing.assert_allclose(y.detach().numpy(), y2.detach().numpy(), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose([float(v) for v in g], [float(v) for v in g2], rtol=1e-9)
    finally:
        tf.set_dtype(torch.float32)


# ---- GPU ---------------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("specialised", [False, True], ids=["interpreter", "specialised"])
def test_lpf_kernels_against_the_reference(dwdf, ref, specialised):
    R1 = dwdf.Resistor(1000.0, True)
    C1 = dwdf.Capacitor(1.0e-6, 48000.0, True)
    circ = dwdf.compile_circuit(dwdf.IdealVoltageSource(), tree=dwdf.Inverter(dwdf.Series(R1, C1)), probe=C1)
    if specialised:
        assert circ.specialize()
    y = circ.forward(torch.from_numpy(ref["lpf_x"]).cuda())
    assert seq_rel_err(y.cpu().numpy(), ref["lpf_y_f64"]) < 1e-5 and seq_rel_err(y.cpu().numpy(), ref["lpf_y_f32"]) < 1e-5
    res = circ.backward(target=torch.from_numpy(ref["lpf_target"]).cuda(), loss="mse")
    g = res["grads"].cpu().numpy()
    got = np.array([g[circ.slot(R1, "R")], g[circ.slot(C1, "C")]])
    assert np.max(np.abs(got / ref["lpf_grad_RC_f64"] - 1)) < 5e-4, (got, ref["lpf_grad_RC_f64"])
    assert abs(float(res["loss"]) / float(ref["lpf_loss_f64"]) - 1) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2x8", "2x16"])
def test_neural_clipper_kernels_against_the_reference(dwdf, ref, name):
    """ClipperModel.forward of clipper_pot.py:103-127 ((B, T, 2) input: signal and per-sample resistance), the loss exactly as
    clipper_pot.py:248 calls it, tape.gradient w.r.t. every kernel and bias."""
    w, sizes = ref[f"clip_{name}_weights"], [int(v) for v in ref[f"clip_{name}_sizes"]]
    fs, Cv, skip = float(ref["clip_fs"]), float(ref["clip_C"]), int(ref["clip_skip"])
    Vs, Cc = dwdf.ResistiveVoltageSource(45.0e3), dwdf.Capacitor(Cv, fs)
    circ = dwdf.compile_circuit(dwdf.DenseRootModel(dwdf.model_io.json_from_weights(w, sizes)), tree=dwdf.Parallel(Vs, Cc), probe=Cc, ordering="python", r_element=Vs)
    x, r, t = (torch.from_numpy(ref[f"clip_{name}_{k}"]).cuda() for k in ("x", "r", "target"))
    y = circ.forward(x, r=r)
    assert seq_rel_err(y.cpu().numpy(), ref[f"clip_{name}_y_f64"]) < 5e-5
    res = circ.backward(target=t, loss="mse+esr_as_called", skip=skip)
    for key, idx in (("loss", "loss"), ("mse", "mse"), ("esr", "esr")):
        assert abs(float(res[idx]) / float(ref[f"clip_{name}_{key}_f64"]) - 1) < 1e-4, key
    gw = ref[f"clip_{name}_grad_w_f64"]
    assert np.max(np.abs(res["grads"].cpu().numpy() - gw)) < 1e-4 * np.max(np.abs(gw))
