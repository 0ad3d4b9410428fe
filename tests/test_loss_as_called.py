"""The reference's training loss AS ITS LOOP CALLS IT (DWDF_LOSS_MSE_ESR_AS_CALLED, loss="mse+esr_as_called").

clipper_pot.py:248 evaluates ``loss_func(outs, train_Y)`` with ``loss_func = lambda target, pred: mse(target, pred) +
esr_loss(target, pred)`` (:177): the model output sits in esr_loss's ``target_y`` slot (:148-156), so the energy in the
denominator is the prediction's and carries its own gradient. ``oracle.torch_wdf.mse_esr_loss(outs, train_Y)`` called the
same way is the restatement; the CPU test pins the closed form of dL/dy against autograd of that call, the GPU tests the
library's composition (batch sums -> dL/dy -> reverse sweep) against the oracle's upstream-mode gradients.
"""
import numpy as np
import pytest
import torch

from conftest import make_inputs
from oracle.cpu import ORDER_PLUGIN, ORDER_PYTHON, ClipperParams

EPS = float(np.finfo(float).eps)


def as_called_ybar(y, t, skip):
    """Closed form of d loss / dy for loss = mse + esr with the energy of y (fp64 numpy); returns (ybar, loss, mse, esr)."""
    y = np.asarray(y, np.float64)
    t = np.asarray(t, np.float64)
    on = np.zeros_like(y)
    on[:, skip:] = 1.0
    e = (y - t) * on
    S, E, N = float(np.sum(e * e)), float(np.sum((y * on) ** 2)) + EPS, float(np.sum(on))
    mse, esr = S / N, np.sqrt(S / E / N)
    c1 = 2.0 / N + 1.0 / (esr * N * E)
    c2 = -esr / E
    return (c1 * e + c2 * y * on), mse + esr, mse, esr


def test_closed_form_matches_autograd_of_the_reference_call():
    from oracle import torch_wdf as tw

    rng = np.random.default_rng(3)
    y = rng.standard_normal((5, 64)) * 0.3
    t = y + 0.05 * rng.standard_normal((5, 64))
    skip = 7
    yt = torch.tensor(y, dtype=torch.float64, requires_grad=True)
    tt = torch.tensor(t, dtype=torch.float64)
    loss = tw.mse_esr_loss(yt[:, skip:], tt[:, skip:])  # the reference's argument order: (outs, train_Y)
    loss.backward()
    ybar, l, _, _ = as_called_ybar(y, t, skip)
    np.testing.assert_allclose(yt.grad.numpy(), ybar, rtol=1e-10, atol=1e-16)
    assert abs(float(loss) / l - 1) < 1e-12
    # and it is NOT the textbook form (energy of the target): the two differ in value and in gradient
    other = tw.mse_esr_loss(tt[:, skip:], yt[:, skip:])
    assert abs(float(other) / l - 1) > 1e-4


def _clipper(dwdf, p, mode, ordering, swapped=False):
    Vs = dwdf.ResistiveVoltageSource(p.R, True)
    C = dwdf.Capacitor(p.C, p.fs, True)
    P1 = dwdf.Parallel(C, Vs) if swapped else dwdf.Parallel(Vs, C)
    dp = dwdf.DiodePair(P1, p.Is, p.Vt, p.nabla, trainable=True, mode=mode)
    circ = dwdf.compile_circuit(dp, probe=C, ordering=ordering)
    return circ, [circ.slot(dp, "Is"), circ.slot(dp, "nabla"), circ.slot(Vs, "R"), circ.slot(C, "C")]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["approx", "exact"])
@pytest.mark.parametrize("ordering,oord", [("plugin", ORDER_PLUGIN), ("python", ORDER_PYTHON)])
@pytest.mark.parametrize("swapped", [False, True], ids=["clipper_kernels", "tree_interpreter"])
def test_analytic_root_against_oracle(dwdf, oracle, mode, ordering, oord, swapped):
    p = ClipperParams()
    B, T, skip = 48, 600, 50
    x = make_inputs(B, T, seed=11)
    target = oracle.clipper_forward(x, ClipperParams(R=p.R * 1.1, C=p.C * 0.9, Is=p.Is * 2, nabla=p.nabla * 1.05), exact=True, ordering=oord)
    circ, order = _clipper(dwdf, p, mode, ordering, swapped)
    assert circ.is_clipper != swapped
    circ.forward(torch.from_numpy(x).cuda())
    res = circ.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr_as_called", skip=skip)
    y_ref = oracle.clipper_forward(x, p, exact=(mode == "exact"), ordering=oord, dtype=np.float64)
    ybar, loss, mse, esr = as_called_ybar(y_ref, target, skip)
    ref = oracle.clipper_grad(x.astype(np.float64), ybar, p, exact=(mode == "exact"), ordering=oord, mode="upstream", dtype=np.float64)
    g = res["grads"].cpu().numpy()[order]
    assert np.max(np.abs(g / ref["grads"] - 1.0)) < 5e-4, (g, ref["grads"])
    assert abs(float(res["loss"]) / loss - 1) < 1e-5 and abs(float(res["mse"]) / mse - 1) < 1e-5 and abs(float(res["esr"]) / esr - 1) < 1e-5
    # the textbook form is a different number: the two kinds are not aliases
    circ.forward(torch.from_numpy(x).cuda())
    other = circ.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr", skip=skip)
    assert abs(float(other["esr"]) / esr - 1) > 1e-4


@pytest.mark.gpu
def test_train_step_and_unsupported_entry_points(dwdf, oracle):
    p = ClipperParams()
    x = torch.from_numpy(make_inputs(64, 512, seed=4)).cuda()
    target = (0.7 * torch.tanh(3.0 * x)).contiguous()
    circ, _ = _clipper(dwdf, p, "approx", "python")
    circ.forward(x)
    want = circ.backward(target=target, loss="mse+esr_as_called", skip=50)["out"].clone()
    circ2, _ = _clipper(dwdf, p, "approx", "python")
    opt = dwdf.Adam(circ2, lr=0.0)  # rate 0: the step leaves the parameters alone, so its result block is comparable
    got = circ2.train_step(x, target, opt, loss="mse+esr_as_called", skip=50)["out"]
    assert torch.allclose(got[:19], want[:19], rtol=1e-9, atol=0)
    with pytest.raises(Exception, match="AS_CALLED"):
        circ.train_pass(x, target, loss="mse+esr_as_called")
    with pytest.raises(ValueError):
        circ.backward(target=target, loss="esr")  # unknown names are refused, not silently MSE


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2x8", "2x16"])
def test_neural_root_composition(dwdf, name):
    """Neural root: the composed loss equals the upstream-mode sweep fed with the closed-form dL/dy of the kernel's own output
    (the upstream mode itself is pinned to fp64 autograd in tests/test_nn_root.py)."""
    import os

    from conftest import GOLDEN
    from test_nn_root import make_circuit

    nnv = np.load(os.path.join(GOLDEN, "nn_vectors.npz"))
    mj = dwdf.model_io.json_from_weights(nnv[f"{name}_weights"], [int(v) for v in nnv[f"{name}_sizes"]])
    circ = make_circuit(dwdf, mj, "python")
    x = torch.from_numpy(make_inputs(40, 300, seed=9)).cuda()
    target = (0.5 * torch.tanh(2.0 * x)).contiguous()
    y = circ.forward(x)
    res = circ.backward(target=target, loss="mse+esr_as_called", skip=50)
    g = res["grads"].clone()
    loss = float(res["loss"])
    ybar, l, _, _ = as_called_ybar(y.cpu().numpy(), target.cpu().numpy(), 50)
    circ.forward(x)
    g_up = circ.backward(gy=torch.from_numpy(ybar.astype(np.float32)).cuda())["grads"]
    scale = float(g_up.abs().max())
    assert float((g - g_up).abs().max()) < 2e-5 * scale
    assert abs(loss / l - 1) < 1e-5
