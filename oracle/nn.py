"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the reference's NEURAL diode-pair root.

Never imported by the product. Follows
  * wdf_py/lib/layers.py:7-39 (DenseLayer: x @ kernel + bias), :42-82 (DenseRootModel: dense layers with
    tanh between them, input (a, log R), output b_nn);
  * wdf_py/diode_clipper/clipper_pot.py:113-124 (per-sample loop: Parallel.reflected -> concat with
    log(P1.R) -> model -> Parallel.incident(-b_nn) -> voltage(C)), optional per-sample resistance channel
    (:116-117);
  * plugin/src/dsp/diode_clipper/DiodePairNeuralModel.h:66-75 (logR = log(next.R); b = -model(a, logR)).
Pinned by tests/test_oracle_nn.py against tests/golden/nn_vectors.npz, which holds outputs of the
reference's own RTNeural + chowdsp_wdf code run on the reference's own trained weight files
(tests/golden/make_golden_nn.py).
"""
from __future__ import annotations

import numpy as np

ORDER_PLUGIN, ORDER_PYTHON = 0, 1


def split_weights(weights, sizes):
    """Flat float vector (kernel in x out row-major, then bias, layer after layer) -> [(W, b), ...]."""
    out, k = [], 0
    for i, o in zip(sizes[:-1], sizes[1:]):
        W = np.asarray(weights[k:k + i * o]).reshape(i, o)
        k += i * o
        b = np.asarray(weights[k:k + o])
        k += o
        out.append((W, b))
    assert k == len(weights), "weight vector does not match the layer sizes"
    return out


def mlp_eval(weights, sizes, a, logR, dtype=np.float32):
    """The bare network (layers.py:72-82): tanh after every layer but the last. Returns model(a, logR)."""
    h = np.stack([np.asarray(a, dtype), np.asarray(logR, dtype) * np.ones_like(np.asarray(a, dtype))], axis=-1)
    layers = split_weights(np.asarray(weights, dtype), list(sizes))
    for k, (W, b) in enumerate(layers):
        h = (h @ W + b).astype(dtype)
        if k + 1 < len(layers):
            h = np.tanh(h).astype(dtype)
    return h[..., 0]


def nn_clipper_forward(x, weights, sizes, fs, R, C, ordering=ORDER_PYTHON, r=None, dtype=np.float32):
    """Parallel(ResistiveVoltageSource R, Capacitor C) closed by the neural root, from reset state.
    r: optional (B, T) per-sample source resistance (clipper_pot.py:116)."""
    x = np.asarray(x, dtype)
    B, T = x.shape
    z = np.zeros(B, dtype)
    y = np.empty((B, T), dtype)
    Gc = dtype(2.0) * dtype(C) * dtype(fs)  # 1 / (1 / (2 C fs)), tf_wdf.py:114-115
    for n in range(T):
        Rv = dtype(R) if r is None else np.asarray(r[:, n], dtype)
        Gv = dtype(1.0) / Rv  # tf_wdf.py:168-177
        G = Gv + Gc
        Rp = dtype(1.0) / G
        gamma = Gv / G
        b_diff = z - x[:, n]  # tf_wdf.py:185-192
        b_temp = -gamma * b_diff
        a = z + b_temp
        b = -mlp_eval(weights, sizes, a, np.log(Rp).astype(dtype), dtype)
        zn = (b + b_temp).astype(dtype)  # tf_wdf.py:179-183 -> Capacitor.incident
        y[:, n] = z if ordering == ORDER_PLUGIN else dtype(0.5) * (zn + z)
        z = zn
    return y


def nn_clipper_grad_torch(x, target, weights, sizes, fs, R, C, ordering=ORDER_PYTHON, r=None, loss="mse", skip=0, gy=None, want_gx=False):
    """Gradient oracle: the same recurrence in fp64 PyTorch, differentiated by autograd (what tf.GradientTape
    does for clipper_pot.py:246-269; trainable variables = the network's kernels and biases, :268).
    Loss: MSE [+ ESR, clipper_pot.py:148-156,177] on samples >= skip (:232,248), or sum(gy * y) when gy is given.
    Returns dict(y, loss, mse, esr, grad_w (flat, same layout as weights)[, gx = dL/dx])."""
    import torch

    dt = torch.float64
    xt = torch.tensor(np.asarray(x, np.float64), dtype=dt, requires_grad=bool(want_gx))
    B, T = xt.shape
    w = torch.tensor(np.asarray(weights, np.float64), dtype=dt, requires_grad=True)
    layers, k = [], 0
    for i, o in zip(sizes[:-1], sizes[1:]):
        W = w[k:k + i * o].reshape(int(i), int(o))
        k += i * o
        b = w[k:k + o]
        k += o
        layers.append((W, b))
    Gc = 2.0 * C * fs
    z = torch.zeros(B, dtype=dt)
    ys = []
    rt = None if r is None else torch.as_tensor(np.asarray(r), dtype=dt)
    for n in range(T):
        Rv = torch.full((B,), float(R), dtype=dt) if rt is None else rt[:, n]
        Gv = 1.0 / Rv
        G = Gv + Gc
        Rp = 1.0 / G
        gamma = Gv / G
        t = gamma * (xt[:, n] - z)
        a = z + t
        h = torch.stack([a, torch.log(Rp)], dim=-1)
        for li, (W, b) in enumerate(layers):
            h = h @ W + b
            if li + 1 < len(layers):
                h = torch.tanh(h)
        zn = -h[:, 0] + t
        ys.append(z if ordering == ORDER_PLUGIN else 0.5 * (zn + z))
        z = zn
    y = torch.stack(ys, dim=1)
    out = {"y": y.detach().numpy()}
    if gy is not None:
        L = (torch.as_tensor(np.asarray(gy), dtype=dt) * y).sum()
        out.update(loss=float(L.detach()), mse=0.0, esr=0.0)
    else:
        tt = torch.as_tensor(np.asarray(target), dtype=dt)[:, skip:]
        e = y[:, skip:] - tt
        N = e.numel()
        mse = (e * e).sum() / N
        L = mse
        esr = torch.zeros((), dtype=dt)
        if loss == "mse+esr":
            esr = torch.sqrt((e * e).sum() / ((tt * tt).sum() + 2.220446049250313e-16) / N)  # clipper_pot.py:145,148-156
            L = mse + esr
        elif loss == "mse+esr_as_called":  # loss_func(outs, train_Y), clipper_pot.py:248: the energy is the model output's
            yy = y[:, skip:]
            esr = torch.sqrt((e * e).sum() / ((yy * yy).sum() + 2.220446049250313e-16) / N)
            L = mse + esr
        out.update(loss=float(L.detach()), mse=float(mse.detach()), esr=float(esr.detach()))
    L.backward()
    out["grad_w"] = w.grad.numpy().copy()
    if want_gx:
        out["gx"] = xt.grad.numpy().copy()
    return out
