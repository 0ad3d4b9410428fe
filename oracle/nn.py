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
