#!/usr/bin/env python
"""Golden vectors from the reference's OWN Python code (run in this container; writes tests/golden/py_reference_vectors.npz).

TensorFlow cannot be installed here, so the unmodified reference sources are executed on `oracle/shim_tf` — a minimal
`tensorflow` look-alike backed by torch (eager ops + autograd behind GradientTape). What runs is the reference's text:
  * wdf_py/lib/tf_wdf.py and wdf_py/lib/layers.py, imported as they are;
  * `class Model` of wdf_py/simple_circuits/lpf.py:20-49 and `class ClipperModel`, `esr_loss`, `loss_func` of
    wdf_py/diode_clipper/clipper_pot.py:94-127,141-177 — cut out of the scripts with `ast` (the scripts themselves load data and train
    at import) and exec'ed unchanged.
Recorded, in float32 (the reference's precision) and in float64 (the same sources with every tf.float32 meaning float64: the gradient
oracle): outputs, the loss exactly as each script's training loop calls it (lpf.py:88-89, clipper_pot.py:247-248) and
tape.gradient(loss, trainable variables). This pins forward AND gradients of the tree path and of the neural root to the reference.
"""
import ast
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/wdf_py"
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim_tf"))
sys.path.insert(0, os.path.join(REF, "lib"))

import tensorflow as tf  # the shim  # noqa: E402
import tf_wdf as wdf  # the reference, unmodified  # noqa: E402
from layers import DenseRootModel  # the reference, unmodified  # noqa: E402


def cut(path, names):
    """Source of the named top-level classes / functions / assignments of a reference script, in file order."""
    src = open(path).read()
    out = []
    for node in ast.parse(src).body:
        key = getattr(node, "name", None)
        if key is None and isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            key = node.targets[0].id
        if key in names:
            out.append(ast.get_source_segment(src, node))
    assert len(out) == len(names), (path, names, len(out))
    return "\n\n".join(out)


def sine_inputs(B, T, fs, seed, amp=(0.1, 2.0)):
    rng = np.random.default_rng(seed)
    n = np.arange(T)
    A = rng.uniform(amp[0], amp[1], B)
    f = np.exp(rng.uniform(np.log(50.0), np.log(5000.0), B))
    return (A[:, None] * np.sin(2 * np.pi * f[:, None] * n[None, :] / fs) + 0.05 * rng.standard_normal((B, T))).astype(np.float32)


out = {}

# ---- lpf.py ------------------------------------------------------------------------------------------------------------
lpf_src = cut(os.path.join(REF, "simple_circuits", "lpf.py"), ["Model"])
FS_LPF = 48000
B, T = 3, 400
x = sine_inputs(B, T, FS_LPF, 1)
target = (0.5 * np.roll(x, 2, axis=1)).astype(np.float32)
out["lpf_x"], out["lpf_target"] = x, target
for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
    tf.set_dtype(dt)
    ns = {"tf": tf, "wdf": wdf, "FS": FS_LPF, "np": np}
    exec(lpf_src, ns)
    model = ns["Model"]()
    loss_func = tf.keras.losses.MeanSquaredError()
    with tf.GradientTape() as tape:
        outs = model.forward(x)[..., 0]  # lpf.py:88: (T, B)
        loss = loss_func(outs, torch.as_tensor(target.T.copy(), dtype=dt))  # lpf.py:89
    gR, gC = tape.gradient(loss, [model.R1.R, model.C1.C])
    out[f"lpf_y_{tag}"] = outs.detach().numpy().T.copy()
    out[f"lpf_loss_{tag}"] = float(loss)
    out[f"lpf_grad_RC_{tag}"] = np.array([float(gR), float(gC)])

# ---- clipper_pot.py ----------------------------------------------------------------------------------------------------------
clip_src = cut(os.path.join(REF, "diode_clipper", "clipper_pot.py"), ["ClipperModel", "pre_emphasis_filter", "eps", "esr_loss", "mse_loss", "loss_func"])
FS_CLIP, C_VAL, SKIP = 50000.0, 4.7e-9, 50  # clipper_pot.py:57 (C_val), :232 (skip_samples); the recordings' rate
for name in ("2x8", "2x16"):
    path = os.path.join(REF, "diode_clipper", "models", "pretrained", f"1N4148 (1U-1D)_{name}_pretrained_model.json")
    model_json = json.load(open(path))
    B, T = 4, 260
    x = sine_inputs(B, T, FS_CLIP, 7, amp=(0.2, 1.5))
    r = np.full((B, T), 10000.0, np.float32)
    r[1], r[2], r[3, T // 2:] = 47000.0, 22000.0, 100000.0
    inp = np.stack([x, r], axis=-1)  # (B, T, 2): clipper_pot.py:61-69
    Y = (0.6 * np.tanh(2.0 * x))[..., None].astype(np.float32)  # (B, T, 1)
    out[f"clip_{name}_x"], out[f"clip_{name}_r"], out[f"clip_{name}_target"] = x, r, Y[..., 0]
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        tf.set_dtype(dt)
        ns = {"tf": tf, "wdf": wdf, "DenseRootModel": DenseRootModel, "FS": FS_CLIP, "C_val": C_VAL, "np": np}
        exec(clip_src, ns)
        model = ns["ClipperModel"](model_json)
        Yt = torch.as_tensor(Y, dtype=dt)
        with tf.GradientTape() as tape:
            outs = tf.transpose(model.forward(inp)[..., 0], perm=[1, 0, 2])  # clipper_pot.py:247: (B, T, 1)
            loss = ns["loss_func"](outs[:, SKIP:, :], Yt[:, SKIP:, :])  # clipper_pot.py:248 — (outs, train_Y): the call as written
        layers = [l for l in model.model.layers if hasattr(l, "kernel")]
        variables = [v for l in layers for v in (l.kernel, l.bias)]
        grads = tape.gradient(loss, variables)
        out[f"clip_{name}_y_{tag}"] = outs.detach().numpy()[..., 0]
        out[f"clip_{name}_loss_{tag}"] = float(loss)
        out[f"clip_{name}_mse_{tag}"] = float(ns["mse_loss"](outs[:, SKIP:, :], Yt[:, SKIP:, :]))
        out[f"clip_{name}_esr_{tag}"] = float(ns["esr_loss"](outs[:, SKIP:, :], Yt[:, SKIP:, :]))
        out[f"clip_{name}_grad_w_{tag}"] = np.concatenate([g.detach().numpy().reshape(-1) for g in grads])
        if tag == "f64":
            out[f"clip_{name}_weights"] = np.concatenate([v.detach().numpy().reshape(-1) for v in variables]).astype(np.float32)
            out[f"clip_{name}_sizes"] = np.array([model_json["in_shape"][-1]] + [l.kernel.shape[-1] for l in layers], np.int32)
out["clip_fs"], out["clip_C"], out["clip_skip"] = FS_CLIP, C_VAL, SKIP
tf.set_dtype(torch.float32)
np.savez_compressed(os.path.join(HERE, "py_reference_vectors.npz"), **out)
for k in sorted(out):
    v = np.asarray(out[k])
    print(k, v.shape, v.dtype, (float(v.reshape(-1)[0]) if v.size else None))
