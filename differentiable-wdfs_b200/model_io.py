"""RTNeural-style JSON weight files: the hand-off format between training and the plugin.

Format (wdf_py/lib/model_utils.py:17-85, clipper_pot.py:298-331): ``{"in_shape": [None, 2], "layers":
[{"type": "dense", "activation": "tanh" | "", "shape": [None, out], "weights": [kernel (in x out), bias]}]}``.
Pre-trained files written by Keras nest the kernel one level deeper for the un-shared DenseLayer
(layers.py:29-35: ``kernel = [weights]``); both nestings are accepted.
"""
from __future__ import annotations

import json
from typing import List, Sequence, Tuple

import numpy as np


def _squeeze_to(a, ndim):
    a = np.asarray(a, np.float32)
    while a.ndim > ndim and a.shape[0] == 1:
        a = a[0]
    return a


def _last_dim(shape) -> int:
    """The feature width of a Keras shape as the reference's files spell it: [None, 2], [[None, 2]] or a bare 2."""
    while isinstance(shape, (list, tuple)):
        if len(shape) == 0:
            raise ValueError("empty shape")
        shape = shape[-1]
    return int(shape)


def layers_from_json(model_json: dict) -> Tuple[List[Tuple[np.ndarray, np.ndarray, str]], List[int]]:
    """-> ([(kernel (in, out), bias (out,), activation)], sizes [in, h1, ..., out])."""
    sizes = [_last_dim(model_json["in_shape"])]
    layers = []
    for layer in model_json["layers"]:
        if layer["type"] != "dense":
            # layers.py:57 builds the model from the dense entries only; the Keras-written pretrained files
            # (models/pretrained/*.json) open with the InputLayer as {"type": "unknown", "weights": []}.
            # A non-dense entry that carries weights (gru, lstm, conv1d ...) would change the function: refuse it.
            if len(layer.get("weights", [])) == 0:
                continue
            raise ValueError(f"layer type {layer['type']!r}: only dense layers make a WDF root model (layers.py:57-70)")
        W = _squeeze_to(layer["weights"][0], 2)
        b = _squeeze_to(layer["weights"][1], 1)
        if W.shape != (sizes[-1], b.shape[0]):
            raise ValueError(f"dense kernel of shape {W.shape} after a layer of width {sizes[-1]} with a bias of {b.shape[0]}")
        layers.append((W, b, layer.get("activation", "")))
        sizes.append(int(b.shape[0]))
    return layers, sizes


def flatten_weights(layers: Sequence[Tuple[np.ndarray, np.ndarray, str]]) -> np.ndarray:
    """The weight vector of include/dwdf.h (dwdf_mlp_desc): kernel row-major then bias, layer after layer."""
    return np.concatenate([np.concatenate([np.asarray(W, np.float32).ravel(), np.asarray(b, np.float32).ravel()]) for W, b, _ in layers])


def unflatten_weights(weights, sizes: Sequence[int]):
    out, k = [], 0
    w = np.asarray(weights, np.float32)
    for i, o in zip(sizes[:-1], sizes[1:]):
        W = w[k:k + i * o].reshape(i, o)
        k += i * o
        b = w[k:k + o]
        k += o
        out.append((W, b))
    if k != w.size:
        raise ValueError("weight vector does not match the layer sizes")
    return out


def json_from_weights(weights, sizes: Sequence[int]) -> dict:
    """The dict clipper_pot.py:298-321 writes: tanh after every layer but the last."""
    layers = []
    pairs = unflatten_weights(weights, sizes)
    for k, (W, b) in enumerate(pairs):
        layers.append({"type": "dense", "shape": [None, int(b.shape[0])], "weights": [W.tolist(), b.tolist()], "activation": "tanh" if k + 1 < len(pairs) else ""})
    return {"in_shape": [None, int(sizes[0])], "layers": layers}


def load_model_json(path: str) -> dict:
    with open(path) as f:
        return json.load(f)


def save_model_json(model_json: dict, path: str) -> None:
    """clipper_pot.py:324-331 (json.dump(..., indent=4))."""
    with open(path, "w") as f:
        json.dump(model_json, f, indent=4)
