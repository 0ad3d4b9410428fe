#!/usr/bin/env python
"""Generates tests/golden/nn_vectors.npz — golden vectors for the NEURAL diode-pair root — by running
the reference itself in this container (needs /root/reference and oracle/_ref/libdwdf_ref_nn.so):

  * the reference's own trained weight files (the five 1N4148 1U-1D models the plugin embeds,
    plugin/src/CMakeLists.txt:20-24) are read as they are; their weights are stored flattened so the
    tests can run where /root/reference does not exist;
  * the bare network (RTNeural ModelT::forward, unmodified) on a grid of (a, log R);
  * the plugin clipper with that root (tree DiodeClipperWDF.h:18-25, loop DiodeClipperWDF.cpp:22-29, root
    DiodePairNeuralModel.h:62-75) on seeded inputs, both probe orderings.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
MODELS = {  # name -> (file, n_hidden_layers, hidden size): plugin/src/CMakeLists.txt:20-24
    "2x4": ("1N4148 (1U-1D)_2x4_training_3.json", 2, 4),
    "2x8": ("1N4148 (1U-1D)_2x8_training_3.json", 2, 8),
    "2x16": ("1N4148 (1U-1D)_2x16_training_2000.json", 2, 16),
    "4x4": ("1N4148 (1U-1D)_4x4_training_1.json", 4, 4),
    "4x8": ("1N4148 (1U-1D)_4x8_training_500.json", 4, 8),
}

from make_golden import make_inputs  # noqa: E402


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdwdf_ref_nn.so"))
    out = {}
    fs, R, Cv = 48000.0, 47000.0, 2.2e-9  # DiodeClipperWDF.h:18-20
    x = make_inputs(6, 512, fs, 4321)
    out["x"] = x
    a = np.tile(np.linspace(-2.5, 2.5, 201, dtype=np.float32), 3)
    logR = np.repeat(np.log(np.array([1.0e3, 4301.5083, 4.0e4], np.float32)), 201).astype(np.float32)
    out["grid_a"], out["grid_logR"] = a, logR
    for name, (fname, nl, h) in MODELS.items():
        path = os.path.join(REF, "wdf_py", "diode_clipper", "models", fname)
        m = json.load(open(path))
        flat = []
        for layer in m["layers"]:  # kernel (in x out, row-major) then bias, layer after layer
            flat += [np.asarray(layer["weights"][0], np.float32).ravel(), np.asarray(layer["weights"][1], np.float32).ravel()]
        out[f"{name}_weights"] = np.concatenate(flat)
        out[f"{name}_sizes"] = np.array([m["in_shape"][-1]] + [layer["shape"][-1] for layer in m["layers"]], np.int32)
        o = np.empty_like(a)
        assert lib.ref_nn_eval(path.encode(), nl, h, P(a), P(logR), P(o), C.c_int64(a.size)) == 0
        out[f"{name}_grid_out"] = o
        for oname, order in (("plugin", 0), ("python", 1)):
            y = np.empty_like(x)
            assert lib.ref_nn_clipper(path.encode(), nl, h, P(x), P(y), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), C.c_float(fs), C.c_float(R), C.c_float(Cv), order) == 0
            out[f"{name}_clip_{oname}"] = y
    np.savez_compressed(os.path.join(HERE, "nn_vectors.npz"), **out)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "nn_vectors.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
