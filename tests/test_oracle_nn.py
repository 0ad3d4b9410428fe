"""Pins the numpy restatement of the neural diode-pair root (oracle/nn.py) against golden vectors
produced by the reference's own code (unmodified RTNeural + chowdsp_wdf, tests/golden/make_golden_nn.py)
on the reference's own trained weight files (the five 1N4148 1U-1D models the plugin embeds)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, seq_rel_err
from oracle import nn

MODELS = ["2x4", "2x8", "2x16", "4x4", "4x8"]
# Tolerances: the trained networks add up terms of magnitude 1..10 that cancel to an output of ~0.1..1, so
# two fp32 evaluations that only differ in summation order (RTNeural's loops vs a matrix product) already
# differ by ~1.5e-6 absolute: the reference does not agree with ITSELF (xsimd vs STL backend) any better.
NET_ATOL = 5e-6  # bare network, outputs O(1)
CLIP_TOL = 5e-5  # max|y - y_ref| / max|y_ref| per sequence (sequences as quiet as 0.1 V peak)


@pytest.fixture(scope="module")
def nnv():
    return np.load(os.path.join(GOLDEN, "nn_vectors.npz"))


@pytest.mark.parametrize("name", MODELS)
def test_bare_network(nnv, name):
    out = nn.mlp_eval(nnv[f"{name}_weights"], nnv[f"{name}_sizes"], nnv["grid_a"], nnv["grid_logR"])
    assert np.max(np.abs(out - nnv[f"{name}_grid_out"])) < NET_ATOL


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("oname,order", [("plugin", nn.ORDER_PLUGIN), ("python", nn.ORDER_PYTHON)])
def test_clipper_with_neural_root(nnv, name, oname, order):
    y = nn.nn_clipper_forward(nnv["x"], nnv[f"{name}_weights"], nnv[f"{name}_sizes"], 48000.0, 47000.0, 2.2e-9, order)
    assert seq_rel_err(y, nnv[f"{name}_clip_{oname}"]) < CLIP_TOL
    y64 = nn.nn_clipper_forward(nnv["x"], nnv[f"{name}_weights"], nnv[f"{name}_sizes"], 48000.0, 47000.0, 2.2e-9, order, dtype=np.float64)
    assert seq_rel_err(y64, nnv[f"{name}_clip_{oname}"]) < CLIP_TOL


def test_sizes_match_the_plugin_types(nnv):
    """DiodePairNeuralModel.h:5-41: 2xH = 2 -> H -> H -> H -> 1, 4xH = 2 -> H x5 -> 1."""
    for name in MODELS:
        n, h = (int(v) for v in name.split("x"))
        assert list(nnv[f"{name}_sizes"]) == [2] + [h] * (n + 1) + [1]
