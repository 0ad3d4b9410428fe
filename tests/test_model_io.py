"""File formats on the way in and out of the engine (no GPU): the RTNeural/Keras JSON weight files and the Digilent
CSV recordings, on the reference's OWN files — tests/golden/pretrained_2x4.json and digilent_excerpt.csv are excerpts
of them (tests/golden/make_golden_io.py), and the full set under /root/reference is swept where it exists.
"""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

REF = "/root/reference"
has_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")


def test_keras_pretrained_layout_loads(dwdf):
    """The pretrained files open with the Keras InputLayer ({"type": "unknown", "weights": []}); layers.py:57 skips every
    non-dense entry, so must we — INTEGRATION.md §1's first lines run on these files."""
    mj = dwdf.model_io.load_model_json(os.path.join(GOLDEN, "pretrained_2x4.json"))
    assert mj["layers"][0]["type"] == "unknown" and mj["layers"][0]["weights"] == []
    layers, sizes = dwdf.model_io.layers_from_json(mj)
    assert sizes == [2, 4, 4, 4, 1]
    assert [a for _, _, a in layers] == ["tanh", "tanh", "tanh", ""]
    root = dwdf.DenseRootModel(mj)
    assert root.sizes == [2, 4, 4, 4, 1]
    w = root.weight_vector()
    assert w.size == 2 * 4 + 4 + 2 * (4 * 4 + 4) + 4 + 1
    # the kernel of the first dense layer sits in the JSON as (in, out) rows
    np.testing.assert_array_equal(w[:8].reshape(2, 4), np.asarray(mj["layers"][1]["weights"][0], np.float32))
    # imperative evaluation = plain numpy
    import torch

    x = torch.tensor([[0.3, 8.1], [-1.2, 9.0]])
    root.incident(x)
    h = x.numpy()
    for W, b, act in layers:
        h = h @ W + b
        if act == "tanh":
            h = np.tanh(h)
    np.testing.assert_allclose(root.reflected().numpy(), h, rtol=1e-6, atol=1e-7)


def test_layers_with_weights_that_are_not_dense_are_refused(dwdf):
    mj = dwdf.model_io.load_model_json(os.path.join(GOLDEN, "pretrained_2x4.json"))
    bad = json.loads(json.dumps(mj))
    bad["layers"].insert(1, {"type": "gru", "activation": "", "shape": [None, 4], "weights": [[[0.0]], [[0.0]], [0.0]]})
    with pytest.raises(ValueError, match="gru"):
        dwdf.model_io.layers_from_json(bad)


@has_ref
def test_every_reference_model_file_loads(dwdf):
    paths = [p for p in sorted(glob.glob(os.path.join(REF, "wdf_py/diode_clipper/models/**/*.json"), recursive=True)) if os.sep + "old" + os.sep not in p]  # models/old: an earlier one-input experiment
    assert len(paths) > 40
    seen = set()
    for p in paths:
        mj = dwdf.model_io.load_model_json(p)
        root = dwdf.DenseRootModel(mj)
        s = root.sizes
        assert s[0] == 2 and s[-1] == 1 and len(set(s[1:-1])) == 1, (p, s)
        seen.add((len(s) - 3, s[1]))
    assert {(2, 4), (2, 8), (2, 16), (4, 4), (4, 8)} <= seen  # the plugin's five network shapes


def test_digilent_csv_excerpt(dwdf, tmp_path):
    """Header parsing and column order on an excerpt of the reference's own recording."""
    p = os.path.join(GOLDEN, "digilent_excerpt.csv")
    fs, n = dwdf.dataimport.read_header(p)
    assert fs == 50000.0 and n == 946909
    data = np.loadtxt(p, delimiter=",", skiprows=11)
    assert data.shape == (400, 2)
    assert dwdf.dataimport.resistance_from_filename("/x/10.0k_4.7nF.csv") == 10.0
    assert dwdf.dataimport.resistance_from_filename("99.1k_4.7nF.csv") == 99.1


@has_ref
def test_reference_dataset_split_and_batches(dwdf):
    """dataimport.py:82-137 on the real 1N4148 1up1down recordings: 5 files x 715000 samples, 45.2k is the validation file."""
    path = dwdf.dataimport.data_path_for_diode(1, 1, REF)
    tr, n_tr, va, n_va, fs = dwdf.dataimport.load_diode_data(path, csv_samples=20000)
    assert fs == 50000.0 and n_tr == 4 * 20000 and n_va == 20000
    assert sorted(set(np.unique(tr[1]).tolist())) == [10000.0, 25200.0, 75000.0, 99100.0]
    assert np.unique(va[1]).tolist() == [45200.0]
    X, Y = dwdf.dataimport.batch_data(tr, 2048)
    assert X.shape == (39, 2048, 2) and Y.shape == (39, 2048, 1)
    one = dwdf.dataimport.create_dataset(os.path.join(path, "10.0k_4.7nF.csv"))
    assert one["num_samples"] == 715000
