"""The ported training scripts run end to end (examples/): clipper_pot_b200.py fine-tunes the reference's pretrained 2x16
network on its measured 1N4148 recordings (the excerpt under examples/data where /root/reference is absent) and evaluates the
reference's own trained JSON the same way; lpf_b200.py recovers the RC low-pass's cutoff as lpf.py does."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def run(script, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout


def test_clipper_pot_port_trains(tmp_path):
    rep, out = tmp_path / "report.json", tmp_path / "model.json"
    run("clipper_pot_b200.py", "--epochs", "40", "--json", str(rep), "--out", str(out), "--csv-samples", "131072")
    r = json.load(open(rep))
    print(json.dumps(r))
    assert r["train_loss_first_last"][1] < r["train_loss_first_last"][0]  # training lowers the training loss
    assert r["fine_tuned_here"]["val_loss"] < r["pretrained"]["val_loss"]  # ... and the validation loss of the pretrained model
    # the reference's own trained model of this shape (the plugin's embedded 2x16), loaded from its JSON and run through the same
    # kernels on the same validation windows: reported next to ours, whatever it scores (measured: val MSE 1.6e-3 against 1.8e-4
    # for the pretrained network and 2.3e-5 after 40 epochs here — the repository does not say which data that file was trained on)
    assert 0.0 < r["reference_trained_json"]["val_mse"] < 0.05
    m = json.load(open(out))
    assert [l["type"] for l in m["layers"]] == ["dense"] * 4 and m["in_shape"][-1] == 2


def test_lpf_port_finds_the_cutoff(tmp_path):
    rep = tmp_path / "report.json"
    run("lpf_b200.py", "--epochs", "100", "--json", str(rep))
    r = json.load(open(rep))
    assert r["loss_first_last"][1] < 0.05 * r["loss_first_last"][0]
    assert abs(r["cutoff_hz"] - r["target_cutoff_hz"]) < abs(r["start_cutoff_hz"] - r["target_cutoff_hz"])  # moved from 159 Hz towards 720 Hz
