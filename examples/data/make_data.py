"""Builds the small data set the examples fall back to where /root/reference does not exist (the GPU boxes), from the
reference's own files, in the build container:

  1N4148_1up1down_excerpt.npz   the first 131072 samples (64 windows of 2048) of the trimmed recording of each of the five
                                CSVs in diode_dataset/1N4148/1up1down (float32 (N, 2) = input, output volts; key = file stem)
  pretrained_2x16.json          wdf_py/diode_clipper/models/pretrained/'1N4148 (1U-1D)_2x16_pretrained_model.json'
  reference_trained_2x16.json   wdf_py/diode_clipper/models/'1N4148 (1U-1D)_2x16_training_2000.json' (the plugin's 2x16 model)

    python examples/data/make_data.py
"""
import importlib
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

if __name__ == "__main__":
    wdf = importlib.import_module("differentiable-wdfs_b200")
    folder = wdf.dataimport.get_data_path_for_diode(wdf.diode_config.diode_1n4148_1u1d, REF)
    out = {}
    for name in sorted(os.listdir(folder)):
        if name.endswith(".csv"):
            d = wdf.dataimport.create_dataset(os.path.join(folder, name))
            out[name[:-4]] = d["dataset"][:131072].astype(np.float32)
            fs = d["FS"]
    np.savez_compressed(os.path.join(HERE, "1N4148_1up1down_excerpt.npz"), FS=np.float64(fs), **out)
    models = os.path.join(REF, "wdf_py/diode_clipper/models")
    shutil.copyfile(os.path.join(models, "pretrained", "1N4148 (1U-1D)_2x16_pretrained_model.json"), os.path.join(HERE, "pretrained_2x16.json"))
    shutil.copyfile(os.path.join(models, "1N4148 (1U-1D)_2x16_training_2000.json"), os.path.join(HERE, "reference_trained_2x16.json"))
    print({k: v.shape for k, v in out.items()}, fs)
