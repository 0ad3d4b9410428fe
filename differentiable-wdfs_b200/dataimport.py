"""Loader for the measured diode-clipper recordings (wdf_py/lib/dataimport.py).

Digilent WaveForms CSV: 9 comment lines (``#Sample rate: 50000Hz``, ``#Samples: N`` among them), a blank
line, a header row, then ``input, output`` volts per line (diode_dataset/1N4148/1up1down/10.0k_4.7nF.csv:1-12).
Same cuts and split as the reference: drop the first 2.5 s, keep 14.3 s (dataimport.py:35-48); the source
resistance comes from the file name (``10.0k_4.7nF.csv`` -> 10 kOhm, :95); files with 36 k <= R <= 73 k are
the validation set, the rest the training set (:98). Batching as clipper_pot.py:58-80: consecutive windows of
``seq_len`` samples, input channels (x, R), target y.
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, Tuple

import numpy as np

TIME_REMOVE_PRE = 2.5  # s, dataimport.py:35
DUR_OF_DATA = 14.3  # s, dataimport.py:38


def read_header(path: str) -> Tuple[float, int]:
    fs, n = None, None
    with open(path) as f:
        for _ in range(9):
            line = f.readline()
            m = re.search(r"#Sample rate:\s*([0-9.eE+-]+)\s*Hz", line)
            if m:
                fs = float(m.group(1))
            m = re.search(r"#Samples:\s*([0-9]+)", line)
            if m:
                n = int(m.group(1))
    if fs is None or n is None:
        raise ValueError(f"{path}: not a Digilent WaveForms acquisition (no '#Sample rate' / '#Samples' header lines)")
    return fs, n


def create_dataset(path: str) -> Dict[str, object]:
    """dataimport.py:25-59 (createDataset): the trimmed (input, output) recording of one file."""
    fs, _ = read_header(path)
    data = np.loadtxt(path, delimiter=",", skiprows=11, dtype=np.float64)  # 9 comment lines, a blank one, the column header (pandas: header=9 after skipping blanks)
    lo = math.floor(TIME_REMOVE_PRE * fs)
    hi = math.ceil((TIME_REMOVE_PRE + DUR_OF_DATA) * fs)
    data = data[lo:hi, :]
    return {"dataset": data, "FS": fs, "num_samples": len(data)}


def resistance_from_filename(path: str) -> float:
    """dataimport.py:95: '10.0k_4.7nF.csv' -> 10.0 (kOhm)."""
    return float(os.path.basename(path).partition("k")[0])


def data_path_for_diode(n_up: int, n_down: int, base_dir: str, family: str = "1N4148", hpf: bool = False) -> str:
    """dataimport.py:62-79."""
    root = os.path.join(base_dir, "diode_dataset")
    if family == "1N4148":
        root = os.path.join(root, "placeholder_data", "HPF") if hpf else os.path.join(root, "1N4148")
    elif family == "OA1154":
        root = os.path.join(root, "OA1154")
    else:
        raise ValueError("No data available for this diode!")
    return os.path.join(root, f"{n_up}up{n_down}down")


def get_data_path_for_diode(diode, BASE_DIR, HPF2: bool = False) -> str:
    """dataimport.py:60-78 with the reference's own signature: the configuration tuple (diode_config.DiodeConfig) picks the folder."""
    family = "1N4148" if "1N4148" in diode.name else "OA1154" if "OA1154" in diode.name else None
    if family is None:
        raise ValueError("No data available for this diode!")
    return data_path_for_diode(int(diode.N_up), int(diode.N_down), str(BASE_DIR), family, hpf=bool(HPF2))


createDataset = create_dataset  # the reference's spelling (dataimport.py:25)


def load_diode_data(data_path, BASE_DIR=None, start_offset: int = 0, csv_samples: int = -1, plot: bool = False, HPF: bool = False):
    """dataimport.py:82-137: -> (train (3, N_train), N_train, val (3, N_val), N_val, FS), rows = (x, R, y_ref).

    Called like the reference, ``load_diode_data(diode, BASE_DIR, start_offset, csv_samples)``, or with the folder of
    one configuration, ``load_diode_data(path, start_offset=..., csv_samples=...)``. (``plot`` is accepted and ignored.)"""
    if hasattr(data_path, "name") and hasattr(data_path, "N_up"):
        if BASE_DIR is None:
            raise ValueError("load_diode_data(diode, BASE_DIR): the dataset's base directory is missing")
        data_path = get_data_path_for_diode(data_path, BASE_DIR, HPF2=HPF)
    elif BASE_DIR is not None and not isinstance(BASE_DIR, (str, os.PathLike)):
        start_offset, BASE_DIR = int(BASE_DIR), None  # load_diode_data(path, start_offset, ...) of this package's first version
    data_path = str(data_path)
    train, val, fs = [], [], 0.0
    for name in sorted(os.listdir(data_path)):
        if not name.endswith(".csv"):
            continue
        path = os.path.join(data_path, name)
        r_val = resistance_from_filename(path)
        raw = create_dataset(path)
        fs = raw["FS"]
        n = raw["num_samples"] if csv_samples < 0 else csv_samples
        d = raw["dataset"]
        x = d[start_offset:start_offset + n, 0].astype(np.float32)
        y = d[start_offset:start_offset + n, 1].astype(np.float32)
        block = np.stack([x, np.ones_like(x) * np.float32(r_val * 1000.0), y])
        (train if (r_val < 36 or r_val > 73) else val).append(block)
    tr = np.concatenate(train, axis=1) if train else np.zeros((3, 0), np.float32)
    va = np.concatenate(val, axis=1) if val else np.zeros((3, 0), np.float32)
    return tr, tr.shape[1], va, va.shape[1], fs


def batch_data(data: np.ndarray, seq_len: int = 2048):
    """clipper_pot.py:61-80: (3, N) rows (x, R, y) -> X (n_seq, seq_len, 2) = [x, R], Y (n_seq, seq_len, 1)."""
    n_seq = data.shape[1] // seq_len
    d = data[:, :n_seq * seq_len]
    X = np.stack([d[0].reshape(n_seq, seq_len), d[1].reshape(n_seq, seq_len)], axis=-1).astype(np.float32)
    Y = d[2].reshape(n_seq, seq_len, 1).astype(np.float32)
    return X, Y
