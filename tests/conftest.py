"""pytest configuration: registers the ``gpu`` marker and shared fixtures.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol export (no compute calls).
`-m gpu`:       parity tests proper — the CUDA path called through the C-ABI, checked by the oracle.
Only tests (and smoke/bench baseline legs) may import ``oracle``; the product never does.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def known():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def diode_configs():
    with open(os.path.join(GOLDEN, "diode_configs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle.cpu import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libdwdf_ref.so); skipped where it was never built."""
    from oracle.cpu import Ref

    try:
        return Ref()
    except (FileNotFoundError, OSError) as e:  # pragma: no cover
        pytest.skip(f"compiled reference unavailable: {e}")


def make_inputs(B, T, fs=48000.0, seed=1234, amp=(0.1, 2.0)):
    """SURVEY.md §8(d) config-2 input: per-sequence sine burst + noise (same law as make_golden.py)."""
    rng = np.random.default_rng(seed)
    n = np.arange(T)
    A = rng.uniform(amp[0], amp[1], B)
    f = np.exp(rng.uniform(np.log(50.0), np.log(5000.0), B))
    x = A[:, None] * np.sin(2 * np.pi * f[:, None] * n[None, :] / fs) + 0.05 * rng.standard_normal((B, T))
    return x.astype(np.float32)


def load_pkg():
    """The product package (its directory name carries the reference's hyphen)."""
    import importlib

    return importlib.import_module("differentiable-wdfs_b200")


@pytest.fixture(scope="session")
def dwdf():
    return load_pkg()


def seq_rel_err(y, y_ref):
    """Parity metric of SURVEY.md §7-4: max|y - y_ref| / max|y_ref| per sequence, worst sequence."""
    y = np.asarray(y, np.float64)
    y_ref = np.asarray(y_ref, np.float64)
    den = np.maximum(np.max(np.abs(y_ref), axis=-1), 1e-30)
    return float(np.max(np.max(np.abs(y - y_ref), axis=-1) / den))
