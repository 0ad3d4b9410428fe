"""CPU-side checks of the drop-in boundary (no GPU, no compute calls):
  * libdwdf.so loads and exports exactly the symbols include/dwdf.h declares (and the ctypes layer binds all of them);
  * the host-only entry points (program validation / lowering, sizes, info) behave, errors are loud;
  * the imperative wdf_py element API (config 1: RC low-pass, 1 x 1024 samples, CPU) reproduces the reference;
  * the compiled path refuses to run without a GPU instead of falling back.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dwdf.h")).read()
    return sorted(set(re.findall(r"DWDF_API\s+[\w\s\*]+?\b(dwdf_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(dwdf):
    names = header_symbols()
    assert len(names) >= 30
    lib = dwdf._lib.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/dwdf.h but not exported by libdwdf.so"
    assert sorted(dwdf._lib.SYMBOLS) == names, (set(names) ^ set(dwdf._lib.SYMBOLS))
    # nothing else leaks out of the library (built with -fvisibility=hidden)
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", dwdf._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if " T " in line and "dwdf_" in line.split()[-1])
    assert exported == names


def test_build_info_and_counters(dwdf):
    info = dwdf.build_info()
    assert "sm_100a" in info and "no-cpu-fallback" in info
    assert dwdf.launch_count() >= 0
    assert dwdf._lib.lib().dwdf_last_error() is not None


def make_nodes(L):
    return (L.Node * 3)(L.Node(L.RESISTIVE_VS, -1, -1, 0), L.Node(L.CAPACITOR, -1, -1, 1), L.Node(L.PARALLEL, 0, 1, -1))


def clipper_desc(L, **kw):
    d = L.CircuitDesc()
    d.root_kind, d.root_mode, d.ordering, d.probe, d.source, d.r_node = L.ROOT_DIODE_PAIR, L.MODE_APPROX, L.ORDER_PYTHON, 1, 0, -1
    d.param_Is, d.param_nabla, d.n_params, d.newton_max_iter = 2, 3, 4, 0
    d.fs, d.Vt, d.n_up, d.n_down, d.newton_tol = 48000.0, 25.85e-3, 1.0, 1.0, 0.0
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def test_program_validation_is_host_only_and_loud(dwdf):
    L = dwdf._lib
    lib = L.lib()
    h = C.c_void_p()
    d = clipper_desc(L)
    assert lib.dwdf_program_create(make_nodes(L), 3, C.byref(d), C.byref(h)) == 0
    assert lib.dwdf_program_is_clipper(h) == 1 and lib.dwdf_program_n_states(h) == 1
    assert lib.dwdf_ckpt_bytes(h, 64, 4096) == 64 * 256 * 4
    assert lib.dwdf_workspace_bytes(h, 64, 4096) > 0
    lib.dwdf_program_destroy(h)
    # the interpreter takes what is not the clipper shape (probe on the source)
    d2 = clipper_desc(L, probe=0)
    assert lib.dwdf_program_create(make_nodes(L), 3, C.byref(d2), C.byref(h)) == 0 and lib.dwdf_program_is_clipper(h) == 0
    lib.dwdf_program_destroy(h)
    for bad, code in ((dict(n_params=0), 1), (dict(fs=0.0), 1), (dict(probe=7), 1), (dict(ordering=5), 1), (dict(root_mode=9), 1), (dict(param_Is=11), 1), (dict(Vt=-1.0), 1), (dict(root_kind=7), 1),
                      (dict(root_mode=L.MODE_APPROX_GOOD, n_down=2.0), 2)):
        db = clipper_desc(L, **bad)
        assert lib.dwdf_program_create(make_nodes(L), 3, C.byref(db), C.byref(h)) == code, bad
        assert len(lib.dwdf_last_error()) > 0
    cyc = (L.Node * 2)(L.Node(L.SERIES, 1, 0, -1), L.Node(L.RESISTOR, -1, -1, 0))  # a child after its parent
    assert lib.dwdf_program_create(cyc, 2, C.byref(clipper_desc(L)), C.byref(h)) == 1
    # neural root: shapes the kernels are built for
    dn = clipper_desc(L, root_kind=L.ROOT_NEURAL, n_params=2, param_Is=-1, param_nabla=-1)
    m = L.MlpDesc(2, 16)
    assert lib.dwdf_mlp_weight_count(C.byref(m)) == 2 * 16 + 16 + 2 * (16 * 16 + 16) + 16 + 1 == 609
    assert lib.dwdf_program_create_neural(make_nodes(L), 3, C.byref(dn), C.byref(m), C.byref(h)) == 0
    assert lib.dwdf_neural_ckpt_bytes(h, 10, 130) == (3 + 1) * 10 * 4
    lib.dwdf_program_destroy(h)
    assert lib.dwdf_program_create_neural(make_nodes(L), 3, C.byref(dn), C.byref(L.MlpDesc(2, 5)), C.byref(h)) == 2  # unsupported width
    with pytest.raises(dwdf.DwdfError):
        L.check(1)


def test_no_gpu_no_result(dwdf):
    """The compiled path is CUDA-only: on a box without a GPU it raises, it never computes on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less build box")
    Vs, Cc = dwdf.ResistiveVoltageSource(47000.0), dwdf.Capacitor(2.2e-9, 48000.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dwdf.compile_circuit(dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9), probe=Cc)
    L = dwdf._lib
    lib = L.lib()
    h = C.c_void_p()
    d = clipper_desc(L)
    assert lib.dwdf_program_create(make_nodes(L), 3, C.byref(d), C.byref(h)) == 0
    x = np.zeros((32, 64), np.float32)
    p = np.array([47000.0, 2.2e-9, 4.352e-9, 1.906], np.float32)
    rc = lib.dwdf_forward_host(h, p.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), None, x.ctypes.data_as(C.c_void_p), 32, 64)
    assert rc in (3, 4) and b"" != lib.dwdf_last_error()  # DWDF_ERR_CUDA / DWDF_ERR_NO_DEVICE
    lib.dwdf_program_destroy(h)


def test_config1_rc_lowpass_imperative_cpu(dwdf, golden):
    """BASELINE config 1: the RC low-pass of lpf.py:20-49 stepped sample by sample through the element API on the
    CPU (1 sequence x 1024 samples, R 1 kOhm, C 1 uF, 48 kHz), against the reference's own C++ output."""
    x = torch.from_numpy(golden["lpf_x"]).reshape(1, -1, 1)  # lpf.py:33: (B, T, 1)
    Vs = dwdf.IdealVoltageSource()
    R1 = dwdf.Resistor(1000.0, True)
    C1 = dwdf.Capacitor(1.0e-6, 48000.0, True)
    S1 = dwdf.Series(R1, C1)
    I1 = dwdf.PolarityInverter(S1)
    I1.calc_impedance()
    out = []
    for i in range(x.shape[1]):  # lpf.py:39-46
        Vs.set_voltage(x[:, i])
        Vs.incident(I1.reflected())
        I1.incident(Vs.reflected())
        out.append(dwdf.voltage(C1))
    y = torch.stack(out).reshape(-1).detach().numpy()
    assert np.max(np.abs(y - golden["lpf_y_f64"])) < 1e-5 * np.max(np.abs(golden["lpf_y_f64"]))  # fp32 element arithmetic like the reference (tf_wdf.py:72,102)
    # the same protocol with a diode pair as root: the imperative reflected() evaluates the law its mode names
    # ('exact': Wright omega; 'approx': omega4 as wdft::DiodePairT) — against the reference C++'s own outputs for each
    for mode, key in (("exact", "clip_plugin_exact_python_f32"), ("approx", "clip_plugin_approx_python_f32")):
        Vr = dwdf.ResistiveVoltageSource(47000.0)
        Cc = dwdf.Capacitor(2.2e-9, 48000.0)
        P1 = dwdf.Parallel(Vr, Cc)
        dp = dwdf.DiodePair(P1, 4.352e-9, 25.85e-3, 1.906, mode=mode)
        P1.calc_impedance()
        xs = torch.from_numpy(golden["clip_x"][:2, :64])
        ys = []
        for i in range(xs.shape[1]):  # clipper_pot.py:113-124
            Vr.set_voltage(xs[:, i:i + 1])
            dp.incident(P1.reflected())
            P1.incident(dp.reflected())
            ys.append(dwdf.voltage(Cc))
        ys = torch.cat(ys, dim=1).numpy()
        ref = golden[key][:2, :64]
        assert np.max(np.abs(ys - ref)) < 1e-5 * np.max(np.abs(ref)), mode
