"""The C ABI called from C++ (no Python in the loop): include/dwdf_clipper.hpp — the twin of the plugin's
DiodeClipperWDF::{prepare, setParameters, process} — is compiled with g++ against include/dwdf.h and linked to libdwdf.so.

not gpu: it compiles, links, and without a device fails loudly (exit code 3) instead of computing anything on the CPU.
gpu:     block-by-block streaming with a cutoff change and model switches in mid-stream, against the UNMODIFIED reference
         (oracle/_ref: chowdsp_wdf DiodePairT, plugin probe ordering) and against the Python path's own streaming calls.
"""
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, make_inputs, seq_rel_err

SRC = os.path.join(ROOT, "tests", "cpp", "clipper_twin_main.cpp")
PKG = os.path.join(ROOT, "differentiable-wdfs_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
TOTAL = 3 * 2048 + 2 * 1000 + 1048


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp") / "clipper_twin_main")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", SRC, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), "-L", PKG, "-ldwdf", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
           f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_twin_builds_and_refuses_to_run_without_a_gpu(twin, tmp_path):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run is covered by the gpu test")
    r = subprocess.run([twin, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), "2"], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)  # DWDF_ERR_NO_DEVICE / CUDA error surfaced; no CPU fallback
    assert not os.path.exists(tmp_path / "out.bin")


@pytest.mark.gpu
def test_cpp_twin_streams_like_the_plugin(twin, tmp_path, dwdf, ref):
    from oracle.cpu import ORDER_PLUGIN, ROOT_APPROX, ClipperParams

    B = 70
    x = make_inputs(B, TOTAL, seed=9, amp=(0.1, 3.0))
    x.tofile(tmp_path / "in.bin")
    nnv = np.load(os.path.join(GOLDEN, "nn_vectors.npz"))
    w = np.asarray(nnv["2x16_weights"], np.float32)
    w.tofile(tmp_path / "w.bin")
    r = subprocess.run([twin, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(B), str(tmp_path / "w.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    y = np.fromfile(tmp_path / "out.bin", np.float32).reshape(B, TOTAL)

    def res(fc):
        return float(np.float32(1.0) / (np.float32(6.283185307179586) * np.float32(fc) * np.float32(2.2e-9)))

    # segment A against the unmodified reference C++ (model 1 = wdft::DiodePairT, fc 1539.3 Hz -> R = 47 k)
    nA = 3 * 2048
    yA = ref.clipper(x[:, :nA], ClipperParams(R=res(1539.3)), root=ROOT_APPROX, ordering=ORDER_PLUGIN)
    assert seq_rel_err(y[:, :nA], yA) < 1e-5
    # every segment against the Python path's streaming calls (same kernels, same state hand-over): bit for bit
    def circuit(mode):
        Vs, Cc = dwdf.ResistiveVoltageSource(47000.0), dwdf.Capacitor(2.2e-9, 48000.0)
        if mode == "nn":
            mj = dwdf.model_io.json_from_weights(w, [int(v) for v in nnv["2x16_sizes"]])
            return dwdf.compile_circuit(dwdf.DenseRootModel(mj), tree=dwdf.Parallel(Vs, Cc), probe=Cc, ordering="plugin"), Vs
        dp = dwdf.DiodePair(dwdf.Parallel(Vs, Cc), 4.352e-9, 25.85e-3, 1.906, mode=mode)
        return dwdf.compile_circuit(dp, probe=Cc, ordering="plugin"), Vs

    xd = torch.from_numpy(x).cuda()
    state = torch.zeros(1, B, device="cuda")
    outs, pos = [], 0
    for mode, fc, blocks in (("approx", 1539.3, [2048] * 3), ("exact", 800.0, [1000] * 2), ("nn", 800.0, [1048])):
        circ, Vs = circuit(mode)
        circ.params[circ.slot(Vs, "R")] = res(fc)
        for n in blocks:
            outs.append(circ.process_block(xd[:, pos:pos + n].contiguous(), state))
            pos += n
    want = torch.cat(outs, 1).cpu().numpy()
    assert np.array_equal(y, want)
