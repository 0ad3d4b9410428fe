"""Run-time specialisation of tree programs (dwdf_program_specialize, csrc/tree_jit.cu).

CPU: the generated source of a few circuits compiles for sm_100a with NVRTC (no device needed); circuits the specialiser
does not cover say so. GPU: the specialised kernels against the interpreter (same circuit, same inputs), the CPU oracle
and fp64 autograd — forward, reverse mode with fused loss and upstream gradients, streaming, ragged shapes (the
direct-access twins) and aligned ones (the TMA kernels).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import make_inputs, seq_rel_err
from oracle.cpu import ORDER_PLUGIN, ORDER_PYTHON, ClipperParams


def _source(lib, h, part):
    n = lib.dwdf_program_specialized_source(h, part, None, 0)
    if n == 0:
        return None
    buf = C.create_string_buffer(n)
    assert lib.dwdf_program_specialized_source(h, part, buf, n) == n
    return buf.value


def _desc(L, **kw):
    d = L.CircuitDesc()
    d.root_kind, d.root_mode, d.ordering, d.probe, d.source, d.r_node = L.ROOT_IDEAL_VS, 0, L.ORDER_PYTHON, 1, -1, -1
    d.param_Is, d.param_nabla, d.n_params, d.newton_max_iter = -1, -1, 2, 0
    d.fs, d.Vt, d.n_up, d.n_down, d.newton_tol = 48000.0, 25.85e-3, 1.0, 1.0, 0.0
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def _programs(L):
    lpf = (L.Node * 4)(L.Node(L.RESISTOR, -1, -1, 0), L.Node(L.CAPACITOR, -1, -1, 1), L.Node(L.SERIES, 0, 1, -1), L.Node(L.INVERTER, 2, -1, -1))  # lpf.py:23-28
    hpf = (L.Node * 5)(L.Node(L.RESISTOR, -1, -1, 0), L.Node(L.RESISTIVE_VS, -1, -1, 1), L.Node(L.CAPACITOR, -1, -1, 2), L.Node(L.SERIES, 1, 2, -1), L.Node(L.PARALLEL, 0, 3, -1))  # HPFDiodeClipper.h:25-37
    return {
        "lpf": (lpf, 4, _desc(L)),
        "hpf_exact_plugin": (hpf, 5, _desc(L, root_kind=L.ROOT_DIODE_PAIR, root_mode=L.MODE_EXACT, probe=2, source=1, param_Is=3, param_nabla=4, n_params=5, ordering=L.ORDER_PLUGIN)),
        "hpf_approx_asym": (hpf, 5, _desc(L, root_kind=L.ROOT_DIODE_PAIR, root_mode=L.MODE_APPROX, probe=0, source=1, param_Is=3, param_nabla=4, n_params=5, n_down=2.0)),
    }


@pytest.mark.parametrize("name", ["lpf", "hpf_exact_plugin", "hpf_approx_asym"])
def test_generated_source_compiles_for_sm_100a(dwdf, name):
    nvrtc = pytest.importorskip("cuda.bindings.nvrtc")
    L = dwdf._lib
    lib = L.lib()
    nodes, n, d = _programs(L)[name]
    h = C.c_void_p()
    assert lib.dwdf_program_create(nodes, n, C.byref(d), C.byref(h)) == 0, lib.dwdf_last_error()
    assert lib.dwdf_program_is_specialized(h) == 0
    main, hm, ht = _source(lib, h, 0), _source(lib, h, 1), _source(lib, h, 2)
    assert b"jit_step_adj" in main and b"jit_tree_adjoint_tma" in main and b"pair_reflect" in hm and b"tma_load_2d" in ht
    err, prog = nvrtc.nvrtcCreateProgram(main, b"dwdf_tree_jit.cu", 2, [hm, ht], [b"dwdf_math.cuh", b"dwdf_tma.cuh"])
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"--fmad=false"]
    (err,) = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, nlog = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * nlog
    nvrtc.nvrtcGetProgramLog(prog, log)
    assert int(err) == 0, log.decode()[:2000]
    _, nb = nvrtc.nvrtcGetCUBINSize(prog)
    assert nb > 10000
    lib.dwdf_program_destroy(h)


def test_unsupported_circuits_say_so(dwdf):
    L = dwdf._lib
    lib = L.lib()
    h = C.c_void_p()
    # the diode clipper has its own kernels; a per-sample resistance channel stays on the interpreter
    clip = (L.Node * 3)(L.Node(L.RESISTIVE_VS, -1, -1, 0), L.Node(L.CAPACITOR, -1, -1, 1), L.Node(L.PARALLEL, 0, 1, -1))
    d = _desc(L, root_kind=L.ROOT_DIODE_PAIR, probe=1, source=0, param_Is=2, param_nabla=3, n_params=4)
    assert lib.dwdf_program_create(clip, 3, C.byref(d), C.byref(h)) == 0
    assert lib.dwdf_program_specialized_source(h, 0, None, 0) == 0 and lib.dwdf_program_specialize(h) == 2
    lib.dwdf_program_destroy(h)
    swapped = (L.Node * 3)(L.Node(L.CAPACITOR, -1, -1, 0), L.Node(L.RESISTIVE_VS, -1, -1, 1), L.Node(L.PARALLEL, 0, 1, -1))
    d = _desc(L, root_kind=L.ROOT_DIODE_PAIR, probe=0, source=1, param_Is=2, param_nabla=3, n_params=4, r_node=1)
    assert lib.dwdf_program_create(swapped, 3, C.byref(d), C.byref(h)) == 0, lib.dwdf_last_error()
    assert lib.dwdf_program_specialized_source(h, 0, None, 0) == 0
    lib.dwdf_program_destroy(h)
    # a current probe is forward only: the generated source carries a stub for the reverse-mode step
    d = _desc(L, root_kind=L.ROOT_DIODE_PAIR, probe=0, source=1, param_Is=2, param_nabla=3, n_params=4, probe_current=1)
    assert lib.dwdf_program_create(swapped, 3, C.byref(d), C.byref(h)) == 0, lib.dwdf_last_error()
    n = lib.dwdf_program_specialized_source(h, 0, None, 0)
    buf = C.create_string_buffer(n)
    lib.dwdf_program_specialized_source(h, 0, buf, n)
    assert b"c.Gprobe" in buf.value and b"jit_step_adj (const JC&, float" in buf.value
    lib.dwdf_program_destroy(h)


# ---- GPU -----------------------------------------------------------------------------------------------------------

def _lpf(dwdf, R=1000.0, Cv=1.0e-6, fs=48000.0):
    R1 = dwdf.Resistor(R, True)
    C1 = dwdf.Capacitor(Cv, fs, True)
    I1 = dwdf.Inverter(dwdf.Series(R1, C1))
    return dwdf.compile_circuit(dwdf.IdealVoltageSource(), tree=I1, probe=C1), R1, C1


@pytest.mark.gpu
@pytest.mark.parametrize("B,T", [(40, 256), (33, 203), (1, 5), (300, 1024)])
def test_rc_lowpass_against_autograd(dwdf, B, T):
    """lpf.py:20-49 on the specialised kernels: output and dL/dR, dL/dC against fp64 autograd over the restatement of tf_wdf.py."""
    from oracle import torch_wdf as tw

    x = make_inputs(B, T, seed=3)
    target = (0.5 * x).astype(np.float32)
    circ, R1, C1 = _lpf(dwdf)
    assert circ.specialize() and circ.is_specialized
    y = circ.forward(torch.from_numpy(x).cuda())
    n = min(B, 8)  # the per-sample torch loop is slow: check a slice against autograd, the rest against the interpreter
    y_ref, leaves = tw.lpf_forward(x[:n])
    assert seq_rel_err(y.cpu().numpy()[:n], y_ref[..., 0].t().detach().numpy()) < 1e-5
    ref, _, _ = _lpf(dwdf)
    y_int = ref.forward(torch.from_numpy(x).cuda())
    assert not ref.is_specialized and seq_rel_err(y.cpu().numpy(), y_int.cpu().numpy()) < 2e-6
    res = circ.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr", skip=min(3, T - 1))
    res_i = ref.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr", skip=min(3, T - 1))
    g, gi = res["grads"].cpu().numpy(), res_i["grads"].cpu().numpy()
    assert np.allclose(g[:2], gi[:2], rtol=2e-4, atol=0), (g, gi)
    assert abs(float(res["loss"]) / float(res_i["loss"]) - 1) < 1e-5
    if B <= 40:
        circ.forward(torch.from_numpy(x[:n]).cuda())
        r2 = circ.backward(target=torch.from_numpy(target[:n]).cuda(), loss="mse")
        loss = torch.mean((y_ref[..., 0].t() - torch.from_numpy(target[:n]).double()) ** 2)  # (T, B, 1) -> (B, T)
        gR, gC = torch.autograd.grad(loss, [leaves["R"], leaves["C"]])
        g2 = r2["grads"].cpu().numpy()
        assert abs(g2[circ.slot(R1, "R")] / float(gR) - 1) < 5e-4 and abs(g2[circ.slot(C1, "C")] / float(gC) - 1) < 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["approx", "exact"])
@pytest.mark.parametrize("ordering,oord", [("plugin", ORDER_PLUGIN), ("python", ORDER_PYTHON)])
def test_swapped_clipper_against_the_clipper_oracle(dwdf, oracle, mode, ordering, oord):
    """Parallel(C, Vs) + DiodePair runs on the tree path; specialised, it must give the diode clipper's output and gradients."""
    p = ClipperParams()
    x = make_inputs(70, 512, seed=31)
    target = oracle.clipper_forward(x, ClipperParams(R=p.R * 1.1, C=p.C * 0.9, Is=p.Is * 2, nabla=p.nabla * 1.05), exact=True, ordering=oord)
    Vs = dwdf.ResistiveVoltageSource(p.R, True)
    Cc = dwdf.Capacitor(p.C, p.fs, True)
    dp = dwdf.DiodePair(dwdf.Parallel(Cc, Vs), p.Is, p.Vt, p.nabla, trainable=True, mode=mode)
    circ = dwdf.compile_circuit(dp, probe=Cc, ordering=ordering)
    assert not circ.is_clipper and circ.specialize()
    y = circ.forward(torch.from_numpy(x).cuda()).cpu().numpy()
    assert seq_rel_err(y, oracle.clipper_forward(x, p, exact=(mode == "exact"), ordering=oord)) < 1e-5
    res = circ.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr", skip=20)
    ref = oracle.clipper_grad(x, target, p, exact=(mode == "exact"), ordering=oord, mode="target", loss="mse+esr", skip=20, dtype=np.float64)
    g = res["grads"].cpu().numpy()[[circ.slot(dp, "Is"), circ.slot(dp, "nabla"), circ.slot(Vs, "R"), circ.slot(Cc, "C")]]
    rel = np.abs(g / ref["grads"] - 1.0)
    # dL/dR is the difference of two nearly cancelling chain-rule terms (gamma and ell both move with R, ~1e4 x cancellation here):
    # the fp32 per-sample terms leave it ~5e-3 accurate on the tree path (interpreter and specialised alike); the others 1e-5
    assert np.max(rel[[0, 1, 3]]) < 5e-4 and rel[2] < 2e-2, (g, ref["grads"])
    assert abs(float(res["loss"]) / ref["loss"] - 1.0) < 1e-4
    # and the interpreter, same circuit: the two tree paths agree much more closely than either agrees with fp64
    Vs2, C2 = dwdf.ResistiveVoltageSource(p.R, True), dwdf.Capacitor(p.C, p.fs, True)
    dp2 = dwdf.DiodePair(dwdf.Parallel(C2, Vs2), p.Is, p.Vt, p.nabla, trainable=True, mode=mode)
    ci = dwdf.compile_circuit(dp2, probe=C2, ordering=ordering)
    ci.forward(torch.from_numpy(x).cuda())
    gi = ci.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr", skip=20)["grads"].cpu().numpy()[[ci.slot(dp2, "Is"), ci.slot(dp2, "nabla"), ci.slot(Vs2, "R"), ci.slot(C2, "C")]]
    assert not ci.is_specialized and np.max(np.abs(g / gi - 1.0)) < 3e-3, (g, gi)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("DWDF_FUZZ_N5", "16"))))
def test_random_trees_specialised_equal_interpreter(dwdf, oracle, seed):
    """Random trees (tests/test_gpu_fuzz.py's generator): the specialised kernels against the interpreter — output, fused-loss
    gradients, upstream gradients, streaming in blocks — on aligned shapes (TMA kernels) and ragged ones (direct twins)."""
    from oracle.cpu import ROOT_DIODE_PAIR, ROOT_IDEAL_VS
    from test_gpu_fuzz import _random_tree

    outs = []
    for spec in (False, True):
        rng = np.random.default_rng(23000 + seed)
        fs = float(rng.choice([44100.0, 48000.0, 96000.0]))
        diode = bool(rng.integers(2))
        mode = str(rng.choice(["approx", "exact"]))
        ordering = str(rng.choice(["plugin", "python"]))
        top, nodes, elems = _random_tree(rng, dwdf, fs, need_source=diode)
        leaves = [i for i, n in enumerate(nodes) if n[1] < 0]
        probe = int(rng.choice(leaves))
        B, T = int(rng.choice([5, 40, 97])), int(rng.choice([64, 203, 512]))
        x = (make_inputs(B, T, fs=fs, seed=seed) * float(rng.choice([0.3, 1.0]))).astype(np.float32)
        p = ClipperParams()
        root = dwdf.DiodePair(top, p.Is, p.Vt, p.nabla, trainable=True, mode=mode) if diode else dwdf.IdealVoltageSource()
        circ = dwdf.compile_circuit(root, tree=top, probe=elems[probe], ordering=ordering)
        if circ.is_clipper:
            return
        if spec:
            assert circ.specialize()
        xd = torch.from_numpy(x).cuda()
        y = circ.forward(xd).clone()
        target = (0.5 * y + 0.01).contiguous()
        rt = circ.backward(target=target, loss="mse+esr", skip=7)
        gt, lt = rt["grads"].cpu().numpy().copy(), float(rt["loss"])
        gy = torch.from_numpy(np.random.default_rng(seed).standard_normal(x.shape).astype(np.float32)).cuda()
        circ.forward(xd)
        gu = circ.backward(gy=gy)["grads"].cpu().numpy().copy()
        st = circ.new_state(B)
        cut = (T // 3) // 4 * 4
        parts = torch.cat([circ.process_block(xd[:, :cut].contiguous(), st), circ.process_block(xd[:, cut:].contiguous(), st)], 1)
        vals = circ.params.cpu().numpy().astype(np.float64)
        outs.append((y.cpu().numpy(), gt, lt, gu, parts.cpu().numpy(), vals))
    (y0, gt0, l0, gu0, s0, vals), (y1, gt1, l1, gu1, s1, _) = outs
    # the circuit's own conditioning in fp32 (as tests/test_gpu_fuzz.py measures it): the oracle's tree executor in fp32 against fp64.
    # Two fp32 evaluations that differ only in where the compiler contracts a product into an FMA differ by about that much.
    oord = ORDER_PLUGIN if ordering == "plugin" else ORDER_PYTHON
    if diode:
        source = [i for i, e in enumerate(elems) if isinstance(e, dwdf.ResistiveVoltageSource)][0]
        args = dict(root_kind=ROOT_DIODE_PAIR, source=source, root_par=[float(mode == "exact"), 0, p.Is, p.Vt, p.nabla, 1, 1])
    else:
        args = dict(root_kind=ROOT_IDEAL_VS, source=-1, root_par=None)
    ref32 = oracle.tree_run(nodes, fs, args["root_kind"], x, probe=probe, source=args["source"], root_par=args["root_par"], ordering=oord)
    ref64 = oracle.tree_run(nodes, fs, args["root_kind"], x, probe=probe, source=args["source"], root_par=args["root_par"], ordering=oord, dtype=np.float64)
    scale = np.maximum(np.max(np.abs(ref64), axis=1, keepdims=True), 1e-3 * np.max(np.abs(x), axis=1, keepdims=True) + 1e-30)
    cond = float(np.max(np.abs(ref32 - ref64) / scale))
    tol = max(5e-6, 5.0 * cond)
    assert np.max(np.abs(y1 - ref64) / scale) < max(1e-5, 5.0 * cond), (nodes, probe, cond)  # the interpreter's own bar
    assert np.max(np.abs(y1 - y0) / scale) < tol, (nodes, probe, cond)
    assert np.max(np.abs(s1 - y1) / scale) < tol  # streaming in two blocks = one long block
    assert abs(l1 / l0 - 1) < max(1e-5, 20.0 * cond)
    gtol = max(5e-4, 200.0 * cond)
    for a, b in ((gt0, gt1), (gu0, gu1)):
        a, b = a[: len(vals)] * vals, b[: len(vals)] * vals  # d/d ln(value): comparable across ohms, farads, amperes
        assert np.max(np.abs(a - b)) < gtol * np.max(np.abs(a)) + 1e-12, (a, b, cond)


@pytest.mark.gpu
def test_auto_specialisation_and_full_size_run(dwdf):
    """forward() specialises a tree program on its own from 2^20 samples on; the RC low-pass at a benchmark-sized batch gives
    the interpreter's numbers (compared on a slice) and a training step moves R and C."""
    B, T = 4096, 1024
    x = torch.from_numpy(make_inputs(B, T, seed=5)).cuda()
    circ, R1, C1 = _lpf(dwdf)
    assert not circ.is_specialized
    y = circ.forward(x)
    assert circ.is_specialized
    ref, _, _ = _lpf(dwdf)
    y_ref = ref.forward(x[:64].contiguous())
    assert seq_rel_err(y[:64].cpu().numpy(), y_ref.cpu().numpy()) < 2e-6
    target = (0.9 * y).contiguous()
    opt = dwdf.Adam(circ, lr={circ.slot(R1, "R"): 1.0, circ.slot(C1, "C"): 1e-9})
    before = circ.params.clone()
    l0 = float(circ.train_step(x, target, opt)["loss"])
    for _ in range(5):
        l1 = float(circ.train_step(x, target, opt)["loss"])
    assert l1 < l0 and not torch.equal(before, circ.params)
