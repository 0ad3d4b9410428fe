"""The neural diode-pair root (inference): host-side pieces on CPU, the fused kernel on
the GPU against golden vectors produced by the reference's own RTNeural + chowdsp_wdf code
(tests/golden/make_golden_nn.py) and against the numpy oracle (oracle/nn.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, make_inputs, seq_rel_err
from oracle import nn

MODELS = ["2x4", "2x8", "2x16", "4x4", "4x8"]
CLIP_TOL = 5e-5  # see tests/test_oracle_nn.py


@pytest.fixture(scope="module")
def nnv():
    return np.load(os.path.join(GOLDEN, "nn_vectors.npz"))


def model_json(dwdf, nnv, name):
    return dwdf.model_io.json_from_weights(nnv[f"{name}_weights"], [int(v) for v in nnv[f"{name}_sizes"]])


# ---- CPU: JSON format, imperative element, dataset loader -----------------------------------------------

@pytest.mark.parametrize("name", MODELS)
def test_json_roundtrip_and_imperative_root(dwdf, nnv, name, tmp_path):
    mj = model_json(dwdf, nnv, name)
    path = tmp_path / "model.json"
    dwdf.model_io.save_model_json(mj, str(path))
    back = dwdf.model_io.load_model_json(str(path))
    layers, sizes = dwdf.model_io.layers_from_json(back)
    assert sizes == [int(v) for v in nnv[f"{name}_sizes"]]
    assert [a for _, _, a in layers] == ["tanh"] * (len(layers) - 1) + [""]
    assert np.array_equal(dwdf.model_io.flatten_weights(layers), nnv[f"{name}_weights"])
    # Keras-style nesting (layers.py:29-35: kernel = [weights]) is accepted too
    nested = json.loads(json.dumps(back))
    for layer in nested["layers"]:
        layer["weights"] = [[layer["weights"][0]], [layer["weights"][1]]]
    assert np.array_equal(dwdf.model_io.flatten_weights(dwdf.model_io.layers_from_json(nested)[0]), nnv[f"{name}_weights"])
    # imperative element = layers.py:72-82
    root = dwdf.DenseRootModel(back)
    x = torch.from_numpy(np.stack([nnv["grid_a"], nnv["grid_logR"]], -1))
    root.incident(x)
    out = root.reflected()[..., 0].numpy()
    assert np.max(np.abs(out - nnv[f"{name}_grid_out"])) < 5e-6


def test_dataset_loader(dwdf, tmp_path):
    """A synthetic Digilent CSV pair (same header layout as diode_dataset/1N4148/1up1down/10.0k_4.7nF.csv:1-12)."""
    fs, n = 1000.0, 20000
    d = tmp_path / "diode_dataset" / "1N4148" / "1up1down"
    d.mkdir(parents=True)
    rng = np.random.default_rng(0)
    data = {}
    for fname in ("10.0k_4.7nF.csv", "47.0k_4.7nF.csv"):
        a = rng.standard_normal((n, 2))
        data[fname] = a
        with open(d / fname, "w") as f:
            f.write("#Digilent WaveForms Oscilloscope Acquisition\n#Device Name: Discovery2\n#Serial Number: SN:0\n#Date Time: 2022-01-12 19:41:59.543\n")
            f.write(f"#Sample rate: {fs:g}Hz\n#Samples: {n}\n#Trigger: x\n#Channel 1: x\n#Channel 2: x\n\nChannel 1 (V),Channel 2 (V)\n")
            for row in a:
                f.write(f"{float(row[0])!r},{float(row[1])!r}\n")
    path = dwdf.dataimport.data_path_for_diode(1, 1, str(tmp_path))
    tr, ntr, va, nva, FS = dwdf.dataimport.load_diode_data(path)
    lo, hi = 2500, 16800  # 2.5 s .. 16.8 s (dataimport.py:35-48)
    assert FS == fs and ntr == hi - lo and nva == hi - lo
    assert np.allclose(tr[0], data["10.0k_4.7nF.csv"][lo:hi, 0].astype(np.float32)) and np.all(tr[1] == 10000.0)  # R < 36 k: training
    assert np.allclose(va[2], data["47.0k_4.7nF.csv"][lo:hi, 1].astype(np.float32)) and np.all(va[1] == 47000.0)  # 36 k..73 k: validation
    X, Y = dwdf.dataimport.batch_data(tr, 2048)
    assert X.shape == (6, 2048, 2) and Y.shape == (6, 2048, 1) and np.array_equal(X[1, :, 0], tr[0, 2048:4096])


# ---- GPU: the fused kernel -------------------------------------------------------------------------------

def make_circuit(dwdf, mj, ordering, with_r=False, R=47000.0, C=2.2e-9, fs=48000.0):
    Vs = dwdf.ResistiveVoltageSource(R)
    Cc = dwdf.Capacitor(C, fs)
    P1 = dwdf.Parallel(Vs, Cc)
    root = dwdf.DenseRootModel(mj)
    return dwdf.compile_circuit(root, tree=P1, probe=Cc, ordering=ordering, r_element=Vs if with_r else None)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("ordering,order", [("plugin", nn.ORDER_PLUGIN), ("python", nn.ORDER_PYTHON)])
def test_forward_against_reference_vectors(dwdf, nnv, name, ordering, order):
    circ = make_circuit(dwdf, model_json(dwdf, nnv, name), ordering)
    assert circ.is_neural
    y = circ.forward(torch.from_numpy(nnv["x"]).cuda()).cpu().numpy()
    assert seq_rel_err(y, nnv[f"{name}_clip_{ordering}"]) < CLIP_TOL
    ref64 = nn.nn_clipper_forward(nnv["x"], nnv[f"{name}_weights"], nnv[f"{name}_sizes"], 48000.0, 47000.0, 2.2e-9, order, dtype=np.float64)
    assert seq_rel_err(y, ref64) < CLIP_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2x8", "4x8", "2x16"])
@pytest.mark.parametrize("B,T", [(1, 5), (7, 333), (64, 256)])
def test_forward_resistance_channel_and_shapes(dwdf, nnv, name, B, T):
    """clipper_pot.py's (B, T, 2) layout: channel 1 sets the source resistance every sample; odd batches, ragged T."""
    x = make_inputs(B, T, seed=B + T)
    r = (np.random.default_rng(2).uniform(1e4, 1e5, (B, 1)) * np.ones((1, T))).astype(np.float32)
    r[:, T // 2:] *= 1.7
    mj = model_json(dwdf, nnv, name)
    circ = make_circuit(dwdf, mj, "python", with_r=True, R=45000.0, C=4.7e-9, fs=50000.0)
    y = circ.forward(torch.from_numpy(x).cuda(), r=torch.from_numpy(r).cuda()).cpu().numpy()
    ref = nn.nn_clipper_forward(x, nnv[f"{name}_weights"], nnv[f"{name}_sizes"], 50000.0, 45000.0, 4.7e-9, nn.ORDER_PYTHON, r=r, dtype=np.float64)
    assert seq_rel_err(y, ref) < CLIP_TOL
    circ0 = make_circuit(dwdf, mj, "python", R=45000.0, C=4.7e-9, fs=50000.0)
    y0 = circ0.forward(torch.from_numpy(x).cuda()).cpu().numpy()
    assert seq_rel_err(y0, nn.nn_clipper_forward(x, nnv[f"{name}_weights"], nnv[f"{name}_sizes"], 50000.0, 45000.0, 4.7e-9, nn.ORDER_PYTHON, dtype=np.float64)) < CLIP_TOL


@pytest.mark.gpu
def test_streaming_and_errors(dwdf, nnv):
    mj = model_json(dwdf, nnv, "2x8")
    circ = make_circuit(dwdf, mj, "plugin")
    x = torch.from_numpy(make_inputs(9, 600, seed=8)).cuda()
    whole = circ.forward(x)
    st = circ.new_state(9)
    parts = [circ.process_block(x[:, a:b].contiguous(), st) for a, b in ((0, 100), (100, 101), (101, 600))]
    assert seq_rel_err(torch.cat(parts, 1).cpu().numpy(), whole.cpu().numpy()) < 2e-5  # (the long block runs time-parallel)
    bad = dwdf.model_io.json_from_weights(np.zeros(2 * 5 + 5 + 5 + 1, np.float32), [2, 5, 1])
    with pytest.raises(Exception):
        make_circuit(dwdf, bad, "plugin")


NN_GRAD_TOL = 1e-4  # relative to the largest gradient entry (fp32 network + fp32 per-lane accumulation vs fp64 autograd; measured 1e-6 .. 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("ordering,order", [("plugin", nn.ORDER_PLUGIN), ("python", nn.ORDER_PYTHON)])
@pytest.mark.parametrize("loss,skip", [("mse", 0), ("mse+esr", 50)])
def test_weight_gradients_against_autograd(dwdf, nnv, name, ordering, order, loss, skip):
    """dL/d(weights) of the hand-written adjoint against fp64 autograd through the same recurrence
    (what tape.gradient computes for clipper_pot.py:246-269)."""
    B, T = 9, 200
    x = make_inputs(B, T, seed=21)
    w, sizes = nnv[f"{name}_weights"], [int(v) for v in nnv[f"{name}_sizes"]]
    target = 0.8 * nn.nn_clipper_forward(x, w, sizes, 48000.0, 47000.0, 2.2e-9, order, dtype=np.float64) + 0.01
    circ = make_circuit(dwdf, model_json(dwdf, nnv, name), ordering)
    circ.forward(torch.from_numpy(x).cuda())
    res = circ.backward(target=torch.from_numpy(target.astype(np.float32)).cuda(), loss=loss, skip=skip)
    ref = nn.nn_clipper_grad_torch(x, target.astype(np.float32), w, sizes, 48000.0, 47000.0, 2.2e-9, order, loss=loss, skip=skip)
    g = res["grads"].cpu().numpy()
    assert g.shape == ref["grad_w"].shape
    assert np.max(np.abs(g - ref["grad_w"])) < NN_GRAD_TOL * np.max(np.abs(ref["grad_w"]))
    assert abs(float(res["loss"]) / ref["loss"] - 1) < 1e-4


@pytest.mark.gpu
def test_weight_gradients_upstream_and_resistance_channel(dwdf, nnv):
    B, T = 5, 150
    x = make_inputs(B, T, seed=22)
    r = (np.random.default_rng(3).uniform(1e4, 1e5, (B, 1)) * np.ones((1, T))).astype(np.float32)
    gy = np.random.default_rng(4).standard_normal((B, T)).astype(np.float32)
    w, sizes = nnv["2x8_weights"], [int(v) for v in nnv["2x8_sizes"]]
    circ = make_circuit(dwdf, model_json(dwdf, nnv, "2x8"), "python", with_r=True, R=45000.0, C=4.7e-9, fs=50000.0)
    circ.forward(torch.from_numpy(x).cuda(), r=torch.from_numpy(r).cuda())
    g = circ.backward(gy=torch.from_numpy(gy).cuda())["grads"].cpu().numpy()
    ref = nn.nn_clipper_grad_torch(x, None, w, sizes, 50000.0, 45000.0, 4.7e-9, nn.ORDER_PYTHON, r=r, gy=gy)
    assert np.max(np.abs(g - ref["grad_w"])) < NN_GRAD_TOL * np.max(np.abs(ref["grad_w"]))


@pytest.mark.gpu
def test_training_reduces_the_loss(dwdf, nnv):
    """A few Adam steps on the network (clipper_pot.py:180: Adam(1e-4, beta_1=0.5)) towards the output of another
    trained model: the loss goes down, run to run identically (fixed-order reductions)."""
    x = torch.from_numpy(make_inputs(64, 512, seed=23)).cuda()
    teacher = make_circuit(dwdf, model_json(dwdf, nnv, "2x16"), "python")
    target = teacher.forward(x, keep_for_backward=False).clone()
    losses = []
    for run in range(2):
        circ = make_circuit(dwdf, model_json(dwdf, nnv, "2x8"), "python")
        opt = dwdf.AdamWeights(circ, lr=1e-3, beta_1=0.5)
        ls = []
        for _ in range(15):
            circ.forward(x)
            ls.append(float(circ.backward(target=target, loss="mse+esr", skip=50)["loss"]))
            opt.apply()
        losses.append(ls)
    assert losses[0] == losses[1]
    assert losses[0][-1] < 0.7 * losses[0][0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2x8", "4x4"])
@pytest.mark.parametrize("with_r", [False, True])
def test_time_parallel_forward_equals_serial(dwdf, nnv, name, with_r):
    """Few long sequences run one lane per (pair, 256-sample chunk) with a speculative warm-up that is verified
    (missed chunks are recomputed). Two trajectories of the neural root never merge bit for bit (the network's own
    fp32 rounding noise, a few 1e-6, keeps them apart), so chunks are accepted within that noise: the output equals
    the one-lane-per-pair kernel's to 2e-5 of the peak, inside the parity budget of 5e-5."""
    B, T = 11, 1500
    x = torch.from_numpy(make_inputs(B, T, seed=71, amp=(0.1, 4.0))).cuda()
    r = None
    if with_r:
        rr = (np.random.default_rng(7).uniform(1e4, 1e5, (B, 1)) * np.ones((1, T))).astype(np.float32)
        rr[:, 700:] *= 0.5
        r = torch.from_numpy(rr).cuda()
    outs = []
    for opts in (0, 8):  # 8 = never time-parallel
        prev = dwdf.set_option(opts)
        try:
            circ = make_circuit(dwdf, model_json(dwdf, nnv, name), "python", with_r=with_r)
            outs.append(circ.forward(x, r=r).clone())
        finally:
            dwdf.set_option(prev)
    assert seq_rel_err(outs[0].cpu().numpy(), outs[1].cpu().numpy()) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["2x8", "4x8", "2x16"])
@pytest.mark.parametrize("ordering", ["plugin", "python"])
def test_time_parallel_adjoint_equals_serial(dwdf, nnv, name, ordering):
    """The two-phase time-parallel adjoint (affine maps per chunk, composed, then accumulation from the true incoming
    adjoint) gives the one-lane-per-pair kernel's weight gradients and loss, and both agree with fp64 autograd."""
    B, T = 7, 1100
    xn = make_inputs(B, T, seed=81)
    x = torch.from_numpy(xn).cuda()
    w, sizes = nnv[f"{name}_weights"], [int(v) for v in nnv[f"{name}_sizes"]]
    order = nn.ORDER_PLUGIN if ordering == "plugin" else nn.ORDER_PYTHON
    target = (0.8 * nn.nn_clipper_forward(xn, w, sizes, 48000.0, 47000.0, 2.2e-9, order, dtype=np.float64) + 0.01).astype(np.float32)
    outs = []
    for opts in (0, 8):
        prev = dwdf.set_option(opts)
        try:
            circ = make_circuit(dwdf, model_json(dwdf, nnv, name), ordering)
            circ.forward(x)
            res = circ.backward(target=torch.from_numpy(target).cuda(), loss="mse+esr", skip=50)
            outs.append((res["grads"].cpu().numpy().copy(), float(res["loss"])))
        finally:
            dwdf.set_option(prev)
    (g_tp, l_tp), (g_se, l_se) = outs
    assert np.max(np.abs(g_tp - g_se)) < 2e-5 * np.max(np.abs(g_se)) and abs(l_tp / l_se - 1) < 1e-5
    ref = nn.nn_clipper_grad_torch(xn, target, w, sizes, 48000.0, 47000.0, 2.2e-9, order, loss="mse+esr", skip=50)
    assert np.max(np.abs(g_tp - ref["grad_w"])) < NN_GRAD_TOL * np.max(np.abs(ref["grad_w"]))


@pytest.mark.gpu
@pytest.mark.parametrize("ordering,order", [("plugin", nn.ORDER_PLUGIN), ("python", nn.ORDER_PYTHON)])
@pytest.mark.parametrize("with_r", [False, True])
@pytest.mark.parametrize("T", [150, 1100], ids=["one_lane_per_pair", "time_parallel"])
def test_dl_dx_against_autograd(dwdf, nnv, ordering, order, with_r, T):
    """dL/dx of the neural-root circuit (what tape.gradient would feed to a layer in front of the clipper) against
    fp64 autograd, on both adjoint kernels (the long sequences run the two-phase time-parallel one)."""
    B = 6
    x = make_inputs(B, T, seed=33)
    gy = np.random.default_rng(5).standard_normal((B, T)).astype(np.float32)
    r = None
    if with_r:
        r = (np.random.default_rng(6).uniform(1e4, 1e5, (B, 1)) * np.ones((1, T))).astype(np.float32)
        r[:, T // 3:] *= 1.5
    w, sizes = nnv["2x8_weights"], [int(v) for v in nnv["2x8_sizes"]]
    circ = make_circuit(dwdf, model_json(dwdf, nnv, "2x8"), ordering, with_r=with_r)
    circ.forward(torch.from_numpy(x).cuda(), r=None if r is None else torch.from_numpy(r).cuda())
    res = circ.backward(gy=torch.from_numpy(gy).cuda(), want_gx=True)
    ref = nn.nn_clipper_grad_torch(x, None, w, sizes, 48000.0, 47000.0, 2.2e-9, order, r=r, gy=gy, want_gx=True)
    gx = res["gx"].cpu().numpy()
    assert np.max(np.abs(gx - ref["gx"])) < 1e-4 * np.max(np.abs(ref["gx"]))
    assert np.max(np.abs(res["grads"].cpu().numpy() - ref["grad_w"])) < NN_GRAD_TOL * np.max(np.abs(ref["grad_w"]))


@pytest.mark.gpu
@pytest.mark.parametrize("loss", ["mse", "mse+esr"])
def test_raw_sums_finalize_and_one_call_step(dwdf, nnv, loss):
    """The multi-GPU split (raw sums -> [all-reduce] -> finalize) equals the fused call, and dwdf_train_step_neural equals
    forward + backward + Adam on the weights."""
    x = torch.from_numpy(make_inputs(40, 300, seed=41)).cuda()
    target = (0.6 * torch.tanh(2.0 * x)).contiguous()
    mj = model_json(dwdf, nnv, "2x8")
    a = make_circuit(dwdf, mj, "python")
    a.forward(x)
    ra = a.backward(target=target, loss=loss, skip=50)
    ga, outa = ra["grads"].clone(), ra["out"].clone()
    b = make_circuit(dwdf, mj, "python")
    b.forward(x)
    b.backward(target=target, skip=50, raw=True)
    assert float(b.out[23]) == 40 * 250
    rb = b.finalize(target=True, loss=loss)
    assert torch.equal(rb["grads"], ga) and torch.equal(rb["out"][16:19], outa[16:19])
    # one-call step
    opt_a = dwdf.AdamWeights(a, lr=1e-3, beta_1=0.5)
    opt_a.apply()
    c = make_circuit(dwdf, mj, "python")
    opt_c = dwdf.AdamWeights(c, lr=1e-3, beta_1=0.5)
    rc = c.train_step(x, target, opt_c, loss=loss, skip=50)
    assert torch.equal(rc["grads"], ga) and torch.equal(rc["out"][16:19], outa[16:19])
    assert torch.equal(c.weights, a.weights) and int(opt_c.step_count) == 1
