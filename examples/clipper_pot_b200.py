# %%
'''Script for training a diode model on actual diode clipper circuit data — wdf_py/diode_clipper/clipper_pot.py on the B200 engine.

Section by section the reference's script; what changes is marked `# B200:` (the imports, the per-sample loop inside
ClipperModel.forward, and GradientTape / apply_gradients). Without /root/reference (a GPU box) it trains on the excerpt of the
same recordings under examples/data/ (examples/data/make_data.py).

    python examples/clipper_pot_b200.py [--epochs 100] [--out model.json] [--json report.json]
'''

import argparse
import importlib
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
from tqdm import tqdm

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
wdf = importlib.import_module("differentiable-wdfs_b200")  # B200: instead of `import tf_wdf as wdf` + layers + model_utils + dataimport
DenseRootModel = wdf.DenseRootModel
load_diode_data = wdf.dataimport.load_diode_data
diode_1n4148_1u1d = wdf.diode_config.diode_1n4148_1u1d

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=100)
ap.add_argument("--out", default=None, help="where to write the trained model (RTNeural JSON, clipper_pot.py:298-331)")
ap.add_argument("--json", default=None, help="write the final metrics as JSON")
ap.add_argument("--csv-samples", type=int, default=-1)
args = ap.parse_args()

HERE = Path(__file__).resolve().parent
BASE_DIR = Path("/root/reference")

# %%
# Model parameter:
n_layers = 2
layer_size = 16
diode = diode_1n4148_1u1d
training_number = 2000

pretrained_model = f"{diode.name}_{n_layers}x{layer_size}_pretrained"
model_name = f"{diode.name}_{n_layers}x{layer_size}_training_{training_number}"

# %%
# Load circuit data:
C_val = 4.7e-9
if BASE_DIR.exists():
    train_data, train_N, val_data, val_N, FS = load_diode_data(diode, BASE_DIR, csv_samples=args.csv_samples)
    models_dir = BASE_DIR / "wdf_py" / "diode_clipper" / "models"
    pretrained_path = models_dir / "pretrained" / f"{pretrained_model}_model.json"
    reference_trained_path = models_dir / f"{model_name}.json"
else:  # B200: the excerpt of the same five recordings; same split (36 k < R < 73 k -> validation, dataimport.py:98)
    z = np.load(HERE / "data" / "1N4148_1up1down_excerpt.npz")
    FS = float(z["FS"])
    blocks = {"train": [], "val": []}
    for key in sorted(k for k in z.files if k != "FS"):
        R_val = float(key.partition("k")[0])
        d = z[key]
        blocks["train" if (R_val < 36 or R_val > 73) else "val"].append(np.stack([d[:, 0], np.full(len(d), R_val * 1000.0, np.float32), d[:, 1]]))
    train_data, val_data = np.concatenate(blocks["train"], 1), np.concatenate(blocks["val"], 1)
    train_N, val_N = train_data.shape[1], val_data.shape[1]
    pretrained_path = HERE / "data" / "pretrained_2x16.json"
    reference_trained_path = HERE / "data" / "reference_trained_2x16.json"

print(train_data.shape)
print(val_data.shape)

# %%
# Batch data:
batch_size = 2048


def batch_data(data, N):
    x = data[0]
    R_data = data[1]
    y_ref = data[2]
    n_batches = N // batch_size

    data_in = np.stack([x, R_data], axis=0).transpose()
    data_in_trim = data_in[: (n_batches * batch_size), :]
    data_in_batched = np.stack(np.array_split(data_in_trim, n_batches))

    data_target = np.transpose(np.array([y_ref]))
    data_target_trim = data_target[: (n_batches * batch_size), :]
    data_target_batched = np.stack(np.array_split(data_target_trim, n_batches))
    return data_in_batched, data_target_batched


train_X, train_Y = batch_data(train_data, train_N)
val_X, val_Y = batch_data(val_data, val_N)
# B200: the batches live on the device, channels split into the two contiguous (B, T) planes the kernels read
dev = torch.device("cuda")
to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)
train_x, train_r, train_y = to_dev(train_X[..., 0]), to_dev(train_X[..., 1]), to_dev(train_Y[..., 0])
val_x, val_r, val_y = to_dev(val_X[..., 0]), to_dev(val_X[..., 1]), to_dev(val_Y[..., 0])


# %%
# Define WDF circuit model:
class ClipperModel:
    def __init__(self, json):
        self.Vs = wdf.ResistiveVoltageSource(45.0e3)
        self.C = wdf.Capacitor(C_val, FS)
        self.P1 = wdf.Parallel(self.Vs, self.C)

        self.model = DenseRootModel(json)
        # B200: the per-sample loop of clipper_pot.py:110-124 (set_voltage / set_resistance / calc_impedance / reflected / root /
        # incident / voltage(C), ~40 eager ops per sample) is compiled once into the fused kernels
        self.circuit = wdf.compile_circuit(self.model, tree=self.P1, probe=self.C, ordering="python", r_element=self.Vs)

    def forward(self, x, r):
        return self.circuit.forward(x, r=r)  # (B, T); .forward_time_major gives the TensorArray layout (T, B, 1)


# %%
# Load pre-trained model:
with open(pretrained_path, "r") as read_file:
    model_json = json.load(read_file)

model = ClipperModel(model_json)

# %%
# Define loss functions: mse_loss + esr_loss (clipper_pot.py:141-177). The training loop calls loss_func(outs, train_Y) (:248): the
# model output sits in esr_loss's `target_y` slot, so the energy in the denominator is the prediction's — loss="mse+esr_as_called"
# reproduces exactly that (value and gradient); loss="mse+esr" would be the textbook error-to-signal ratio (energy of the target)
LOSS = "mse+esr_as_called"
optimizer = wdf.AdamWeights(model.circuit, lr=0.0001, beta_1=0.5, beta_2=0.999)  # B200: was tf.keras.optimizers.Adam(0.0001, 0.5, 0.999)

# %%
# Set up training history:
skip_samples = 50  # skip the first few samples to let state build up
history = {"loss": [], "mse": [], "esr": [], "val_loss": [], "val_mse": [], "val_esr": []}


def evaluate(m, x, r, y):
    """loss / mse / esr of a model on a data set, through the same fused loss the training uses"""
    m.forward(x, r)
    res = m.circuit.backward(target=y, loss=LOSS, skip=skip_samples)
    return float(res["loss"]), float(res["mse"]), float(res["esr"])


# %%
# Training loop:
for epoch in tqdm(range(args.epochs)):
    outs = model.forward(train_x, train_r)  # B200: was `with tf.GradientTape() as tape:` + the eager loop
    res = model.circuit.backward(target=train_y, loss=LOSS, skip=skip_samples)  # B200: was tape.gradient(loss, model.trainable_variables)
    history["loss"].append(float(res["loss"]))
    history["mse"].append(float(res["mse"]))
    history["esr"].append(float(res["esr"]))
    optimizer.apply()  # B200: was optimizer.apply_gradients(zip(grads, model.trainable_variables))

    if epoch % 5 == 0 or epoch == args.epochs - 1:
        val_loss, val_mse, val_esr = evaluate(model, val_x, val_r, val_y)
        history["val_loss"].append(val_loss)
        history["val_mse"].append(val_mse)
        history["val_esr"].append(val_esr)
        if epoch % 25 == 0:
            print(f"\nCheckpoint (Epoch = {epoch}):")
            print(f"    Loss: {history['loss'][-1]}")
            print(f"    Val Loss: {val_loss}")

print(f"\nFinal Results:")
print(f"    Loss: {history['loss'][-1]}")

# %%
# Next to the reference's own result: its trained model of the same shape, evaluated the same way on the same validation data
with open(reference_trained_path, "r") as read_file:
    reference_model = ClipperModel(json.load(read_file))
pre = ClipperModel(model_json)
report = {
    "data": "reference recordings" if BASE_DIR.exists() else "excerpt (examples/data)", "train_windows": int(train_x.shape[0]), "val_windows": int(val_x.shape[0]), "epochs": args.epochs,
    "pretrained": dict(zip(("val_loss", "val_mse", "val_esr"), evaluate(pre, val_x, val_r, val_y))),
    "fine_tuned_here": dict(zip(("val_loss", "val_mse", "val_esr"), evaluate(model, val_x, val_r, val_y))),
    "reference_trained_json": dict(zip(("val_loss", "val_mse", "val_esr"), evaluate(reference_model, val_x, val_r, val_y))),
    "train_loss_first_last": [history["loss"][0], history["loss"][-1]],
}
print(json.dumps(report, indent=2))
if args.json:
    with open(args.json, "w") as f:
        json.dump(report, f, indent=2)

# %%
# Save final model weights (clipper_pot.py:298-331: the RTNeural JSON the plugin embeds):
if args.out:
    wdf.model_io.save_model_json(wdf.model_io.json_from_weights(model.circuit.weights.cpu().numpy(), model.model.sizes), args.out)
