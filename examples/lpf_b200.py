# %%
'''Example script for using differentiable WDFs to determine the parameters of an RC lowpass filter —
wdf_py/simple_circuits/lpf.py on the B200 engine. What changes against the reference is marked `# B200:`.

    python examples/lpf_b200.py [--epochs 100] [--json report.json]
'''

import argparse
import importlib
import json
import sys
from pathlib import Path

import numpy as np
import scipy.signal as signal
import torch
import tqdm as tqdm

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
wdf = importlib.import_module("differentiable-wdfs_b200")  # B200: instead of `import tf_wdf as wdf`

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=100)
ap.add_argument("--json", default=None)
args = ap.parse_args()

FS = 48000


# %%
# Construct Differentiable WDF circuit model:
class Model:
    def __init__(self):
        self.Vs = wdf.IdealVoltageSource()
        self.R1 = wdf.Resistor(1000, True)
        self.C1 = wdf.Capacitor(1.0e-6, FS, True)

        self.S1 = wdf.Series(self.R1, self.C1)
        self.I1 = wdf.Inverter(self.S1)
        # B200: lpf.py:30-49's loop (calc_impedance, then per sample set_voltage / incident / reflected / voltage(C1)) compiled once
        self.circuit = wdf.compile_circuit(self.Vs, tree=self.I1, probe=self.C1)

    def forward(self, input):
        return self.circuit.forward_time_major(input)  # (T, B, 1): what output_sequence.stack() returns

    @property
    def trainable_variables(self):  # lpf.py:98-99 indexes grads[0] -> C, grads[1] -> R
        return [self.circuit.slot(self.C1, "C"), self.circuit.slot(self.R1, "R")]


# %%
# Generate data: (audio_dspy 0.0.4 is not installable here; the two helpers it provides are restated)
batch_size = 256
n_batches = 5
freq = 720


def sweep_log(f0, f1, duration, fs):  # adsp.sweep_log: exponential sine sweep
    N = int(duration * fs)
    n = np.arange(N)
    beta = N / np.log(f1 / f0)
    phase = 2 * np.pi * beta * f0 * (np.power(f1 / f0, n / N) - 1.0)
    return np.sin((phase + np.pi / 180) / fs)


def design_LPF1(fc, fs):  # adsp.design_LPF1: first-order lowpass, bilinear transform with pre-warping
    wc = 2 * np.pi * fc
    c = wc / np.tan(wc / (2.0 * fs))
    a0 = c + wc
    return np.array([wc / a0, wc / a0]), np.array([1.0, (wc - c) / a0])


sweep = sweep_log(100, 10000, (batch_size * n_batches) / FS, FS)[: batch_size * n_batches]
b, a = design_LPF1(720, FS)
sweep_filt = signal.lfilter(b, a, sweep)
data_in = np.array([sweep])
data_target = np.transpose(np.array([sweep_filt]))
# B200: on the device, (B, T) float32
x_dev = torch.from_numpy(data_in.astype(np.float32)).cuda()
t_dev = torch.from_numpy(data_target[:, 0].astype(np.float32))[None, :].cuda().contiguous()

# %%
# Training loop:
model = Model()
iC, iR = model.trainable_variables
# B200: two Keras Adam optimizers with their own rates (lpf.py:79-80) = one Adam with a rate per slot
optimizer = wdf.Adam(model.circuit, lr={iR: 25.0, iC: 10.0e-9})

Rs, Cs, losses = [], [], []
for epoch in tqdm.tqdm(range(args.epochs)):
    outs = model.forward(x_dev)[..., 0]  # B200: was inside `with tf.GradientTape() as tape:`
    res = model.circuit.backward(target=t_dev, loss="mse")  # B200: was loss_func(outs, data_target); tape.gradient(loss, model.trainable_variables)
    grads = [res["grads"][iC], res["grads"][iR]]

    if epoch % 25 == 0:
        print(f"\nCheckpoint (Epoch = {epoch}):")
        print(f"    Loss: {float(res['loss'])}")
        print(f"    Grads: {[float(g) for g in grads]}")
        print(f"    Trainables: {[float(model.circuit.params[i]) for i in (iC, iR)]}")

    optimizer.apply()  # B200: was R_optimizer.apply_gradients / C_optimizer.apply_gradients

    Rs.append(float(model.circuit.params[iR]))
    Cs.append(float(model.circuit.params[iC]))
    losses.append(float(res["loss"]))

# %%
# Print results:
final_freq = 1.0 / (2 * np.pi * Rs[-1] * Cs[-1])
print(final_freq)
report = {"epochs": args.epochs, "loss_first_last": [losses[0], losses[-1]], "R": Rs[-1], "C": Cs[-1], "cutoff_hz": final_freq, "target_cutoff_hz": freq, "start_cutoff_hz": 1.0 / (2 * np.pi * 1000 * 1.0e-6)}
print(json.dumps(report, indent=2))
if args.json:
    with open(args.json, "w") as f:
        json.dump(report, f, indent=2)
