"""TEST INFRASTRUCTURE ONLY — torch restatement of ``wdf_py/lib/tf_wdf.py`` (the Python reference).

TensorFlow 2.5 (``requirements.txt:1``) cannot be installed here (Python 3.12, no network), so the
Python half of the reference cannot be imported. This module restates its element/adaptor classes
on torch tensors, method for method, and drives them with the scripts' own per-sample Python loops
(``lpf.py:38-46``, ``clipper_pot.py:103-127``). ``torch.autograd`` over those loops stands in for
``tf.GradientTape`` and is the *gradient oracle* (independent of the analytic adjoint in
``wdf_oracle.c``). It is also the "wdf_py TensorFlow CPU path" stand-in that bench.py times
 — labelled as a stand-in everywhere it is reported.

Gradients w.r.t. (Is, nabla, R, C): parity unpinned by the reference (it never asserts any
gradient); the derivative of the Wright-omega function is defined as w/(1+w).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.special import wrightomega as _scipy_wrightomega

from . import cpu as _cpu


class WrightOmega(torch.autograd.Function):
    """omega(x) with custom gradient omega/(1+omega).

    kind "exact": scipy.special.wrightomega in fp64 (what diode_pretraining.py:8,57-58 calls; same
    TOMS-917 algorithm as modules/toms917). kind "approx": omega4 of omega.h:172-177 evaluated by the
    C oracle in the tensor's own precision.
    """

    @staticmethod
    def forward(ctx, x, kind="exact"):
        xn = x.detach().cpu().numpy()
        if kind == "exact":
            w = np.real(_scipy_wrightomega(xn.astype(np.float64))).astype(xn.dtype)
        else:
            w = _oracle().omega("omega4", xn, dtype=xn.dtype.type).reshape(xn.shape)
        w = torch.from_numpy(np.asarray(w)).to(x.dtype)
        ctx.save_for_backward(w)
        return w

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        return g * w / (1 + w), None


_ORACLE = None


def _oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = _cpu.Oracle()
    return _ORACLE


def voltage(wdf):
    """tf_wdf.py:8-10"""
    return (wdf.a + wdf.b) * 0.5


class IdealVoltageSource:
    """tf_wdf.py:13-28"""

    def __init__(self):
        self.a = torch.zeros(1)
        self.b = torch.zeros(1)

    def set_voltage(self, voltage):
        self.Vs = voltage

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = -self.a + 2.0 * self.Vs
        return self.b


class ResistiveVoltageSource:
    """tf_wdf.py:31-59"""

    def __init__(self, initial_R=1.0e-9, trainable=False, dtype=torch.float64):
        self.a = torch.zeros(1, dtype=dtype)
        self.b = torch.zeros(1, dtype=dtype)
        self.R = torch.tensor(initial_R, dtype=dtype, requires_grad=trainable)

    def calc_impedance(self):
        pass

    def reset(self):
        self.a = torch.zeros(1, dtype=self.a.dtype)

    def set_voltage(self, voltage):
        self.Vs = voltage

    def set_resistance(self, resistance):
        self.R = resistance

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = self.Vs * torch.ones_like(self.a)
        return self.b


class Resistor:
    """tf_wdf.py:62-88"""

    def __init__(self, initial_R, trainable=False, dtype=torch.float64):
        self.a = torch.zeros(1, dtype=dtype)
        self.b = torch.zeros(1, dtype=dtype)
        self.R = torch.tensor(float(initial_R), dtype=dtype, requires_grad=trainable)

    def calc_impedance(self):
        pass

    def set_resistance(self, resistance):
        self.R = resistance

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = torch.zeros_like(self.a)
        return self.b


class Capacitor:
    """tf_wdf.py:91-126"""

    def __init__(self, initial_C, FS, trainable=False, dtype=torch.float64):
        self.a = torch.zeros(1, dtype=dtype)
        self.b = torch.zeros(1, dtype=dtype)
        self.FS = FS
        self.C = torch.tensor(float(initial_C), dtype=dtype, requires_grad=trainable)
        self.R = torch.tensor(1.0 / (2.0 * initial_C * FS), dtype=dtype)
        self.z = torch.zeros(1, dtype=dtype)

    def calc_impedance(self):
        self.R = torch.reciprocal(self.C * (2.0 * self.FS))

    def reset(self):
        self.z = torch.zeros(1, dtype=self.a.dtype)

    def incident(self, x):
        self.a = x
        self.z = self.a

    def reflected(self):
        self.b = self.z
        return self.b


class Series:
    """tf_wdf.py:129-155"""

    def __init__(self, P1, P2):
        self.a = torch.zeros(1)
        self.b = torch.zeros(1)
        self.P1 = P1
        self.P2 = P2

    def calc_impedance(self):
        self.P1.calc_impedance()
        self.P2.calc_impedance()
        self.R = self.P1.R + self.P2.R
        self.p1R = self.P1.R / self.R
        self.p2R = self.P2.R / self.R

    def incident(self, x):
        b1 = self.P1.b - self.p1R * (x + self.P1.b + self.P2.b)
        self.P1.incident(b1)
        self.P2.incident(-(x + b1))
        self.a = x

    def reflected(self):
        self.b = -(self.P1.reflected() + self.P2.reflected())
        return self.b


class Parallel:
    """tf_wdf.py:158-192"""

    def __init__(self, P1, P2):
        self.a = torch.zeros(1)
        self.b = torch.zeros(1)
        self.P1 = P1
        self.P2 = P2

    def calc_impedance(self):
        self.P1.calc_impedance()
        self.P2.calc_impedance()
        G1 = 1.0 / self.P1.R
        G2 = 1.0 / self.P2.R
        G = G1 + G2
        self.R = 1.0 / G
        self.p1R = G1 / G

    def incident(self, x):
        b2 = x + self.b_temp
        self.P1.incident(self.b_diff + b2)
        self.P2.incident(b2)
        self.a = x

    def reflected(self):
        b1 = self.P1.reflected()
        b2 = self.P2.reflected()
        self.b_diff = b2 - b1
        self.b_temp = -self.p1R * self.b_diff
        self.b = b2 + self.b_temp
        return self.b


class Inverter:
    """tf_wdf.py:195-214"""

    def __init__(self, P1):
        self.a = torch.zeros(1)
        self.b = torch.zeros(1)
        self.P1 = P1

    def calc_impedance(self):
        self.P1.calc_impedance()
        self.R = self.P1.R

    def incident(self, x):
        self.P1.incident(-x)
        self.a = x

    def reflected(self):
        self.b = -self.P1.reflected()
        return self.b


class DiodePair:
    """Analytic diode-pair root (new Python surface).

    Law: diode_pretraining.py:39-60 (Werner eq. 45; reduces to eq. 39 = wdf_t.h:917-924 when
    N_up == N_down == 1). Constructor modelled on DiodePairT (wdf_t.h:868-872) + DiodeConfig
    (diode_config.py:5-9).
    """

    def __init__(self, next, Is, Vt=25.85e-3, nabla=1.0, N_up=1, N_down=1, trainable=False, mode="exact", dtype=torch.float64):
        self.next = next
        self.Is = torch.tensor(float(Is), dtype=dtype, requires_grad=trainable)
        self.nabla = torch.tensor(float(nabla), dtype=dtype, requires_grad=trainable)
        self.Vt0 = float(Vt)
        self.N_up = float(N_up)
        self.N_down = float(N_down)
        self.mode = mode
        self.a = torch.zeros(1, dtype=dtype)
        self.b = torch.zeros(1, dtype=dtype)

    def incident(self, x):
        self.a = x

    def reflected(self):
        a = self.a
        Vt = self.Vt0 * self.nabla
        R_Is_overVt = self.Is * self.next.R / Vt
        pos = a >= 0
        mu0 = torch.where(pos, torch.full_like(a, self.N_down), torch.full_like(a, self.N_up))
        mu1 = torch.where(pos, torch.full_like(a, self.N_up), torch.full_like(a, self.N_down))
        lamb = torch.sign(a)
        u0 = torch.log(R_Is_overVt / mu0) + lamb * a / (mu0 * Vt)
        u1 = torch.log(R_Is_overVt / mu1) - lamb * a / (mu1 * Vt)
        w0 = WrightOmega.apply(u0, self.mode)
        w1 = WrightOmega.apply(u1, self.mode)
        self.b = a - 2 * Vt * lamb * (mu0 * w0 - mu1 * w1)
        return self.b


# ---- the scripts' forward loops ------------------------------------------------------------------


def lpf_forward(x, R=1000.0, C=1.0e-6, FS=48000.0, dtype=torch.float64, probe="C1"):
    """lpf.py:20-49 (Model + forward). x: (B, T). Returns (y (T, B, 1), dict of trainables)."""
    Vs = IdealVoltageSource()
    R1 = Resistor(R, True, dtype)
    C1 = Capacitor(C, FS, True, dtype)
    S1 = Series(R1, C1)
    I1 = Inverter(S1)
    inp = torch.as_tensor(x, dtype=dtype).unsqueeze(-1)
    outs = []
    I1.calc_impedance()
    for i in range(inp.shape[1]):
        Vs.set_voltage(inp[:, i])
        Vs.incident(I1.reflected())
        I1.incident(Vs.reflected())
        outs.append(voltage(C1 if probe == "C1" else R1))
    return torch.stack(outs), {"R": R1.R, "C": C1.C}


def clipper_forward(x, p: _cpu.ClipperParams = _cpu.ClipperParams(), mode="exact", ordering=_cpu.ORDER_PYTHON, dtype=torch.float64, r_in=None):
    """ClipperModel.forward of clipper_pot.py:94-127 with the analytic DiodePair as root.

    x: (B, T); r_in: optional (B, T) per-sample resistance channel (clipper_pot.py:116).
    Returns (y (B, T), dict of leaf tensors Is, nabla, R, C with requires_grad).
    """
    Vs = ResistiveVoltageSource(p.R, True, dtype)
    C = Capacitor(p.C, p.fs, True, dtype)
    P1 = Parallel(Vs, C)
    dp = DiodePair(P1, p.Is, p.Vt, p.nabla, p.n_up, p.n_down, True, mode, dtype)
    inp = torch.as_tensor(x, dtype=dtype).unsqueeze(-1)
    Vs.reset()
    C.reset()
    outs = []
    if r_in is None:
        P1.calc_impedance()
    for i in range(inp.shape[1]):
        Vs.set_voltage(inp[:, i])
        if r_in is not None:
            Vs.set_resistance(torch.as_tensor(r_in[:, i], dtype=dtype).unsqueeze(-1))
            P1.calc_impedance()
        dp.incident(P1.reflected())
        if ordering == _cpu.ORDER_PLUGIN:
            outs.append(voltage(C) if i > 0 else torch.zeros_like(inp[:, 0]))
        P1.incident(dp.reflected())
        if ordering != _cpu.ORDER_PLUGIN:
            outs.append(voltage(C))
    y = torch.stack(outs, dim=1)[..., 0]
    return y, {"Is": dp.Is, "nabla": dp.nabla, "R": Vs.R, "C": C.C}


def mse_esr_loss(target, pred, with_esr=True):
    """clipper_pot.py:141-156,176-177: MeanSquaredError + esr_loss (eps = float64 eps, :145), arguments in esr_loss's own
    order: the energy is the FIRST argument's. The reference's training loop calls ``loss_func(outs, train_Y)`` (:248), i.e.
    with the model output first — ``mse_esr_loss(outs, train_Y)`` restates that call (DWDF_LOSS_MSE_ESR_AS_CALLED),
    ``mse_esr_loss(train_Y, outs)`` the textbook error-to-signal ratio (DWDF_LOSS_MSE_ESR)."""
    mse = torch.mean((target - pred) ** 2)
    if not with_esr:
        return mse
    eps = np.finfo(float).eps
    sse = torch.sum((target - pred) ** 2)
    energy = torch.sum(target ** 2)
    N = target.shape[0] * target.shape[1]
    return mse + torch.sqrt(sse / (energy + eps) / N)
