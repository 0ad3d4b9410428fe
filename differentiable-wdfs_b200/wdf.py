"""The wdf_py element/adaptor API (wdf_py/lib/tf_wdf.py) over the B200 engine.

Same class names, constructor signatures, attributes (``a``, ``b``, ``R``, ``P1``, ``P2``, ``Vs``,
``z``, ``C``, ``FS``, ``p1R``) and methods (``incident``, ``reflected``, ``calc_impedance``,
``set_voltage``, ``set_resistance``, ``reset``) as the reference, plus the analytic ``DiodePair``
root the north-star adds (semantics: wdf_t.h:859-985, Toms917DiodePair.h, diode_pretraining.py:39-60)
and ``PolarityInverter`` as the C++ name of ``Inverter`` (wdf_t.h:558).

Two ways to run a circuit built from these objects:

* **imperative**, exactly like the reference scripts — ``root.incident(tree.reflected());
  tree.incident(root.reflected())`` once per sample in a Python loop. The element methods are the
  reference's own one-line wave equations on torch tensors; this is the API surface the scripts are
  written against (config 1, RC low-pass plumbing) and is as slow as the reference's eager loop.
* **compiled** — ``compile_circuit(root, tree, probe)`` lowers the object graph to the flat
  post-order program of ``include/dwdf.h`` and returns a :class:`CompiledCircuit` whose
  ``forward`` / ``backward`` / ``train_pass`` run whole (B, T) batches in the fused sm_100a kernels.
  That path is CUDA-only and raises if the extension or the GPU is missing — it never falls back
  to the imperative code.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Optional

import torch

from . import _lib as L
from . import model_io


def voltage(wdf):
    """tf_wdf.py:8-10"""
    return (wdf.a + wdf.b) * 0.5


def current(wdf):
    """wdf_t.h:1119-1123: the current through a node, (a - b) / (2 R)"""
    return (wdf.a - wdf.b) * (0.5 / wdf.R)


def _scalar(v, trainable=False):
    t = torch.tensor(float(v), dtype=torch.float32)
    t.requires_grad_(bool(trainable))
    return t


class _Element:
    trainable = False

    def __init__(self):
        self.a = torch.zeros(1)
        self.b = torch.zeros(1)


class IdealVoltageSource(_Element):
    """tf_wdf.py:13-28 (root): b = -a + 2 Vs"""

    def __init__(self):
        super().__init__()
        self.Vs = torch.zeros(1)

    def set_voltage(self, voltage):
        self.Vs = voltage

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = -self.a + 2.0 * self.Vs
        return self.b


class ResistiveVoltageSource(_Element):
    """tf_wdf.py:31-59: b = Vs, port resistance R (settable per sample, clipper_pot.py:116)"""

    def __init__(self, initial_R=1.0e-9, trainable=False):
        super().__init__()
        self.trainable = trainable
        self.R = _scalar(initial_R, trainable)
        self.Vs = torch.zeros(1)

    def calc_impedance(self):
        pass

    def reset(self):
        self.a = torch.zeros(1)

    def set_voltage(self, voltage):
        self.Vs = voltage

    def set_resistance(self, resistance):
        self.R = resistance

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = self.Vs * torch.ones_like(self.a) if torch.is_tensor(self.Vs) else torch.full_like(self.a, float(self.Vs))
        return self.b


class Resistor(_Element):
    """tf_wdf.py:62-88: b = 0; R trainable, clipped to [180, 1e6] after each optimizer step (:74)"""

    clip = (180.0, 1.0e6)

    def __init__(self, initial_R, trainable=False):
        super().__init__()
        self.trainable = trainable
        self.R = _scalar(initial_R, trainable)

    def calc_impedance(self):
        pass

    def set_resistance(self, resistance):
        self.R = resistance

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = torch.zeros_like(self.a)
        return self.b


class Capacitor(_Element):
    """tf_wdf.py:91-126: b = z, z <- a; R = 1/(2 C FS); C trainable, clipped to [1e-13, 1] (:104)"""

    clip = (1.0e-13, 1.0)

    def __init__(self, initial_C, FS, trainable=False):
        super().__init__()
        self.trainable = trainable
        self.FS = FS
        self.C = _scalar(initial_C, trainable)
        self.R = torch.tensor(1.0 / (2.0 * float(initial_C) * FS), dtype=torch.float32)
        self.z = torch.zeros(1)

    def calc_impedance(self):
        self.R = torch.reciprocal(self.C * (2.0 * self.FS))

    def reset(self):
        self.z = torch.zeros(1)

    def incident(self, x):
        self.a = x
        self.z = self.a

    def reflected(self):
        self.b = self.z
        return self.b


class Series(_Element):
    """tf_wdf.py:129-155"""

    def __init__(self, P1, P2):
        super().__init__()
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


class Parallel(_Element):
    """tf_wdf.py:158-192 (``incident`` uses the b_temp / b_diff left by the preceding ``reflected``)"""

    def __init__(self, P1, P2):
        super().__init__()
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


class Inverter(_Element):
    """tf_wdf.py:195-214"""

    def __init__(self, P1):
        super().__init__()
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


PolarityInverter = Inverter  # wdf_t.h:558


# ---- the remaining chowdsp_wdf elements (wdf_t.h): same protocol, C++ constructor signatures ---------------------

class Inductor(_Element):
    """wdf_t.h:280-348 (InductorT): b = -z, z <- a; R = 2 L FS. Trainable like the Capacitor (reverse mode implemented)."""

    def __init__(self, initial_L, FS, trainable=False):
        super().__init__()
        self.trainable = trainable
        self.FS = FS
        self.L = _scalar(initial_L, trainable)
        self.R = torch.tensor(2.0 * float(initial_L) * FS, dtype=torch.float32)
        self.z = torch.zeros(1)

    def calc_impedance(self):
        self.R = self.L * (2.0 * self.FS)

    def reset(self):
        self.z = torch.zeros(1)

    def incident(self, x):
        self.a = x
        self.z = self.a

    def reflected(self):
        self.b = -self.z
        return self.b


class CapacitorAlpha(_Element):
    """wdf_t.h:190-276 (CapacitorAlphaT): alpha transform, alpha = 1 bilinear ... 0 backward Euler.
    R = 1 / ((1 + alpha) C FS); b = b_coef b_prev + a_coef z."""

    def __init__(self, initial_C, FS, alpha=1.0):
        super().__init__()
        self.FS = FS
        self.C = _scalar(initial_C)
        self.alpha = float(alpha)
        self.z = torch.zeros(1)
        self.calc_impedance()

    def set_alpha(self, alpha):
        self.alpha = float(alpha)
        self.calc_impedance()

    def calc_impedance(self):
        self.b_coef = (1.0 - self.alpha) / 2.0
        self.a_coef = (1.0 + self.alpha) / 2.0
        self.R = torch.reciprocal(self.C * ((1.0 + self.alpha) * self.FS))

    def reset(self):
        self.z = torch.zeros(1)
        self.b = torch.zeros(1)

    def incident(self, x):
        self.a = x
        self.z = self.a

    def reflected(self):
        self.b = self.b_coef * self.b + self.a_coef * self.z
        return self.b


class InductorAlpha(_Element):
    """wdf_t.h:352-444 (InductorAlphaT): R = (1 + alpha) L FS; b = b_coef b_prev - a_coef z."""

    def __init__(self, initial_L, FS, alpha=1.0):
        super().__init__()
        self.FS = FS
        self.L = _scalar(initial_L)
        self.alpha = float(alpha)
        self.z = torch.zeros(1)
        self.calc_impedance()

    def set_alpha(self, alpha):
        self.alpha = float(alpha)
        self.calc_impedance()

    def calc_impedance(self):
        self.b_coef = (1.0 - self.alpha) / 2.0
        self.a_coef = (1.0 + self.alpha) / 2.0
        self.R = self.L * ((1.0 + self.alpha) * self.FS)

    def reset(self):
        self.z = torch.zeros(1)
        self.b = torch.zeros(1)

    def incident(self, x):
        self.a = x
        self.z = self.a

    def reflected(self):
        self.b = self.b_coef * self.b - self.a_coef * self.z
        return self.b


class ResistiveCurrentSource(_Element):
    """wdf_t.h:786-846 (ResistiveCurrentSourceT): current source with a parallel resistance, b = R Is."""

    def __init__(self, initial_R=1.0e9):
        super().__init__()
        self.R = _scalar(initial_R)
        self.Is = torch.zeros(1)

    def calc_impedance(self):
        pass

    def set_current(self, current):
        self.Is = current

    set_voltage = set_current  # (the compiled path drives whichever source the circuit has with x[n])

    def set_resistance(self, resistance):
        self.R = resistance

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = self.R * self.Is * torch.ones_like(self.a) if torch.is_tensor(self.Is) else self.R * float(self.Is) * torch.ones_like(self.a)
        return self.b


class YParameter(_Element):
    """wdf_t.h:597-654 (YParameterT): two-port described by its short-circuit admittances, port 1 = ``P1``."""

    def __init__(self, P1, y11, y12, y21, y22):
        super().__init__()
        self.P1 = P1
        self.y = (float(y11), float(y12), float(y21), float(y22))

    def calc_impedance(self):
        self.P1.calc_impedance()
        y11, y12, y21, y22 = self.y
        R1 = self.P1.R
        den = y22 + R1 * y11 * y22 - R1 * y12 * y21
        self.R = (R1 * y11 + 1.0) / den
        self.A = (-y22 * R1 * R1 * y11 * y11 + y12 * y21 * R1 * R1 * y11 + y22) / (den * (R1 * y11 + 1.0))
        self.B = -R1 * y12 / (R1 * y11 + 1.0)
        self.C = -y21 / den

    def incident(self, x):
        self.a = x
        self.P1.incident(self.A * self.P1.b + self.B * x)

    def reflected(self):
        self.b = self.C * self.P1.reflected()
        return self.b


class IdealCurrentSource(_Element):
    """wdf_t.h:746-784 (root): b = 2 R Is + a, R = the port resistance of ``next``."""

    def __init__(self, next=None):
        super().__init__()
        self.next = next
        self.Is = torch.zeros(1)

    def set_current(self, current):
        self.Is = current

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = 2.0 * self.next.R * self.Is + self.a
        return self.b


class Diode(_Element):
    """wdf_t.h:987-1072 (DiodeT, root): a single diode, Werner eq. 10 with omega4 (both of the reference's qualities):
    b = a + 2 R Is - 2 Vt omega4(ln(R Is / Vt) + a / Vt + R Is / Vt), Vt := nDiodes Vt."""

    def __init__(self, next, Is, Vt=25.85e-3, nDiodes=1.0):
        super().__init__()
        self.next = next
        self.Is = _scalar(Is)
        self.nabla = _scalar(nDiodes)
        self.Vt = float(Vt)

    def incident(self, x):
        self.a = x

    def reflected(self):
        V = self.Vt * self.nabla
        RIs = self.next.R * self.Is
        self.b = self.a + 2.0 * RIs - 2.0 * V * omega4_approx(torch.log(RIs / V) + self.a / V + RIs / V)
        return self.b


class Switch(_Element):
    """wdf_t.h:1076-1106 (SwitchT, root): b = -a closed (short), b = a open."""

    def __init__(self, next=None, closed=True):
        super().__init__()
        self.next = next
        self.closed = bool(closed)

    def set_closed(self, closed):
        self.closed = bool(closed)

    def incident(self, x):
        self.a = x

    def reflected(self):
        self.b = -self.a if self.closed else self.a
        return self.b


def _exp_approx(x):
    """omega.h:99-116 (exp_approx with pow2_approx :83-92) on float32 tensors"""
    xp = torch.clamp(x.float() * 1.442695040888963, min=-126.0)
    l = torch.floor(xp)
    f = xp - l
    p = 1.0 + f * (0.6931471805599453 + f * (0.2274112777602189 + f * 0.07944154167983575))
    return torch.ldexp(p, l.to(torch.int32))


def _log_approx(x):
    """omega.h:49-63 (log_approx with log2_approx :33-42): exponent + cubic on the mantissa"""
    m, e = torch.frexp(x.float())  # x = m 2^e, m in [0.5, 1)
    m, e = m * 2.0, (e - 1).float()
    p = -2.213475204444817 + m * (3.148297929334117 + m * (-1.098865286222744 + m * 0.1640425613334452))
    return 0.693147180559945 * (e + p)


def omega4_approx(x):
    """omega.h:159-177 (omega3 + one Newton step): the approximation wdft::DiodePairT / DiodeT evaluate"""
    x = torch.as_tensor(x, dtype=torch.float32)
    cubic = 6.313183464296682e-1 + x * (3.631952663804445e-1 + x * (4.775931364975583e-2 - x * 1.314293149877800e-3))
    y = torch.where(x < -3.341459552768620, torch.zeros_like(x), torch.where(x < 8.0, cubic, x - _log_approx(torch.clamp(x, min=1.0))))
    return y - (y - _exp_approx(x - y)) / (y + 1.0)



def wright_omega(x: torch.Tensor, iters: int = 3) -> torch.Tensor:
    """Wright omega on the real axis for the imperative API: start from e^x below -2 (the cubic of
    omega.h:160-167 touches zero at -3.34 and is useless as a start near there), the cubic up to 8 and
    x - ln x above, then Fritsch-Shafer-Crowley iterations on w + ln w = x (round-off in 3)."""
    xs = torch.clamp(x, min=-20.0)  # below: omega(x) = e^x (1 - e^x + ...) to round-off; the iteration's residual would cancel
    w = torch.where(xs < -2.0, torch.exp(xs), torch.where(xs < 8.0, 0.6313183464296682 + xs * (0.3631952663804445 + xs * (0.04775931364975583 - xs * 0.0013142931498778)),
                                                                       xs - torch.log(torch.clamp(xs, min=1.0))))
    for _ in range(iters):
        r = xs - w - torch.log(w)
        wp1 = w + 1.0
        q = 2.0 * wp1 * (wp1 + (2.0 / 3.0) * r)
        w = w * (1.0 + (r / wp1) * (q - r) / (q - 2.0 * r))
    e = torch.exp(torch.clamp(x, max=-20.0))
    return torch.where(x < -20.0, e - e * e, w)


class DiodePair(_Element):
    """Analytic antiparallel diode-pair root (new Python surface).

    Law: Werner eq. 45 as in diode_pretraining.py:39-60 (``N_up``/``N_down`` diodes per branch); with
    ``N_up == N_down == 1`` it is eq. 39 = wdf_t.h:917-924 / Toms917DiodePair.h:51-59. Constructor
    modelled on ``DiodePairT(next, Is, Vt, nDiodes)`` (wdf_t.h:868-872) and ``DiodeConfig``
    (diode_config.py:5-9: ``nabla`` multiplies Vt). ``mode``: 'approx' (omega4, omega.h:172-177),
    'exact' (Wright omega to fp32 round-off) or 'approx_good' (eq. 18, wdf_t.h:907-913) — the mode
    selects the fused kernels' root, and the imperative ``reflected()`` below evaluates the same law.
    """

    def __init__(self, next, Is, Vt=25.85e-3, nabla=1.0, N_up=1, N_down=1, trainable=False, mode="approx", newton_max_iter=0, newton_tol=0.0):
        super().__init__()
        if mode not in ("approx", "exact", "approx_good"):
            raise ValueError(f"unknown DiodePair mode {mode!r}")
        self.next = next
        self.trainable = trainable
        self.Is = _scalar(Is, trainable)
        self.nabla = _scalar(nabla, trainable)
        self.Vt = float(Vt)
        self.N_up = float(N_up)
        self.N_down = float(N_down)
        self.mode = mode
        self.newton_max_iter = int(newton_max_iter)
        self.newton_tol = float(newton_tol)

    def incident(self, x):
        self.a = x

    def reflected(self):
        a = self.a
        V = self.Vt * self.nabla
        k = self.Is * self.next.R / V
        pos = a >= 0
        mu0 = torch.where(pos, torch.full_like(a, self.N_down), torch.full_like(a, self.N_up))
        mu1 = torch.where(pos, torch.full_like(a, self.N_up), torch.full_like(a, self.N_down))
        lam = torch.sign(a)
        if self.mode == "approx_good":  # eq. 18, wdf_t.h:907-913
            RIs = self.Is * self.next.R
            self.b = a + 2 * lam * (RIs - V * omega4_approx(torch.log(k) + lam * a / V + k))
            return self.b
        omega = wright_omega if self.mode == "exact" else omega4_approx  # 'approx': what wdft::DiodePairT evaluates (omega.h:172-177)
        w0 = omega(torch.log(k / mu0) + lam * a / (mu0 * V))
        w1 = omega(torch.log(k / mu1) - lam * a / (mu1 * V))
        self.b = a - 2 * V * lam * (mu0 * w0 - mu1 * w1)
        return self.b


class DenseLayer:
    """layers.py:7-39: ``input @ kernel + bias`` (kernel (in, out), bias (out,))."""

    def __init__(self, in_size, out_size):
        self.kernel = torch.zeros(in_size, out_size)
        self.bias = torch.zeros(out_size)

    def set_weights(self, json_weights):
        W, b = json_weights
        self.kernel = torch.as_tensor(model_io._squeeze_to(W, 2))
        self.bias = torch.as_tensor(model_io._squeeze_to(b, 1))

    def __call__(self, input):
        return input @ self.kernel + self.bias


class DenseRootModel(_Element):
    """layers.py:42-82: the neural WDF root. Built from the RTNeural-style JSON dict (dense layers, tanh
    between them); ``incident(x)`` takes the (..., 2) tensor [a, log R] the scripts assemble
    (clipper_pot.py:119-120), ``reflected()`` returns the network output b_nn (the scripts feed
    ``-b_nn`` to the tree, :121). In a compiled circuit the same weights run in the fused kernel."""

    def __init__(self, json):
        super().__init__()
        self.json = json
        parsed, self.sizes = model_io.layers_from_json(json)
        self.layers = []
        self.activations = []
        for W, b, act in parsed:
            layer = DenseLayer(*W.shape)
            layer.set_weights([W, b])
            self.layers.append(layer)
            self.activations.append(act)
        self.model_in = None

    def weight_vector(self):
        return model_io.flatten_weights([(l.kernel.numpy(), l.bias.numpy(), a) for l, a in zip(self.layers, self.activations)])

    def incident(self, x):
        self.a = x[..., 0]
        self.model_in = x

    def reflected(self):
        x = self.model_in
        for layer, act in zip(self.layers, self.activations):
            x = layer(x)
            if act == "tanh":
                x = torch.tanh(x)
            elif act == "relu":
                x = torch.relu(x)
        self.b = x
        return self.b


# ==================================================================================================
# compiled path
# ==================================================================================================
_KIND = {Resistor: L.RESISTOR, Capacitor: L.CAPACITOR, ResistiveVoltageSource: L.RESISTIVE_VS, Series: L.SERIES, Parallel: L.PARALLEL, Inverter: L.INVERTER, Inductor: L.INDUCTOR,
         CapacitorAlpha: L.CAPACITOR_ALPHA, InductorAlpha: L.INDUCTOR_ALPHA, ResistiveCurrentSource: L.RESISTIVE_CS, YParameter: L.Y_PARAMETER}
_MODE = {"approx": L.MODE_APPROX, "exact": L.MODE_EXACT, "approx_good": L.MODE_APPROX_GOOD}


_LOSS_KINDS = {"mse": L.LOSS_MSE, "mse+esr": L.LOSS_MSE_ESR, "mse+esr_as_called": L.LOSS_MSE_ESR_AS_CALLED}


def _loss_kind(loss: str) -> int:
    """'mse' (clipper_pot.py:176); 'mse+esr': MSE + the error-to-signal ratio as esr_loss's signature reads, normalised by the
    TARGET's energy (clipper_pot.py:148-156,177); 'mse+esr_as_called': the same loss as the reference's training loop calls it,
    ``loss_func(outs, train_Y)`` (clipper_pot.py:248) — the arguments are swapped there, so the energy is the PREDICTION's and
    has a gradient of its own. The last one is what reproduces the reference's optimisation trajectory."""
    try:
        return _LOSS_KINDS[loss]
    except KeyError:
        raise ValueError(f"loss must be one of {sorted(_LOSS_KINDS)}, got {loss!r}") from None


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _on_device(method):
    """Runs a CompiledCircuit / optimizer method with the circuit's device current: the library launches on the
    caller's stream, allocates stream-ordered scratch and builds TMA descriptors for the CURRENT device."""

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        dev = self.device if hasattr(self, "device") else self.c.device
        if torch.cuda.current_device() == dev.index:
            return method(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return method(self, *args, **kwargs)

    return wrapper


class CompiledCircuit:
    """A circuit lowered to the flat program of include/dwdf.h, bound to one CUDA device.

    ``params`` (float32, on the device) is the parameter vector the kernels read: one slot per leaf
    in post-order, then ``Is`` and ``nabla`` for a DiodePair root. ``slots`` maps (element, attribute)
    to the slot; ``trainable`` lists the slots whose elements were built with ``trainable=True``.
    """

    def __init__(self, root, tree=None, probe=None, ordering="python", r_element=None, device=None, fs=None, probe_kind="voltage"):
        if probe_kind not in ("voltage", "current"):
            raise ValueError("probe_kind is 'voltage' ((a + b) / 2, tf_wdf.py:8-10) or 'current' ((a - b) / (2 R), wdf_t.h:1119-1123)")
        if not torch.cuda.is_available():
            raise RuntimeError("differentiable-wdfs_b200: the compiled path is CUDA-only and no CUDA device is visible (there is no CPU fallback)")
        self.lib = L.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("the compiled path is CUDA-only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.root = root
        self.tree = tree if tree is not None else getattr(root, "next", None)
        if self.tree is None:
            raise ValueError("compile_circuit needs the tree the root closes (tree=...)")
        if ordering not in ("python", "plugin"):
            raise ValueError("ordering is 'python' (probe after tree.incident) or 'plugin' (probe between the sweeps)")
        self.ordering = ordering
        nodes, self.elements, self.slots, values = [], [], {}, []
        fs_seen = []

        def visit(e):
            kind = _KIND.get(type(e))
            if kind is None:
                raise TypeError(f"{type(e).__name__} is not a WDF tree element")
            c1 = c2 = -1
            if kind in (L.SERIES, L.PARALLEL):
                c1, c2 = visit(e.P1), visit(e.P2)
            elif kind in (L.INVERTER, L.Y_PARAMETER):
                c1 = visit(e.P1)
            slot = -1
            if kind in (L.INDUCTOR, L.INDUCTOR_ALPHA):
                slot = len(values)
                values.append(float(e.L.detach()))
                self.slots[(id(e), "L")] = slot
                fs_seen.append(float(e.FS))
            if kind in (L.CAPACITOR_ALPHA, L.INDUCTOR_ALPHA):
                if kind == L.CAPACITOR_ALPHA:
                    slot = len(values)
                    values.append(float(e.C.detach()))
                    self.slots[(id(e), "C")] = slot
                    fs_seen.append(float(e.FS))
                values.append(float(e.alpha))  # the constants of an element follow its value slot
            elif kind == L.Y_PARAMETER:
                slot = len(values)
                values.extend(e.y)
            elif kind == L.RESISTIVE_CS:
                slot = len(values)
                values.append(float(e.R.detach()) if torch.is_tensor(e.R) else float(e.R))
                self.slots[(id(e), "R")] = slot
            if kind in (L.RESISTOR, L.RESISTIVE_VS):
                slot = len(values)
                values.append(float(e.R.detach()) if torch.is_tensor(e.R) else float(e.R))
                self.slots[(id(e), "R")] = slot
            elif kind == L.CAPACITOR:
                slot = len(values)
                values.append(float(e.C.detach()) if torch.is_tensor(e.C) else float(e.C))
                self.slots[(id(e), "C")] = slot
                fs_seen.append(float(e.FS))
            nodes.append(L.Node(kind, c1, c2, slot))
            self.elements.append(e)
            return len(nodes) - 1

        visit(self.tree)
        if len(nodes) > L.MAX_NODES:
            raise ValueError(f"circuit has {len(nodes)} nodes; the engine handles up to {L.MAX_NODES}")
        index = {id(e): i for i, e in enumerate(self.elements)}
        if probe is None or id(probe) not in index:
            raise ValueError("probe must be an element of the tree (the node whose voltage is the output)")
        d = L.CircuitDesc()
        d.probe = index[id(probe)]
        d.probe_current = 1 if probe_kind == "current" else 0
        d.ordering = L.ORDER_PYTHON if ordering == "python" else L.ORDER_PLUGIN
        d.source = -1
        d.r_node = index[id(r_element)] if r_element is not None else -1
        d.param_Is = d.param_nabla = -1
        d.Vt, d.n_up, d.n_down = 25.85e-3, 1.0, 1.0
        if isinstance(root, DiodePair):
            d.root_kind = L.ROOT_DIODE_PAIR
            d.root_mode = _MODE[root.mode]
            sources = [i for i, e in enumerate(self.elements) if isinstance(e, (ResistiveVoltageSource, ResistiveCurrentSource))]
            if len(sources) != 1:
                raise ValueError("a DiodePair circuit is driven through exactly one ResistiveVoltageSource (or ResistiveCurrentSource)")
            d.source = sources[0]
            d.param_Is = len(values)
            values.append(float(root.Is.detach()))
            self.slots[(id(root), "Is")] = d.param_Is
            d.param_nabla = len(values)
            values.append(float(root.nabla.detach()))
            self.slots[(id(root), "nabla")] = d.param_nabla
            d.Vt, d.n_up, d.n_down = root.Vt, root.N_up, root.N_down
            d.newton_max_iter, d.newton_tol = root.newton_max_iter, root.newton_tol
        elif isinstance(root, IdealVoltageSource):
            d.root_kind = L.ROOT_IDEAL_VS
        elif isinstance(root, IdealCurrentSource):
            d.root_kind = L.ROOT_IDEAL_CS
        elif isinstance(root, (Diode, Switch)):
            sources = [i for i, e in enumerate(self.elements) if isinstance(e, (ResistiveVoltageSource, ResistiveCurrentSource))]
            if len(sources) != 1:
                raise ValueError("a Diode / Switch circuit is driven through exactly one ResistiveVoltageSource or ResistiveCurrentSource")
            d.source = sources[0]
            if isinstance(root, Switch):
                d.root_kind, d.root_mode = L.ROOT_SWITCH, 1 if root.closed else 0
            else:
                d.root_kind = L.ROOT_DIODE
                d.param_Is = len(values)
                values.append(float(root.Is.detach()))
                self.slots[(id(root), "Is")] = d.param_Is
                d.param_nabla = len(values)
                values.append(float(root.nabla.detach()))
                self.slots[(id(root), "nabla")] = d.param_nabla
                d.Vt = root.Vt
        elif isinstance(root, DenseRootModel):
            d.root_kind = L.ROOT_NEURAL
            sources = [i for i, e in enumerate(self.elements) if isinstance(e, ResistiveVoltageSource)]
            if len(sources) != 1:
                raise ValueError("a neural-root circuit is driven through exactly one ResistiveVoltageSource")
            d.source = sources[0]
            hidden = root.sizes[1:-1]
            if root.sizes[0] != 2 or root.sizes[-1] != 1 or len(set(hidden)) != 1 or any(a != "tanh" for a in root.activations[:-1]) or root.activations[-1] not in ("", None):
                raise ValueError(f"network {root.sizes} / {root.activations}: the fused kernel runs the reference's shapes, 2 -> H (tanh) x (n+1) -> 1")
            self.mlp = L.MlpDesc(len(hidden) - 1, hidden[0])
        else:
            raise TypeError("root must be an IdealVoltageSource, an IdealCurrentSource, a DiodePair, a Diode, a Switch or a DenseRootModel")
        if fs is None:
            if not fs_seen:
                fs = 48000.0
            elif len(set(fs_seen)) != 1:
                raise ValueError("capacitors disagree on the sample rate")
            else:
                fs = fs_seen[0]
        d.fs = float(fs)
        d.n_params = len(values)
        self.desc = d
        self.n_params = len(values)
        arr = (L.Node * len(nodes))(*nodes)
        handle = C.c_void_p()
        self.is_neural = isinstance(root, DenseRootModel)
        with torch.cuda.device(self.device):  # the library prepares its per-device state for the device current at creation
            if self.is_neural:
                L.check(self.lib.dwdf_program_create_neural(arr, len(nodes), C.byref(d), C.byref(self.mlp), C.byref(handle)))
                w = root.weight_vector()
                assert w.size == self.lib.dwdf_mlp_weight_count(C.byref(self.mlp))
                self.weights = torch.from_numpy(w).to(self.device)
            else:
                L.check(self.lib.dwdf_program_create(arr, len(nodes), C.byref(d), C.byref(handle)))
        self.handle = handle
        self.is_clipper = bool(self.lib.dwdf_program_is_clipper(handle))
        self.is_specialized = False  # tree programs: run-time specialised kernels instead of the interpreter (specialize())
        self._specialize_error = None
        self.n_states = int(self.lib.dwdf_program_n_states(handle))
        self.params = torch.tensor(values, dtype=torch.float32, device=self.device)
        self.trainable = sorted(s for (eid, _), s in self.slots.items() if getattr(self._owner(eid), "trainable", False))
        lo = [-float("inf")] * self.n_params
        hi = [float("inf")] * self.n_params
        for (eid, attr), s in self.slots.items():
            e = self._owner(eid)
            # Keras applies a variable's constraint inside apply_gradients, i.e. to the TRAINABLE variables only
            # (tf_wdf.py:74,104): a frozen element outside the range keeps its value
            if attr in ("R", "C") and hasattr(e, "clip") and getattr(e, "trainable", False):
                lo[s], hi[s] = e.clip
        self.clip_lo = torch.tensor(lo, dtype=torch.float32, device=self.device)
        self.clip_hi = torch.tensor(hi, dtype=torch.float32, device=self.device)
        self.out = torch.zeros(L.OUT_LEN, dtype=torch.float64, device=self.device)
        self._ckpt = None
        self._work = None
        self._last = None

    def _owner(self, eid):
        if id(self.root) == eid:
            return self.root
        for e in self.elements:
            if id(e) == eid:
                return e
        raise KeyError(eid)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dwdf_program_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def slot(self, element, attr) -> int:
        return self.slots[(id(element), attr)]

    # ---- run-time specialisation of tree programs ----------------------------------------------------
    AUTO_SPECIALIZE_SAMPLES = 1 << 20  # forward() / train_step() specialise a tree program on their own from this batch size on

    @_on_device
    def specialize(self, quiet=False) -> bool:
        """Generates, compiles (NVRTC, about a second) and loads kernels written for THIS circuit (dwdf_program_specialize):
        straight-line wave arithmetic with every state in registers, TMA-tiled data movement, reverse mode without a tape.
        Trees the specialiser does not cover (alpha-transform / Y-parameter / current-source elements, diode / switch roots,
        current probe, resistance channel) keep running on the interpreter: returns False (or raises unless ``quiet``)."""
        if self.is_specialized:
            return True
        if self.is_clipper or self.is_neural:
            return False
        if self._specialize_error is None:
            try:
                L.check(self.lib.dwdf_program_specialize(self.handle))
                self.is_specialized = True
                self._last = None  # a forward pass of the interpreter left no checkpoints for the specialised reverse sweep
                self._ckpt = self._work = None
                return True
            except L.DwdfError as e:
                self._specialize_error = e
        if not quiet:
            raise self._specialize_error
        return False

    def _maybe_specialize(self, B, T):
        if not (self.is_clipper or self.is_neural or self.is_specialized) and self._specialize_error is None and B * T >= self.AUTO_SPECIALIZE_SAMPLES:
            self.specialize(quiet=True)

    # ---- buffers -------------------------------------------------------------------------------
    def _check_xy(self, x, name="x", shape=None):
        """Every companion buffer of a launch is checked against x's (B, T): the kernels index all of them with it."""
        if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous float32 CUDA tensor of shape (B, T)")
        if x.device != self.device:
            raise ValueError(f"{name} lives on {x.device}, the circuit on {self.device}")
        if shape is not None and tuple(x.shape) != tuple(shape):
            raise ValueError(f"{name} has shape {tuple(x.shape)}, expected {tuple(shape)} (the shape of x)")
        return x.shape

    @staticmethod
    def _check_host(t, name, shape=None):
        if not (torch.is_tensor(t) and not t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous float32 CPU tensor of shape (B, T)")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)} (the shape of x)")
        return t.shape

    def _scratch(self, attr, nbytes):
        buf = getattr(self, attr)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
            setattr(self, attr, buf)
        return buf

    # ---- forward / backward ----------------------------------------------------------------------
    @_on_device
    def forward(self, x, r=None, out=None, keep_for_backward=True):
        """y = circuit(x): (B, T) float32 CUDA tensors, every sequence from reset state."""
        B, T = self._check_xy(x)
        if r is not None:
            self._check_xy(r, "r", (B, T))
        y = torch.empty_like(x) if out is None else out
        self._check_xy(y, "out", (B, T))
        if B * T == 0:
            self._last = None
            return y
        if self.is_neural:
            ck = self._scratch("_ckpt", self.lib.dwdf_neural_ckpt_bytes(self.handle, B, T)) if keep_for_backward else None
            L.check(self.lib.dwdf_forward_neural(self.handle, _ptr(self.params), _ptr(self.weights), _ptr(x), _ptr(r), _ptr(y), None, _ptr(ck), B, T, _stream_ptr(self.device)))
            self._last = (x, r, y, B, T, y._version) if keep_for_backward else None
            return y
        self._maybe_specialize(B, T)
        ck = None
        if keep_for_backward and (self.is_clipper or self.is_specialized) and B * T > 0:
            ck = self._scratch("_ckpt", self.lib.dwdf_ckpt_bytes(self.handle, B, T))
        L.check(self.lib.dwdf_forward(self.handle, _ptr(self.params), _ptr(x), _ptr(r), _ptr(y), _ptr(ck), B, T, _stream_ptr(self.device)))
        self._last = (x, r, y, B, T, y._version) if keep_for_backward else None
        return y

    def forward_time_major(self, x, r=None):
        """The reference's output convention: (T, B, 1) as stacked by TensorArray (lpf.py:48, clipper_pot.py:126)."""
        return self.forward(x, r).t().unsqueeze(-1)

    @_on_device
    def backward(self, gy=None, target=None, loss="mse", skip=0, want_gx=False, raw=False):
        """Gradients of the last ``forward``: upstream ``gy = dL/dy`` or a fused loss against ``target``.
        The adjoint kernel recovers the states from the output tensor that ``forward`` returned (no tape, no replay):
        an in-place torch operation on that tensor between the two calls is detected (version counter) and raises.

        Returns a dict with ``grads`` (float64 tensor, one per parameter slot, on the device), ``loss``,
        ``mse``, ``esr`` (0-d device tensors; target mode) and ``gx`` if requested.
        """
        if self._last is None:
            raise RuntimeError("backward() needs a preceding forward(keep_for_backward=True)")
        if (gy is None) == (target is None):
            raise ValueError("give exactly one of gy (upstream gradient) or target (fused loss)")
        x, r, y, B, T, y_version = self._last
        if y._version != y_version:
            raise RuntimeError("the tensor forward() returned was modified in place before backward(): the adjoint recovers the circuit's states from it (clone it before modifying, or call forward() again)")
        g = gy if gy is not None else target
        self._check_xy(g, "gy/target", (B, T))
        if self.is_neural:
            work = self._scratch("_work", self.lib.dwdf_neural_workspace_bytes(self.handle, B, T))
            if getattr(self, "grad_w", None) is None:
                self.grad_w = torch.zeros(self.weights.numel(), dtype=torch.float64, device=self.device)
            gx = torch.empty_like(x) if want_gx else None
            mode = L.GRAD_UPSTREAM if gy is not None else L.GRAD_TARGET
            if raw:  # this rank's sums, before the loss's scale: all-reduce grad_w and out, then finalize()
                L.check(self.lib.dwdf_backward_neural_raw(self.handle, _ptr(self.params), _ptr(self.weights), _ptr(x), _ptr(r), _ptr(y), _ptr(self._ckpt), _ptr(g), mode, int(skip), _ptr(gx), _ptr(self.grad_w),
                                                          _ptr(self.out), _ptr(work), work.numel(), B, T, _stream_ptr(self.device)))
            else:
                L.check(self.lib.dwdf_backward_neural(self.handle, _ptr(self.params), _ptr(self.weights), _ptr(x), _ptr(r), _ptr(y), _ptr(self._ckpt), _ptr(g), mode, _loss_kind(loss), int(skip), _ptr(gx),
                                                      _ptr(self.grad_w), _ptr(self.out), _ptr(work), work.numel(), B, T, _stream_ptr(self.device)))
            res = self._result(gx)
            res["grads"] = self.grad_w
            return res
        gx = torch.empty_like(x) if want_gx else None
        nbytes = self.lib.dwdf_workspace_bytes(self.handle, B, T)
        work = self._scratch("_work", nbytes)
        ck = self._ckpt if (self.is_clipper or self.is_specialized) else None
        mode = L.GRAD_UPSTREAM if gy is not None else L.GRAD_TARGET
        if raw:
            L.check(self.lib.dwdf_backward_raw(self.handle, _ptr(self.params), _ptr(x), _ptr(r), _ptr(y), _ptr(ck), _ptr(g), mode, int(skip), _ptr(gx), _ptr(self.out), _ptr(work), work.numel(), B, T,
                                               _stream_ptr(self.device)))
        else:
            L.check(self.lib.dwdf_backward(self.handle, _ptr(self.params), _ptr(x), _ptr(r), _ptr(y), _ptr(ck), _ptr(g), mode, _loss_kind(loss), int(skip), _ptr(gx),
                                           _ptr(self.out), _ptr(work), work.numel(), B, T, _stream_ptr(self.device)))
        return self._result(gx)

    @_on_device
    def finalize(self, target=True, loss="mse"):
        """Turns the (all-reduced) raw sums in ``self.out`` (neural root: and ``self.grad_w``) into gradients and loss, in place."""
        if self.is_neural:
            L.check(self.lib.dwdf_finalize_neural(self.handle, L.GRAD_TARGET if target else L.GRAD_UPSTREAM, _loss_kind(loss), _ptr(self.grad_w), _ptr(self.out), _stream_ptr(self.device)))
            res = self._result(None)
            res["grads"] = self.grad_w
            return res
        L.check(self.lib.dwdf_finalize(self.handle, _ptr(self.params), L.GRAD_TARGET if target else L.GRAD_UPSTREAM, _loss_kind(loss), _ptr(self.out),
                                       _stream_ptr(self.device)))
        return self._result(None)

    @_on_device
    def train_pass(self, x, target, loss="mse", skip=0, y=None, raw=False):
        """Fused forward + loss + parameter gradients in one sweep (diode clipper only)."""
        B, T = self._check_xy(x)
        self._check_xy(target, "target", (B, T))
        if y is not None:
            self._check_xy(y, "y", (B, T))
        work = self._scratch("_work", self.lib.dwdf_workspace_bytes(self.handle, B, T))
        if raw:
            L.check(self.lib.dwdf_train_pass_raw(self.handle, _ptr(self.params), _ptr(x), None, _ptr(target), int(skip), _ptr(y), _ptr(self.out), _ptr(work), work.numel(), B, T,
                                                 _stream_ptr(self.device)))
        else:
            L.check(self.lib.dwdf_train_pass(self.handle, _ptr(self.params), _ptr(x), None, _ptr(target), _loss_kind(loss), int(skip), _ptr(y), _ptr(self.out),
                                             _ptr(work), work.numel(), B, T, _stream_ptr(self.device)))
        return self._result(None)

    @_on_device
    def train_step(self, x, target, optimizer: "Adam", loss="mse", skip=0, out=None, comm=None, r=None, engine="adjoint"):
        """forward + adjoint (fused loss) + Adam in one library call, nothing synchronises: capturable in a
        ``torch.cuda.CUDAGraph`` and replayable (small batches are launch-bound). Returns the result dict of
        ``backward`` (device tensors, overwritten by the next step). ``comm`` (a ``data_parallel.PeerComm``): x and
        target are THIS RANK'S shard; the step's reduction kernel exchanges the raw sums with the other ranks over peer
        memory, loss and gradients are those of the global batch and every rank applies the same update.
        ``engine="tangent"`` (diode clipper, no resistance channel): the same step on the one-sweep kernel that carries the
        parameter sensitivities forward with the recurrence (``train_pass``: 12 instead of 20 bytes of traffic per sample, same
        gradients to fp32 summation order) followed by the exchange, the chain rule and Adam."""
        if engine not in ("adjoint", "tangent"):
            raise ValueError("engine is 'adjoint' (forward + reverse sweep) or 'tangent' (one sweep with forward-mode sensitivities)")
        if engine == "tangent":
            if not self.is_clipper or self.is_neural or r is not None:
                raise ValueError("engine='tangent' exists for the diode-clipper program without a resistance channel")
            many = comm is not None and comm.world_size > 1
            self.train_pass(x, target, loss=loss, skip=skip, y=out, raw=many)
            if many:
                comm.all_reduce_(self.out)
                self.finalize(target=True, loss=loss)
            if optimizer is not None:
                optimizer.apply()
            return self._result(None)
        B, T = self._check_xy(x)
        self._check_xy(target, "target", (B, T))
        y = torch.empty_like(x) if out is None else out
        self._check_xy(y, "out", (B, T))
        if comm is not None and comm.device != self.device:
            raise ValueError(f"the communicator lives on {comm.device}, the circuit on {self.device}")
        if self.is_neural:
            # the network's kernels and biases are the trainable variables (clipper_pot.py:246-269); optimizer: AdamWeights or None
            if r is not None:
                self._check_xy(r, "r", (B, T))
            ck = self._scratch("_ckpt", self.lib.dwdf_neural_ckpt_bytes(self.handle, B, T))
            work = self._scratch("_work", self.lib.dwdf_neural_workspace_bytes(self.handle, B, T))
            if getattr(self, "grad_w", None) is None:
                self.grad_w = torch.zeros(self.weights.numel(), dtype=torch.float64, device=self.device)
            o = optimizer
            L.check(self.lib.dwdf_train_step_neural(self.handle, comm.handle if comm is not None else None, _ptr(self.params), _ptr(self.weights), _ptr(x), _ptr(r), _ptr(target), _loss_kind(loss), int(skip), _ptr(y),
                                                    _ptr(ck), _ptr(self.grad_w), _ptr(self.out), _ptr(work), work.numel(), _ptr(o.m) if o is not None else None, _ptr(o.v) if o is not None else None,
                                                    _ptr(o.step_count) if o is not None else None, float(o.lr) if o is not None else 0.0, float(o.beta_1) if o is not None else 0.0,
                                                    float(o.beta_2) if o is not None else 0.0, float(o.epsilon) if o is not None else 0.0, B, T, _stream_ptr(self.device)))
            self._last = (x, r, y, B, T, y._version)
            res = self._result(None)
            res["grads"] = self.grad_w
            return res
        if r is not None:
            raise ValueError("train_step takes the resistance channel for the neural root only; use forward(x, r) + backward for the analytic root")
        self._maybe_specialize(B, T)
        ck = self._scratch("_ckpt", self.lib.dwdf_ckpt_bytes(self.handle, B, T))
        work = self._scratch("_work", self.lib.dwdf_workspace_bytes(self.handle, B, T))
        o = optimizer
        if comm is not None:
            L.check(self.lib.dwdf_train_step_dp(self.handle, comm.handle, _ptr(self.params), _ptr(x), None, _ptr(target), _loss_kind(loss), int(skip), _ptr(y), _ptr(ck), _ptr(self.out),
                                                _ptr(work), work.numel(), _ptr(o.m), _ptr(o.v), _ptr(o.step_count), 0.0, _ptr(o.lr), float(o.beta_1), float(o.beta_2), float(o.epsilon), _ptr(self.clip_lo),
                                                _ptr(self.clip_hi), B, T, _stream_ptr(self.device)))
        else:
            L.check(self.lib.dwdf_train_step(self.handle, _ptr(self.params), _ptr(x), None, _ptr(target), _loss_kind(loss), int(skip), _ptr(y), _ptr(ck), _ptr(self.out),
                                             _ptr(work), work.numel(), _ptr(o.m), _ptr(o.v), _ptr(o.step_count), 0.0, _ptr(o.lr), float(o.beta_1), float(o.beta_2), float(o.epsilon), _ptr(self.clip_lo),
                                             _ptr(self.clip_hi), B, T, _stream_ptr(self.device)))
        self._last = (x, None, y, B, T, y._version)
        return self._result(None)

    def _result(self, gx):
        o = self.out
        return {"grads": o[: self.n_params], "loss": o[L.OUT_LOSS], "mse": o[L.OUT_MSE], "esr": o[L.OUT_ESR], "gx": gx, "out": o}

    @_on_device
    def process_block(self, x, state, r=None, out=None):
        """Streaming twin of DiodeClipperWDF::process: continues from ``state`` ((n_states, B) float32) and updates it."""
        B, T = self._check_xy(x)
        if not (torch.is_tensor(state) and state.is_cuda and state.dtype == torch.float32 and state.numel() == self.n_states * B and state.is_contiguous()):
            raise ValueError(f"state must be a contiguous float32 CUDA tensor with {self.n_states} x B elements")
        if r is not None:
            self._check_xy(r, "r", (B, T))
        y = torch.empty_like(x) if out is None else out
        self._check_xy(y, "out", (B, T))
        if self.is_neural:
            L.check(self.lib.dwdf_forward_neural(self.handle, _ptr(self.params), _ptr(self.weights), _ptr(x), _ptr(r), _ptr(y), _ptr(state), None, B, T, _stream_ptr(self.device)))
            return y
        L.check(self.lib.dwdf_process_block(self.handle, _ptr(self.params), _ptr(x), _ptr(r), _ptr(y), _ptr(state), B, T, _stream_ptr(self.device)))
        return y

    def new_state(self, B):
        return torch.zeros(self.n_states, B, dtype=torch.float32, device=self.device)

    # ---- end-to-end with host (numpy / pinned torch) buffers ------------------------------------------
    @_on_device
    def forward_host(self, x_host: torch.Tensor, y_host: torch.Tensor, params_host: Optional[torch.Tensor] = None):
        B, T = self._check_host(x_host, "x_host")
        self._check_host(y_host, "y_host", (B, T))
        p = self._host_params(params_host)
        L.check(self.lib.dwdf_forward_host(self.handle, _ptr(p), _ptr(x_host), None, _ptr(y_host), B, T))
        return y_host

    @_on_device
    def grad_host(self, x_host, g_host, out_host, y_host=None, params_host=None, target=True, loss="mse", skip=0):
        B, T = self._check_host(x_host, "x_host")
        self._check_host(g_host, "g_host", (B, T))
        if y_host is not None:
            self._check_host(y_host, "y_host", (B, T))
        if not (torch.is_tensor(out_host) and not out_host.is_cuda and out_host.dtype == torch.float64 and out_host.is_contiguous() and out_host.numel() >= L.OUT_LEN):
            raise ValueError(f"out_host must be a contiguous float64 CPU tensor of at least {L.OUT_LEN} elements")
        p = self._host_params(params_host)
        L.check(self.lib.dwdf_grad_host(self.handle, _ptr(p), _ptr(x_host), None, _ptr(g_host), L.GRAD_TARGET if target else L.GRAD_UPSTREAM, _loss_kind(loss),
                                        int(skip), _ptr(y_host), _ptr(out_host), B, T))
        return out_host

    def _host_params(self, params_host):
        if params_host is None:
            return self.params.cpu()
        if not (torch.is_tensor(params_host) and not params_host.is_cuda and params_host.dtype == torch.float32 and params_host.is_contiguous() and params_host.numel() == self.n_params):
            raise ValueError(f"params_host must be a contiguous float32 CPU tensor of {self.n_params} elements")
        return params_host

    # ---- parameters ------------------------------------------------------------------------------------
    def sync_from_elements(self):
        """Re-reads every element's current value into the device parameter vector."""
        vals = [0.0] * self.n_params
        for (eid, attr), s in self.slots.items():
            vals[s] = float(torch.as_tensor(getattr(self._owner(eid), attr)).detach())
        self.params.copy_(torch.tensor(vals, dtype=torch.float32))

    def sync_to_elements(self):
        """Writes the device parameter vector back into the element objects (R, C, Is, nabla)."""
        vals = self.params.detach().cpu()
        for (eid, attr), s in self.slots.items():
            e = self._owner(eid)
            setattr(e, attr, _scalar(float(vals[s]), getattr(e, "trainable", False)))


def compile_circuit(root, tree=None, probe=None, ordering="python", r_element=None, device=None, fs=None, probe_kind="voltage") -> CompiledCircuit:
    return CompiledCircuit(root, tree, probe, ordering, r_element, device, fs, probe_kind)


class AdamWeights:
    """tf.keras.optimizers.Adam on the weight vector of a neural-root circuit (clipper_pot.py:180,269)."""

    def __init__(self, circuit: "CompiledCircuit", lr=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.c = circuit
        n = circuit.weights.numel()
        self.m = torch.zeros(n, dtype=torch.float32, device=circuit.device)
        self.v = torch.zeros(n, dtype=torch.float32, device=circuit.device)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=circuit.device)
        self.lr, self.beta_1, self.beta_2, self.epsilon = float(lr), float(beta_1), float(beta_2), float(epsilon)

    @_on_device
    def apply(self, grad_scale=1.0):
        c = self.c
        L.check(c.lib.dwdf_adam_step_vec(_ptr(c.weights), _ptr(c.grad_w), _ptr(self.m), _ptr(self.v), _ptr(self.step_count), c.weights.numel(), self.lr, self.beta_1, self.beta_2, self.epsilon,
                                         float(grad_scale), _stream_ptr(c.device)))


class Adam:
    """tf.keras.optimizers.Adam for a CompiledCircuit (clipper_pot.py:180,269). ``lr`` is one rate for
    every trainable slot or a dict slot -> rate (lpf.py:79-80 trains R and C with different rates);
    slots that are not trainable get rate 0. One device kernel per step, moments and step counter on
    the device, then the reference's clip constraints (tf_wdf.py:74,104)."""

    def __init__(self, circuit: CompiledCircuit, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, slots=None):
        self.c = circuit
        dev = circuit.device
        self.m = torch.zeros(circuit.n_params, dtype=torch.float32, device=dev)
        self.v = torch.zeros(circuit.n_params, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        rates = [0.0] * circuit.n_params
        if isinstance(lr, dict):
            for s, v in lr.items():
                rates[int(s)] = float(v)
        else:
            for s in (circuit.trainable if slots is None else slots):
                rates[int(s)] = float(lr)
        self.lr = torch.tensor(rates, dtype=torch.float32, device=dev)

    @_on_device
    def apply(self, grad_scale=1.0):
        c = self.c
        L.check(c.lib.dwdf_adam_step(_ptr(c.params), _ptr(c.out), _ptr(self.m), _ptr(self.v), _ptr(self.step_count), c.n_params, 0.0, _ptr(self.lr), float(self.beta_1), float(self.beta_2),
                                     float(self.epsilon), float(grad_scale), _ptr(c.clip_lo), _ptr(c.clip_hi), _stream_ptr(c.device)))
