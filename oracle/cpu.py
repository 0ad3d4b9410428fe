"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU oracle and the compiled reference.

``Oracle``  -> oracle/_ref/libdwdf_oracle.so  (plain-C restatement, oracle/wdf_oracle.c)
``Ref``     -> oracle/_ref/libdwdf_ref.so     (the reference's own C++ compiled in place,
               oracle/ref_harness.cpp; ``fast=True`` picks the -O3/AVX2 build used for timing)

Both are built by ``oracle/Makefile`` (``__graft_entry__.build()`` runs it). Nothing in the
product path imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, "_ref")

# enums shared with wdf_oracle.c / ref_harness.cpp
RESISTOR, CAPACITOR, RESVS, SERIES, PARALLEL, INVERTER, INDUCTOR, CAPACITOR_ALPHA, INDUCTOR_ALPHA, RESCS, YPARAM = range(11)
ROOT_IDEAL_VS, ROOT_DIODE_PAIR, ROOT_IDEAL_CS, ROOT_DIODE, ROOT_SWITCH = 0, 1, 3, 4, 5
ORDER_PLUGIN, ORDER_PYTHON = 0, 1
ROOT_APPROX, ROOT_EXACT, ROOT_APPROX_GOOD = 0, 1, 2

OMEGA_KINDS = {"omega1": 0, "omega2": 1, "omega3": 2, "omega4": 3, "log_approx": 4, "exp_approx": 5, "log2_approx": 6, "pow2_approx": 7}


def build(ref: bool = True) -> None:
    """Compile the oracle (and, when /root/reference exists, the reference harness)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", _HERE, *targets], check=True, stdout=subprocess.DEVNULL)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _np(a, dtype):
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


@dataclass
class ClipperParams:
    """Diode clipper constants. Defaults = plugin values (DiodeClipperWDF.h:16-25), 1N4148 1U-1D."""

    fs: float = 48000.0
    R: float = 47000.0
    C: float = 2.2e-9
    Is: float = 4.352e-9
    Vt: float = 25.85e-3
    nabla: float = 1.906
    n_up: float = 1.0
    n_down: float = 1.0


class Oracle:
    def __init__(self, path: str | None = None):
        path = path or os.path.join(_OUT, "libdwdf_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        self.lib.ow_toms917_real.restype = C.c_double
        self.lib.ow_toms917_real.argtypes = [C.c_double]

    # ---- scalar functions -------------------------------------------------------------------
    def omega(self, kind: str, x, dtype=np.float32):
        x = _np(x, dtype)
        out = np.empty_like(x)
        fn = self.lib.ow_omega_f32 if dtype == np.float32 else self.lib.ow_omega_f64
        fn(C.c_int(OMEGA_KINDS[kind]), _ptr(x), _ptr(out), C.c_int64(x.size))
        return out

    def toms917(self, x):
        x = _np(x, np.float64)
        out = np.empty_like(x)
        self.lib.ow_toms917_real_n(_ptr(x), _ptr(out), C.c_int64(x.size))
        return out

    def diode_pair(self, a, Rp, p: ClipperParams = ClipperParams(), exact=False, good=False, dtype=np.float32):
        a = _np(a, dtype)
        b = np.empty_like(a)
        par = _np([float(exact), float(good), p.Is, p.Vt, p.nabla, p.n_up, p.n_down], dtype)
        if dtype == np.float32:
            self.lib.ow_diode_pair_f32(_ptr(par), C.c_float(Rp), _ptr(a), _ptr(b), C.c_int64(a.size))
        else:
            self.lib.ow_diode_pair_f64(_ptr(par), C.c_double(Rp), _ptr(a), _ptr(b), C.c_int64(a.size))
        return b

    # ---- generic tree interpreter --------------------------------------------------------------
    def tree_run(self, nodes, fs, root_kind, x, probe, source=-1, root_par=None, ordering=ORDER_PYTHON, r_in=None, r_node=-1, dtype=np.float32):
        """nodes: list of (kind, c1, c2, value) in post-order; x: (B, T)."""
        x = _np(x, dtype)
        assert x.ndim == 2
        B, T = x.shape
        kind = _np([n[0] for n in nodes], np.int32)
        c1 = _np([n[1] for n in nodes], np.int32)
        c2 = _np([n[2] for n in nodes], np.int32)
        val = _np([n[3] for n in nodes], dtype)
        rp = _np(root_par if root_par is not None else [0] * 7, dtype)
        y = np.empty_like(x)
        r = None if r_in is None else _np(r_in, dtype)
        fn = self.lib.ow_tree_run_f32 if dtype == np.float32 else self.lib.ow_tree_run_f64
        ct = C.c_float if dtype == np.float32 else C.c_double
        rc = fn(C.c_int(len(nodes)), _ptr(kind), _ptr(c1), _ptr(c2), _ptr(val), ct(fs), C.c_int(root_kind), _ptr(rp), C.c_int(source), C.c_int(probe), C.c_int(ordering), C.c_int(r_node), _ptr(x), _ptr(r) if r is not None else None, _ptr(y), C.c_int64(B), C.c_int64(T))
        assert rc == 0
        return y

    def tree_run_ext(self, nodes, fs, root_kind, x, probe, source=-1, root_par=None, ordering=ORDER_PYTHON, probe_current=False, dtype=np.float32):
        """The full chowdsp_wdf element set. nodes: (kind, c1, c2, value[, aux...]) in post-order, aux = alpha for the
        alpha-transform leaves, (y12, y21, y22) for the Y-parameter (value = y11); x: (B, T)."""
        x = _np(x, dtype)
        B, T = x.shape
        kind = _np([n[0] for n in nodes], np.int32)
        c1 = _np([n[1] for n in nodes], np.int32)
        c2 = _np([n[2] for n in nodes], np.int32)
        val = _np([n[3] for n in nodes], dtype)
        aux = np.zeros((len(nodes), 3), dtype)
        for i, n in enumerate(nodes):
            aux[i, :len(n) - 4] = n[4:]
        rp = _np(root_par if root_par is not None else [0] * 7, dtype)
        y = np.empty_like(x)
        fn = self.lib.ow_tree_run_ext_f32 if dtype == np.float32 else self.lib.ow_tree_run_ext_f64
        ct = C.c_float if dtype == np.float32 else C.c_double
        rc = fn(C.c_int(len(nodes)), _ptr(kind), _ptr(c1), _ptr(c2), _ptr(val), _ptr(aux), ct(fs), C.c_int(root_kind), _ptr(rp), C.c_int(source), C.c_int(probe), C.c_int(int(probe_current)), C.c_int(ordering), _ptr(x), _ptr(y),
                C.c_int64(B), C.c_int64(T))
        assert rc == 0
        return y

    # ---- diode clipper -------------------------------------------------------------------------
    def clipper_forward(self, x, p: ClipperParams = ClipperParams(), exact=False, good=False, ordering=ORDER_PYTHON, dtype=np.float32, threads=1):
        x = _np(x, dtype)
        B, T = x.shape
        y = np.empty_like(x)
        ct = C.c_float if dtype == np.float32 else C.c_double
        fn = self.lib.ow_clipper_forward_f32 if dtype == np.float32 else self.lib.ow_clipper_forward_f64
        fn(C.c_int(int(exact)), C.c_int(int(good)), C.c_int(ordering), ct(p.fs), ct(p.R), ct(p.C), ct(p.Is), ct(p.Vt), ct(p.nabla), ct(p.n_up), ct(p.n_down), _ptr(x), _ptr(y), C.c_int64(B), C.c_int64(T), C.c_int(threads))
        return y

    def clipper_grad(self, x, gy_or_target, p: ClipperParams = ClipperParams(), exact=False, ordering=ORDER_PYTHON, mode="target", loss="mse", skip=0, dtype=np.float64, threads=1, want_gx=False):
        """Returns dict(grads=[dIs,dnabla,dR,dC], loss, mse, esr, y, gx, raw)."""
        x = _np(x, dtype)
        g = _np(gy_or_target, dtype)
        B, T = x.shape
        y = np.empty_like(x)
        gx = np.empty_like(x) if want_gx else None
        out = np.zeros(8, np.float64)
        raw = np.zeros(8, np.float64)
        ct = C.c_float if dtype == np.float32 else C.c_double
        fn = self.lib.ow_clipper_grad_f32 if dtype == np.float32 else self.lib.ow_clipper_grad_f64
        fn(C.c_int(int(exact)), C.c_int(ordering), ct(p.fs), ct(p.R), ct(p.C), ct(p.Is), ct(p.Vt), ct(p.nabla), ct(p.n_up), ct(p.n_down), _ptr(x), _ptr(g), C.c_int(1 if mode == "target" else 0), C.c_int(1 if loss == "mse+esr" else 0), C.c_int64(skip), _ptr(y), _ptr(gx) if gx is not None else None, _ptr(out), _ptr(raw), C.c_int64(B), C.c_int64(T), C.c_int(threads))
        return dict(grads=out[:4].copy(), loss=out[4], mse=out[5], esr=out[6], y=y, gx=gx, raw=raw[:6].copy())


class Ref:
    """The reference's own C++ (chowdsp_wdf + toms917), compiled in place. Absent => FileNotFoundError."""

    def __init__(self, fast: bool = False):
        path = os.path.join(_OUT, "libdwdf_ref_fast.so" if fast else "libdwdf_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference"):
                build(ref=True)
            if not os.path.exists(path):
                raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ref_standalone_test.restype = C.c_float
        self.lib.ref_clipper_port_impedance_f32.restype = C.c_float
        self.lib.ref_hardware_threads.restype = C.c_int

    @staticmethod
    def available(fast: bool = False) -> bool:
        return os.path.exists(os.path.join(_OUT, "libdwdf_ref_fast.so" if fast else "libdwdf_ref.so")) or os.path.isdir("/root/reference")

    def omega(self, kind: str, x, dtype=np.float32):
        x = _np(x, dtype)
        out = np.empty_like(x)
        fn = self.lib.ref_omega_f32 if dtype == np.float32 else self.lib.ref_omega_f64
        fn(C.c_int(OMEGA_KINDS[kind]), _ptr(x), _ptr(out), C.c_int64(x.size))
        return out

    def toms917(self, x):
        x = _np(x, np.float64)
        out = np.empty_like(x)
        self.lib.ref_toms917_real(_ptr(x), _ptr(out), C.c_int64(x.size))
        return out

    def diode_pair(self, a, Rp, p: ClipperParams = ClipperParams(), root=ROOT_APPROX, dtype=np.float32):
        a = _np(a, dtype)
        b = np.empty_like(a)
        ct = C.c_float if dtype == np.float32 else C.c_double
        fn = self.lib.ref_diode_pair_f32 if dtype == np.float32 else self.lib.ref_diode_pair_f64
        fn(C.c_int(root), _ptr(a), _ptr(b), C.c_int64(a.size), ct(Rp), ct(p.Is), ct(p.Vt), ct(p.nabla))
        return b

    def clipper(self, x, p: ClipperParams = ClipperParams(), root=ROOT_APPROX, ordering=ORDER_PLUGIN, dtype=np.float32, threads=1, out=None):
        x = _np(x, dtype)
        B, T = x.shape
        y = np.empty_like(x) if out is None else out
        ct = C.c_float if dtype == np.float32 else C.c_double
        fn = self.lib.ref_clipper_f32 if dtype == np.float32 else self.lib.ref_clipper_f64
        fn(C.c_int(root), C.c_int(ordering), _ptr(x), _ptr(y), C.c_int64(B), C.c_int64(T), ct(p.fs), ct(p.R), ct(p.C), ct(p.Is), ct(p.Vt), ct(p.nabla), C.c_int(threads))
        return y

    def port_impedance(self, p: ClipperParams = ClipperParams()):
        return float(self.lib.ref_clipper_port_impedance_f32(C.c_float(p.fs), C.c_float(p.R), C.c_float(p.C)))

    def rc_lowpass(self, x, fs, R, Cval, probe=0, dtype=np.float64):
        x = _np(x, dtype).ravel()
        y = np.empty_like(x)
        ct = C.c_float if dtype == np.float32 else C.c_double
        fn = self.lib.ref_rc_lowpass_f32 if dtype == np.float32 else self.lib.ref_rc_lowpass_f64
        fn(_ptr(x), _ptr(y), C.c_int64(x.size), ct(fs), ct(R), ct(Cval), C.c_int(probe))
        return y

    def voltage_divider(self, x, Ra, Rb):
        x = _np(x, np.float64).ravel()
        y = np.empty_like(x)
        self.lib.ref_voltage_divider_f64(_ptr(x), _ptr(y), C.c_int64(x.size), C.c_double(Ra), C.c_double(Rb))
        return y

    def standalone_test(self) -> float:
        return float(self.lib.ref_standalone_test())

    def static_wdf_test(self, quality: int, fs: float = 44100.0):
        out = np.zeros(5, np.float64)
        self.lib.ref_static_wdf_test(C.c_int(quality), C.c_double(fs), _ptr(out))
        return out

    def hardware_threads(self) -> int:
        return int(self.lib.ref_hardware_threads())


class RefElements:
    """oracle/_ref/libdwdf_ref_elements.so: the unmodified chowdsp_wdf classes of the remaining elements in the circuits of
    the reference's own tests (oracle/ref_elements_harness.cpp). One mono signal in, one out."""

    def __init__(self):
        path = os.path.join(_OUT, "libdwdf_ref_elements.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    def _run(self, fn, x, *args):
        x = _np(x, np.float32).ravel()
        y = np.empty_like(x)
        fn(_ptr(x), _ptr(y), C.c_int64(x.size), *args)
        return y

    def current_divider(self, x, R1, R2):
        return self._run(self.lib.ref_el_current_divider, x, C.c_float(R1), C.c_float(R2))

    def current_switch(self, x, R1, Rs, closed):
        return self._run(self.lib.ref_el_current_switch, x, C.c_float(R1), C.c_float(Rs), C.c_int(int(closed)))

    def rlc_highpass(self, x, fs, R, Cv, L, alpha=-1.0):
        return self._run(self.lib.ref_el_rlc_highpass, x, C.c_float(fs), C.c_float(R), C.c_float(Cv), C.c_float(L), C.c_float(alpha))

    def y_parameter(self, x, R, y11, y12, y21, y22, probe):
        return self._run(self.lib.ref_el_y_parameter, x, C.c_float(R), C.c_float(y11), C.c_float(y12), C.c_float(y21), C.c_float(y22), C.c_int(probe))

    def diode(self, x, fs, Rs, Cv, Is, Vt, n_diodes, probe):
        return self._run(self.lib.ref_el_diode, x, C.c_float(fs), C.c_float(Rs), C.c_float(Cv), C.c_float(Is), C.c_float(Vt), C.c_float(n_diodes), C.c_int(probe))
