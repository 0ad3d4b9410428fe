"""ctypes binding of ``libdwdf.so`` (the C ABI declared in ``include/dwdf.h``).

The library is CUDA-only. Loading it needs no GPU (so the ABI can be inspected on a build box), but
every compute entry point fails with ``DWDF_ERR_NO_DEVICE`` / ``DWDF_ERR_CUDA`` when there is none:
there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DWDF_LIBRARY") or os.path.join(_HERE, "libdwdf.so")  # (DWDF_LIBRARY: a side-by-side build variant, see csrc/Makefile)

# enums of include/dwdf.h
RESISTOR, CAPACITOR, RESISTIVE_VS, SERIES, PARALLEL, INVERTER, INDUCTOR, CAPACITOR_ALPHA, INDUCTOR_ALPHA, RESISTIVE_CS, Y_PARAMETER = range(11)
ROOT_IDEAL_VS, ROOT_DIODE_PAIR, ROOT_NEURAL, ROOT_IDEAL_CS, ROOT_DIODE, ROOT_SWITCH = range(6)
MODE_APPROX, MODE_EXACT, MODE_APPROX_GOOD = 0, 1, 2
ORDER_PLUGIN, ORDER_PYTHON = 0, 1
GRAD_UPSTREAM, GRAD_TARGET = 0, 1
LOSS_MSE, LOSS_MSE_ESR, LOSS_MSE_ESR_AS_CALLED = 0, 1, 2
MAX_PARAMS, MAX_NODES = 16, 16
OUT_LOSS, OUT_MSE, OUT_ESR, OUT_LEN = 16, 17, 18, 24

STATUS = {0: "ok", 1: "invalid argument", 2: "unsupported", 3: "CUDA error", 4: "no CUDA device", 5: "workspace too small"}

# every symbol include/dwdf.h declares (tests check the export list against this and the header)
SYMBOLS = (
    "dwdf_program_create", "dwdf_program_destroy", "dwdf_program_is_clipper", "dwdf_program_n_states", "dwdf_ckpt_bytes", "dwdf_workspace_bytes",
    "dwdf_forward", "dwdf_backward", "dwdf_train_pass", "dwdf_adam_step", "dwdf_forward_host", "dwdf_grad_host", "dwdf_process_block",
    "dwdf_backward_raw", "dwdf_train_pass_raw", "dwdf_finalize", "dwdf_mlp_weight_count", "dwdf_program_create_neural", "dwdf_forward_neural", "dwdf_backward_neural", "dwdf_neural_ckpt_bytes", "dwdf_neural_workspace_bytes", "dwdf_adam_step_vec", "dwdf_last_error", "dwdf_build_info", "dwdf_train_step", "dwdf_launch_count", "dwdf_set_tma", "dwdf_set_option", "dwdf_time_parallel_redone",
    "dwdf_comm_create", "dwdf_comm_handle_bytes", "dwdf_comm_get_handle", "dwdf_comm_connect", "dwdf_comm_set_timeout", "dwdf_comm_destroy", "dwdf_allreduce_sum", "dwdf_train_step_dp", "dwdf_profile_begin", "dwdf_profile_end",
    "dwdf_backward_neural_raw", "dwdf_finalize_neural", "dwdf_train_step_neural",
    "dwdf_program_specialize", "dwdf_program_is_specialized", "dwdf_program_specialized_source",
)


class Node(C.Structure):
    _fields_ = [("kind", C.c_int32), ("child1", C.c_int32), ("child2", C.c_int32), ("param", C.c_int32)]


class CircuitDesc(C.Structure):
    _fields_ = [
        ("root_kind", C.c_int32), ("root_mode", C.c_int32), ("ordering", C.c_int32), ("probe", C.c_int32), ("source", C.c_int32), ("r_node", C.c_int32),
        ("param_Is", C.c_int32), ("param_nabla", C.c_int32), ("n_params", C.c_int32), ("newton_max_iter", C.c_int32),
        ("fs", C.c_float), ("Vt", C.c_float), ("n_up", C.c_float), ("n_down", C.c_float), ("newton_tol", C.c_float), ("probe_current", C.c_int32),
    ]


class MlpDesc(C.Structure):
    _fields_ = [("n_hidden", C.c_int32), ("hidden", C.c_int32)]


class DwdfError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libdwdf: {STATUS.get(status, status)}: {message}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """Loads libdwdf.so once. Raises (loudly) if it was never built: run ``__graft_entry__.build()``."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: the CUDA extension was not built (python -c 'import __graft_entry__ as g; g.build()'). "
                          "This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    L.dwdf_program_create.argtypes = [C.POINTER(Node), i32, C.POINTER(CircuitDesc), C.POINTER(vp)]
    L.dwdf_program_destroy.argtypes = [vp]
    L.dwdf_program_is_clipper.argtypes = [vp]
    L.dwdf_program_n_states.argtypes = [vp]
    L.dwdf_program_specialize.argtypes = [vp]
    L.dwdf_program_is_specialized.argtypes = [vp]
    L.dwdf_program_specialized_source.argtypes = [vp, i32, C.c_char_p, sz]
    L.dwdf_program_specialized_source.restype = sz
    L.dwdf_ckpt_bytes.argtypes = [vp, i64, i64]
    L.dwdf_ckpt_bytes.restype = sz
    L.dwdf_workspace_bytes.argtypes = [vp, i64, i64]
    L.dwdf_workspace_bytes.restype = sz
    L.dwdf_forward.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, vp]
    L.dwdf_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, vp, vp, vp, sz, i64, i64, vp]
    L.dwdf_train_pass.argtypes = [vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, sz, i64, i64, vp]
    L.dwdf_backward_raw.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, sz, i64, i64, vp]
    L.dwdf_train_pass_raw.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, sz, i64, i64, vp]
    L.dwdf_finalize.argtypes = [vp, vp, i32, i32, vp, vp]
    L.dwdf_adam_step.argtypes = [vp, vp, vp, vp, vp, i32, C.c_float, vp, C.c_float, C.c_float, C.c_float, C.c_double, vp, vp, vp]
    L.dwdf_train_step.argtypes = [vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, sz, vp, vp, vp, C.c_float, vp, C.c_float, C.c_float, C.c_float, vp, vp, i64, i64, vp]
    L.dwdf_comm_create.argtypes = [i32, i32, C.POINTER(vp)]
    L.dwdf_comm_handle_bytes.restype = sz
    L.dwdf_comm_get_handle.argtypes = [vp, vp]
    L.dwdf_comm_connect.argtypes = [vp, vp]
    L.dwdf_comm_set_timeout.argtypes = [vp, C.c_double]
    L.dwdf_comm_destroy.argtypes = [vp]
    L.dwdf_allreduce_sum.argtypes = [vp, vp, i64, vp]
    L.dwdf_train_step_dp.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, sz, vp, vp, vp, C.c_float, vp, C.c_float, C.c_float, C.c_float, vp, vp, i64, i64, vp]
    L.dwdf_profile_begin.argtypes = [i32]
    L.dwdf_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i32)]
    L.dwdf_forward_host.argtypes = [vp, vp, vp, vp, vp, i64, i64]
    L.dwdf_grad_host.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64, vp, vp, i64, i64]
    L.dwdf_process_block.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, vp]
    L.dwdf_mlp_weight_count.argtypes = [C.POINTER(MlpDesc)]
    L.dwdf_mlp_weight_count.restype = sz
    L.dwdf_program_create_neural.argtypes = [C.POINTER(Node), i32, C.POINTER(CircuitDesc), C.POINTER(MlpDesc), C.POINTER(vp)]
    L.dwdf_forward_neural.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp]
    L.dwdf_backward_neural.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, vp, vp, vp, vp, sz, i64, i64, vp]
    L.dwdf_backward_neural_raw.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, sz, i64, i64, vp]
    L.dwdf_finalize_neural.argtypes = [vp, i32, i32, vp, vp, vp]
    L.dwdf_train_step_neural.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, sz, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, i64, i64, vp]
    L.dwdf_neural_ckpt_bytes.argtypes = [vp, i64, i64]
    L.dwdf_neural_ckpt_bytes.restype = sz
    L.dwdf_neural_workspace_bytes.argtypes = [vp, i64, i64]
    L.dwdf_neural_workspace_bytes.restype = sz
    L.dwdf_adam_step_vec.argtypes = [vp, vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double, vp]
    L.dwdf_last_error.restype = C.c_char_p
    L.dwdf_build_info.restype = C.c_char_p
    L.dwdf_launch_count.restype = i64
    L.dwdf_set_tma.argtypes = [C.c_int]
    L.dwdf_set_option.argtypes = [C.c_int]
    L.dwdf_time_parallel_redone.restype = i64
    _lib = L
    return L


def check(status: int) -> None:
    if status != 0:
        raise DwdfError(status, lib().dwdf_last_error().decode())


def launch_count() -> int:
    return int(lib().dwdf_launch_count())


def build_info() -> str:
    return lib().dwdf_build_info().decode()


def set_tma(enable: bool) -> bool:
    return bool(lib().dwdf_set_tma(1 if enable else 0))


def set_option(bits: int) -> int:
    return int(lib().dwdf_set_option(int(bits)))


def time_parallel_redone() -> int:
    return int(lib().dwdf_time_parallel_redone())


def profile_begin(max_steps: int) -> None:
    check(lib().dwdf_profile_begin(int(max_steps)))


def profile_end():
    """-> (ms_forward, ms_adjoint, ms_tail, steps): mean device time per phase of the training steps since profile_begin."""
    f, a, t, n = C.c_double(), C.c_double(), C.c_double(), C.c_int32()
    check(lib().dwdf_profile_end(C.byref(f), C.byref(a), C.byref(t), C.byref(n)))
    return f.value, a.value, t.value, n.value
