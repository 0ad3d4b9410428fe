"""Diode configurations the reference's scripts select by name (wdf_py/diode_clipper/diode_config.py:5-38).

Same tuple type and field order (``name, Is, nabla, Vt, N_up, N_down``) and the same six 1N4148 series / anti-parallel
arrangements, so ``from diode_config import diode_1n4148_1u1d`` keeps working against this package
(``importlib.import_module("differentiable-wdfs_b200").diode_config``). ``as_diode_pair`` builds the analytic root from one.
"""
from collections import namedtuple

DiodeConfig = namedtuple("DiodeConfig", ["name", "Is", "nabla", "Vt", "N_up", "N_down"], defaults=["", 1.0e-9, 1.0, 25.85e-3, 1, 1])

default_diode = DiodeConfig("DefaultDiode")

# 1N4148: Is = 4.352 nA, ideality 1.906 (diode_config.py:14-16, the SPICE model the reference cites)
_IS_1N4148, _NABLA_1N4148 = 4.352e-9, 1.906


def _cfg(n_up, n_down):
    return DiodeConfig(f"1N4148 ({n_up}U-{n_down}D)", Is=_IS_1N4148, nabla=_NABLA_1N4148, N_up=n_up, N_down=n_down)


diode_1n4148_1u1d = _cfg(1, 1)
diode_1n4148_1u2d = _cfg(1, 2)
diode_1n4148_1u3d = _cfg(1, 3)
diode_1n4148_2u2d = _cfg(2, 2)
diode_1n4148_2u3d = _cfg(2, 3)
diode_1n4148_3u3d = _cfg(3, 3)


def as_diode_pair(diode: DiodeConfig, next, trainable=False, mode="exact"):
    """The analytic root of a configuration: diode_pretraining.py:39-60's law with this tuple's constants."""
    from .wdf import DiodePair

    return DiodePair(next, diode.Is, diode.Vt, diode.nabla, diode.N_up, diode.N_down, trainable=trainable, mode=mode)
