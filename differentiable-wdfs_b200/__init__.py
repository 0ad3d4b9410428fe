"""differentiable-wdfs_b200 — B200-native differentiable wave digital filters.

Import with ``importlib.import_module("differentiable-wdfs_b200")`` (the directory name carries the
reference's hyphen). The element API mirrors ``wdf_py/lib/tf_wdf.py``; the compiled path runs the
fused sm_100a kernels of ``csrc/`` through the C ABI of ``include/dwdf.h`` (``libdwdf.so``).
"""
from . import _lib
from ._lib import DwdfError, build_info, launch_count, set_option, set_tma, time_parallel_redone
from . import dataimport, diode_config, model_io
from .wdf import (Adam, AdamWeights, Capacitor, CapacitorAlpha, CompiledCircuit, Diode, IdealCurrentSource, Inductor, InductorAlpha, ResistiveCurrentSource, Switch, YParameter, current, omega4_approx, DenseLayer, DenseRootModel, DiodePair, IdealVoltageSource, Inverter, Parallel, PolarityInverter, ResistiveVoltageSource, Resistor, Series, compile_circuit, voltage,
                  wright_omega)
from .data_parallel import DataParallelTrainer, PeerComm, shard_rows

__all__ = ["Adam", "AdamWeights", "Capacitor", "CapacitorAlpha", "Diode", "IdealCurrentSource", "Inductor", "InductorAlpha", "ResistiveCurrentSource", "Switch", "YParameter", "current", "omega4_approx", "CompiledCircuit", "DataParallelTrainer", "PeerComm", "DenseLayer", "DenseRootModel", "DiodePair", "dataimport", "diode_config", "model_io", "DwdfError", "IdealVoltageSource", "Inverter", "Parallel", "PolarityInverter", "ResistiveVoltageSource",
           "Resistor", "Series", "build_info", "compile_circuit", "launch_count", "set_tma", "shard_rows", "voltage", "wright_omega"]
