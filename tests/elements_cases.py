"""The circuits of tests/golden/ref_elements.npz (make_golden_elements.py), once as the oracle's node lists and once as
builders over the product's element API. Each entry: golden key -> dict(fs, nodes, root_kind, root_par, source, probe,
probe_current, build(dwdf) -> (root, tree, probe element, probe_kind))."""
from oracle.cpu import (CAPACITOR, CAPACITOR_ALPHA, INDUCTOR, INDUCTOR_ALPHA, INVERTER, PARALLEL, RESCS, RESISTOR, RESVS, ROOT_DIODE, ROOT_IDEAL_CS, ROOT_IDEAL_VS, ROOT_SWITCH, SERIES, YPARAM)


def _rlc(alpha):
    def build(w):
        r1 = w.Resistor(300.0)
        c1 = w.Capacitor(1.0e-6, 44100.0) if alpha is None else w.CapacitorAlpha(1.0e-6, 44100.0, alpha)
        l1 = w.Inductor(0.022, 44100.0) if alpha is None else w.InductorAlpha(0.022, 44100.0, alpha)
        top = w.Inverter(w.Series(w.Series(r1, c1), l1))
        return w.IdealVoltageSource(), top, l1, "voltage"
    ck, lk = (CAPACITOR, INDUCTOR) if alpha is None else (CAPACITOR_ALPHA, INDUCTOR_ALPHA)
    extra = () if alpha is None else (alpha,)
    nodes = [(RESISTOR, -1, -1, 300.0), (ck, -1, -1, 1.0e-6) + extra, (SERIES, 0, 1, 0.0), (lk, -1, -1, 0.022) + extra, (SERIES, 2, 3, 0.0), (INVERTER, 4, -1, 0.0)]
    return dict(fs=44100.0, nodes=nodes, root_kind=ROOT_IDEAL_VS, root_par=None, source=-1, probe=3, probe_current=False, build=build)


def _divider():
    def build(w):
        r1, r2 = w.Resistor(10000.0), w.Resistor(4700.0)
        p1 = w.Parallel(r1, r2)
        return w.IdealCurrentSource(p1), p1, r2, "current"
    return dict(fs=48000.0, nodes=[(RESISTOR, -1, -1, 10000.0), (RESISTOR, -1, -1, 4700.0), (PARALLEL, 0, 1, 0.0)], root_kind=ROOT_IDEAL_CS, root_par=None, source=-1, probe=1, probe_current=True, build=build)


def _switch(closed):
    def build(w):
        r1, cs = w.Resistor(10000.0), w.ResistiveCurrentSource(1.0e9)
        s1 = w.Series(r1, cs)
        return w.Switch(s1, closed=bool(closed)), s1, r1, "current"
    return dict(fs=48000.0, nodes=[(RESISTOR, -1, -1, 10000.0), (RESCS, -1, -1, 1.0e9), (SERIES, 0, 1, 0.0)], root_kind=ROOT_SWITCH, root_par=[closed, 0, 0, 0, 0, 0, 0], source=1, probe=0, probe_current=True, build=build)


def _ypar(current):
    def build(w):
        res = w.Resistor(10000.0)
        yp = w.YParameter(res, 0.11, 0.22, 0.33, 0.44)
        return w.IdealVoltageSource(), yp, (yp if current else res), ("current" if current else "voltage")
    return dict(fs=48000.0, nodes=[(RESISTOR, -1, -1, 10000.0), (YPARAM, 0, -1, 0.11, 0.22, 0.33, 0.44)], root_kind=ROOT_IDEAL_VS, root_par=None, source=-1, probe=1 if current else 0, probe_current=bool(current), build=build)


def _shockley():
    def build(w):
        vs = w.ResistiveVoltageSource(1.0e-9)
        i1 = w.Inverter(vs)
        return w.Diode(i1, 1.0e-7, 25.85e-3, 1.0), i1, vs, "voltage"
    return dict(fs=48000.0, nodes=[(RESVS, -1, -1, 1.0e-9), (INVERTER, 0, -1, 0.0)], root_kind=ROOT_DIODE, root_par=[0, 0, 1.0e-7, 25.85e-3, 1, 1, 1], source=0, probe=0, probe_current=False, build=build)


def _rectifier(n_diodes):
    def build(w):
        vs, c1 = w.ResistiveVoltageSource(4700.0), w.Capacitor(47.0e-9, 48000.0)
        p1 = w.Parallel(vs, c1)
        return w.Diode(p1, 2.52e-9, 25.85e-3, n_diodes), p1, c1, "voltage"
    return dict(fs=48000.0, nodes=[(RESVS, -1, -1, 4700.0), (CAPACITOR, -1, -1, 47.0e-9), (PARALLEL, 0, 1, 0.0)], root_kind=ROOT_DIODE, root_par=[0, 0, 2.52e-9, 25.85e-3, n_diodes, 1, 1], source=0, probe=1, probe_current=False, build=build)


CASES = {
    "current_divider": _divider(),
    "current_switch_closed": _switch(1),
    "current_switch_open": _switch(0),
    "rlc_plain": _rlc(None),
    "rlc_alpha_1.0": _rlc(1.0),
    "rlc_alpha_0.5": _rlc(0.5),
    "rlc_alpha_0.1": _rlc(0.1),
    "ypar_voltage": _ypar(False),
    "ypar_current": _ypar(True),
    "shockley_voltage": _shockley(),
    "rectifier_voltage": _rectifier(1.0),
    "rectifier_2diodes": _rectifier(2.0),
}
