"""Golden vectors for the remaining chowdsp_wdf elements (inductor, alpha-transform C / L, current sources, Y-parameter,
single diode, switch), produced in the build container by the UNMODIFIED reference classes wired as in the reference's own
tests (oracle/ref_elements_harness.cpp -> oracle/_ref/libdwdf_ref_elements.so):

    python tests/golden/make_golden_elements.py        ->  tests/golden/ref_elements.npz

Each case stores its input signal, the reference's output and the circuit as the node list the oracle / the engine take.
Known answers the reference's tests assert are checked here too (current divider 0.5 A, switch -1 A / 0 A, the Y-parameter
port equations, the Shockley current, the alpha-transform passband gains).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.cpu import RefElements  # noqa: E402

if __name__ == "__main__":
    ref = RefElements()
    rng = np.random.default_rng(20)
    T = 2048
    n = np.arange(T)
    x = (0.6 * np.sin(2 * np.pi * 997.0 * n / 44100.0) + 0.3 * rng.standard_normal(T)).astype(np.float32)
    out = {"x": x}
    out["current_divider"] = ref.current_divider(x, 10000.0, 4700.0)
    out["current_switch_closed"] = ref.current_switch(x, 10000.0, 1.0e9, 1)
    out["current_switch_open"] = ref.current_switch(x, 10000.0, 1.0e9, 0)
    out["rlc_plain"] = ref.rlc_highpass(x, 44100.0, 300.0, 1.0e-6, 0.022)
    for a in (1.0, 0.5, 0.1):
        out[f"rlc_alpha_{a}"] = ref.rlc_highpass(x, 44100.0, 300.0, 1.0e-6, 0.022, a)
    out["ypar_voltage"] = ref.y_parameter(x, 10000.0, 0.11, 0.22, 0.33, 0.44, 0)
    out["ypar_current"] = ref.y_parameter(x, 10000.0, 0.11, 0.22, 0.33, 0.44, 1)
    out["shockley_voltage"] = ref.diode(x, 48000.0, 1.0e-9, 0.0, 1.0e-7, 25.85e-3, 1.0, 0)
    out["rectifier_voltage"] = ref.diode(x, 48000.0, 4700.0, 47.0e-9, 2.52e-9, 25.85e-3, 1.0, 0)
    out["rectifier_2diodes"] = ref.diode(x, 48000.0, 4700.0, 47.0e-9, 2.52e-9, 25.85e-3, 2.0, 0)

    # the reference's own assertions on these circuits
    one = np.ones(4, np.float32)
    assert np.all(ref.current_divider(one, 10000.0, 10000.0) == 0.5)  # StaticWDFTest.cpp:53-54
    assert np.all(np.abs(ref.current_switch(one, 10000.0, 1.0e9, 1) + 1.0) < 1e-3)  # :70-71
    assert np.all(ref.current_switch(one, 10000.0, 1.0e9, 0) == 0.0)  # :78-79
    v = 2.0 * one  # WDFTest.cpp:62-73: -I(res) = y11 V(res) + y12 V, I(port 2) = y21 V(res) + y22 V
    vres, i2 = ref.y_parameter(v, 10000.0, 0.11, 0.22, 0.33, 0.44, 0), ref.y_parameter(v, 10000.0, 0.11, 0.22, 0.33, 0.44, 1)
    assert np.all(np.abs(i2 - (0.33 * vres + 0.44 * 2.0)) < 1e-3)
    # (WDFTest.cpp:76-94's Shockley current is asserted by the reference in double only — in float (a - b) / (2 R) with the
    #  source's 1e-9 ohm has no digits left; tests/test_oracle.py checks it through the oracle's double instantiation)
    sine = np.sin(2 * np.pi * 10.0e3 * np.arange(44100) / 44100.0).astype(np.float32)  # StaticWDFTest.cpp:172-212: passband gains

    def mag_db(y):
        return 20 * np.log10(np.max(np.abs(y[1000:-1000])))
    ref_mag = mag_db(ref.rlc_highpass(sine, 44100.0, 300.0, 1.0e-6, 0.022))
    assert abs(ref_mag) < 0.1 and abs(mag_db(ref.rlc_highpass(sine, 44100.0, 300.0, 1.0e-6, 0.022, 1.0)) - ref_mag) < 1e-6
    assert abs(mag_db(ref.rlc_highpass(sine, 44100.0, 300.0, 1.0e-6, 0.022, 0.1)) - (ref_mag - 1.1)) < 0.1
    out["known_alpha_passband_db"] = np.array([ref_mag, mag_db(ref.rlc_highpass(sine, 44100.0, 300.0, 1.0e-6, 0.022, 0.1))])
    np.savez_compressed(os.path.join(HERE, "ref_elements.npz"), **out)
    print({k: (v.shape, float(np.max(np.abs(v)))) for k, v in out.items()})
