#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

Sources of truth, none of them restated here:
  * oracle/_ref/libdwdf_ref.so  — the unmodified reference C++ (chowdsp_wdf templates, omega.h,
    Toms917DiodePair.h, modules/toms917) compiled in place by oracle/Makefile (-O2, no contraction);
  * wdf_py/diode_clipper/diode_pretraining.py:39-60 ``diode_pair_func`` — extracted from the
    reference file with ``ast`` and executed as is (the module itself imports TensorFlow and cannot
    be imported; the function only needs numpy + scipy.special.wrightomega);
  * wdf_py/diode_clipper/diode_config.py — imported as is;
  * the 41-entry Wright-omega table of wdf_tests/OmegaTest.cpp:6-48 — parsed out of the file.
The GPU box has no /root/reference: tests read only the files written here.
"""
import ast
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle.cpu import ORDER_PLUGIN, ORDER_PYTHON, ROOT_APPROX, ROOT_APPROX_GOOD, ROOT_EXACT, ClipperParams, Ref  # noqa: E402


def make_inputs(B, T, fs, seed):
    """SURVEY.md §8(d) config-2 style input: per-sequence sine burst + noise."""
    rng = np.random.default_rng(seed)
    n = np.arange(T)
    A = rng.uniform(0.1, 2.0, B)
    f = np.exp(rng.uniform(np.log(50.0), np.log(5000.0), B))
    x = A[:, None] * np.sin(2 * np.pi * f[:, None] * n[None, :] / fs) + 0.05 * rng.standard_normal((B, T))
    return x.astype(np.float32)


def main():
    ref = Ref()
    out = {}

    # ---- (1) omega known-answer table, OmegaTest.cpp:6-48 ------------------------------------
    src = open(f"{REF}/modules/chowdsp_utils/tests/chowdsp_utils_test/dsp_tests/wdf_tests/OmegaTest.cpp").read()
    table = re.findall(r"\{\s*(-?[0-9.]+),\s*([0-9.e+-]+)\s*\}", src.split("WO_vals")[1].split("};")[0])
    omega_table = [[float(a), float(b)] for a, b in table]
    assert len(omega_table) == 41
    known = {
        "omega_table": omega_table,
        "omega_table_source": "wdf_tests/OmegaTest.cpp:6-48",
        # tolerances asserted by the reference test, OmegaTest.cpp:112-164
        "omega_tolerances": {"log2_approx": 0.008, "log_approx": 0.005, "pow2_approx": 0.001, "exp_approx": 0.005, "omega1": 2.1, "omega2": 2.1, "omega3": 0.3, "omega4": 0.05},
    }

    # ---- (2) the reference's own fixtures, run here ----------------------------------------------
    known["standalone_test"] = {"value": ref.standalone_test(), "expected": 4.77, "tol": 0.1, "source": "wdf_standalone_test.cpp:16-36"}
    known["static_wdf_test"] = {"inputs": [1.0, 0.5, 0.0, -0.5, -1.0], "fs": 44100.0, "C": 47.0e-9, "R": 4700.0, "Is": 2.52e-9, "best": ref.static_wdf_test(1).tolist(), "good": ref.static_wdf_test(0).tolist(), "source": "StaticWDFTest.cpp:216-271"}
    p = ClipperParams()
    imp = np.zeros((1, 16), np.float32)
    imp[0, 0] = 1.0
    known["plugin_impulse_response_toms"] = ref.clipper(imp, p, ROOT_EXACT, ORDER_PLUGIN)[0].tolist()
    known["plugin_impulse_response_omega4"] = ref.clipper(imp, p, ROOT_APPROX, ORDER_PLUGIN)[0].tolist()
    known["port_impedance_f32"] = ref.port_impedance(p)
    spot = np.array([-2.0, 0.5, 1.0], np.float32)
    known["pair_law_spot"] = {
        "a": spot.tolist(),
        "Rp": known["port_impedance_f32"],
        "toms_f32": ref.diode_pair(spot, known["port_impedance_f32"], p, ROOT_EXACT).tolist(),
        "omega4_f32": ref.diode_pair(spot, known["port_impedance_f32"], p, ROOT_APPROX).tolist(),
    }

    # RC low-pass magnitudes, WDFTest.cpp:96-140 / StaticWDFTest.cpp:83-128: fs 44.1k, fc 500, C 1uF
    fs, fc, Cv = 44100.0, 500.0, 1.0e-6
    Rv = 1.0 / (2 * np.pi * fc * Cv)
    mags = {}
    for name, f in (("2fc", 2 * fc), ("fc", fc), ("fc/2", fc / 2)):
        n = np.arange(int(fs))
        y = ref.rc_lowpass(np.sin(2 * np.pi * f * n / fs), fs, Rv, Cv)
        mags[name] = float(20 * np.log10(np.max(np.abs(y[len(y) // 2 :]))))
    known["rc_lowpass_mag_db"] = {"fs": fs, "fc": fc, "C": Cv, "R": Rv, "measured": mags, "expected": {"2fc": -7.0, "fc": -3.0, "fc/2": -1.0}, "tol": 0.1, "source": "WDFTest.cpp:96-140"}
    known["divider"] = {"value": float(ref.voltage_divider(np.array([10.0]), 10000.0, 10000.0)[0]), "expected": 5.0, "source": "CommonWDFTests.h:6-23"}

    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known, f, indent=1)

    # ---- (3) scalar function grids ---------------------------------------------------------------
    xg = np.concatenate([np.linspace(-60, 60, 4801), np.array([-3.341459552768620, 8.0, -126.0 / 1.442695040888963, -3.0, -2.0, 0.0, 1.0, 1 + np.pi])]).astype(np.float32)
    out["omega_x"] = xg
    out["omega3_f32"] = ref.omega("omega3", xg)
    out["omega4_f32"] = ref.omega("omega4", xg)
    out["omega4_f64"] = ref.omega("omega4", xg.astype(np.float64), np.float64)
    out["exp_approx_f32"] = ref.omega("exp_approx", xg)
    xpos = np.abs(xg) + np.float32(1e-3)
    out["log_x"] = xpos
    out["log_approx_f32"] = ref.omega("log_approx", xpos)
    out["toms917_f64"] = ref.toms917(xg.astype(np.float64))

    # ---- (4) diode-pair laws -----------------------------------------------------------------------
    a = np.concatenate([np.linspace(-3, 3, 1201), [0.0, -0.0, 1e-20, -1e-20]]).astype(np.float32)
    out["pair_a"] = a
    for Rp in (4301.5083, 100.0, 1.0e6):
        tag = f"Rp{Rp:g}"
        out[f"pair_best_f32_{tag}"] = ref.diode_pair(a, Rp, p, ROOT_APPROX)
        out[f"pair_good_f32_{tag}"] = ref.diode_pair(a, Rp, p, ROOT_APPROX_GOOD)
        out[f"pair_toms_f32_{tag}"] = ref.diode_pair(a, Rp, p, ROOT_EXACT)
        out[f"pair_toms_f64_{tag}"] = ref.diode_pair(a.astype(np.float64), Rp, p, ROOT_EXACT, np.float64)

    # eq. (45), the Python reference's own function executed as is
    sys.path.insert(0, f"{REF}/wdf_py/diode_clipper")
    import diode_config  # noqa: E402
    from scipy.special import wrightomega  # noqa: E402

    tree = ast.parse(open(f"{REF}/wdf_py/diode_clipper/diode_pretraining.py").read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "diode_pair_func"][0]
    ns = {"np": np, "wrightomega": wrightomega}
    exec(compile(ast.Module([fn], []), "diode_pretraining.py", "exec"), ns)
    diode_pair_func = ns["diode_pair_func"]
    a45 = np.linspace(-2.5, 2.5, 501)
    R45 = 10.0 ** np.linspace(1, 9, 9)
    out["eq45_a"] = a45
    out["eq45_R"] = R45
    names = ["diode_1n4148_1u1d", "diode_1n4148_1u2d", "diode_1n4148_1u3d", "diode_1n4148_2u2d", "diode_1n4148_2u3d", "diode_1n4148_3u3d"]
    cfgs = {}
    for nm in names:
        d = getattr(diode_config, nm)
        cfgs[nm] = dict(Is=d.Is, nabla=d.nabla, Vt=d.Vt, N_up=d.N_up, N_down=d.N_down)
        out[f"eq45_{nm}"] = np.array([[diode_pair_func(av, Rv_, d) for av in a45] for Rv_ in R45], np.float32)
    with open(os.path.join(HERE, "diode_configs.json"), "w") as f:
        json.dump(cfgs, f, indent=1)

    # ---- (5) clipper trajectories ------------------------------------------------------------------
    B, T = 8, 512
    x = make_inputs(B, T, 48000.0, 1234)
    out["clip_x"] = x
    cases = {"plugin": ClipperParams(), "training": ClipperParams(R=45.0e3, C=4.7e-9)}
    for cname, cp in cases.items():
        for rname, root in (("approx", ROOT_APPROX), ("exact", ROOT_EXACT)):
            for oname, order in (("plugin", ORDER_PLUGIN), ("python", ORDER_PYTHON)):
                out[f"clip_{cname}_{rname}_{oname}_f32"] = ref.clipper(x, cp, root, order)
                out[f"clip_{cname}_{rname}_{oname}_f64"] = ref.clipper(x.astype(np.float64), cp, root, order, np.float64)
    # a loud one: 10 V peaks (hard conduction, |A| -> 1)
    xl = (make_inputs(4, 512, 48000.0, 99) * 5).astype(np.float32)
    out["clip_loud_x"] = xl
    out["clip_loud_exact_python_f32"] = ref.clipper(xl, ClipperParams(), ROOT_EXACT, ORDER_PYTHON)
    out["clip_loud_approx_python_f32"] = ref.clipper(xl, ClipperParams(), ROOT_APPROX, ORDER_PYTHON)

    # ---- (6) RC low-pass + divider (config 1) ----------------------------------------------------
    rng = np.random.default_rng(0)
    xr = (0.5 * rng.standard_normal(1024)).astype(np.float32)
    out["lpf_x"] = xr
    out["lpf_y_f64"] = ref.rc_lowpass(xr.astype(np.float64), 48000.0, 1000.0, 1.0e-6, 0)
    out["lpf_y_f32"] = ref.rc_lowpass(xr, 48000.0, 1000.0, 1.0e-6, 0, np.float32)
    out["lpf_vr_f64"] = ref.rc_lowpass(xr.astype(np.float64), 48000.0, 1000.0, 1.0e-6, 1)
    out["divider_y_f64"] = ref.voltage_divider(xr.astype(np.float64), 2000.0, 100.0)

    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "ref_vectors.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
