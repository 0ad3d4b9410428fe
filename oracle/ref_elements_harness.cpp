// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// oracle/_ref/libdwdf_ref_elements.so: the UNMODIFIED chowdsp_wdf element classes (wdf_t.h) of the remaining one-ports,
// the two-port and the non-adaptable roots — InductorT, CapacitorAlphaT, InductorAlphaT, ResistiveCurrentSourceT,
// YParameterT, IdealCurrentSourceT, DiodeT, SwitchT — wired into the circuits of the reference's own tests
// (StaticWDFTest.cpp / WDFTest.cpp) and driven sample by sample with a caller-supplied signal. Nothing here restates
// reference arithmetic: only the harness loops (which elements, which probe) are ours, each citing the test it follows.
// tests/golden/make_golden_elements.py runs it in the build container to produce tests/golden/ref_elements.npz.
#include <cstdint>
#include <cmath>

#include <pch.h> // oracle/shim/pch.h (-> <wdf_t.h>)

using namespace chowdsp;

extern "C" {

// StaticWDFTest.cpp:39-55 (currentDividerTest): Parallel(r1, r2) closed by an ideal current source; out = current(r2)
void ref_el_current_divider (const float* x, float* y, int64_t n, float R1, float R2)
{
    wdft::ResistorT<float> r1 (R1), r2 (R2);
    auto p1 = wdft::makeParallel<float> (r1, r2);
    wdft::IdealCurrentSourceT<float, decltype (p1)> is { p1 };
    for (int64_t i = 0; i < n; ++i)
    {
        is.setCurrent (x[i]);
        is.incident (p1.reflected());
        p1.incident (is.reflected());
        y[i] = wdft::current<float> (r2);
    }
}

// StaticWDFTest.cpp:57-80 (currentSwitchTest): Series(r1, ResistiveCurrentSource) closed by a switch; out = current(r1)
void ref_el_current_switch (const float* x, float* y, int64_t n, float R1, float Rs, int closed)
{
    wdft::ResistorT<float> r1 (R1);
    wdft::ResistiveCurrentSourceT<float> Is (Rs);
    auto s1 = wdft::makeSeries<float> (r1, Is);
    wdft::SwitchT<float, decltype (s1)> sw { s1 };
    sw.setClosed (closed != 0);
    for (int64_t i = 0; i < n; ++i)
    {
        Is.setCurrent (x[i]);
        sw.incident (s1.reflected());
        s1.incident (sw.reflected());
        y[i] = wdft::current<float> (r1);
    }
}

// StaticWDFTest.cpp:140-215 (alphaTransformTest): 2nd-order highpass Inverter(Series(Series(R, C), L)) closed by an ideal
// voltage source, probe = voltage(l1). alpha < 0: the plain CapacitorT / InductorT (the test's reference filter).
void ref_el_rlc_highpass (const float* x, float* y, int64_t n, float fs, float R, float C, float L, float alpha)
{
    if (alpha < 0.0f)
    {
        wdft::CapacitorT<float> c1 (C);
        wdft::ResistorT<float> r1 (R);
        wdft::InductorT<float> l1 (L);
        c1.prepare (fs);
        l1.prepare (fs);
        auto s1 = wdft::makeSeries<float> (r1, c1);
        auto s2 = wdft::makeSeries<float> (s1, l1);
        auto p1 = wdft::makeInverter<float> (s2);
        wdft::IdealVoltageSourceT<float, decltype (p1)> vs { p1 };
        for (int64_t i = 0; i < n; ++i)
        {
            vs.setVoltage (x[i]);
            vs.incident (p1.reflected());
            p1.incident (vs.reflected());
            y[i] = wdft::voltage<float> (l1);
        }
        return;
    }
    wdft::CapacitorAlphaT<float> c1 (C);
    wdft::ResistorT<float> r1 (R);
    wdft::InductorAlphaT<float> l1 (L);
    auto s1 = wdft::makeSeries<float> (r1, c1);
    auto s2 = wdft::makeSeries<float> (s1, l1);
    auto p1 = wdft::makeInverter<float> (s2);
    wdft::IdealVoltageSourceT<float, decltype (p1)> vs { p1 };
    c1.prepare (fs);
    c1.setAlpha (alpha);
    l1.prepare (fs);
    l1.setAlpha (alpha);
    for (int64_t i = 0; i < n; ++i)
    {
        vs.setVoltage (x[i]);
        vs.incident (p1.reflected());
        p1.incident (vs.reflected());
        y[i] = wdft::voltage<float> (l1);
    }
}

// WDFTest.cpp:55-74 (yParameterTest) with the compile-time classes: Resistor behind a Y-parameter two-port, ideal voltage
// source; out = voltage(res) (probe 0) or current(yParam) (probe 1)
void ref_el_y_parameter (const float* x, float* y, int64_t n, float R, float y11, float y12, float y21, float y22, int probe)
{
    wdft::ResistorT<float> res (R);
    wdft::YParameterT<float, decltype (res)> yParam { res, y11, y12, y21, y22 };
    wdft::IdealVoltageSourceT<float, decltype (yParam)> vs { yParam };
    for (int64_t i = 0; i < n; ++i)
    {
        vs.setVoltage (x[i]);
        vs.incident (yParam.reflected());
        yParam.incident (vs.reflected());
        y[i] = probe == 0 ? wdft::voltage<float> (res) : wdft::current<float> (yParam);
    }
}

// WDFTest.cpp:76-94 (shockleyDiodeTest) with the compile-time classes: Inverter(ResistiveVoltageSource) closed by a single
// diode; out = voltage(Vs) (probe 0) or the diode's current (probe 1). A capacitor in parallel (C > 0) makes it a rectifier.
void ref_el_diode (const float* x, float* y, int64_t n, float fs, float Rs, float C, float Is, float Vt, float nDiodes, int probe)
{
    if (C > 0.0f)
    {
        wdft::ResistiveVoltageSourceT<float> Vs (Rs);
        wdft::CapacitorT<float> c1 (C);
        c1.prepare (fs);
        auto p1 = wdft::makeParallel<float> (Vs, c1);
        wdft::DiodeT<float, decltype (p1)> d1 { p1, Is, Vt, nDiodes };
        for (int64_t i = 0; i < n; ++i)
        {
            Vs.setVoltage (x[i]);
            d1.incident (p1.reflected());
            p1.incident (d1.reflected());
            y[i] = probe == 0 ? wdft::voltage<float> (c1) : wdft::current<float> (d1);
        }
        return;
    }
    wdft::ResistiveVoltageSourceT<float> Vs (Rs);
    auto i1 = wdft::makeInverter<float> (Vs);
    wdft::DiodeT<float, decltype (i1)> d1 { i1, Is, Vt, nDiodes };
    for (int64_t i = 0; i < n; ++i)
    {
        Vs.setVoltage (x[i]);
        d1.incident (i1.reflected());
        i1.incident (d1.reflected());
        y[i] = probe == 0 ? wdft::voltage<float> (Vs) : wdft::current<float> (d1);
    }
}

} // extern "C"
