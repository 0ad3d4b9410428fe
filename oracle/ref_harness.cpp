// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// oracle/_ref/libdwdf_ref.so: the UNMODIFIED reference C++ (chowdsp_wdf templates, omega.h,
// Toms917DiodePair.h, modules/toms917/toms917.cpp) compiled in place from /root/reference and
// exposed through a flat C ABI so that tests/ and bench.py's reference arm can call it with ctypes.
// Nothing here restates reference arithmetic: every number comes out of the reference's own
// headers. Only the harness loops (which element to build, which loop to run) are ours; each cites
// the reference call site it reproduces.
//
// Build: see oracle/Makefile (g++ only; the reference's CMake/JUCE build is not used).
#include <cstdint>
#include <cstring>
#include <cmath>
#include <complex>
#include <thread>
#include <vector>
#include <algorithm>

#include <pch.h>                                     // oracle/shim/pch.h  (-> <wdf_t.h>)
#include <wdf.h>                                     // run-time API, used by wdf_standalone_test
#include <toms917.hpp>                               // modules/toms917
#include "dsp/diode_clipper/Toms917DiodePair.h"      // plugin/src

namespace
{
using namespace chowdsp;

enum RootKind
{
    kRootApprox = 0, // wdft::DiodePairT<..., Best>  (omega4)      plugin model 1
    kRootExact = 1, // Toms917DiodePairT                           plugin model 0
    kRootApproxGood = 2, // wdft::DiodePairT<..., Good> (eq. 18)
};

enum Ordering
{
    kOrderPlugin = 0, // probe between root.incident and tree.incident  (DiodeClipperWDF.cpp:26-28)
    kOrderPython = 1, // probe after tree.incident                      (clipper_pot.py:121-123)
};

// Tree of plugin/src/dsp/diode_clipper/DiodeClipperWDF.h:18-25 with the constants lifted to arguments,
// loop of DiodeClipperWDF.cpp:22-29 (processDiodeClipper). One fresh circuit (zero state) per sequence.
template <typename T, template <typename, typename, wdft::DiodeQuality> class Pair, wdft::DiodeQuality Q>
void clipperRows (const T* x, T* y, int64_t rowBegin, int64_t rowEnd, int64_t nT, T fs, T R, T C_, T Is, T Vt, T nD, int ordering, T* zOut)
{
    for (int64_t r = rowBegin; r < rowEnd; ++r)
    {
        wdft::ResistiveVoltageSourceT<T> Vs { R };
        wdft::CapacitorT<T> C { C_ };
        wdft::WDFParallelT<T, decltype (Vs), decltype (C)> P1 { Vs, C };
        Pair<T, decltype (P1), Q> dp { P1, Is, Vt, nD };

        C.prepare (fs); // DiodeClipperWDF.cpp:5
        P1.connectToParent (&dp); // DiodeClipperWDF.cpp:38-39
        dp.calcImpedance();

        const T* xr = x + r * nT;
        T* yr = y + r * nT;
        if (ordering == kOrderPlugin)
        {
            for (int64_t n = 0; n < nT; ++n)
            {
                Vs.setVoltage (xr[n]);
                dp.incident (P1.reflected());
                yr[n] = wdft::voltage<T> (C);
                P1.incident (dp.reflected());
            }
        }
        else
        {
            for (int64_t n = 0; n < nT; ++n)
            {
                Vs.setVoltage (xr[n]);
                dp.incident (P1.reflected());
                P1.incident (dp.reflected());
                yr[n] = wdft::voltage<T> (C);
            }
        }
        if (zOut != nullptr)
            zOut[r] = C.reflected(); // final capacitor state
    }
}

template <typename T>
void clipperDispatch (int root, const T* x, T* y, int64_t b0, int64_t b1, int64_t nT, T fs, T R, T C, T Is, T Vt, T nD, int ordering, T* zOut)
{
    if (root == kRootApprox)
        clipperRows<T, wdft::DiodePairT, wdft::DiodeQuality::Best> (x, y, b0, b1, nT, fs, R, C, Is, Vt, nD, ordering, zOut);
    else if (root == kRootApproxGood)
        clipperRows<T, wdft::DiodePairT, wdft::DiodeQuality::Good> (x, y, b0, b1, nT, fs, R, C, Is, Vt, nD, ordering, zOut);
    else
        clipperRows<T, Toms917DiodePairT, wdft::DiodeQuality::Best> (x, y, b0, b1, nT, fs, R, C, Is, Vt, nD, ordering, zOut);
}

template <typename T>
void clipperThreaded (int root, const T* x, T* y, int64_t nB, int64_t nT, T fs, T R, T C, T Is, T Vt, T nD, int ordering, int nThreads)
{
    nThreads = std::max (1, nThreads);
    if (nThreads == 1)
    {
        clipperDispatch<T> (root, x, y, 0, nB, nT, fs, R, C, Is, Vt, nD, ordering, nullptr);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < nThreads; ++t)
    {
        int64_t b0 = nB * t / nThreads, b1 = nB * (t + 1) / nThreads; // one contiguous range per thread
        pool.emplace_back ([=] { clipperDispatch<T> (root, x, y, b0, b1, nT, fs, R, C, Is, Vt, nD, ordering, nullptr); });
    }
    for (auto& th : pool)
        th.join();
}

// A stub "next" port so the root elements can be evaluated stand-alone at a given port impedance.
template <typename T>
struct FixedPort : public wdft::BaseWDF
{
    explicit FixedPort (T R)
    {
        wdf.R = R;
        wdf.G = (T) 1 / R;
    }
    void calcImpedance() override {}
    wdft::WDFMembers<T> wdf;
};

template <typename T, typename Pair>
void pairLaw (const T* a, T* b, int64_t n, T Rp, T Is, T Vt, T nD)
{
    FixedPort<T> port { Rp };
    Pair dp { port, Is, Vt, nD };
    for (int64_t i = 0; i < n; ++i)
    {
        dp.incident (a[i]);
        b[i] = dp.reflected();
    }
}
} // namespace

extern "C" {

// --- omega.h / toms917 scalar functions -------------------------------------------------------
// kind: 0 omega1, 1 omega2, 2 omega3, 3 omega4, 4 log_approx, 5 exp_approx, 6 log2_approx, 7 pow2_approx
void ref_omega_f32 (int kind, const float* x, float* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
    {
        float v = x[i];
        switch (kind)
        {
            case 0: out[i] = Omega::omega1<float> (v); break;
            case 1: out[i] = Omega::omega2<float> (v); break;
            case 2: out[i] = Omega::omega3<float> (v); break;
            case 3: out[i] = Omega::omega4<float> (v); break;
            case 4: out[i] = Omega::log_approx<float> (v); break;
            case 5: out[i] = Omega::exp_approx<float> (v); break;
            case 6: out[i] = Omega::log2_approx<float> (v); break;
            default: out[i] = Omega::pow2_approx<float> (v); break;
        }
    }
}

void ref_omega_f64 (int kind, const double* x, double* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
    {
        double v = x[i];
        switch (kind)
        {
            case 0: out[i] = Omega::omega1<double> (v); break;
            case 1: out[i] = Omega::omega2<double> (v); break;
            case 2: out[i] = Omega::omega3<double> (v); break;
            case 3: out[i] = Omega::omega4<double> (v); break;
            case 4: out[i] = Omega::log_approx<double> (v); break;
            case 5: out[i] = Omega::exp_approx<double> (v); break;
            case 6: out[i] = Omega::log2_approx<double> (v); break;
            default: out[i] = Omega::pow2_approx<double> (v); break;
        }
    }
}

// modules/toms917/toms917.cpp:21 on the real axis (what Toms917DiodePair.h:64-67 calls)
void ref_toms917_real (const double* x, double* out, int64_t n)
{
    for (int64_t i = 0; i < n; ++i)
        out[i] = std::real (wrightomega (std::complex<double> (x[i])));
}

int ref_signum_f32 (float v) { return chowdsp::signum (v); }

// --- diode-pair root laws at a fixed port impedance -------------------------------------------
void ref_diode_pair_f32 (int root, const float* a, float* b, int64_t n, float Rp, float Is, float Vt, float nD)
{
    if (root == kRootApprox)
        pairLaw<float, wdft::DiodePairT<float, FixedPort<float>, wdft::DiodeQuality::Best>> (a, b, n, Rp, Is, Vt, nD);
    else if (root == kRootApproxGood)
        pairLaw<float, wdft::DiodePairT<float, FixedPort<float>, wdft::DiodeQuality::Good>> (a, b, n, Rp, Is, Vt, nD);
    else
        pairLaw<float, Toms917DiodePairT<float, FixedPort<float>>> (a, b, n, Rp, Is, Vt, nD);
}

void ref_diode_pair_f64 (int root, const double* a, double* b, int64_t n, double Rp, double Is, double Vt, double nD)
{
    if (root == kRootApprox)
        pairLaw<double, wdft::DiodePairT<double, FixedPort<double>, wdft::DiodeQuality::Best>> (a, b, n, Rp, Is, Vt, nD);
    else if (root == kRootApproxGood)
        pairLaw<double, wdft::DiodePairT<double, FixedPort<double>, wdft::DiodeQuality::Good>> (a, b, n, Rp, Is, Vt, nD);
    else
        pairLaw<double, Toms917DiodePairT<double, FixedPort<double>>> (a, b, n, Rp, Is, Vt, nD);
}

// --- the diode clipper (DiodeClipperWDF.h:18-25 + DiodeClipperWDF.cpp:22-29) -------------------
void ref_clipper_f32 (int root, int ordering, const float* x, float* y, int64_t nB, int64_t nT, float fs, float R, float C, float Is, float Vt, float nD, int nThreads)
{
    clipperThreaded<float> (root, x, y, nB, nT, fs, R, C, Is, Vt, nD, ordering, nThreads);
}

void ref_clipper_f64 (int root, int ordering, const double* x, double* y, int64_t nB, int64_t nT, double fs, double R, double C, double Is, double Vt, double nD, int nThreads)
{
    clipperThreaded<double> (root, x, y, nB, nT, fs, R, C, Is, Vt, nD, ordering, nThreads);
}

// Port impedance the root sees (P1.wdf.R) for given R, C, fs — float arithmetic of wdf_t.h:166-170,465-470
float ref_clipper_port_impedance_f32 (float fs, float R, float C)
{
    wdft::ResistiveVoltageSourceT<float> Vs { R };
    wdft::CapacitorT<float> Cc { C };
    wdft::WDFParallelT<float, decltype (Vs), decltype (Cc)> P1 { Vs, Cc };
    Cc.prepare (fs);
    return P1.wdf.R;
}

// --- RC low-pass of wdf_py/simple_circuits/lpf.py:23-28,38-46 built from the C++ templates ------
// IdealVoltageSource root over Inverter(Series(R1, C1)); probe voltage(C1) after tree.incident.
// probe: 0 -> voltage(C1), 1 -> voltage(R1)
void ref_rc_lowpass_f64 (const double* x, double* y, int64_t nT, double fs, double R, double C, int probe)
{
    wdft::ResistorT<double> R1 { R };
    wdft::CapacitorT<double> C1 { C, fs };
    auto S1 = wdft::makeSeries<double> (R1, C1);
    auto I1 = wdft::makeInverter<double> (S1);
    wdft::IdealVoltageSourceT<double, decltype (I1)> Vs { I1 };
    for (int64_t n = 0; n < nT; ++n)
    {
        Vs.setVoltage (x[n]);
        Vs.incident (I1.reflected());
        I1.incident (Vs.reflected());
        y[n] = probe == 0 ? wdft::voltage<double> (C1) : wdft::voltage<double> (R1);
    }
}

void ref_rc_lowpass_f32 (const float* x, float* y, int64_t nT, float fs, float R, float C, int probe)
{
    wdft::ResistorT<float> R1 { R };
    wdft::CapacitorT<float> C1 { C, fs };
    auto S1 = wdft::makeSeries<float> (R1, C1);
    auto I1 = wdft::makeInverter<float> (S1);
    wdft::IdealVoltageSourceT<float, decltype (I1)> Vs { I1 };
    for (int64_t n = 0; n < nT; ++n)
    {
        Vs.setVoltage (x[n]);
        Vs.incident (I1.reflected());
        I1.incident (Vs.reflected());
        y[n] = probe == 0 ? wdft::voltage<float> (C1) : wdft::voltage<float> (R1);
    }
}

// Voltage divider of wdf_py/simple_circuits/voltage_divider.py:19-24,33-41 (probe voltage(R1))
void ref_voltage_divider_f64 (const double* x, double* y, int64_t nT, double Ra, double Rb)
{
    wdft::ResistorT<double> R1 { Ra };
    wdft::ResistorT<double> R2 { Rb };
    auto S1 = wdft::makeSeries<double> (R1, R2);
    auto I1 = wdft::makeInverter<double> (S1);
    wdft::IdealVoltageSourceT<double, decltype (I1)> Vs { I1 };
    for (int64_t n = 0; n < nT; ++n)
    {
        Vs.setVoltage (x[n]);
        Vs.incident (I1.reflected());
        I1.incident (Vs.reflected());
        y[n] = wdft::voltage<double> (R1);
    }
}

// --- the reference's own known-answer fixtures -------------------------------------------------
// modules/chowdsp_utils/tests/wdf_standalone_test/wdf_standalone_test.cpp:16-36 (expects 4.77 +- 0.1)
float ref_standalone_test()
{
    using namespace chowdsp::WDF;
    Resistor<float> R1 { 1.0e3f };
    ResistiveVoltageSource<float> Vin { 1.0e3f };
    PolarityInverter<float> I1 { &Vin };
    WDFSeries<float> S1 { &R1, &I1 };
    DiodePair<float> D1 { &S1, 1.0e-10f };
    Vin.setVoltage (10.0f);
    D1.incident (S1.reflected());
    S1.incident (D1.reflected());
    return R1.voltage();
}

// wdf_tests/StaticWDFTest.cpp:216-271, "static" half (fs = 44.1 kHz there). quality: 0 Good, 1 Best.
void ref_static_wdf_test (int quality, double fs, double* out5)
{
    using T = double;
    double data[5] = { 1.0, 0.5, 0.0, -0.5, -1.0 };
    wdft::ResistiveVoltageSourceT<T> Vs {};
    wdft::ResistorT<T> R1 { 4700.0 };
    wdft::CapacitorT<T> C1 { 47.0e-9, fs };
    auto S1 = wdft::makeSeries<T> (Vs, R1);
    auto P1 = wdft::makeParallel<T> (S1, C1);
    auto I1 = wdft::makeInverter<T> (P1);
    auto run = [&] (auto& dp) {
        for (double& v : data)
        {
            Vs.setVoltage (v);
            dp.incident (P1.reflected());
            v = wdft::voltage<T> (C1);
            P1.incident (dp.reflected());
        }
    };
    if (quality == 0)
    {
        wdft::DiodePairT<T, decltype (I1), wdft::DiodeQuality::Good> dp { I1, 2.52e-9 };
        run (dp);
    }
    else
    {
        wdft::DiodePairT<T, decltype (I1), wdft::DiodeQuality::Best> dp { I1, 2.52e-9 };
        run (dp);
    }
    std::memcpy (out5, data, sizeof (data));
}

int ref_hardware_threads() { return (int) std::thread::hardware_concurrency(); }

} // extern "C"
