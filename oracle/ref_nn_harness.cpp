// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product path.
//
// oracle/_ref/libdwdf_ref_nn.so: the reference's neural diode-pair root, i.e. the UNMODIFIED RTNeural
// (modules/RTNeural, STL backend) running the reference's own JSON weight files inside the UNMODIFIED
// chowdsp_wdf clipper tree, behind a flat C ABI for ctypes. plugin/src/dsp/diode_clipper/
// DiodePairNeuralModel.h itself needs JUCE (String, BinaryData, MemoryInputStream); its three
// one-line methods are reproduced below with the line they come from. The model types are the ones of
// DiodePairNeuralModel.h:5-41 (2xH: 2->H->H->H->1, 4xH: 2->H x5->1, tanh).
#include <cstdint>
#include <cmath>
#include <fstream>
#include <string>

#include <pch.h> // oracle/shim/pch.h (-> <wdf_t.h>)
#include <RTNeural.h>

namespace
{
template <int N, int H>
struct ModelType;
template <int H>
struct ModelType<2, H>
{
    using type = RTNeural::ModelT<float, 2, 1, RTNeural::DenseT<float, 2, H>, RTNeural::TanhActivationT<float, H>, RTNeural::DenseT<float, H, H>, RTNeural::TanhActivationT<float, H>,
                                  RTNeural::DenseT<float, H, H>, RTNeural::TanhActivationT<float, H>, RTNeural::DenseT<float, H, 1>>;
};
template <int H>
struct ModelType<4, H>
{
    using type = RTNeural::ModelT<float, 2, 1, RTNeural::DenseT<float, 2, H>, RTNeural::TanhActivationT<float, H>, RTNeural::DenseT<float, H, H>, RTNeural::TanhActivationT<float, H>,
                                  RTNeural::DenseT<float, H, H>, RTNeural::TanhActivationT<float, H>, RTNeural::DenseT<float, H, H>, RTNeural::TanhActivationT<float, H>, RTNeural::DenseT<float, H, H>,
                                  RTNeural::TanhActivationT<float, H>, RTNeural::DenseT<float, H, 1>>;
};

template <typename Next, int N, int H>
struct NeuralRoot : public wdft::RootWDF
{
    NeuralRoot (Next& n, const nlohmann::json& j) : next (n) { model.parseJson (j, false); } // DiodePairNeuralModel.h:62-63
    void calcImpedance() override { logR = std::log (next.wdf.R); } // :66
    inline void incident (float x) { wdf.a = x; } // :68
    inline float reflected() // :70-75
    {
        const float inData alignas (16)[] = { wdf.a, logR };
        wdf.b = -model.template forward (inData);
        return wdf.b;
    }
    wdft::WDFMembers<float> wdf;
    const Next& next;
    float logR = 1.0f;
    typename ModelType<N, H>::type model;
};

// tree of DiodeClipperWDF.h:18-25, loop of DiodeClipperWDF.cpp:22-29; ordering 0 = plugin probe, 1 = python probe
template <int N, int H>
int run (const nlohmann::json& j, const float* x, float* y, int64_t B, int64_t T, float fs, float R, float Cv, int ordering)
{
    for (int64_t r = 0; r < B; ++r)
    {
        wdft::ResistiveVoltageSourceT<float> Vs { R };
        wdft::CapacitorT<float> C { Cv };
        wdft::WDFParallelT<float, decltype (Vs), decltype (C)> P1 { Vs, C };
        NeuralRoot<decltype (P1), N, H> dp { P1, j };
        C.prepare (fs);
        P1.connectToParent (&dp);
        dp.calcImpedance();
        for (int64_t n = 0; n < T; ++n)
        {
            Vs.setVoltage (x[r * T + n]);
            dp.incident (P1.reflected());
            if (ordering == 0)
                y[r * T + n] = wdft::voltage<float> (C);
            P1.incident (dp.reflected());
            if (ordering != 0)
                y[r * T + n] = wdft::voltage<float> (C);
        }
    }
    return 0;
}
} // namespace

extern "C" int ref_nn_clipper (const char* json_path, int n_layers, int hidden, const float* x, float* y, int64_t B, int64_t T, float fs, float R, float C, int ordering)
{
    std::ifstream f (json_path);
    if (! f)
        return 1;
    const auto j = nlohmann::json::parse (f);
#define CASE(N, H) \
    if (n_layers == N && hidden == H) \
        return run<N, H> (j, x, y, B, T, fs, R, C, ordering);
    CASE (2, 4) CASE (2, 8) CASE (2, 16) CASE (4, 4) CASE (4, 8)
#undef CASE
    return 2;
}

// the bare network: out[i] = model.forward({a[i], logR[i]}) (no sign flip)
extern "C" int ref_nn_eval (const char* json_path, int n_layers, int hidden, const float* a, const float* logR, float* out, int64_t n)
{
    std::ifstream f (json_path);
    if (! f)
        return 1;
    const auto j = nlohmann::json::parse (f);
#define CASE(N, H) \
    if (n_layers == N && hidden == H) \
    { \
        typename ModelType<N, H>::type m; \
        m.parseJson (j, false); \
        for (int64_t i = 0; i < n; ++i) \
        { \
            const float in alignas (16)[] = { a[i], logR[i] }; \
            out[i] = m.template forward (in); \
        } \
        return 0; \
    }
    CASE (2, 4) CASE (2, 8) CASE (2, 16) CASE (4, 4) CASE (4, 8)
#undef CASE
    return 2;
}
