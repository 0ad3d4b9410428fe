// dwdf_clipper.hpp — C++ twin of the plugin's DiodeClipperWDF over the C ABI of dwdf.h.
//
// Replaces plugin/src/dsp/diode_clipper/DiodeClipperWDF.{h,cpp}:
//   prepare (sampleRate)                 DiodeClipperWDF.cpp:3-8    (C.prepare, forget the active model)
//   setParameters (cutoffHz, modelIndex) DiodeClipperWDF.cpp:10-16  (Vs.setResistanceValue (1 / (2 pi fc C)) -> impedance change
//                                                                     propagates to the root, wdf_t.h:705-713; model choice)
//   process (buffer)                     DiodeClipperWDF.cpp:18-166 (model switch -> connectToParent + calcImpedance, then the
//                                                                     per-sample loop Vs.setVoltage / dp.incident / voltage(C) / P1.incident)
// Same circuit and constants (DiodeClipperWDF.h:18-25: Vs 47 k, C 2.2 nF, 1N4148 Is 4.352e-9, Vt 25.85e-3, nDiodes 1.906), same
// probe position (between root.incident and tree.incident), the capacitor state carried from block to block and across model
// switches (the reference's models share one CapacitorT). What differs is the shape of the work: `channels` independent mono
// streams are processed at once, each block is ONE kernel launch on the caller's stream, and buffers live on the device.
//
// Header-only, C++17, needs only dwdf.h, the CUDA runtime and libdwdf.so. Every method returns a dwdf_status (0 = ok; the text is
// in dwdf_last_error ()): like the reference's methods nothing throws.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <vector>

#include "dwdf.h"

namespace dwdf
{

class DiodeClipperB200
{
public:
    // model indices of DiodeClipperWDF::process (DiodeClipperWDF.cpp:32-166)
    enum Model : int
    {
        kToms917 = 0, // Toms917DiodePairT: exact Wright omega
        kOmega4 = 1, // wdft::DiodePairT: D'Angelo's omega4 approximation
        kFirstNeural = 2 // 2 .. 11: DiodePairNeuralModel<..>; weights come from loadNeuralModel ()
    };
    static constexpr int kNumModels = 12;
    static constexpr float capVal = 2.2e-9f; // DiodeClipperWDF.h:17

    DiodeClipperB200 () = default;
    DiodeClipperB200 (const DiodeClipperB200&) = delete;
    DiodeClipperB200& operator= (const DiodeClipperB200&) = delete;
    ~DiodeClipperB200 () { release (); }

    // DiodeClipperWDF::prepare, for `channels` independent mono streams: builds the programs and zeroes the capacitor states
    int prepare (double sampleRate, int64_t channels = 1)
    {
        release ();
        fs = sampleRate;
        n_channels = channels;
        if (channels < 1)
            return DWDF_ERR_INVALID;
        const float init[4] = { 47000.0f, capVal, 4.352e-9f, 1.906f }; // { R, C, Is, nDiodes }: DiodeClipperWDF.h:19-24
        if (cudaMalloc ((void**) &d_params, sizeof (init)) != cudaSuccess || cudaMalloc ((void**) &d_state, sizeof (float) * (size_t) channels) != cudaSuccess)
            return DWDF_ERR_CUDA;
        if (cudaMemcpy (d_params, init, sizeof (init), cudaMemcpyHostToDevice) != cudaSuccess || cudaMemset (d_state, 0, sizeof (float) * (size_t) channels) != cudaSuccess)
            return DWDF_ERR_CUDA;
        if (int rc = make_program (DWDF_ROOT_DIODE_PAIR, DWDF_MODE_EXACT, nullptr, &programs[kToms917]))
            return rc;
        if (int rc = make_program (DWDF_ROOT_DIODE_PAIR, DWDF_MODE_APPROX, nullptr, &programs[kOmega4]))
            return rc;
        for (int m = kFirstNeural; m < kNumModels; ++m) // re-create the neural programs at the new sample rate
            if (! host_weights[m].empty ())
                if (int rc = loadNeuralModel (m, mlp[m].n_hidden, mlp[m].hidden, host_weights[m].data (), host_weights[m].size ()))
                    return rc;
        prevModelChoice = -1;
        return DWDF_OK;
    }

    // Plugin models 2-11 (DiodePairNeuralModel.h:62-75): the flattened weight vector of the model's JSON file (layer by layer:
    // kernel (in x out) row-major, then bias — model_io.flatten_weights), e.g. n_hidden = 2, hidden = 16 for "2x16"
    int loadNeuralModel (int modelIndex, int n_hidden, int hidden, const float* weights, size_t n_weights)
    {
        if (modelIndex < kFirstNeural || modelIndex >= kNumModels || weights == nullptr)
            return DWDF_ERR_INVALID;
        const dwdf_mlp_desc desc { n_hidden, hidden };
        if (n_weights != dwdf_mlp_weight_count (&desc))
            return DWDF_ERR_INVALID;
        if (host_weights[modelIndex].data () != weights)
            host_weights[modelIndex].assign (weights, weights + n_weights);
        mlp[modelIndex] = desc;
        if (d_params == nullptr)
            return DWDF_OK; // prepare () will build it
        if (programs[modelIndex] != nullptr)
            dwdf_program_destroy (programs[modelIndex]);
        programs[modelIndex] = nullptr;
        if (d_weights[modelIndex] != nullptr)
            cudaFree (d_weights[modelIndex]);
        d_weights[modelIndex] = nullptr;
        if (cudaMalloc ((void**) &d_weights[modelIndex], sizeof (float) * n_weights) != cudaSuccess
            || cudaMemcpy (d_weights[modelIndex], host_weights[modelIndex].data (), sizeof (float) * n_weights, cudaMemcpyHostToDevice) != cudaSuccess)
            return DWDF_ERR_CUDA;
        return make_program (DWDF_ROOT_NEURAL, 0, &desc, &programs[modelIndex]);
    }

    // DiodeClipperWDF::setParameters: the source resistance from the cutoff, and which root closes the tree
    int setParameters (float cutoffFreqHz, int modelIndex, cudaStream_t stream = nullptr)
    {
        if (d_params == nullptr || modelIndex < 0 || modelIndex >= kNumModels || programs[modelIndex] == nullptr || ! (cutoffFreqHz > 0.0f))
            return DWDF_ERR_INVALID;
        const float twoPi = 6.283185307179586476925286766559f; // MathConstants<float>::twoPi
        resVal = 1.0f / (twoPi * cutoffFreqHz * capVal);
        if (cudaMemcpyAsync (d_params, &resVal, sizeof (float), cudaMemcpyHostToDevice, stream) != cudaSuccess) // slot 0 = R; ordered before the next block on `stream`
            return DWDF_ERR_CUDA;
        modelChoice = modelIndex;
        return DWDF_OK;
    }

    // DiodeClipperWDF::process for `channels` streams at once: d_x, d_y are (channels, numSamples) device buffers, batch-major
    // (d_y == d_x is NOT allowed: the kernels read ahead of what they write). The port impedances follow the parameters at every
    // launch, so the reference's "model switch -> calcImpedance" needs no extra step here; prevModelChoice is kept for parity.
    int process (const float* d_x, float* d_y, int64_t numSamples, cudaStream_t stream = nullptr)
    {
        if (d_params == nullptr || programs[modelChoice] == nullptr)
            return DWDF_ERR_INVALID;
        prevModelChoice = modelChoice;
        if (modelChoice >= kFirstNeural)
            return dwdf_forward_neural (programs[modelChoice], d_params, d_weights[modelChoice], d_x, nullptr, d_y, d_state, nullptr, n_channels, numSamples, stream);
        return dwdf_process_block (programs[modelChoice], d_params, d_x, nullptr, d_y, d_state, n_channels, numSamples, stream);
    }

    // AudioBuffer-style convenience: host buffers in and out, synchronous (one mono block: channels x numSamples floats)
    int processHost (const float* x, float* y, int64_t numSamples)
    {
        const size_t bytes = sizeof (float) * (size_t) n_channels * (size_t) numSamples;
        if (bytes > staging_bytes)
        {
            if (d_in != nullptr)
                cudaFree (d_in);
            if (d_out != nullptr)
                cudaFree (d_out);
            d_in = d_out = nullptr;
            staging_bytes = 0;
            if (cudaMalloc ((void**) &d_in, bytes) != cudaSuccess || cudaMalloc ((void**) &d_out, bytes) != cudaSuccess)
                return DWDF_ERR_CUDA;
            staging_bytes = bytes;
        }
        if (cudaMemcpy (d_in, x, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
            return DWDF_ERR_CUDA;
        if (int rc = process (d_in, d_out, numSamples, nullptr))
            return rc;
        return cudaMemcpy (y, d_out, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? DWDF_OK : DWDF_ERR_CUDA;
    }

    void reset ()
    {
        if (d_state != nullptr)
            cudaMemset (d_state, 0, sizeof (float) * (size_t) n_channels);
    }

    float sourceResistance () const { return resVal; }
    int activeModel () const { return modelChoice; }
    int64_t channels () const { return n_channels; }

private:
    int make_program (int root_kind, int root_mode, const dwdf_mlp_desc* net, dwdf_program** out)
    {
        const dwdf_node nodes[3] = { { DWDF_RESISTIVE_VS, -1, -1, 0 }, { DWDF_CAPACITOR, -1, -1, 1 }, { DWDF_PARALLEL, 0, 1, -1 } }; // P1 { Vs, C }, DiodeClipperWDF.h:19-21
        dwdf_circuit_desc d {};
        d.root_kind = root_kind;
        d.root_mode = root_mode;
        d.ordering = DWDF_ORDER_PLUGIN; // x[n] = voltage (C) between dp.incident and P1.incident, DiodeClipperWDF.cpp:26-28
        d.probe = 1;
        d.source = 0;
        d.r_node = -1;
        d.param_Is = 2;
        d.param_nabla = 3;
        d.n_params = 4;
        d.fs = (float) fs;
        d.Vt = 25.85e-3f;
        d.n_up = d.n_down = 1.0f;
        return net != nullptr ? dwdf_program_create_neural (nodes, 3, &d, net, out) : dwdf_program_create (nodes, 3, &d, out);
    }

    void release ()
    {
        for (int m = 0; m < kNumModels; ++m)
        {
            if (programs[m] != nullptr)
                dwdf_program_destroy (programs[m]);
            programs[m] = nullptr;
            if (d_weights[m] != nullptr)
                cudaFree (d_weights[m]);
            d_weights[m] = nullptr;
        }
        for (float** p : { &d_params, &d_state, &d_in, &d_out })
        {
            if (*p != nullptr)
                cudaFree (*p);
            *p = nullptr;
        }
        staging_bytes = 0;
    }

    double fs = 48000.0;
    int64_t n_channels = 0;
    dwdf_program* programs[kNumModels] = {};
    float* d_weights[kNumModels] = {};
    std::vector<float> host_weights[kNumModels];
    dwdf_mlp_desc mlp[kNumModels] = {};
    float *d_params = nullptr, *d_state = nullptr, *d_in = nullptr, *d_out = nullptr;
    size_t staging_bytes = 0;
    float resVal = 47000.0f;
    int modelChoice = 0;
    int prevModelChoice = -1;
};

} // namespace dwdf
