// Drives include/dwdf_clipper.hpp (the C++ twin of the plugin's DiodeClipperWDF) the way the plugin drives the original:
// prepare -> setParameters -> process block by block, with a cutoff change and model switches in mid-stream.
//   clipper_twin_main <in.bin> <out.bin> <channels> [weights_2x16.bin]
// in.bin: channels x 9192 float32 (batch-major). Segments: A = 3 blocks of 2048, model 1 (omega4), fc 1539.3 Hz;
// B = 2 blocks of 1000, model 0 (TOMS-917), fc 800 Hz; C = 1 block of 1048, model 4 (neural 2x16) if weights are given, else model 1.
// Exit codes: 0 ok, 2 usage / IO, 3 no CUDA device (the library has no CPU fallback), 4 library error.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dwdf_clipper.hpp"

static int die (int code, const char* what)
{
    std::fprintf (stderr, "clipper_twin_main: %s: %s\n", what, dwdf_last_error ());
    return code;
}

int main (int argc, char** argv)
{
    if (argc < 4)
        return die (2, "usage: clipper_twin_main in.bin out.bin channels [weights_2x16.bin]");
    const int64_t B = std::atoll (argv[3]);
    const int64_t total = 3 * 2048 + 2 * 1000 + 1048;
    int n_dev = 0;
    if (cudaGetDeviceCount (&n_dev) != cudaSuccess || n_dev == 0)
    {
        dwdf::DiodeClipperB200 probe;
        const int rc = probe.prepare (48000.0, 1); // must fail loudly, not fall back
        std::fprintf (stderr, "clipper_twin_main: no CUDA device (prepare returned %d)\n", rc);
        return rc != DWDF_OK ? 3 : 4;
    }
    std::vector<float> x ((size_t) (B * total)), y ((size_t) (B * total));
    FILE* f = std::fopen (argv[1], "rb");
    if (f == nullptr || std::fread (x.data (), sizeof (float), x.size (), f) != x.size ())
        return die (2, "cannot read the input file");
    std::fclose (f);
    std::vector<float> w;
    if (argc > 4)
    {
        const dwdf_mlp_desc d { 2, 16 };
        w.resize (dwdf_mlp_weight_count (&d));
        f = std::fopen (argv[4], "rb");
        if (f == nullptr || std::fread (w.data (), sizeof (float), w.size (), f) != w.size ())
            return die (2, "cannot read the weight file");
        std::fclose (f);
    }

    dwdf::DiodeClipperB200 clipper;
    if (! w.empty () && clipper.loadNeuralModel (4, 2, 16, w.data (), w.size ()) != DWDF_OK) // before prepare: built there
        return die (4, "loadNeuralModel");
    if (clipper.prepare (48000.0, B) != DWDF_OK)
        return die (4, "prepare");

    // one block = the (channels, n) slice [pos, pos + n) of every stream; the twin takes batch-major blocks
    std::vector<float> bx, by;
    int64_t pos = 0;
    auto run = [&] (int64_t n) -> int {
        bx.resize ((size_t) (B * n));
        by.resize ((size_t) (B * n));
        for (int64_t b = 0; b < B; ++b)
            for (int64_t i = 0; i < n; ++i)
                bx[(size_t) (b * n + i)] = x[(size_t) (b * total + pos + i)];
        if (int rc = clipper.processHost (bx.data (), by.data (), n))
            return rc;
        for (int64_t b = 0; b < B; ++b)
            for (int64_t i = 0; i < n; ++i)
                y[(size_t) (b * total + pos + i)] = by[(size_t) (b * n + i)];
        pos += n;
        return DWDF_OK;
    };
    if (clipper.setParameters (1539.3f, dwdf::DiodeClipperB200::kOmega4) != DWDF_OK)
        return die (4, "setParameters A");
    for (int k = 0; k < 3; ++k)
        if (run (2048) != DWDF_OK)
            return die (4, "process A");
    if (clipper.setParameters (800.0f, dwdf::DiodeClipperB200::kToms917) != DWDF_OK)
        return die (4, "setParameters B");
    for (int k = 0; k < 2; ++k)
        if (run (1000) != DWDF_OK)
            return die (4, "process B");
    if (clipper.setParameters (800.0f, w.empty () ? 1 : 4) != DWDF_OK)
        return die (4, "setParameters C");
    if (run (1048) != DWDF_OK)
        return die (4, "process C");
    if (clipper.setParameters (800.0f, 7) == DWDF_OK) // a model that was never loaded must be refused
        return die (4, "setParameters accepted an unloaded model");

    f = std::fopen (argv[2], "wb");
    if (f == nullptr || std::fwrite (y.data (), sizeof (float), y.size (), f) != y.size ())
        return die (2, "cannot write the output file");
    std::fclose (f);
    std::printf ("clipper_twin_main: %lld channels x %lld samples, R(800 Hz) = %.3f ohm\n", (long long) B, (long long) total, (double) clipper.sourceResistance ());
    return 0;
}
