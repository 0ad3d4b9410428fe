// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
// Stand-in for the reference's JUCE pre-compiled header (plugin/src/pch.h:8,17) so that
// plugin/src/dsp/diode_clipper/Toms917DiodePair.h can be compiled in place without JUCE.
#pragma once
#include <variant>
#include <cstdint>
#include <wdf_t.h>
namespace wdft = chowdsp::WDFT;
