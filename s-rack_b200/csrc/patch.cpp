// Module catalog facts and patch bookkeeping (no CUDA).
#include "patch.hpp"

#include <cstdio>
#include <random>

namespace srk {

namespace {
// clang-format off
const KindInfo kKinds[SRK_KIND_COUNT] = {
  // Output: output.rs -- `channels` inputs, no outputs, labels Ok(None)
  {"Output",      -1, 0, {nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, 0, {0, 0, 0, 0}, {false, false, false, false}},
  // Oscillator: oscillator.rs:99-106 (outputs), :172-178 (inputs), :32,38 (defaults)
  {"Oscillator",   2, 3, {"CV", "Sync", nullptr, nullptr}, {"Sine", "Square", "Sawtooth"}, 2, {0.0f, 1.0f, 0, 0}, {false, true, false, false}},
  // Noise: oscillator.rs:344-375
  {"Noise",        0, 1, {nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, 0, {0, 0, 0, 0}, {false, false, false, false}},
  // ADSR: adsr.rs:109-132, defaults :39-42
  {"ADSR",         1, 1, {"Gate", nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, 4, {0.0f, 0.5f, 0.25f, 0.5f}, {false, false, false, false}},
  // VCA: vca.rs (labels "Audio"/"CV"), negative=false (:24)
  {"VCA",          2, 1, {"Audio", "CV", nullptr, nullptr}, {nullptr, nullptr, nullptr}, 1, {0.0f, 0, 0, 0}, {true, false, false, false}},
  // Moog filter: filter.rs:154-180, defaults :36-38
  {"Moog Filter",  2, 3, {"Audio", "CV", nullptr, nullptr}, {nullptr, nullptr, nullptr}, 3, {0.2f, 0.5f, 0.5f, 0}, {false, false, false, false}},
  // Mono mixer: mixer.rs:19-20 (4 inputs, gain 1.0), labels Ok(None)
  {"Mono Mixer",   4, 1, {nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, 4, {1.0f, 1.0f, 1.0f, 1.0f}, {false, false, false, false}},
  // Math: math.rs:113-137, constant 0.0 (:32)
  {"Add",          2, 1, {"In1", "In2", nullptr, nullptr}, {nullptr, nullptr, nullptr}, 1, {0.0f, 0, 0, 0}, {false, false, false, false}},
  {"Subtract",     2, 1, {"In1", "In2", nullptr, nullptr}, {nullptr, nullptr, nullptr}, 1, {0.0f, 0, 0, 0}, {false, false, false, false}},
  {"Multiply",     2, 1, {"In1", "In2", nullptr, nullptr}, {nullptr, nullptr, nullptr}, 1, {0.0f, 0, 0, 0}, {false, false, false, false}},
  // Non-Linear: math.rs:266-290, constant 1.0 (:194)
  {"Non-Linear",   2, 1, {"In1", "In2", nullptr, nullptr}, {nullptr, nullptr, nullptr}, 1, {1.0f, 0, 0, 0}, {false, false, false, false}},
  // Grid sequencer: sequencer.rs:262-307 (labels), steps_per_octave 12 (:44)
  {"Grid Sequencer", 2, 3, {"Step", "Sync", nullptr, nullptr}, {"CV", "Gate", "Sync"}, 1, {12.0f, 0, 0, 0}, {true, false, false, false}},
  // Pattern sequencer: sequencer.rs:551-596: 8 gate rows labelled "0".."7", then "Sync"
  {"Pattern Sequencer", 2, 9, {"Step", "Sync", nullptr, nullptr}, {"0", "1", "2", "3", "4", "5", "6", "7", "Sync"}, 0, {0, 0, 0, 0}, {false, false, false, false}},
  // Sample: sample.rs:163-190 (inputs "Gate", "CV"; one unlabelled output), no numeric parameters
  {"Sample",       2, 1, {"Gate", "CV", nullptr, nullptr}, {nullptr, nullptr, nullptr}, 0, {0, 0, 0, 0}, {false, false, false, false}},
};
// clang-format on
}  // namespace

const KindInfo& kind_info(int kind) { return kKinds[kind]; }

std::string make_uuid_v4() {
  static thread_local std::mt19937_64 rng{std::random_device{}()};
  uint64_t a = rng(), b = rng();
  a = (a & 0xFFFFFFFFFFFF0FFFull) | 0x0000000000004000ull;  // version 4
  b = (b & 0x3FFFFFFFFFFFFFFFull) | 0x8000000000000000ull;  // variant 1
  char buf[40];
  std::snprintf(buf, sizeof buf, "%08x-%04x-%04x-%04x-%012llx", (unsigned)(a >> 32), (unsigned)((a >> 16) & 0xFFFF),
                (unsigned)(a & 0xFFFF), (unsigned)(b >> 48), (unsigned long long)(b & 0xFFFFFFFFFFFFull));
  return buf;
}

}  // namespace srk

int srk_module::n_outputs() const { return srk::kind_info(kind).n_outputs; }

srk_module* srk_patch::find_output() const {
  for (srk_module* m : modules)
    if (m->kind == SRK_KIND_OUTPUT) return m;
  return nullptr;
}
