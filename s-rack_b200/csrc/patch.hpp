// Host-side patch graph: the module list, wiring and parameters behind the C ABI.
// Mirrors the reference's SynthModule surface (src/synth.rs:222-263) and the
// workspace that owns `modules` / `plan` (src/ui.rs:51-60).  No CUDA in here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/srack_b200.h"

namespace srk {

constexpr int kMaxParams = 4;

// Static facts per module kind (port counts, labels, parameter defaults), each
// taken from the reference file cited in include/srack_b200.h.
struct KindInfo {
  const char* name;
  int n_inputs;   // -1: `channels` (Output)
  int n_outputs;
  const char* in_labels[4];   // nullptr = the reference's Ok(None)
  const char* out_labels[9];
  int n_params;
  float param_default[kMaxParams];
  bool param_uniform_only[kMaxParams];
};

const KindInfo& kind_info(int kind);

struct Engine;  // device side (engine.cu)

}  // namespace srk

struct srk_module {
  srk_patch* patch = nullptr;
  int kind = -1;
  std::string id;
  // input i -> (source module, source port); source == nullptr when unconnected
  std::vector<std::pair<srk_module*, uint8_t>> inputs;
  float param[srk::kMaxParams] = {0, 0, 0, 0};
  std::vector<float> param_pv[srk::kMaxParams];  // per-voice override (global voice index) or empty
  // Sequencers: the step table (sequencer.rs:18,341), -1 = None; rows x seq_steps for the pattern
  std::vector<int32_t> sequence;
  size_t seq_steps = 0;
  // Sample: the WaveBox (sample.rs:14-20): decoded channel-0 samples, the file's rate, `new`
  std::vector<float> wave;
  float wave_rate = 0.0f;
  bool wave_new = false;
  // DSP state a loaded .srk file carried (device state words, program.hpp); empty = X::new().  Every voice starts
  // from it and srk_reset() returns to it.
  std::vector<uint32_t> init_state;
  uint16_t osc_sample_rate = 0;  // Oscillator / Sample: follows set_audio_config (oscillator.rs:83-84, sample.rs:120-123)
  float adsr_sample_rate = 0;    // ADSR: fixed at construction (adsr.rs:47,69-71)
  int n_outputs() const;
};

struct srk_patch {
  srk_audio_config cfg{};
  uint64_t seed = 0x5EED5EEDull;
  int device = -1;
  std::vector<std::unique_ptr<srk_module>> owned;
  std::vector<srk_module*> modules;  // `all_modules` order (creation order unless reordered)
  // result of the last srk_plan()
  bool planned = false;
  std::vector<srk_module*> plan;
  std::vector<std::pair<srk_module*, srk_module*>> cuts;  // (reader, writer)
  uint64_t wiring_epoch = 1;  // bumped by every wiring / list change
  uint64_t param_epoch = 1;   // bumped by every parameter change
  uint64_t table_epoch = 1;   // bumped by every sequence-table change (program image, not state)
  uint64_t wave_epoch = 1;    // bumped by every Sample table change (device copy of the waves)
  std::string last_error;
  std::vector<std::pair<std::string, std::pair<float, float>>> positions;  // GUI positions of a loaded .srk, written back on save
  std::vector<unsigned char> saved;  // last srk_patch_save_srk() image
  std::string fused_source;          // last srk_fused_source() text
  std::string kernel_id;             // last srk_kernel_id() text
  std::string tune_report;           // last srk_schedule_report() text
  std::vector<unsigned char> saved_state;  // last srk_state_export() blob
  size_t co_resident_voices = 0;     // srk_set_co_resident_voices(): voices other patches render on this device concurrently
  std::unique_ptr<srk::Engine, void (*)(srk::Engine*)> engine{nullptr, nullptr};

  int index_of(const srk_module* m) const {
    for (size_t i = 0; i < modules.size(); ++i)
      if (modules[i] == m) return (int)i;
    return -1;
  }
  srk_module* find_output() const;  // ui.rs:84-96
};

namespace srk {

// plan_execution, src/synth.rs:128-212.  `deps[m]` lists the connected sources of
// module m (indices into the module list) in input order, duplicates kept.
// Returns the plan as list indices; `cuts` receives (reader, writer) index pairs.
void plan_execution(int output, const std::vector<std::vector<int>>& deps, std::vector<int>& plan,
                    std::vector<std::pair<int, int>>& cuts);

std::string make_uuid_v4();

}  // namespace srk
