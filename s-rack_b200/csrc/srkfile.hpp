// .srk patch files (the reference's FileFormat, src/ui.rs:578-586, rmp-serde MessagePack).  Host only.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/srack_b200.h"

namespace srk {

struct SrkModule {
  std::string variant;            // SynthModuleType variant name in the file
  int kind = -1;                  // srk_kind, -1 = outside the hot path (Freeverb)
  std::string id;
  float param[4] = {0, 0, 0, 0};
  std::vector<int32_t> sequence;  // sequencers (srk_set_sequence cell encoding)
  size_t seq_steps = 0;
  std::vector<float> wave;        // Sample: WaveBox.samples / sample_rate
  float wave_rate = 0.0f;
  float adsr_sample_rate = 0.0f;  // ADSR: the rate it was built with travels with the file (adsr.rs:17,69-71)
  bool has_adsr_rate = false;
  // The DSP state the file carries, as device state words (program.hpp "Per-voice state words"); empty = X::new()
  std::vector<uint32_t> state;
};
struct SrkConnection {
  std::string src_id, sink_id;
  uint8_t src_port = 0, sink_port = 0;
};
struct SrkPosition {
  std::string id;
  float x = 0, y = 0;
};
struct SrkFile {
  std::vector<SrkModule> modules;          // file order
  std::vector<SrkConnection> connections;  // file order
  std::vector<SrkPosition> positions;
};

bool srk_file_decode(const void* bytes, size_t n_bytes, SrkFile& out, std::string& err);
// Writes what the reference would save for modules with these settings: zeroed port buffers of `buffer_size`,
// and the DSP state in SrkModule::state (X::new() state when that is empty).
void srk_file_encode(const SrkFile& f, size_t buffer_size, uint16_t sample_rate, uint8_t channels,
                     std::vector<unsigned char>& out);

}  // namespace srk
