// WAV decode (WaveBox::load, src/synth/sample.rs:32-69) and WAV export.  Host only.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace srk {

enum WavStatus {
  WAV_OK = 0,
  WAV_BAD_HEADER,   // hound's Err before the reference touches the WaveBox: `samples` untouched
  WAV_UNSUPPORTED,  // the reference's DecodeError / a failed read after `samples.clear()`: `samples` emptied
};

WavStatus wav_decode(const void* bytes, size_t n_bytes, std::vector<float>& samples, float& sample_rate,
                     std::string& err);
bool wav_write(const char* path, const float* planar, unsigned channels, size_t n_samples, uint32_t sample_rate,
               int bits, std::string& err);

}  // namespace srk
