// WAV decode for the Sample module and WAV export of a render.  Host only.
//
// Decode follows WaveBox::load, src/synth/sample.rs:32-69, whose byte-level parsing is the
// hound 3.5.1 crate (Cargo.lock; not under /root/reference): `WavReader::new` walks the RIFF
// chunks up to `data`, `spec()` gives (channels, sample_rate, bits_per_sample, sample_format),
// `into_samples()` yields the interleaved samples.  The reference keeps channel 0 only
// (`idx % channels == 0`, :42,60) and converts 8 / 16 / 24-bit integers by dividing by 2^(bits-1)
// (:49-52); 32-bit float is taken as is (:38-47); any other integer width is a DecodeError raised
// AFTER `self.samples.clear()` (:36,53).
#include "wav.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace srk {

namespace {

uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// KSDATAFORMAT_SUBTYPE_{PCM, IEEE_FLOAT}: {0000000x-0000-0010-8000-00aa00389b71}, bytes 2..15
const unsigned char kGuidTail[14] = {0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71};

}  // namespace

WavStatus wav_decode(const void* bytes, size_t n_bytes, std::vector<float>& samples, float& sample_rate,
                     std::string& err) {
  const unsigned char* d = static_cast<const unsigned char*>(bytes);
  if (n_bytes < 12 || std::memcmp(d, "RIFF", 4) != 0 || std::memcmp(d + 8, "WAVE", 4) != 0) {
    err = "no RIFF/WAVE header";
    return WAV_BAD_HEADER;
  }
  size_t off = 12;
  bool have_fmt = false, is_float = false;
  unsigned channels = 0, bits = 0, bytes_per_sample = 0;
  uint32_t rate = 0;
  for (;;) {
    if (off + 8 > n_bytes) { err = "no data chunk"; return WAV_BAD_HEADER; }
    const unsigned char* kind = d + off;
    const size_t size = rd32(d + off + 4);
    off += 8;
    if (std::memcmp(kind, "fmt ", 4) == 0) {
      if (size < 16 || off + size > n_bytes) { err = "short fmt chunk"; return WAV_BAD_HEADER; }
      const unsigned tag = rd16(d + off);
      channels = rd16(d + off + 2);
      rate = rd32(d + off + 4);
      const unsigned align = rd16(d + off + 12);
      bits = rd16(d + off + 14);
      if (channels == 0) { err = "zero channels"; return WAV_BAD_HEADER; }
      bytes_per_sample = align / channels;
      if (tag == 1) {
        is_float = false;
      } else if (tag == 3) {
        if (bits != 32) { err = "IEEE float must be 32 bit"; return WAV_BAD_HEADER; }
        is_float = true;
      } else if (tag == 0xFFFE) {
        if (size < 40) { err = "short extensible fmt chunk"; return WAV_BAD_HEADER; }
        const unsigned valid_bits = rd16(d + off + 18);
        const unsigned char* guid = d + off + 24;
        const bool pcm = guid[0] == 1 && guid[1] == 0, flt = guid[0] == 3 && guid[1] == 0;
        if (std::memcmp(guid + 2, kGuidTail, 14) != 0 || !(pcm || flt)) { err = "unknown sub-format"; return WAV_BAD_HEADER; }
        if (valid_bits != 8 * bytes_per_sample) { err = "valid bits differ from the container size"; return WAV_BAD_HEADER; }
        bits = valid_bits;
        is_float = flt;
        if (is_float && bits != 32) { err = "IEEE float must be 32 bit"; return WAV_BAD_HEADER; }
      } else {
        err = "unsupported format tag";
        return WAV_BAD_HEADER;
      }
      if ((bits != 8 && bits != 16 && bits != 24 && bits != 32) || bytes_per_sample * 8 != bits) {
        err = "unsupported sample size";
        return WAV_BAD_HEADER;
      }
      have_fmt = true;
    } else if (std::memcmp(kind, "data", 4) == 0) {
      if (!have_fmt) { err = "data before fmt"; return WAV_BAD_HEADER; }
      // from here on the reference has already cleared the WaveBox's samples (sample.rs:36)
      samples.clear();
      if (!is_float && bits == 32) { err = "32-bit integer PCM (sample.rs:53 DecodeError)"; return WAV_UNSUPPORTED; }
      if (off + size > n_bytes) { err = "data chunk is truncated"; return WAV_UNSUPPORTED; }
      const size_t n = size / bytes_per_sample;
      samples.reserve(n / channels + 1);
      const unsigned char* p = d + off;
      for (size_t i = 0; i < n; i += channels) {  // channel 0 of every frame
        const unsigned char* q = p + i * bytes_per_sample;
        float x;
        if (is_float) {
          const uint32_t u = rd32(q);
          std::memcpy(&x, &u, 4);
        } else if (bits == 8) {
          x = (float)((int)q[0] - 128) / 128.0f;   // hound: u8 - 128; i8::MAX as f32 + 1.0
        } else if (bits == 16) {
          x = (float)(int16_t)rd16(q) / 32768.0f;
        } else {
          int32_t v = (int32_t)(q[0] | (q[1] << 8) | (q[2] << 16));
          if (v & 0x800000) v -= 0x1000000;
          x = (float)v / 8388608.0f;               // cpal I24::to_float_sample
        }
        samples.push_back(x);
      }
      sample_rate = (float)rate;
      return WAV_OK;
    }
    off += size + (size & 1);
  }
}

bool wav_write(const char* path, const float* planar, unsigned channels, size_t n_samples, uint32_t sample_rate,
               int bits, std::string& err) {
  if (!path || (!planar && n_samples) || channels == 0 || channels > 65535 || (bits != 16 && bits != 24 && bits != 32)) {
    err = "bad argument";
    return false;
  }
  const unsigned bps = (unsigned)bits / 8;
  const uint64_t data_bytes = (uint64_t)n_samples * channels * bps;
  if (data_bytes > 0xFFFFFFFFull - 44) { err = "render too long for a RIFF file"; return false; }
  std::FILE* f = std::fopen(path, "wb");
  if (!f) { err = "cannot open file"; return false; }
  unsigned char h[44];
  auto w16 = [](unsigned char* p, unsigned v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); };
  auto w32 = [](unsigned char* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (unsigned char)(v >> (8 * i)); };
  std::memcpy(h, "RIFF", 4); w32(h + 4, (uint32_t)(36 + data_bytes)); std::memcpy(h + 8, "WAVEfmt ", 8);
  w32(h + 16, 16); w16(h + 20, bits == 32 ? 3 : 1); w16(h + 22, channels); w32(h + 24, sample_rate);
  w32(h + 28, sample_rate * channels * bps); w16(h + 32, channels * bps); w16(h + 34, (unsigned)bits);
  std::memcpy(h + 36, "data", 4); w32(h + 40, (uint32_t)data_bytes);
  bool ok = std::fwrite(h, 1, 44, f) == 44;
  std::vector<unsigned char> row((size_t)4096 * channels * bps);
  const float scale = bits == 16 ? 32768.0f : 8388608.0f;
  for (size_t n0 = 0; ok && n0 < n_samples; n0 += 4096) {
    const size_t cnt = n_samples - n0 < 4096 ? n_samples - n0 : 4096;
    unsigned char* o = row.data();
    for (size_t i = 0; i < cnt; ++i)
      for (unsigned c = 0; c < channels; ++c) {
        const float x = planar[(size_t)c * n_samples + n0 + i];
        if (bits == 32) {
          uint32_t u;
          std::memcpy(&u, &x, 4);
          w32(o, u);
        } else {
          float y = std::nearbyintf(x * scale);
          if (!(y >= -scale)) y = -scale;          // also NaN -> most negative
          if (y > scale - 1.0f) y = scale - 1.0f;
          const int32_t v = (int32_t)y;
          for (unsigned b = 0; b < bps; ++b) o[b] = (unsigned char)((uint32_t)v >> (8 * b));
        }
        o += bps;
      }
    ok = std::fwrite(row.data(), 1, cnt * channels * bps, f) == cnt * channels * bps;
  }
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) err = "write failed";
  return ok;
}

}  // namespace srk
