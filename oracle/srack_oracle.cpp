// ============================================================================
// srack_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain, scalar, single-patch-instance-per-voice C++17 restatement of the
// s-rack module-graph tick (reference: src/synth.rs + src/synth/*.rs, commit
// 20e549b).  It exists so the CUDA path can be checked sample by sample and so
// bench.py has a "reference CPU path" to time.  Nothing under s-rack_b200/ may
// include, link, import or call it: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs do.
//
// It executes the reference's way: block based (`execute` = every module's
// `calc()` once per `buffer_size` samples, src/synth.rs:97-101), one f32
// buffer per output port, a cycle-cutting planner (src/synth.rs:107-212), f64
// exactly where the reference uses f64 (oscillator phase/frequency/sin), f32
// everywhere else, no FMA contraction (build with -ffp-contract=off), glibc
// sin/exp2/fmod/powf.
//
// PARITY PINNING.  The Rust reference cannot be compiled in this environment
// (no rustc/cargo, crates not vendored), so the oracle is pinned against the
// reference's own tests only:
//   * dco_tests::produces_440        (src/synth/oscillator.rs:284-305)
//   * tests::topological_sort        (src/synth.rs:537-613)
// Both are restated in tests/test_oracle_kat.py.  The reference has no test
// for MoogFilter, ADSR, VCA, MonoMixer, Math, NonLinear, Output or Noise, so
// for those modules this oracle is a line-by-line restatement with
// **parity unpinned** by any reference vector.  NoiseModule additionally uses
// an unseeded OS RNG in the reference (src/synth/oscillator.rs:385) and can
// not be reproduced even in principle: this oracle (and the CUDA path) define
// a seeded counter-based generator (Philox4x32-10) with the reference's
// u32 -> [0,1) -> (r-0.5)*2 mapping instead.
//
// Third-party arithmetic the reference reaches through Rust std / crates and
// that is not under /root/reference: platform libm (sin, pow/exp2, fmod, powf)
// -- pinned here to glibc 2.39; rand 0.8.5 / rand_chacha 0.3.1 (Cargo.lock) --
// replaced as described above.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace {

// ---- kinds / params: numeric ids deliberately equal to include/srack_b200.h
enum Kind {
  K_OUTPUT = 0, K_OSC = 1, K_NOISE = 2, K_ADSR = 3, K_VCA = 4, K_MOOG = 5,
  K_MIXER = 6, K_ADD = 7, K_SUB = 8, K_MUL = 9, K_NONLIN = 10, K_GRIDSEQ = 11, K_PATSEQ = 12, K_SAMPLE = 13, K_COUNT
};

// src/synth.rs:20-25
struct AudioConfig {
  uint16_t sample_rate;
  size_t buffer_size;
  uint8_t channels;
};

// src/synth.rs:276-298 -- `last` starts true (:283)
struct TransitionDetector {
  bool last = true;
  bool is_transition(float val) {
    bool above = val > 0.0f;
    bool t = above && !last;
    last = above;
    return t;
  }
};

inline float word_f32(uint32_t w) { float f; std::memcpy(&f, &w, 4); return f; }

// ---- Philox4x32-10 (Salmon et al., SC'11), our seeded stand-in for rand::random
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

struct Module;
using Input = std::optional<std::pair<Module*, uint8_t>>;

// trait SynthModule, src/synth.rs:222-263 (GUI / serde parts omitted)
struct Module {
  int kind;
  int index = -1;                      // position in the patch's module list
  std::vector<Input> inputs;           // get_input / set_input
  std::vector<std::vector<float>> outs;  // AudioBuffer(Some), zero-initialised (:32)
  Module(int k, int n_in, int n_out, size_t B) : kind(k), inputs(n_in), outs(n_out) {
    for (auto& o : outs) o.assign(B, 0.0f);
  }
  virtual ~Module() {}
  virtual void calc() = 0;
  virtual void reset() = 0;
  virtual bool set_param(int pid, float v) = 0;
  // DSP state from a deserialized module, as the device's state words (s-rack_b200/csrc/program.hpp)
  virtual bool set_state(const uint32_t*, size_t) { return false; }
  // resolve_input, src/synth.rs:249-254: None -> AudioBuffer(None)
  const float* resolve(int i) const {
    if (!inputs[i]) return nullptr;
    return inputs[i]->first->outs[inputs[i]->second].data();
  }
  size_t B() const { return outs.empty() ? 0 : outs[0].size(); }
};

// ---- src/synth/oscillator.rs:9-158
struct Oscillator : Module {
  float val = 0.0f;            // :32
  bool antialiasing = true;    // :38
  uint16_t sample_rate;
  double pos = 0.0;            // :37
  TransitionDetector sync_detector;
  Oscillator(const AudioConfig& c) : Module(K_OSC, 2, 3, c.buffer_size), sample_rate(c.sample_rate) {}
  void reset() override { pos = 0.0; sync_detector = TransitionDetector(); for (auto& o : outs) std::fill(o.begin(), o.end(), 0.0f); }
  bool set_param(int pid, float v) override {
    if (pid == 0) { val = v; return true; }
    if (pid == 1) { antialiasing = v != 0.0f; return true; }
    return false;
  }
  bool set_state(const uint32_t* w, size_t n) override {  // pos (f64 lo, hi), sync detector
    if (n != 3) return false;
    const uint64_t b = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    std::memcpy(&pos, &b, 8);
    sync_detector.last = w[2] != 0;
    return true;
  }
  // :43-48.  2.0_f64.powf(x): LLVM folds pow(2.0, x) to exp2(x) in optimised builds.
  double freq_hz(const float* cv, size_t i) const {
    if (cv) return 440.0 * std::exp2((double)cv[i] + (double)val);
    return 440.0 * std::exp2((double)val);
  }
  // :50-67
  static double poly_blep(double t, double dt) {
    if (dt == 0.0) return 0.0;
    if (t < dt) {
      t /= dt;
      return ((t + t) - (t * t)) - 1.0;
    } else if (t > 1.0 - dt) {
      t = (t - 1.0) / dt;
      return (((t * t) + t) + t) + 1.0;
    }
    return 0.0;
  }
  // :108-158
  void calc() override {
    const float* cv = resolve(0);
    const float* sync_in = resolve(1);
    float* sine = outs[0].data();
    float* square = outs[1].data();
    float* saw = outs[2].data();
    const size_t n = outs[0].size();
    for (size_t i = 0; i < n; ++i) {
      float sync_val = sync_in ? sync_in[i] : 0.0f;
      if (sync_detector.is_transition(sync_val)) pos = 0.0;
      double delta = freq_hz(cv, i) / (double)sample_rate;
      sine[i] = (float)std::sin((pos * M_PI) * 2.0);
      float sq_base = pos < 0.5 ? -1.0f : 1.0f;
      float sq_corr = antialiasing
                          ? (float)(poly_blep(pos, delta) - poly_blep(std::fmod(pos + 0.5, 1.0), delta))
                          : 0.0f;
      square[i] = sq_base - sq_corr;
      float saw_corr = antialiasing ? (float)poly_blep(pos, delta) : 0.0f;
      saw[i] = ((float)pos * 2.0f - 1.0f) - saw_corr;
      pos += delta;
      pos = std::fmod(pos, 1.0);
    }
  }
};

// ---- src/synth/oscillator.rs:308-393; RNG replaced (see header)
struct Noise : Module {
  uint64_t seed = 0;
  uint64_t voice = 0;   // global voice index
  uint64_t n = 0;       // samples generated since reset
  Noise(const AudioConfig& c) : Module(K_NOISE, 0, 1, c.buffer_size) {}
  void reset() override { n = 0; std::fill(outs[0].begin(), outs[0].end(), 0.0f); }
  bool set_param(int, float) override { return false; }
  void calc() override {
    float* out = outs[0].data();
    for (size_t i = 0; i < outs[0].size(); ++i, ++n) {
      uint64_t blk = n >> 2;
      uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)voice, (uint32_t)index};
      philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
      uint32_t r = c[n & 3];
      // rand 0.8.5 Standard for f32: (u32 >> 8) * 2^-24 in [0,1)
      float u = (float)(r >> 8) * (1.0f / 16777216.0f);
      out[i] = (u - 0.5f) * 2.0f;  // :385
    }
  }
};

// ---- src/synth/filter.rs:48-92 (InternalMoogFilterState) and :182-221
struct MoogFilter : Module {
  float freq = 0.2f, res = 0.5f, exp_amt = 0.5f;  // :36-38
  struct State { float f = 0, p = 0, q = 0, b[5] = {0, 0, 0, 0, 0}, freq = 0, res = 0; } st;
  MoogFilter(const AudioConfig& c) : Module(K_MOOG, 2, 3, c.buffer_size) {}
  void reset() override { st = State(); for (auto& o : outs) std::fill(o.begin(), o.end(), 0.0f); }
  bool set_param(int pid, float v) override {
    if (pid == 0) { freq = v; return true; }
    if (pid == 1) { res = v; return true; }
    if (pid == 2) { exp_amt = v; return true; }
    return false;
  }
  bool set_state(const uint32_t* w, size_t n) override {  // f, p, q, b[0..4], freq, res
    if (n != 10) return false;
    st.f = word_f32(w[0]); st.p = word_f32(w[1]); st.q = word_f32(w[2]);
    for (int k = 0; k < 5; ++k) st.b[k] = word_f32(w[3 + k]);
    st.freq = word_f32(w[8]); st.res = word_f32(w[9]);
    return true;
  }
  // :60-83; returns (b4, in - b4, 3*(b3-b4))
  void state_calc(float input, float frequency, float resonance, float& r0, float& r1, float& r2) {
    if (frequency != st.freq || resonance != st.res) {
      st.freq = frequency;
      st.res = resonance;
      st.q = 1.0f - st.freq;
      st.p = st.freq + (0.8f * st.freq) * st.q;
      st.f = st.p * 2.0f - 1.0f;
      st.q = st.res * (1.0f + (0.5f * st.q) * ((1.0f - st.q) + (5.6f * st.q) * st.q));
    }
    input = input - (st.q * st.b[4]);
    float t1 = st.b[1];
    st.b[1] = (input + st.b[0]) * st.p - st.b[1] * st.f;
    float t2 = st.b[2];
    st.b[2] = (st.b[1] + t1) * st.p - st.b[2] * st.f;
    t1 = st.b[3];
    st.b[3] = (st.b[2] + t2) * st.p - st.b[3] * st.f;
    st.b[4] = (st.b[3] + t1) * st.p - st.b[4] * st.f;
    st.b[4] = st.b[4] - ((st.b[4] * st.b[4]) * st.b[4]) * 0.166667f;  // powi(3)
    st.b[0] = input;
    for (float& x : st.b) x = std::fmax(std::fmin(x, 1.0f), -1.0f);  // :86-91
    r0 = st.b[4];
    r1 = input - st.b[4];
    r2 = 3.0f * (st.b[3] - st.b[4]);
  }
  void calc() override {
    const float* audio_in = resolve(0);
    const float* cv_in = resolve(1);
    float* lowpass = outs[0].data();
    float* bandpass = outs[1].data();
    float* highpass = outs[2].data();
    for (size_t i = 0; i < outs[0].size(); ++i) {
      float audio = audio_in ? audio_in[i] : 0.0f;
      float cv = cv_in ? cv_in[i] : 0.0f;
      // :211  (lowpass, highpass, bandpass) = state.calc(...)
      state_calc(audio, std::fmin(std::fmax(freq + cv * exp_amt, 0.0f), 0.9f),
                 std::fmin(std::fmax(res, 0.0f), 1.0f), lowpass[i], highpass[i], bandpass[i]);
    }
  }
};

// ---- src/synth/adsr.rs:8-53, :134-217
struct ADSR : Module {
  enum Mode { Attack, Decay, Sustain, Release, None };
  float a_sec = 0.0f, d_sec = 0.5f, s_val = 0.25f, r_sec = 0.5f;  // :39-42
  float phase = 0.0f;
  Mode mode = None;
  float r_val = 0.0f, from_a_val = 0.0f;
  float sample_rate;  // fixed at construction (:47)
  TransitionDetector det;
  ADSR(const AudioConfig& c) : Module(K_ADSR, 1, 1, c.buffer_size), sample_rate((float)c.sample_rate) {}
  void reset() override {
    phase = 0; mode = None; r_val = 0; from_a_val = 0; det = TransitionDetector();
    std::fill(outs[0].begin(), outs[0].end(), 0.0f);
  }
  bool set_param(int pid, float v) override {
    switch (pid) {
      case 0: a_sec = v; return true;
      case 1: d_sec = v; return true;
      case 2: s_val = v; return true;
      case 3: r_sec = v; return true;
      case 100: sample_rate = v; return true;  // a deserialized ADSR keeps the rate it was saved with (adsr.rs:17,69-71)
    }
    return false;
  }
  bool set_state(const uint32_t* w, size_t n) override {  // phase, r_val, from_a_val, mode | gate_last << 8
    if (n != 4 || (w[3] & 0xFF) > 4) return false;
    phase = word_f32(w[0]); r_val = word_f32(w[1]); from_a_val = word_f32(w[2]);
    mode = (Mode)(w[3] & 0xFF);
    det.last = (w[3] >> 8) & 1;
    return true;
  }
  void calc() override {
    const float* gate = resolve(0);
    float* out = outs[0].data();
    for (size_t i = 0; i < outs[0].size(); ++i) {
      bool tr = det.is_transition(gate ? gate[i] : 0.0f);
      switch (mode) {
        case None:
          if (gate && gate[i] > 0.0f) { phase = 0.0f; mode = Attack; }
          break;
        case Attack:
          phase += 1.0f / (sample_rate * a_sec);
          if (phase >= 1.0f) { phase = 0.0f; mode = Decay; }
          else if (tr) { phase = 0.0f; r_val = from_a_val; }
          break;
        case Decay:
          phase += 1.0f / (sample_rate * d_sec);
          if (phase >= 1.0f) { phase = 0.0f; mode = Sustain; }
          if (tr) { phase = 0.0f; mode = Attack; }
          break;
        case Sustain:
          if (!gate || gate[i] <= 0.0f) { phase = 0.0f; mode = Release; }
          if (tr) { phase = 0.0f; mode = Attack; }
          break;
        case Release:
          if (gate && gate[i] > 0.0f) { phase = 0.0f; mode = Attack; }
          phase += 1.0f / (sample_rate * r_sec);
          if (phase >= 1.0f) { phase = 0.0f; r_val = 0.0f; mode = None; }
          break;
      }
      float o = 0.0f;
      switch (mode) {
        case None: o = 0.0f; break;
        case Attack: o = r_val + (1.0f - r_val) * phase; break;
        case Decay: o = s_val + (1.0f - s_val) * (1.0f - phase); break;
        case Sustain: o = s_val; break;
        case Release: o = s_val * (1.0f - phase); break;
      }
      out[i] = o;
      if (mode != Attack) r_val = o; else from_a_val = o;
    }
  }
};

// ---- src/synth/vca.rs:117-148
struct VCA : Module {
  bool negative = false;
  VCA(const AudioConfig& c) : Module(K_VCA, 2, 1, c.buffer_size) {}
  void reset() override { std::fill(outs[0].begin(), outs[0].end(), 0.0f); }
  bool set_param(int pid, float v) override { if (pid == 0) { negative = v != 0.0f; return true; } return false; }
  void calc() override {
    const float* audio = resolve(0);
    const float* cv = resolve(1);
    float* out = outs[0].data();
    if (audio && cv) {
      for (size_t i = 0; i < outs[0].size(); ++i)
        out[i] = (negative || cv[i] > 0.0f) ? audio[i] * cv[i] : 0.0f;
    } else {
      std::fill(outs[0].begin(), outs[0].end(), 0.0f);
    }
  }
};

// ---- src/synth/mixer.rs:101-122
struct MonoMixer : Module {
  float gain[4] = {1.0f, 1.0f, 1.0f, 1.0f};
  MonoMixer(const AudioConfig& c) : Module(K_MIXER, 4, 1, c.buffer_size) {}
  void reset() override { std::fill(outs[0].begin(), outs[0].end(), 0.0f); }
  bool set_param(int pid, float v) override { if (pid >= 0 && pid < 4) { gain[pid] = v; return true; } return false; }
  void calc() override {
    // NB: inputs are resolved (and, in Rust, read-locked) before the output is
    // filled, but the data are read after the fill; only a self-loop could
    // observe the difference and that deadlocks in the reference.
    float* out = outs[0].data();
    const float* in[4] = {resolve(0), resolve(1), resolve(2), resolve(3)};
    std::fill(outs[0].begin(), outs[0].end(), 0.0f);
    for (int k = 0; k < 4; ++k) {
      if (!in[k]) continue;
      for (size_t i = 0; i < outs[0].size(); ++i) out[i] += in[k][i] * gain[k];
    }
  }
};

// ---- src/synth/math.rs:14-53, :139-160 and :177-206, :292-313
struct Math : Module {
  float constant;
  Math(const AudioConfig& c, int kind) : Module(kind, 2, 1, c.buffer_size), constant(kind == K_NONLIN ? 1.0f : 0.0f) {}
  void reset() override { std::fill(outs[0].begin(), outs[0].end(), 0.0f); }
  bool set_param(int pid, float v) override { if (pid == 0) { constant = v; return true; } return false; }
  float op(float a, float b) const {
    switch (kind) {
      case K_ADD: return a + b;
      case K_SUB: return a - b;
      case K_MUL: return a * b;
      default: return a > 0.0f ? std::pow(a, b) : -std::pow(-a, b);  // math.rs:203-205 (f32 powf)
    }
  }
  void calc() override {
    const float* i1 = resolve(0);
    const float* i2 = resolve(1);
    float* out = outs[0].data();
    for (size_t i = 0; i < outs[0].size(); ++i) {
      if (i1 && i2) out[i] = op(i1[i], i2[i]);
      else if (i1) out[i] = op(i1[i], constant);
      else if (i2) out[i] = op(0.0f, i2[i]);
      else out[i] = op(0.0f, constant);
    }
  }
};

// ---- src/synth/sequencer.rs:12-61 (struct, defaults), :190-246 (calc).
// Inputs 0 = Step (clock), 1 = Sync; outputs 0 = CV, 1 = Gate, 2 = Sync (:262-307).
// A cell is Option<(u16 val, bool hold)>; here -1 = None, else val | hold << 16.
struct GridSequencer : Module {
  std::vector<int32_t> sequence = std::vector<int32_t>(64, -1);  // vec![None; 64] (:40)
  uint16_t steps_per_octave = 12;                                  // (:44)
  uint16_t current_step = 0;
  TransitionDetector det, sync_det;
  float last = 0.0f;
  GridSequencer(const AudioConfig& c) : Module(K_GRIDSEQ, 2, 3, c.buffer_size) {}
  void reset() override {
    current_step = 0; det = TransitionDetector(); sync_det = TransitionDetector(); last = 0.0f;
    for (auto& o : outs) std::fill(o.begin(), o.end(), 0.0f);
  }
  bool set_param(int pid, float v) override { if (pid == 0) { steps_per_octave = (uint16_t)v; return true; } return false; }
  bool set_state(const uint32_t* w, size_t n) override {  // step | step_last << 16 | sync_last << 17, last cv
    if (n != 2) return false;
    current_step = (uint16_t)(w[0] & 0xFFFF); det.last = (w[0] >> 16) & 1; sync_det.last = (w[0] >> 17) & 1;
    last = word_f32(w[1]);
    return true;
  }
  void calc() override {
    const float* step_buf = resolve(0);
    const float* sync_buf = resolve(1);
    float* cv_out = outs[0].data(); float* gate_out = outs[1].data(); float* sync_out = outs[2].data();
    for (size_t idx = 0; idx < outs[0].size(); ++idx) {
      const float step_in = step_buf ? step_buf[idx] : 0.0f;
      const float sync_in = sync_buf ? sync_buf[idx] : 0.0f;
      if (det.is_transition(step_in)) current_step += 1;
      if (sync_det.is_transition(sync_in)) current_step = 0;
      size_t cur = current_step;
      if (cur >= sequence.size()) { current_step = 0; cur = 0; }
      const int32_t cell = sequence[cur];
      if (cell >= 0) {
        cv_out[idx] = (float)(uint16_t)(cell & 0xFFFF) * (1.0f / (float)steps_per_octave);
        gate_out[idx] = (cell >> 16) & 1 ? 1.0f : step_in;
      } else {
        cv_out[idx] = last;
        gate_out[idx] = 0.0f;
      }
      sync_out[idx] = cur == 0 ? 1.0f : 0.0f;
      last = cv_out[idx];
    }
  }
};

// ---- src/synth/sequencer.rs:336-366 (struct, defaults), :482-533 (calc).
// Inputs 0 = Step, 1 = Sync; outputs 0..7 = gate rows, 8 = Sync (:577-596).
// A cell is Option<bool>; here -1 = None, 0 = Some(false), 1 = Some(true); rows x steps.
struct PatternSequencer : Module {
  std::vector<std::vector<int32_t>> sequence = std::vector<std::vector<int32_t>>(8, std::vector<int32_t>(64, -1));
  uint16_t current_step = 0;
  TransitionDetector det, sync_det;
  PatternSequencer(const AudioConfig& c) : Module(K_PATSEQ, 2, 9, c.buffer_size) {}
  void reset() override {
    current_step = 0; det = TransitionDetector(); sync_det = TransitionDetector();
    for (auto& o : outs) std::fill(o.begin(), o.end(), 0.0f);
  }
  bool set_param(int, float) override { return false; }
  bool set_state(const uint32_t* w, size_t n) override {
    if (n != 1) return false;
    current_step = (uint16_t)(w[0] & 0xFFFF); det.last = (w[0] >> 16) & 1; sync_det.last = (w[0] >> 17) & 1;
    return true;
  }
  void calc() override {
    const float* step_buf = resolve(0);
    const float* sync_buf = resolve(1);
    for (size_t idx = 0; idx < outs[0].size(); ++idx) {
      const float step_in = step_buf ? step_buf[idx] : 0.0f;
      const float sync_in = sync_buf ? sync_buf[idx] : 0.0f;
      if (det.is_transition(step_in)) current_step += 1;
      if (sync_det.is_transition(sync_in)) current_step = 0;
      size_t cur = current_step;
      if (cur >= sequence[0].size()) { current_step = 0; cur = 0; }
      for (size_t row = 0; row < 8; ++row) {
        const int32_t cell = sequence[row][cur];
        outs[row][idx] = cell < 0 ? 0.0f : (cell ? 1.0f : step_in);
      }
      outs[8][idx] = cur == 0 ? 1.0f : 0.0f;
    }
  }
};

// ---- src/synth/sample.rs:14-20 (WaveBox), :72-84 (struct), :192-240 (calc).
// Inputs 0 = Gate, 1 = CV; output 0.  The WaveBox (decoded channel-0 samples + the file's sample
// rate) is shared by all voices of the patch; `is_new` restates WaveBox.new per instance: the
// first calc() after a load rewinds (:212-216).  The try_lock-failed arm (:205-210) only runs
// while the GUI thread is decoding a file and is not restated.  WaveBox::load itself (hound WAV
// decode, :32-69) is restated in oracle/wav.py.
struct WaveBox {
  std::vector<float> samples;
  float sample_rate = 0.0f;  // Default (:14)
};

// Rust `f32 as usize`: truncates toward zero, saturates, NaN -> 0.
inline size_t f32_as_usize(float x) {
  if (!(x > 0.0f)) return 0;                       // NaN, negatives, -0.0, +0.0
  if (x >= 18446744073709551616.0f) return SIZE_MAX;
  return (size_t)x;
}

struct Sample : Module {
  std::shared_ptr<const WaveBox> wavebox = std::make_shared<WaveBox>();
  bool is_new = false;
  TransitionDetector det;
  float pos = 0.0f;
  bool playing = false;
  float sample_rate;
  Sample(const AudioConfig& c) : Module(K_SAMPLE, 2, 1, c.buffer_size), sample_rate((float)c.sample_rate) {}
  void reset() override {
    det = TransitionDetector(); pos = 0.0f; playing = false;
    std::fill(outs[0].begin(), outs[0].end(), 0.0f);
  }
  bool set_param(int, float) override { return false; }
  bool set_state(const uint32_t* w, size_t n) override {  // pos, playing | gate_last << 1
    if (n != 2) return false;
    pos = word_f32(w[0]); playing = w[1] & 1; det.last = (w[1] >> 1) & 1;
    is_new = false;  // (a file whose WaveBox.new was set arrives here already rewound, srk_file.state_words)
    return true;
  }
  void calc() override {
    const float* gate_in = resolve(0);
    const float* cv_in = resolve(1);
    float* output = outs[0].data();
    const WaveBox& wb = *wavebox;
    if (is_new) { pos = 0.0f; playing = false; is_new = false; }   // :212-216
    for (size_t idx = 0; idx < outs[0].size(); ++idx) {
      const bool trigger = det.is_transition(gate_in ? gate_in[idx] : 0.0f);
      if (trigger) { pos = 0.0f; playing = true; }
      if (f32_as_usize(pos) >= wb.samples.size()) { pos = 0.0f; playing = false; }
      output[idx] = !wb.samples.empty() ? wb.samples[f32_as_usize(pos)] : 0.0f;
      if (playing) {
        // `wavebox.sample_rate / self.sample_rate * 2.0_f32.powf(cv)` (:234-235); LLVM folds
        // pow(2.0f, x) to exp2f(x) in optimised builds -> glibc exp2f.
        const float e = exp2f(cv_in ? cv_in[idx] : 0.0f);
        pos += (wb.sample_rate / sample_rate) * e;
      }
    }
  }
};

// ---- src/synth/output.rs:46-60.  `bufs` are kept as outs[] so they can be read.
struct Output : Module {
  Output(const AudioConfig& c) : Module(K_OUTPUT, c.channels, c.channels, c.buffer_size) {}
  void reset() override { for (auto& o : outs) std::fill(o.begin(), o.end(), 0.0f); }
  bool set_param(int, float) override { return false; }
  void calc() override {
    for (size_t c = 0; c < inputs.size(); ++c) {
      const float* in = resolve((int)c);
      if (in) std::memcpy(outs[c].data(), in, outs[c].size() * sizeof(float));
      else std::fill(outs[c].begin(), outs[c].end(), 0.0f);
    }
  }
};

int kind_num_outputs(int kind) {
  switch (kind) {
    case K_OUTPUT: return 0;  // get_num_outputs() == 0 (output.rs:62)
    case K_OSC: return 3;
    case K_MOOG: return 3;
    case K_GRIDSEQ: return 3;
    case K_PATSEQ: return 9;
    default: return 1;
  }
}

std::unique_ptr<Module> make_module(int kind, const AudioConfig& cfg) {
  switch (kind) {
    case K_OUTPUT: return std::make_unique<Output>(cfg);
    case K_OSC: return std::make_unique<Oscillator>(cfg);
    case K_NOISE: return std::make_unique<Noise>(cfg);
    case K_ADSR: return std::make_unique<ADSR>(cfg);
    case K_VCA: return std::make_unique<VCA>(cfg);
    case K_MOOG: return std::make_unique<MoogFilter>(cfg);
    case K_MIXER: return std::make_unique<MonoMixer>(cfg);
    case K_ADD: case K_SUB: case K_MUL: case K_NONLIN: return std::make_unique<Math>(cfg, kind);
    case K_GRIDSEQ: return std::make_unique<GridSequencer>(cfg);
    case K_PATSEQ: return std::make_unique<PatternSequencer>(cfg);
    case K_SAMPLE: return std::make_unique<Sample>(cfg);
  }
  return nullptr;
}

// ---- src/synth.rs:107-126.  Returns the node whose dependency list contains
// `module` (reached from `module` through `edges`), or nullptr.
Module* is_loop(Module* module, std::unordered_map<Module*, std::vector<Module*>>& edges) {
  std::vector<Module*> to_search{module};
  std::vector<Module*> to_add;
  std::unordered_set<Module*> visited;
  for (;;) {
    Module* current = nullptr;
    for (Module* m : to_search)
      if (!visited.count(m)) { current = m; break; }
    if (!current) return nullptr;
    visited.insert(current);
    for (Module* dep : edges.at(current)) {
      if (dep == module) return current;
      to_add.push_back(dep);
    }
    to_search.insert(to_search.end(), to_add.begin(), to_add.end());
    to_add.clear();
  }
}

// ---- src/synth.rs:128-212
void plan_execution(Module* output, const std::vector<Module*>& all_modules, std::vector<Module*>& plan,
                    std::vector<std::pair<Module*, Module*>>* cuts /* (reader, writer) */) {
  std::unordered_map<Module*, std::vector<Module*>> edges;  // K: sink, V: sources
  std::unordered_set<Module*> visited;
  std::vector<Module*> to_search = all_modules;
  to_search.push_back(output);
  while (!to_search.empty()) {  // create all edges
    Module* m = to_search.back();
    to_search.pop_back();
    if (!visited.insert(m).second) continue;
    std::vector<Module*> deps;
    for (auto& in : m->inputs)
      if (in) {
        to_search.push_back(in->first);
        deps.push_back(in->first);
      }
    edges[m] = deps;
  }
  to_search = all_modules;
  plan.clear();
  visited.clear();
  to_search.push_back(output);
  while (!to_search.empty()) {  // remove cycles
    Module* m = to_search.back();
    to_search.pop_back();
    if (!visited.insert(m).second) continue;
    for (Module* dep : edges.at(m)) to_search.push_back(dep);
    while (Module* from = is_loop(m, edges)) {
      auto& deps = edges.at(from);
      deps.erase(std::remove(deps.begin(), deps.end(), m), deps.end());
      if (cuts) cuts->push_back({from, m});
    }
  }
  visited.clear();
  for (;;) {  // first unvisited module (in all_modules order) with all deps visited
    Module* node = nullptr;
    for (Module* m : all_modules) {
      if (visited.count(m)) continue;
      bool ready = true;
      for (Module* d : edges.at(m))
        if (!visited.count(d)) { ready = false; break; }
      if (ready) { node = m; break; }
    }
    if (!node) break;
    visited.insert(node);
    plan.push_back(node);
  }
}

// ---- src/synth.rs:97-101
void execute(const std::vector<Module*>& plan) {
  for (Module* m : plan) m->calc();
}

// One patch instance (what the reference app holds exactly one of).
struct Instance {
  std::vector<std::unique_ptr<Module>> modules;
  std::vector<Module*> plan;
  Module* output = nullptr;
};

struct ParamSetting {
  int module, pid;
  bool per_voice;
  float value;
  std::vector<float> values;
};

struct Conn { int sink, in_idx, src, port; bool connected; };

struct Patch {
  AudioConfig cfg;
  uint64_t seed = 0x5EED5EEDull;
  std::vector<int> kinds;
  std::vector<std::vector<std::optional<std::pair<int, int>>>> wiring;  // [module][input] -> (src, port)
  std::vector<ParamSetting> params;  // applied in order
  std::unordered_map<int, std::vector<int32_t>> sequences;  // module -> cells (rows x steps for the pattern sequencer)
  std::unordered_map<int, std::shared_ptr<const WaveBox>> waves;  // Sample module -> its WaveBox
  std::unordered_map<int, std::vector<uint32_t>> states;          // module -> deserialized DSP state (device words)
  std::vector<int> order;            // all_modules order (module indices); empty = creation order
  // voice bank
  std::vector<Instance> voices;
  size_t bank_offset = 0;
  bool bank_dirty = true;

  int n_inputs(int m) const {
    switch (kinds[m]) {
      case K_OUTPUT: return cfg.channels;
      case K_OSC: case K_VCA: case K_MOOG: case K_ADD: case K_SUB: case K_MUL: case K_NONLIN: return 2;
      case K_GRIDSEQ: case K_PATSEQ: case K_SAMPLE: return 2;
      case K_NOISE: return 0;
      case K_ADSR: return 1;
      case K_MIXER: return 4;
    }
    return 0;
  }

  void build_instance(Instance& inst, size_t voice) const {
    inst.modules.clear();
    for (size_t m = 0; m < kinds.size(); ++m) {
      auto mod = make_module(kinds[m], cfg);
      mod->index = (int)m;
      if (kinds[m] == K_SAMPLE) {
        auto it = waves.find((int)m);
        if (it != waves.end()) { static_cast<Sample*>(mod.get())->wavebox = it->second; static_cast<Sample*>(mod.get())->is_new = true; }
      }
      if (kinds[m] == K_NOISE) {
        auto* nz = static_cast<Noise*>(mod.get());
        nz->seed = seed;
        nz->voice = voice;
      }
      inst.modules.push_back(std::move(mod));
    }
    for (size_t m = 0; m < kinds.size(); ++m)
      for (size_t i = 0; i < wiring[m].size(); ++i)
        if (wiring[m][i])
          inst.modules[m]->inputs[i] = std::make_pair(inst.modules[wiring[m][i]->first].get(), (uint8_t)wiring[m][i]->second);
    apply_params(inst, voice);
    for (const auto& kv : states) inst.modules[kv.first]->set_state(kv.second.data(), kv.second.size());
    inst.output = nullptr;
    std::vector<Module*> all;
    if (order.empty())
      for (auto& m : inst.modules) all.push_back(m.get());
    else
      for (int idx : order) all.push_back(inst.modules[idx].get());
    for (Module* m : all)  // find_output, ui.rs:84-96: first Output in list order
      if (m->kind == K_OUTPUT) { inst.output = m; break; }
    inst.plan.clear();
    if (inst.output) plan_execution(inst.output, all, inst.plan, nullptr);
  }

  void apply_params(Instance& inst, size_t voice) const {
    for (const auto& p : params) {
      float v = p.per_voice ? p.values[voice] : p.value;
      inst.modules[p.module]->set_param(p.pid, v);
    }
    for (const auto& kv : sequences) {
      Module* m = inst.modules[kv.first].get();
      if (m->kind == K_GRIDSEQ) {
        static_cast<GridSequencer*>(m)->sequence = kv.second;
      } else if (m->kind == K_PATSEQ) {
        auto* ps = static_cast<PatternSequencer*>(m);
        const size_t steps = kv.second.size() / 8;
        for (size_t row = 0; row < 8; ++row)
          ps->sequence[row].assign(kv.second.begin() + row * steps, kv.second.begin() + (row + 1) * steps);
      }
    }
  }
};

}  // namespace

// ============================================================================
// C API (ctypes-friendly).  Module handles are indices into the module list.
// ============================================================================
extern "C" {

void* orc_patch_create(uint16_t sample_rate, size_t buffer_size, uint8_t channels) {
  auto* p = new Patch();
  p->cfg = AudioConfig{sample_rate, buffer_size, channels};
  return p;
}

void orc_patch_destroy(void* h) { delete static_cast<Patch*>(h); }

void orc_set_seed(void* h, uint64_t seed) {
  auto* p = static_cast<Patch*>(h);
  p->seed = seed;
  p->bank_dirty = true;
}

int orc_module_create(void* h, int kind) {
  auto* p = static_cast<Patch*>(h);
  if (kind < 0 || kind >= K_COUNT) return -1;
  p->kinds.push_back(kind);
  p->wiring.emplace_back(p->n_inputs((int)p->kinds.size() - 1));
  p->bank_dirty = true;
  return (int)p->kinds.size() - 1;
}

int orc_num_inputs(void* h, int m) { return static_cast<Patch*>(h)->n_inputs(m); }
int orc_num_outputs(void* h, int m) { return kind_num_outputs(static_cast<Patch*>(h)->kinds[m]); }

// set_input: Err(()) on a bad input index (e.g. oscillator.rs:199).  The source
// port is validated here too (the reference would panic later in resolve_input).
int orc_connect(void* h, int sink, int in_idx, int src, int src_port) {
  auto* p = static_cast<Patch*>(h);
  int n = (int)p->kinds.size();
  if (sink < 0 || sink >= n || src < 0 || src >= n) return 1;
  if (in_idx < 0 || in_idx >= p->n_inputs(sink)) return 2;
  if (src_port < 0 || src_port >= kind_num_outputs(p->kinds[src])) return 2;
  if (sink == src) return 3;  // self-loop: deadlocks in the reference (RwLock write then read)
  p->wiring[sink][in_idx] = std::make_pair(src, src_port);
  p->bank_dirty = true;
  return 0;
}

int orc_disconnect(void* h, int sink, int in_idx) {
  auto* p = static_cast<Patch*>(h);
  if (sink < 0 || sink >= (int)p->kinds.size()) return 1;
  if (in_idx < 0 || in_idx >= p->n_inputs(sink)) return 2;
  p->wiring[sink][in_idx].reset();
  p->bank_dirty = true;
  return 0;
}

int orc_set_param(void* h, int module, int pid, float value) {
  auto* p = static_cast<Patch*>(h);
  if (module < 0 || module >= (int)p->kinds.size()) return 1;
  auto probe = make_module(p->kinds[module], p->cfg);
  if (!probe->set_param(pid, value)) return 2;
  p->params.push_back(ParamSetting{module, pid, false, value, {}});
  for (size_t v = 0; v < p->voices.size(); ++v) p->voices[v].modules[module]->set_param(pid, value);
  return 0;
}

int orc_set_param_per_voice(void* h, int module, int pid, const float* values, size_t n) {
  auto* p = static_cast<Patch*>(h);
  if (module < 0 || module >= (int)p->kinds.size()) return 1;
  auto probe = make_module(p->kinds[module], p->cfg);
  if (!probe->set_param(pid, 0.0f)) return 2;
  p->params.push_back(ParamSetting{module, pid, true, 0.0f, std::vector<float>(values, values + n)});
  for (size_t v = 0; v < p->voices.size(); ++v)
    if (p->bank_offset + v < n) p->voices[v].modules[module]->set_param(pid, values[p->bank_offset + v]);
  return 0;
}

// The sequence table a sequencer's ui() edits (sequencer.rs:98-188, :388-470): n_steps cells for the
// grid sequencer, 8 x n_steps (row major) for the pattern sequencer; 1 <= n_steps <= 64.
int orc_set_sequence(void* h, int module, const int32_t* cells, size_t n_steps) {
  auto* p = static_cast<Patch*>(h);
  if (module < 0 || module >= (int)p->kinds.size()) return 1;
  const int kind = p->kinds[module];
  if (kind != K_GRIDSEQ && kind != K_PATSEQ) return 2;
  if (n_steps < 1 || n_steps > 64) return 3;
  const size_t n = kind == K_GRIDSEQ ? n_steps : 8 * n_steps;
  p->sequences[module] = std::vector<int32_t>(cells, cells + n);
  // like the reference's ui(): the table changes under the running module, state is kept
  for (size_t v = 0; v < p->voices.size(); ++v) p->apply_params(p->voices[v], p->bank_offset + v);
  return 0;
}

// What WaveBox::load leaves behind (sample.rs:32-69): channel-0 samples as f32, the file's sample
// rate, new = true.  Every voice shares the table; each instance rewinds at its next calc().
int orc_set_sample(void* h, int module, const float* samples, size_t n, float sample_rate) {
  auto* p = static_cast<Patch*>(h);
  if (module < 0 || module >= (int)p->kinds.size()) return 1;
  if (p->kinds[module] != K_SAMPLE) return 2;
  auto wb = std::make_shared<WaveBox>();
  wb->samples.assign(samples, samples + n);
  wb->sample_rate = sample_rate;
  p->waves[module] = wb;
  for (auto& inst : p->voices) {
    auto* sm = static_cast<Sample*>(inst.modules[module].get());
    sm->wavebox = wb;
    sm->is_new = true;
  }
  return 0;
}

// glibc exp2f, exposed so the device restatement of it can be checked value by value.
float orc_exp2f(float x) { return exp2f(x); }

// all_modules order for plan_execution (a permutation of module indices); n == 0 restores creation order
void orc_set_module_order(void* h, const int* order, int n) {
  auto* p = static_cast<Patch*>(h);
  p->order.assign(order, order + n);
  p->bank_dirty = true;
}

// Runs plan_execution on a fresh instance; writes module indices in plan order and
// the cut wires as (reader, writer) pairs.  Returns 0, or 1 when there is no Output.
int orc_plan(void* h, int* out_plan, int* out_n, int* out_cuts, int* out_n_cuts) {
  auto* p = static_cast<Patch*>(h);
  Instance inst;
  // build without planning twice: build_instance plans; redo with cut recording
  p->build_instance(inst, 0);
  if (!inst.output) { *out_n = 0; if (out_n_cuts) *out_n_cuts = 0; return 1; }
  std::vector<Module*> all;
  if (p->order.empty()) for (auto& m : inst.modules) all.push_back(m.get());
  else for (int idx : p->order) all.push_back(inst.modules[idx].get());
  std::vector<std::pair<Module*, Module*>> cuts;
  std::vector<Module*> plan;
  plan_execution(inst.output, all, plan, &cuts);
  *out_n = (int)plan.size();
  for (size_t i = 0; i < plan.size(); ++i) out_plan[i] = plan[i]->index;
  if (out_n_cuts) {
    *out_n_cuts = (int)cuts.size();
    for (size_t i = 0; i < cuts.size(); ++i) {
      out_cuts[2 * i] = cuts[i].first->index;
      out_cuts[2 * i + 1] = cuts[i].second->index;
    }
  }
  return 0;
}

void orc_reset(void* h) {
  auto* p = static_cast<Patch*>(h);
  for (auto& inst : p->voices) {
    for (auto& m : inst.modules) m->reset();
    for (const auto& kv : p->states) inst.modules[kv.first]->set_state(kv.second.data(), kv.second.size());  // back to the loaded state
  }
}

// The DSP state a deserialized module carries (enum_to_sharedsynthmodule, synth.rs:325-348, keeps it), given as
// the device's per-voice state words; every voice starts from it.
int orc_set_state(void* h, int module, const uint32_t* words, size_t n) {
  auto* p = static_cast<Patch*>(h);
  if (module < 0 || module >= (int)p->kinds.size()) return 1;
  auto probe = make_module(p->kinds[module], p->cfg);
  if (!probe->set_state(words, n)) return 2;
  p->states[module] = std::vector<uint32_t>(words, words + n);
  p->bank_dirty = true;
  return 0;
}

// Render n_samples for voices [voice_offset, voice_offset + n_voices).  Runs
// ceil(n_samples / B) whole blocks per voice (the reference cannot stop inside
// a block); only the first n_samples are reported.  State persists across
// calls (at block granularity).
//   stems: [channels][n_samples][n_voices] f32, may be NULL
//   mix:   [channels][n_samples] f64 (sum over the rendered voices in voice order), may be NULL
int orc_render(void* h, size_t n_voices, size_t voice_offset, size_t n_samples, float* stems, double* mix,
               int n_threads) {
  auto* p = static_cast<Patch*>(h);
  if (p->bank_dirty || p->voices.size() != n_voices || p->bank_offset != voice_offset) {
    for (const auto& ps : p->params)
      if (ps.per_voice && ps.values.size() < voice_offset + n_voices) return 4;
    p->voices.clear();
    p->voices.resize(n_voices);
    p->bank_offset = voice_offset;
    for (size_t v = 0; v < n_voices; ++v) p->build_instance(p->voices[v], voice_offset + v);
    p->bank_dirty = false;
  }
  if (n_voices == 0 || n_samples == 0) return 0;
  if (!p->voices[0].output) return 1;
  const size_t B = p->cfg.buffer_size;
  const size_t C = p->cfg.channels;
  const size_t n_blocks = (n_samples + B - 1) / B;
  if (n_threads < 1) n_threads = 1;
  if ((size_t)n_threads > n_voices) n_threads = (int)n_voices;
  std::vector<std::vector<double>> partial(n_threads);
  auto work = [&](int t) {
    size_t v0 = n_voices * t / n_threads, v1 = n_voices * (t + 1) / n_threads;
    if (mix) partial[t].assign(C * n_samples, 0.0);
    for (size_t v = v0; v < v1; ++v) {
      Instance& inst = p->voices[v];
      for (size_t blk = 0; blk < n_blocks; ++blk) {
        execute(inst.plan);
        size_t n0 = blk * B, cnt = std::min(B, n_samples - n0);
        for (size_t c = 0; c < C; ++c) {
          const float* buf = inst.output->outs[c].data();  // OutputModule.bufs[c], main.rs:66-75
          if (stems)
            for (size_t i = 0; i < cnt; ++i) stems[(c * n_samples + n0 + i) * n_voices + v] = buf[i];
          if (mix)
            for (size_t i = 0; i < cnt; ++i) partial[t][c * n_samples + n0 + i] += (double)buf[i];
        }
      }
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  if (mix) {
    std::fill(mix, mix + C * n_samples, 0.0);
    for (int t = 0; t < n_threads; ++t)
      for (size_t i = 0; i < C * n_samples; ++i) mix[i] += partial[t][i];
  }
  return 0;
}

// --- debugging hooks used by the known-answer tests -------------------------
// Run one module's calc() on one voice of the bank (bank is built by a
// zero-sample render first).
int orc_debug_prepare(void* h, size_t n_voices) { return orc_render(h, n_voices, 0, 0, nullptr, nullptr, 1); }

int orc_debug_calc(void* h, size_t voice, int module) {
  auto* p = static_cast<Patch*>(h);
  if (voice >= p->voices.size() || module < 0 || module >= (int)p->kinds.size()) return 1;
  p->voices[voice].modules[module]->calc();
  return 0;
}

int orc_debug_execute(void* h, size_t voice) {
  auto* p = static_cast<Patch*>(h);
  if (voice >= p->voices.size()) return 1;
  execute(p->voices[voice].plan);
  return 0;
}

int orc_debug_output(void* h, size_t voice, int module, int port, float* out) {
  auto* p = static_cast<Patch*>(h);
  if (voice >= p->voices.size() || module < 0 || module >= (int)p->kinds.size()) return 1;
  Module* m = p->voices[voice].modules[module].get();
  if (port < 0 || port >= (int)m->outs.size()) return 2;
  std::memcpy(out, m->outs[port].data(), m->outs[port].size() * sizeof(float));
  return 0;
}

// Test hook: overwrite one output buffer of one module of one voice (the next calc() of a module wired to that port
// then reads arbitrary input, NaN and infinities included; tests/test_oracle_independent.py).
int orc_debug_set_output(void* h, size_t voice, int module, int port, const float* data) {
  auto* p = static_cast<Patch*>(h);
  if (voice >= p->voices.size() || module < 0 || module >= (int)p->kinds.size()) return 1;
  Module* m = p->voices[voice].modules[module].get();
  if (port < 0 || port >= (int)m->outs.size()) return 2;
  std::memcpy(m->outs[port].data(), data, m->outs[port].size() * sizeof(float));
  return 0;
}

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox4x32_10(c, key[0], key[1]);
  std::memcpy(out, c, sizeof(c));
}

const char* orc_version() { return "srack-oracle 1 (restates sharph/s-rack @20e549b src/synth.rs + src/synth/*.rs)"; }

}  // extern "C"
