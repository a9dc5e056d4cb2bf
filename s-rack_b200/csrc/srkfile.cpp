// .srk patch files: the reference's FileFormat (src/ui.rs:578-586) in MessagePack, as written by
// `container.serialize(&mut rmp_serde::Serializer::new(&mut buf))` (ui.rs:112) and read back by
// SynthModuleWorkspaceImpl::deserialize (ui.rs:115-134).  Host only.
//
// The byte layout is rmp-serde 1.3.0's (Cargo.toml:31; the crate is not under /root/reference):
// structs are arrays of their non-skipped fields in declaration order, newtype structs and
// Arc / RwLock / Mutex / Box are transparent, Option is nil or the value, an enum variant carrying
// data is a one-entry map {variant name: data}, a unit variant is its name as a string, tuples and
// fixed arrays are arrays, f32 is 0xca, f64 0xcb, integers take the shortest encoding.  The reader is
// lenient where other rmp-serde versions / configurations differ: a struct may also be a map keyed by
// field name, a variant may also be keyed by its index.
//
//   FileFormat { modules: Vec<SynthModuleType>, connections: Vec<(src_id, src_port, sink_id, sink_port)>,
//                positions: Vec<(id, (x, y))> }
// Module structs: src/synth/{output.rs:7, oscillator.rs:10,309, sequencer.rs:13,337,628, adsr.rs:8,
// vca.rs:7, filter.rs:12,49,252, mixer.rs:7, sample.rs:16,73, math.rs:7,14,177, freeverb.rs:8};
// the enum: src/synth.rs:300-317.
//
// What a load keeps: the module list (in the order the reference ends up with: unpack_modules pops
// from the back, ui.rs:652-660, so the list is the file's REVERSED), ids, parameters, sequencer
// tables, the Sample module's WaveBox, the connections (applied back to front like
// unpack_connections, ui.rs:662-681: unknown ids and bad sink ports are skipped silently), the DSP
// state the modules were saved with (oscillator phase, filter memory, envelope stage, step counters,
// play position, detectors: every voice starts from it, as the reference's deserialized modules do)
// and the GUI positions (kept only to be written back).  What it drops: the serialized port buffers
// (they only matter as the first block of history of a wire the cycle breaker cut).
#include "srkfile.hpp"

#include <algorithm>
#include <cstring>
#include <map>

namespace srk {

namespace {

// ---------------------------------------------------------------- MessagePack reader
struct Val {
  enum Type { NIL, BOOL, INT, FLOAT, STR, ARR, FARR, MAP } type = NIL;
  bool b = false;
  int64_t i = 0;
  double f = 0.0;
  bool is_f32 = false;
  std::string s;
  std::vector<Val> arr;                      // ARR; MAP: key, value, key, value ...
  std::vector<float> farr;                   // FARR: an array whose elements are all f32 (port buffers, wave tables)
  size_t size() const { return type == FARR ? farr.size() : type == ARR ? arr.size() : 0; }
  bool is_seq() const { return type == ARR || type == FARR; }
  double num() const { return type == INT ? (double)i : f; }
};

struct Reader {
  const unsigned char* p;
  const unsigned char* end;
  std::string err;
  int depth = 0;
  // Memory is bounded by the bytes actually consumed, not by the counts a header declares: containers grow element by
  // element (every element costs at least one input byte) and the tree may hold at most kMaxNodes values in total
  // (a Val is ~112 bytes: 4M nodes = 450 MB is far above any real patch; f32 arrays -- port buffers, wave tables -- do not
  // count, they are stored flat).
  static constexpr size_t kMaxNodes = 4u << 20;
  size_t nodes = 0;

  bool need(size_t n) {
    if ((size_t)(end - p) < n) { err = "truncated MessagePack"; return false; }
    return true;
  }
  uint64_t be(int n) {
    uint64_t v = 0;
    for (int k = 0; k < n; ++k) v = (v << 8) | *p++;
    return v;
  }
  bool read_str(size_t n, Val& v) {
    if (!need(n)) return false;
    v.type = Val::STR;
    v.s.assign(reinterpret_cast<const char*>(p), n);
    p += n;
    return true;
  }
  bool read_arr(size_t n, Val& v) {
    if (n > (size_t)(end - p)) { err = "array longer than the file"; return false; }
    if (n > 0 && *p == 0xca && n * 5 <= (size_t)(end - p)) {  // all-f32 fast path
      bool all = true;
      for (size_t k = 0; k < n; ++k)
        if (p[k * 5] != 0xca) { all = false; break; }
      if (all) {
        v.type = Val::FARR;
        v.farr.resize(n);
        for (size_t k = 0; k < n; ++k) {
          ++p;
          const uint32_t u = (uint32_t)be(4);
          std::memcpy(&v.farr[k], &u, 4);
        }
        return true;
      }
    }
    v.type = Val::ARR;
    return read_items(n, v);
  }
  bool read_items(size_t n, Val& v) {
    v.arr.reserve(std::min<size_t>(n, 64));
    for (size_t k = 0; k < n; ++k) {
      v.arr.emplace_back();
      if (!read(v.arr.back())) return false;
    }
    return true;
  }
  bool read_map(size_t n, Val& v) {
    if (n > (size_t)(end - p) / 2) { err = "map longer than the file"; return false; }
    v.type = Val::MAP;
    return read_items(2 * n, v);
  }
  bool read(Val& v) {
    if (++nodes > kMaxNodes) { err = "MessagePack document has too many values"; return false; }
    if (++depth > 64) { err = "MessagePack nested too deep"; return false; }
    const bool ok = read1(v);
    --depth;
    return ok;
  }
  bool read1(Val& v) {
    if (!need(1)) return false;
    const unsigned c = *p++;
    if (c <= 0x7f) { v.type = Val::INT; v.i = c; return true; }
    if (c >= 0xe0) { v.type = Val::INT; v.i = (int8_t)c; return true; }
    if (c >= 0xa0 && c <= 0xbf) return read_str(c & 0x1f, v);
    if (c >= 0x90 && c <= 0x9f) return read_arr(c & 0x0f, v);
    if (c >= 0x80 && c <= 0x8f) return read_map(c & 0x0f, v);
    switch (c) {
      case 0xc0: v.type = Val::NIL; return true;
      case 0xc2: case 0xc3: v.type = Val::BOOL; v.b = c == 0xc3; return true;
      case 0xca: {
        if (!need(4)) return false;
        const uint32_t u = (uint32_t)be(4);
        float x;
        std::memcpy(&x, &u, 4);
        v.type = Val::FLOAT; v.f = x; v.is_f32 = true;
        return true;
      }
      case 0xcb: {
        if (!need(8)) return false;
        const uint64_t u = be(8);
        std::memcpy(&v.f, &u, 8);
        v.type = Val::FLOAT;
        return true;
      }
      case 0xcc: if (!need(1)) return false; v.type = Val::INT; v.i = (int64_t)be(1); return true;
      case 0xcd: if (!need(2)) return false; v.type = Val::INT; v.i = (int64_t)be(2); return true;
      case 0xce: if (!need(4)) return false; v.type = Val::INT; v.i = (int64_t)be(4); return true;
      case 0xcf: if (!need(8)) return false; v.type = Val::INT; v.i = (int64_t)be(8); return true;
      case 0xd0: if (!need(1)) return false; v.type = Val::INT; v.i = (int8_t)be(1); return true;
      case 0xd1: if (!need(2)) return false; v.type = Val::INT; v.i = (int16_t)be(2); return true;
      case 0xd2: if (!need(4)) return false; v.type = Val::INT; v.i = (int32_t)be(4); return true;
      case 0xd3: if (!need(8)) return false; v.type = Val::INT; v.i = (int64_t)be(8); return true;
      case 0xd9: if (!need(1)) return false; return read_str((size_t)be(1), v);
      case 0xda: if (!need(2)) return false; return read_str((size_t)be(2), v);
      case 0xdb: if (!need(4)) return false; return read_str((size_t)be(4), v);
      case 0xc4: if (!need(1)) return false; return read_str((size_t)be(1), v);  // bin: treated as bytes-in-a-string
      case 0xc5: if (!need(2)) return false; return read_str((size_t)be(2), v);
      case 0xc6: if (!need(4)) return false; return read_str((size_t)be(4), v);
      case 0xdc: if (!need(2)) return false; return read_arr((size_t)be(2), v);
      case 0xdd: if (!need(4)) return false; return read_arr((size_t)be(4), v);
      case 0xde: if (!need(2)) return false; return read_map((size_t)be(2), v);
      case 0xdf: if (!need(4)) return false; return read_map((size_t)be(4), v);
    }
    err = "unsupported MessagePack type byte";
    return false;
  }
};

// A struct's field: by position when the struct is an array, by name when it is a map.
const Val* field(const Val& st, size_t idx, const char* name) {
  if (st.type == Val::ARR) return idx < st.arr.size() ? &st.arr[idx] : nullptr;
  if (st.type == Val::MAP)
    for (size_t k = 0; k + 1 < st.arr.size(); k += 2)
      if (st.arr[k].type == Val::STR && st.arr[k].s == name) return &st.arr[k + 1];
  return nullptr;
}

bool get_num(const Val* v, double& out) {
  if (!v || (v->type != Val::INT && v->type != Val::FLOAT)) return false;
  out = v->num();
  return true;
}
bool get_bool(const Val* v, bool& out) {
  if (!v || v->type != Val::BOOL) return false;
  out = v->b;
  return true;
}
bool get_str(const Val* v, std::string& out) {
  if (!v || v->type != Val::STR) return false;
  out = v->s;
  return true;
}
// element k of a sequence as a number
bool seq_num(const Val& seq, size_t k, double& out) {
  if (seq.type == Val::FARR) { if (k >= seq.farr.size()) return false; out = seq.farr[k]; return true; }
  if (seq.type == Val::ARR) return k < seq.arr.size() && get_num(&seq.arr[k], out);
  return false;
}

// SynthModuleType, src/synth.rs:300-317, in declaration order (variant index = position)
const char* const kVariants[] = {
    "OutputModuleV0", "OscillatorModuleV0", "NoiseModuleV0", "GridSequencerModuleV0", "GridSequencerModuleV1",
    "PatternSequencerModuleV0", "ADSRModuleV0", "VCAModuleV0", "MoogFilterModuleV0", "MoogFilterModuleV1",
    "MonoMixerModuleV0", "SampleModuleV0", "MathModuleV0", "NonLinearModuleV0", "FreeverbModuleV0"};
constexpr int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));

int variant_index(const Val& key) {
  if (key.type == Val::INT) return key.i >= 0 && key.i < kNumVariants ? (int)key.i : -1;
  if (key.type == Val::STR)
    for (int k = 0; k < kNumVariants; ++k)
      if (key.s == kVariants[k]) return k;
  return -1;
}

uint32_t f32_bits(double x) {
  const float f = (float)x;
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
// TransitionDetector { last } (synth.rs:277): [bool] or {"last": bool}
bool get_detector(const Val* v, bool& last) { return v && get_bool(field(*v, 0, "last"), last); }

#define SRK_FIELD(expr, what)                                 \
  do {                                                        \
    if (!(expr)) { err = std::string("bad field: ") + what; return false; } \
  } while (0)

// InternalMoogFilterState { f, p, q, b: [f32; 5], freq, res } (filter.rs:49-56) -> the ten device words
bool decode_moog_state(const Val* sv, std::vector<uint32_t>& out) {
  if (!sv) return false;
  double f, p, q, fr, rs, b;
  const Val* bv = field(*sv, 3, "b");
  if (!get_num(field(*sv, 0, "f"), f) || !get_num(field(*sv, 1, "p"), p) || !get_num(field(*sv, 2, "q"), q) || !bv ||
      !bv->is_seq() || bv->size() != 5 || !get_num(field(*sv, 4, "freq"), fr) || !get_num(field(*sv, 5, "res"), rs))
    return false;
  out = {f32_bits(f), f32_bits(p), f32_bits(q)};
  for (size_t k = 0; k < 5; ++k) {
    if (!seq_num(*bv, k, b)) return false;
    out.push_back(f32_bits(b));
  }
  out.push_back(f32_bits(fr));
  out.push_back(f32_bits(rs));
  return true;
}

bool decode_module(int variant, const Val& st, SrkModule& m, std::string& err) {
  double x = 0;
  SRK_FIELD(st.type == Val::ARR || st.type == Val::MAP, "module struct");
  SRK_FIELD(get_str(field(st, 0, "id"), m.id), "id");
  auto num = [&](size_t idx, const char* name, float& out) {
    if (!get_num(field(st, idx, name), x)) return false;
    out = (float)x;
    return true;
  };
  switch (variant) {
    case 0:  // OutputModule { id, bufs }
      m.kind = SRK_KIND_OUTPUT;
      return true;
    case 1: {  // OscillatorModule { id, val, sample_rate, sine, square, saw, pos, antialiasing, sync_detector }
      m.kind = SRK_KIND_OSCILLATOR;
      SRK_FIELD(num(1, "val", m.param[SRK_OSC_VAL]), "Oscillator.val");
      bool aa = true;
      SRK_FIELD(get_bool(field(st, 7, "antialiasing"), aa), "Oscillator.antialiasing");
      m.param[SRK_OSC_ANTIALIASING] = aa ? 1.0f : 0.0f;
      bool last = true;
      SRK_FIELD(get_num(field(st, 6, "pos"), x) && get_detector(field(st, 8, "sync_detector"), last), "Oscillator state");
      uint64_t pb;
      std::memcpy(&pb, &x, 8);
      m.state = {(uint32_t)pb, (uint32_t)(pb >> 32), last ? 1u : 0u};
      return true;
    }
    case 2:  // NoiseModule { id, out }
      m.kind = SRK_KIND_NOISE;
      return true;
    case 3: case 4: {  // GridSequencerModule{,V0} { id, cv_out, gate_out, sync_out, sequence, octaves, steps_per_octave, ... }
      m.kind = SRK_KIND_GRID_SEQUENCER;
      const Val* seq = field(st, 4, "sequence");
      SRK_FIELD(seq && seq->is_seq() && seq->size() >= 1 && seq->size() <= SRK_SEQ_MAX_STEPS, "GridSequencer.sequence");
      SRK_FIELD(seq->type == Val::ARR, "GridSequencer.sequence");
      for (const Val& c : seq->arr) {
        if (c.type == Val::NIL) { m.sequence.push_back(SRK_SEQ_NONE); continue; }
        if (variant == 3) {  // V0: Option<u16>, becomes (v, false) (sequencer.rs:651-655)
          SRK_FIELD(get_num(&c, x) && x >= 0 && x <= 65535, "GridSequencerV0 cell");
          m.sequence.push_back(SRK_GRID_CELL((int)x, false));
        } else {             // Option<(u16, bool)>
          bool hold = false;
          SRK_FIELD(c.type == Val::ARR && c.arr.size() == 2 && get_num(&c.arr[0], x) && x >= 0 && x <= 65535 &&
                        get_bool(&c.arr[1], hold), "GridSequencer cell");
          m.sequence.push_back(SRK_GRID_CELL((int)x, hold));
        }
      }
      m.seq_steps = m.sequence.size();
      SRK_FIELD(num(6, "steps_per_octave", m.param[SRK_GRIDSEQ_STEPS_PER_OCTAVE]), "GridSequencer.steps_per_octave");
      bool l1 = true, l2 = true;
      double step = 0, lastcv = 0;
      SRK_FIELD(get_num(field(st, 7, "current_step"), step) && step >= 0 && step <= 65535 &&
                    get_detector(field(st, 8, "transition_detector"), l1) &&
                    get_detector(field(st, 9, "sync_transition_detector"), l2) && get_num(field(st, 10, "last"), lastcv),
                "GridSequencer state");
      m.state = {(uint32_t)step | (l1 ? 1u << 16 : 0u) | (l2 ? 1u << 17 : 0u), f32_bits(lastcv)};
      return true;
    }
    case 5: {  // PatternSequencerModule { id, gate_outs, sync_out, sequence: Vec<Vec<Option<bool>>>, ... }
      m.kind = SRK_KIND_PATTERN_SEQUENCER;
      const Val* rows = field(st, 3, "sequence");
      SRK_FIELD(rows && rows->type == Val::ARR && rows->arr.size() == SRK_PATTERN_ROWS, "PatternSequencer.sequence");
      const size_t steps = rows->arr[0].size();
      SRK_FIELD(steps >= 1 && steps <= SRK_SEQ_MAX_STEPS, "PatternSequencer.sequence length");
      for (const Val& row : rows->arr) {
        SRK_FIELD(row.type == Val::ARR && row.arr.size() == steps, "PatternSequencer row");
        for (const Val& c : row.arr) {
          SRK_FIELD(c.type == Val::NIL || c.type == Val::BOOL, "PatternSequencer cell");
          m.sequence.push_back(c.type == Val::NIL ? SRK_SEQ_NONE : (c.b ? 1 : 0));
        }
      }
      m.seq_steps = steps;
      bool l1 = true, l2 = true;
      double step = 0;
      SRK_FIELD(get_num(field(st, 4, "current_step"), step) && step >= 0 && step <= 65535 &&
                    get_detector(field(st, 5, "transition_detector"), l1) &&
                    get_detector(field(st, 6, "sync_transition_detector"), l2), "PatternSequencer state");
      m.state = {(uint32_t)step | (l1 ? 1u << 16 : 0u) | (l2 ? 1u << 17 : 0u)};
      return true;
    }
    case 6:  // ADSRModule { id, a_sec, d_sec, s_val, r_sec, phase, mode, r_val, from_a_val, sample_rate, ... }
      m.kind = SRK_KIND_ADSR;
      SRK_FIELD(num(1, "a_sec", m.param[SRK_ADSR_A_SEC]) && num(2, "d_sec", m.param[SRK_ADSR_D_SEC]) &&
                    num(3, "s_val", m.param[SRK_ADSR_S_VAL]) && num(4, "r_sec", m.param[SRK_ADSR_R_SEC]), "ADSR times");
      SRK_FIELD(num(9, "sample_rate", m.adsr_sample_rate), "ADSR.sample_rate");  // kept: adsr.rs:69-71 never updates it
      m.has_adsr_rate = true;
      {
        double phase = 0, r_val = 0, from_a = 0;
        bool last = true;
        const Val* mode = field(st, 6, "mode");
        int mi = -1;  // ADSRMode (adsr.rs:27-33) in declaration order == the device encoding
        if (mode && mode->type == Val::STR) {
          static const char* const names[] = {"Attack", "Decay", "Sustain", "Release", "None"};
          for (int k = 0; k < 5; ++k)
            if (mode->s == names[k]) mi = k;
        } else if (mode && mode->type == Val::INT) {
          mi = (int)mode->i;
        }
        SRK_FIELD(mi >= 0 && mi <= 4 && get_num(field(st, 5, "phase"), phase) && get_num(field(st, 7, "r_val"), r_val) &&
                      get_num(field(st, 8, "from_a_val"), from_a) && get_detector(field(st, 10, "transition_detector"), last),
                  "ADSR state");
        m.state = {f32_bits(phase), f32_bits(r_val), f32_bits(from_a), (uint32_t)mi | (last ? 1u << 8 : 0u)};
      }
      return true;
    case 7: {  // VCAModule { id, buf, negative }
      m.kind = SRK_KIND_VCA;
      bool neg = false;
      SRK_FIELD(get_bool(field(st, 2, "negative"), neg), "VCA.negative");
      m.param[SRK_VCA_NEGATIVE] = neg ? 1.0f : 0.0f;
      return true;
    }
    case 8:  // MoogFilterModuleV0 { id, buf, freq, res, exp_amt, state }
      m.kind = SRK_KIND_MOOG_FILTER;
      SRK_FIELD(num(2, "freq", m.param[SRK_MOOG_FREQ]) && num(3, "res", m.param[SRK_MOOG_RES]) &&
                    num(4, "exp_amt", m.param[SRK_MOOG_EXP_AMT]), "MoogFilterV0 parameters");
      SRK_FIELD(decode_moog_state(field(st, 5, "state"), m.state), "MoogFilterV0 state");
      return true;
    case 9:  // MoogFilterModule { id, lowpass, bandpass, highpass, freq, res, exp_amt, state }
      m.kind = SRK_KIND_MOOG_FILTER;
      SRK_FIELD(num(4, "freq", m.param[SRK_MOOG_FREQ]) && num(5, "res", m.param[SRK_MOOG_RES]) &&
                    num(6, "exp_amt", m.param[SRK_MOOG_EXP_AMT]), "MoogFilter parameters");
      SRK_FIELD(decode_moog_state(field(st, 7, "state"), m.state), "MoogFilter state");
      return true;
    case 10: {  // MonoMixerModule { id, gain: Vec<f32>, buf }
      m.kind = SRK_KIND_MONO_MIXER;
      const Val* g = field(st, 1, "gain");
      SRK_FIELD(g && g->is_seq() && g->size() == 4, "MonoMixer.gain");
      for (size_t k = 0; k < 4; ++k) {
        SRK_FIELD(seq_num(*g, k, x), "MonoMixer.gain");
        m.param[k] = (float)x;
      }
      return true;
    }
    case 11: {  // SampleModule { id, transition_detector, pos, buf, wavebox { samples, sample_rate, new }, playing, sample_rate }
      m.kind = SRK_KIND_SAMPLE;
      const Val* wb = field(st, 4, "wavebox");
      SRK_FIELD(wb && (wb->type == Val::ARR || wb->type == Val::MAP), "Sample.wavebox");
      const Val* smp = field(*wb, 0, "samples");
      SRK_FIELD(smp && smp->is_seq(), "Sample.wavebox.samples");
      if (smp->type == Val::FARR) {
        m.wave = smp->farr;
      } else {
        for (size_t k = 0; k < smp->arr.size(); ++k) {
          SRK_FIELD(seq_num(*smp, k, x), "Sample.wavebox.samples");
          m.wave.push_back((float)x);
        }
      }
      SRK_FIELD(get_num(field(*wb, 1, "sample_rate"), x), "Sample.wavebox.sample_rate");
      m.wave_rate = (float)x;
      bool is_new = false, playing = false, last = true;
      double pos = 0;
      SRK_FIELD(get_bool(field(*wb, 2, "new"), is_new) && get_detector(field(st, 1, "transition_detector"), last) &&
                    get_num(field(st, 2, "pos"), pos) && get_bool(field(st, 5, "playing"), playing), "Sample state");
      if (is_new) { pos = 0; playing = false; }  // the first calc() after the load rewinds (sample.rs:212-216)
      m.state = {f32_bits(pos), (playing ? 1u : 0u) | (last ? 2u : 0u)};
      return true;
    }
    case 12: {  // MathModule { id, buf, constant, operation }
      SRK_FIELD(num(2, "constant", m.param[SRK_MATH_CONSTANT]), "Math.constant");
      const Val* op = field(st, 3, "operation");
      int which = -1;
      if (op && op->type == Val::STR) which = op->s == "Add" ? 0 : op->s == "Subtract" ? 1 : op->s == "Multiply" ? 2 : -1;
      else if (op && op->type == Val::INT) which = (int)op->i;
      else if (op && op->type == Val::MAP && op->arr.size() == 2) which = op->arr[0].type == Val::INT ? (int)op->arr[0].i : -1;
      SRK_FIELD(which >= 0 && which <= 2, "Math.operation");
      m.kind = SRK_KIND_ADD + which;
      return true;
    }
    case 13:  // NonLinearModule { id, buf, constant }
      m.kind = SRK_KIND_NON_LINEAR;
      SRK_FIELD(num(2, "constant", m.param[SRK_MATH_CONSTANT]), "NonLinear.constant");
      return true;
    case 14:
      m.kind = -1;  // Freeverb: outside the hot path
      return true;
  }
  err = "unknown module variant";
  return false;
}

// ---------------------------------------------------------------- MessagePack writer (rmp-serde's choices)
struct Writer {
  std::vector<unsigned char>& o;
  void raw(uint64_t v, int n) { for (int k = n - 1; k >= 0; --k) o.push_back((unsigned char)(v >> (8 * k))); }
  void nil() { o.push_back(0xc0); }
  void boolean(bool b) { o.push_back(b ? 0xc3 : 0xc2); }
  void uint(uint64_t v) {  // rmp::encode::write_uint: the shortest encoding
    if (v < 128) o.push_back((unsigned char)v);
    else if (v < 256) { o.push_back(0xcc); raw(v, 1); }
    else if (v < 65536) { o.push_back(0xcd); raw(v, 2); }
    else if (v < (1ull << 32)) { o.push_back(0xce); raw(v, 4); }
    else { o.push_back(0xcf); raw(v, 8); }
  }
  void f32(float x) { uint32_t u; std::memcpy(&u, &x, 4); o.push_back(0xca); raw(u, 4); }
  void f64(double x) { uint64_t u; std::memcpy(&u, &x, 8); o.push_back(0xcb); raw(u, 8); }
  void str(const std::string& s) {
    const size_t n = s.size();
    if (n < 32) o.push_back((unsigned char)(0xa0 | n));
    else if (n < 256) { o.push_back(0xd9); raw(n, 1); }
    else if (n < 65536) { o.push_back(0xda); raw(n, 2); }
    else { o.push_back(0xdb); raw(n, 4); }
    o.insert(o.end(), s.begin(), s.end());
  }
  void array(size_t n) {
    if (n < 16) o.push_back((unsigned char)(0x90 | n));
    else if (n < 65536) { o.push_back(0xdc); raw(n, 2); }
    else { o.push_back(0xdd); raw(n, 4); }
  }
  void variant(const char* name) { o.push_back(0x81); str(name); }
  void buffer(size_t n) { array(n); for (size_t k = 0; k < n; ++k) f32(0.0f); }  // AudioBuffer::new(Some(n)): zeros
  void detector(bool last = true) { array(1); boolean(last); }                    // TransitionDetector { last }; new(): true
};

}  // namespace

bool srk_file_decode(const void* bytes, size_t n_bytes, SrkFile& out, std::string& err) {
  Reader rd{static_cast<const unsigned char*>(bytes), static_cast<const unsigned char*>(bytes) + n_bytes, "", 0};
  Val root;
  if (!rd.read(root)) { err = rd.err; return false; }
  if (root.type != Val::ARR && root.type != Val::MAP) { err = "not a FileFormat"; return false; }
  const Val* modules = field(root, 0, "modules");
  const Val* conns = field(root, 1, "connections");
  const Val* poss = field(root, 2, "positions");
  if (!modules || !modules->is_seq() || !conns || !conns->is_seq() || !poss || !poss->is_seq()) { err = "not a FileFormat"; return false; }
  out = SrkFile();
  if (modules->type == Val::ARR)
    for (const Val& mv : modules->arr) {
      if (mv.type != Val::MAP || mv.arr.size() != 2) { err = "module is not an enum variant"; return false; }
      const int variant = variant_index(mv.arr[0]);
      if (variant < 0) { err = "unknown SynthModuleType variant"; return false; }
      SrkModule m;
      m.variant = kVariants[variant];
      if (!decode_module(variant, mv.arr[1], m, err)) { err = m.variant + ": " + err; return false; }
      out.modules.push_back(std::move(m));
    }
  if (conns->type == Val::ARR)
    for (const Val& c : conns->arr) {
      SrkConnection k;
      double sp = 0, dp = 0;
      if (c.type != Val::ARR || c.arr.size() != 4 || !get_str(&c.arr[0], k.src_id) || !get_num(&c.arr[1], sp) ||
          !get_str(&c.arr[2], k.sink_id) || !get_num(&c.arr[3], dp) || sp < 0 || sp > 255 || dp < 0 || dp > 255) {
        err = "bad connection";
        return false;
      }
      k.src_port = (uint8_t)sp;
      k.sink_port = (uint8_t)dp;
      out.connections.push_back(k);
    }
  if (poss->type == Val::ARR)
    for (const Val& c : poss->arr) {
      SrkPosition k;
      double x = 0, y = 0;
      if (c.type != Val::ARR || c.arr.size() != 2 || !get_str(&c.arr[0], k.id) || !c.arr[1].is_seq() || c.arr[1].size() != 2 ||
          !seq_num(c.arr[1], 0, x) || !seq_num(c.arr[1], 1, y)) {
        err = "bad position";
        return false;
      }
      k.x = (float)x;
      k.y = (float)y;
      out.positions.push_back(k);
    }
  return true;
}

void srk_file_encode(const SrkFile& f, size_t buffer_size, uint16_t sample_rate, uint8_t channels,
                     std::vector<unsigned char>& o) {
  Writer w{o};
  auto wf = [](uint32_t u) { float x; std::memcpy(&x, &u, 4); return x; };  // state word -> f32
  const size_t B = buffer_size;
  w.array(3);
  w.array(f.modules.size());
  for (const SrkModule& m : f.modules) {
    switch (m.kind) {
      case SRK_KIND_OUTPUT:
        w.variant("OutputModuleV0"); w.array(2); w.str(m.id);
        w.array(channels); for (unsigned c = 0; c < channels; ++c) w.buffer(B);
        break;
      case SRK_KIND_OSCILLATOR:
        w.variant("OscillatorModuleV0"); w.array(9); w.str(m.id); w.f32(m.param[SRK_OSC_VAL]); w.uint(sample_rate);
        w.buffer(B); w.buffer(B); w.buffer(B);
        {
          double pos = 0.0;
          bool last = true;
          if (m.state.size() == 3) {
            const uint64_t b = (uint64_t)m.state[0] | ((uint64_t)m.state[1] << 32);
            std::memcpy(&pos, &b, 8);
            last = m.state[2] != 0;
          }
          w.f64(pos); w.boolean(m.param[SRK_OSC_ANTIALIASING] != 0.0f); w.detector(last);
        }
        break;
      case SRK_KIND_NOISE:
        w.variant("NoiseModuleV0"); w.array(2); w.str(m.id); w.buffer(B);
        break;
      case SRK_KIND_GRID_SEQUENCER:
        w.variant("GridSequencerModuleV1"); w.array(12); w.str(m.id); w.buffer(B); w.buffer(B); w.buffer(B);
        w.array(m.sequence.size());
        for (int32_t c : m.sequence) {
          if (c < 0) { w.nil(); continue; }
          w.array(2); w.uint((uint32_t)c & 0xFFFF); w.boolean((c >> 16) & 1);
        }
        w.uint(2);  // octaves (sequencer.rs:39; GUI only)
        { const float spo = m.param[SRK_GRIDSEQ_STEPS_PER_OCTAVE]; w.uint(!(spo > 0.0f) ? 0u : spo >= 65535.0f ? 65535u : (unsigned)spo); }
        if (m.state.size() == 2) {  // current_step, detectors, last, ui_dirty
          w.uint(m.state[0] & 0xFFFF); w.detector((m.state[0] >> 16) & 1); w.detector((m.state[0] >> 17) & 1); w.f32(wf(m.state[1]));
        } else {
          w.uint(0); w.detector(); w.detector(); w.f32(0.0f);
        }
        w.boolean(true);
        break;
      case SRK_KIND_PATTERN_SEQUENCER:
        w.variant("PatternSequencerModuleV0"); w.array(8); w.str(m.id);
        w.array(SRK_PATTERN_ROWS); for (int r = 0; r < SRK_PATTERN_ROWS; ++r) w.buffer(B);
        w.buffer(B);
        w.array(SRK_PATTERN_ROWS);
        for (int r = 0; r < SRK_PATTERN_ROWS; ++r) {
          w.array(m.seq_steps);
          for (size_t s = 0; s < m.seq_steps; ++s) {
            const int32_t c = m.sequence[r * m.seq_steps + s];
            if (c < 0) w.nil(); else w.boolean(c != 0);
          }
        }
        if (m.state.size() == 1) {
          w.uint(m.state[0] & 0xFFFF); w.detector((m.state[0] >> 16) & 1); w.detector((m.state[0] >> 17) & 1);
        } else {
          w.uint(0); w.detector(); w.detector();
        }
        w.boolean(true);
        break;
      case SRK_KIND_ADSR:
        w.variant("ADSRModuleV0"); w.array(13); w.str(m.id);
        for (int k = 0; k < 4; ++k) w.f32(m.param[k]);
        if (m.state.size() == 4) {  // phase, mode, r_val, from_a_val
          static const char* const modes[] = {"Attack", "Decay", "Sustain", "Release", "None"};
          w.f32(wf(m.state[0])); w.str(modes[(m.state[3] & 0xFF) <= 4 ? (m.state[3] & 0xFF) : 4]); w.f32(wf(m.state[1])); w.f32(wf(m.state[2]));
          w.f32(m.adsr_sample_rate); w.detector((m.state[3] >> 8) & 1);
        } else {
          w.f32(0.0f); w.str("None"); w.f32(0.0f); w.f32(0.0f); w.f32(m.adsr_sample_rate); w.detector();
        }
        w.buffer(B); w.boolean(true);
        break;
      case SRK_KIND_VCA:
        w.variant("VCAModuleV0"); w.array(3); w.str(m.id); w.buffer(B); w.boolean(m.param[SRK_VCA_NEGATIVE] != 0.0f);
        break;
      case SRK_KIND_MOOG_FILTER:
        w.variant("MoogFilterModuleV1"); w.array(8); w.str(m.id); w.buffer(B); w.buffer(B); w.buffer(B);
        for (int k = 0; k < 3; ++k) w.f32(m.param[k]);
        w.array(6);
        if (m.state.size() == 10) {
          w.f32(wf(m.state[0])); w.f32(wf(m.state[1])); w.f32(wf(m.state[2]));
          w.array(5); for (int k = 0; k < 5; ++k) w.f32(wf(m.state[3 + k]));
          w.f32(wf(m.state[8])); w.f32(wf(m.state[9]));
        } else {
          w.f32(0); w.f32(0); w.f32(0); w.array(5); for (int k = 0; k < 5; ++k) w.f32(0); w.f32(0); w.f32(0);
        }
        break;
      case SRK_KIND_MONO_MIXER:
        w.variant("MonoMixerModuleV0"); w.array(3); w.str(m.id);
        w.array(4); for (int k = 0; k < 4; ++k) w.f32(m.param[k]);
        w.buffer(B);
        break;
      case SRK_KIND_SAMPLE:
        {
          const bool st = m.state.size() == 2;  // pos, playing | gate_last << 1
          w.variant("SampleModuleV0"); w.array(7); w.str(m.id); w.detector(st ? (m.state[1] >> 1) & 1 : true);
          w.f32(st ? wf(m.state[0]) : 0.0f); w.buffer(B);
          w.array(3); w.array(m.wave.size()); for (float x : m.wave) w.f32(x);
          w.f32(m.wave_rate); w.boolean(!st && !m.wave.empty());  // new: a fresh table rewinds at its first calc()
          w.boolean(st ? m.state[1] & 1 : false); w.f32((float)sample_rate);
        }
        break;
      case SRK_KIND_ADD: case SRK_KIND_SUBTRACT: case SRK_KIND_MULTIPLY: {
        static const char* const names[] = {"Add", "Subtract", "Multiply"};
        w.variant("MathModuleV0"); w.array(4); w.str(m.id); w.buffer(B); w.f32(m.param[SRK_MATH_CONSTANT]);
        w.str(names[m.kind - SRK_KIND_ADD]);
        break;
      }
      case SRK_KIND_NON_LINEAR:
        w.variant("NonLinearModuleV0"); w.array(3); w.str(m.id); w.buffer(B); w.f32(m.param[SRK_MATH_CONSTANT]);
        break;
    }
  }
  w.array(f.connections.size());
  for (const SrkConnection& c : f.connections) {
    w.array(4); w.str(c.src_id); w.uint(c.src_port); w.str(c.sink_id); w.uint(c.sink_port);
  }
  w.array(f.positions.size());
  for (const SrkPosition& p : f.positions) {
    w.array(2); w.str(p.id); w.array(2); w.f32(p.x); w.f32(p.y);
  }
}

}  // namespace srk
