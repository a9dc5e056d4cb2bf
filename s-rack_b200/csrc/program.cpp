// Patch compiler (host): planned module graph -> flat, scheduled device program.
//
// Wire semantics follow block execution in the reference (src/synth.rs:97-101):
// a reader placed AFTER its source in the plan sees the source's samples of the
// same block (zero delay -> a plain wire); a reader placed BEFORE its source
// (only possible across a wire the cycle breaker removed, synth.rs:168-192) sees
// the source's previous block, i.e. exactly buffer_size samples of delay with
// zero initial history (synth.rs:32) -> a per-voice ring in HBM.
//
// Scheduling: see program.hpp.  The reference's block-based execute() only needs
// block c of a module's inputs to compute block c of its outputs, so modules of
// one voice group can run as a software pipeline over chunks, one warp each.
#include "program.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "patch.hpp"

namespace srk {

namespace {

struct Pending {
  Instr ins;
  int in_vw[4];
  int out_vw[3];
  int cost;
};

// Rough cycles per sample of a warp that runs only this instruction (measured on B200,
// profiles/r01j): used to balance warps over the 4 SM sub-partitions and to decide
// which oscillators to time-split.
int osc_phase_cost(const Pending& p) {  // the recurrence (+ 2^cv / sr when the CV is converted here)
  if (p.ins.op == OP_OSC_DELTA || p.ins.op == OP_OSC_SHAPE) return 0;
  if (p.ins.op == OP_OSC_PHASE) return 30 + (p.ins.n_ch ? 10 : 0);
  return 25 + (p.in_vw[0] >= 0 ? 100 : 0) + (p.ins.n_ch ? 10 : 0);
}
int osc_shape_cost(const Pending& p) {  // the stateless part a time-split copy shares
  if (p.ins.op == OP_OSC_DELTA) return 125;
  if (p.ins.op == OP_OSC_PHASE) return 0;
  return (p.out_vw[0] >= 0 ? 90 : 0) + (p.out_vw[1] >= 0 ? 90 : 0) + (p.out_vw[2] >= 0 ? 100 : 0);
}

int op_cost(const Pending& p) {
  switch (p.ins.op) {
    case OP_MOOG: return p.ins.flags & F_MOOG_EXT_COEF ? 95 : 115;
    case OP_MOOG_COEF: return 45;
    case OP_GRIDSEQ: case OP_PATSEQ: return 20;
    case OP_SAMPLE: return p.in_vw[1] >= 0 ? 70 : 40;
    case OP_OSC: case OP_OSC_DELTA: case OP_OSC_PHASE: case OP_OSC_SHAPE: {
      const int n = std::max(1, p.ins.flags >> 4);  // time-split copies share the shaping work
      return osc_phase_cost(p) + osc_shape_cost(p) / n;
    }
    case OP_ADSR: return 60;
    case OP_NOISE: return 25;
    case OP_OUTPUT: return 10;
    case OP_MIX: return 8;
    case OP_MIXER: return 10;
    case OP_MATH: return p.ins.flags == F_MATH_NONLIN ? 150 : 4;
    default: return 4;
  }
}

uint16_t pow2_ceil(int x) {
  uint16_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace

// `steps_per_octave` is a u16 in the reference (sequencer.rs:22); the parameter arrives as f32 through the ABI or a file:
// NaN and values below 0 become 0, values above 65535 saturate (a float -> u16 cast outside the range is undefined in C++).
static uint16_t steps_per_octave_u16(float v) {
  if (!(v > 0.0f)) return 0;
  return v >= 65535.0f ? (uint16_t)65535 : (uint16_t)v;
}

int compile_program(const srk_patch& patch, int max_warps, Program& prog, std::string& err) {
  prog = Program();
  const int n = (int)patch.modules.size();
  const srk_module* output = patch.find_output();
  if (!output) { err = "patch has no Output module"; return SRK_ERR_NO_OUTPUT; }
  std::vector<int> pos(n, -1);
  for (size_t i = 0; i < patch.plan.size(); ++i) {
    int idx = patch.index_of(patch.plan[i]);
    if (idx < 0) { err = "plan refers to a module that left the patch"; return SRK_ERR_NOT_PLANNED; }
    pos[idx] = (int)i;
  }
  for (int m = 0; m < n; ++m)
    if (pos[m] < 0) { err = "module missing from plan"; return SRK_ERR_NOT_PLANNED; }

  prog.channels = patch.cfg.channels;
  prog.ring_len = (uint32_t)patch.cfg.buffer_size;

  // ---- virtual wires -----------------------------------------------------
  struct VWire { int def = -1, last_use = -1, slot = -1; };
  std::vector<VWire> vw;
  std::map<std::pair<int, int>, int> port_wire;   // (module, port) -> vwire written by the module
  std::map<std::pair<int, int>, int> ring_of;     // (module, port) -> ring id
  std::vector<int> ring_load_wire;                // ring id -> vwire produced by its RING_LOAD
  auto wire_for_port = [&](int m, int port) {
    auto key = std::make_pair(m, port);
    auto it = port_wire.find(key);
    if (it != port_wire.end()) return it->second;
    vw.push_back(VWire());
    port_wire[key] = (int)vw.size() - 1;
    return (int)vw.size() - 1;
  };
  // Pass 1: classify every connection as direct or delayed.
  std::vector<std::vector<int>> in_wire(n);
  for (int m = 0; m < n; ++m) {
    const srk_module* mod = patch.modules[m];
    in_wire[m].assign(mod->inputs.size(), -1);
    bool live_sink = mod->kind != SRK_KIND_OUTPUT || mod == output;  // other Outputs are never read
    for (size_t i = 0; i < mod->inputs.size(); ++i) {
      const srk_module* src = mod->inputs[i].first;
      if (!src || !live_sink) continue;
      int s = patch.index_of(src);
      int port = mod->inputs[i].second;
      if (pos[s] < pos[m]) {
        in_wire[m][i] = wire_for_port(s, port);
      } else {
        auto key = std::make_pair(s, port);
        auto it = ring_of.find(key);
        if (it == ring_of.end()) {
          ring_of[key] = (int)ring_load_wire.size();
          vw.push_back(VWire());
          ring_load_wire.push_back((int)vw.size() - 1);
          wire_for_port(s, port);  // the writer must materialise the port for RING_STORE
          it = ring_of.find(key);
        }
        in_wire[m][i] = ring_load_wire[it->second];
      }
    }
  }
  prog.n_rings = (uint32_t)ring_load_wire.size();

  // ---- instruction emission (plan order) -------------------------------------
  std::vector<Pending> code;
  auto blank = [] {
    Pending p{};
    p.ins.op = OP_END;
    for (int i = 0; i < 4; ++i) { p.ins.in[i] = -1; p.in_vw[i] = -1; }
    for (int i = 0; i < 3; ++i) { p.ins.out[i] = -1; p.out_vw[i] = -1; }
    return p;
  };
  for (auto& kv : ring_of) {
    Pending p = blank();
    p.ins.op = OP_RING_LOAD;
    p.ins.aux = (uint16_t)kv.second;
    p.out_vw[0] = ring_load_wire[kv.second];
    code.push_back(p);
  }
  const srk_module* state_owner = nullptr;  // module being emitted: a loaded file's state overrides X::new()
  auto alloc_state = [&](int words, std::initializer_list<uint32_t> init) {
    uint16_t first = (uint16_t)prog.state_init.size();
    std::vector<uint32_t> v(init);
    v.resize(words, 0u);
    if (state_owner && (int)state_owner->init_state.size() == words) v = state_owner->init_state;
    prog.state_init.insert(prog.state_init.end(), v.begin(), v.end());
    return first;
  };
  auto alloc_params = [&](int module, std::initializer_list<int> pids) {
    uint16_t first = (uint16_t)prog.param_src.size();
    for (int pid : pids) prog.param_src.push_back(ParamSource{module, pid});
    return first;
  };
  for (const srk_module* mod : patch.plan) {
    const int m = patch.index_of(mod);
    state_owner = mod;
    Pending p = blank();
    for (size_t i = 0; i < mod->inputs.size() && i < 4; ++i) p.in_vw[i] = in_wire[m][i];
    for (int port = 0; port < mod->n_outputs() && port < 3; ++port) {
      auto it = port_wire.find({m, port});
      if (it != port_wire.end()) p.out_vw[port] = it->second;
    }
    switch (mod->kind) {
      case SRK_KIND_OSCILLATOR:
        p.ins.op = OP_OSC;
        p.ins.state = alloc_state(kStateOsc, {0u, 0u, 1u});  // pos = 0.0, sync detector last = true
        p.ins.param = alloc_params(m, {SRK_OSC_VAL, -1, -2});
        if (mod->param[SRK_OSC_ANTIALIASING] == 0.0f) p.ins.flags |= F_OSC_NO_ANTIALIASING;
        p.ins.imm = (float)mod->osc_sample_rate;
        break;
      case SRK_KIND_NOISE:
        p.ins.op = OP_NOISE;
        p.ins.state = alloc_state(kStateNoise, {0u, 0u});
        p.ins.aux = (uint16_t)m;
        break;
      case SRK_KIND_MOOG_FILTER:
        p.ins.op = OP_MOOG;
        p.ins.state = alloc_state(kStateMoog, {});
        p.ins.param = alloc_params(m, {SRK_MOOG_FREQ, SRK_MOOG_RES, SRK_MOOG_EXP_AMT});
        break;
      case SRK_KIND_ADSR:
        p.ins.op = OP_ADSR;
        p.ins.state = alloc_state(kStateAdsr, {0u, 0u, 0u, ADSR_NONE | (1u << 8)});  // gate detector last = true
        p.ins.param = alloc_params(m, {SRK_ADSR_A_SEC, SRK_ADSR_D_SEC, SRK_ADSR_S_VAL, SRK_ADSR_R_SEC});
        p.ins.imm = mod->adsr_sample_rate;
        break;
      case SRK_KIND_VCA:
        p.ins.op = OP_VCA;
        if (mod->param[SRK_VCA_NEGATIVE] != 0.0f) p.ins.flags |= F_VCA_NEGATIVE;
        break;
      case SRK_KIND_MONO_MIXER:
        p.ins.op = OP_MIXER;
        p.ins.param = alloc_params(m, {SRK_MIXER_GAIN0, SRK_MIXER_GAIN1, SRK_MIXER_GAIN2, SRK_MIXER_GAIN3});
        break;
      case SRK_KIND_ADD: case SRK_KIND_SUBTRACT: case SRK_KIND_MULTIPLY: case SRK_KIND_NON_LINEAR:
        p.ins.op = OP_MATH;
        p.ins.flags = (uint8_t)(mod->kind - SRK_KIND_ADD);
        p.ins.param = alloc_params(m, {SRK_MATH_CONSTANT});
        break;
      case SRK_KIND_GRID_SEQUENCER:
        p.ins.op = OP_GRIDSEQ;
        p.ins.state = alloc_state(kStateGridSeq, {(1u << 16) | (1u << 17), 0u});  // step 0, both detectors last = true
        p.ins.aux = (uint16_t)prog.tables.size();
        p.ins.n_ch = (uint8_t)mod->seq_steps;
        // `val as f32 * (1.0 / self.steps_per_octave as f32)` (sequencer.rs:233-234): the f32 reciprocal
        p.ins.imm = 1.0f / (float)steps_per_octave_u16(mod->param[SRK_GRIDSEQ_STEPS_PER_OCTAVE]);
        prog.tables.insert(prog.tables.end(), mod->sequence.begin(), mod->sequence.end());
        break;
      case SRK_KIND_SAMPLE: {
        p.ins.op = OP_SAMPLE;
        p.ins.state = alloc_state(kStateSample, {0u, 1u << 1});  // pos 0.0, not playing, gate detector last = true
        p.ins.aux = (uint16_t)prog.tables.size();
        WaveDesc d;
        d.offset = (uint32_t)prog.wave_total;
        d.len = (uint32_t)mod->wave.size();
        d.ratio = mod->wave_rate / (float)mod->osc_sample_rate;  // f32 division, as sample.rs:234
        d.is_new = mod->wave_new ? 1u : 0u;
        int32_t words[4];
        std::memcpy(words, &d, sizeof d);
        prog.tables.insert(prog.tables.end(), words, words + 4);
        prog.wave_modules.push_back(m);
        prog.wave_total += mod->wave.size();
        break;
      }
      case SRK_KIND_PATTERN_SEQUENCER: {
        // 9 output ports, 3 per instruction; every instruction keeps its own copy of the step
        // counter (identical evolution), triples nobody reads are not emitted
        const uint16_t table = (uint16_t)prog.tables.size();
        prog.tables.insert(prog.tables.end(), mod->sequence.begin(), mod->sequence.end());
        for (int first = 0; first < 9; first += 3) {
          Pending q = p;
          bool any = false;
          for (int k = 0; k < 3; ++k) {
            auto it = port_wire.find({m, first + k});
            q.out_vw[k] = it != port_wire.end() ? it->second : -1;
            any |= q.out_vw[k] >= 0;
          }
          if (!any) continue;
          q.ins.op = OP_PATSEQ;
          q.ins.flags = (uint8_t)first;
          q.ins.state = alloc_state(kStatePatSeq, {(1u << 16) | (1u << 17)});
          q.ins.aux = table;
          q.ins.n_ch = (uint8_t)mod->seq_steps;
          code.push_back(q);
        }
        for (int port = 0; port < mod->n_outputs(); ++port) {
          auto it = ring_of.find({m, port});
          if (it == ring_of.end()) continue;
          Pending st = blank();
          st.ins.op = OP_RING_STORE;
          st.ins.aux = (uint16_t)it->second;
          st.in_vw[0] = port_wire.at({m, port});
          code.push_back(st);
        }
        continue;
      }
      case SRK_KIND_OUTPUT: {
        if (mod != output) continue;  // only the first Output's bufs are ever read (ui.rs:84-96, main.rs:66)
        for (size_t c0 = 0; c0 < mod->inputs.size(); c0 += kOutputChannelsPerInstr) {
          Pending q = blank();
          q.ins.op = OP_OUTPUT;
          q.ins.aux = (uint16_t)c0;
          q.ins.n_ch = (uint8_t)std::min<size_t>(kOutputChannelsPerInstr, mod->inputs.size() - c0);
          for (int j = 0; j < q.ins.n_ch; ++j) q.in_vw[j] = in_wire[m][c0 + j];
          code.push_back(q);
          q.ins.op = OP_MIX;  // same wires, summed over the group's voices
          code.push_back(q);
        }
        continue;
      }
      default:
        err = "unsupported module kind in plan";
        return SRK_ERR_UNSUPPORTED;
    }
    code.push_back(p);
    for (int port = 0; port < mod->n_outputs(); ++port) {
      auto it = ring_of.find({m, port});
      if (it == ring_of.end()) continue;
      Pending s = blank();
      s.ins.op = OP_RING_STORE;
      s.ins.aux = (uint16_t)it->second;
      s.in_vw[0] = port_wire.at({m, port});
      code.push_back(s);
    }
  }
  // (table offsets ride in the 16-bit Instr::aux: a pattern sequencer takes 512 words)
  if (code.size() > 4000 || prog.state_init.size() > 60000 || prog.param_src.size() > 60000 || prog.tables.size() > 65535) {
    err = "patch too large";
    return SRK_ERR_LIMIT;
  }

  // ---- liveness: drop output ports nobody reads --------------------------------
  for (size_t i = 0; i < code.size(); ++i)
    for (int k = 0; k < 4; ++k)
      if (code[i].in_vw[k] >= 0) vw[code[i].in_vw[k]].last_use = std::max(vw[code[i].in_vw[k]].last_use, (int)i);
  for (size_t i = 0; i < code.size(); ++i)
    for (int k = 0; k < 3; ++k) {
      int w = code[i].out_vw[k];
      if (w >= 0 && vw[w].last_use < 0) code[i].out_vw[k] = -1;
    }
  for (const Pending& p : code) prog.sum_cost += (uint32_t)op_cost(p);  // before any splitting: one warp does it all
  const bool pipelined = max_warps > 1 && code.size() > 1;
  if (pipelined) {
    // Time-split heavy oscillators.  An oscillator without a CV input spends ~25 cycles per
    // sample on its phase recurrence and 4x that on shaping the outputs (sin, polyBLEP with an
    // f64 division); the shaping is stateless.  n copies of the instruction on n warps all run
    // the recurrence (identical state, bit for bit) and copy i shapes only chunks with
    // chunk % n == i, so the slowest pipeline stage shrinks from phase + shape to
    // phase + shape / n.  Copy 0 alone stores the state back.
    int spare = std::min(max_warps, kMaxWarps) - (int)code.size();
    // SRK_MOOG_SPLIT=1 (experiment, see below): the coefficient warps are reserved BEFORE the oscillators take
    // the spare warps, and one more warp is left unused so that the ladder keeps sub-partition 0 to itself.
    const char* split_env = std::getenv("SRK_MOOG_SPLIT");
    const bool split_coef = split_env && split_env[0] == '1';
    int reserved_coef = 0;
    if (split_coef) {
      for (const Pending& p : code)
        if (p.ins.op == OP_MOOG && p.in_vw[1] >= 0 && spare > 0) { ++reserved_coef; --spare; }
      const int isolate_cap = 1 + 3 * (std::min(max_warps, kMaxWarps) / 4);
      spare = std::min(spare, std::max(0, isolate_cap - (int)code.size() - reserved_coef));
    }
    // First take the V/oct conversion out of CV-driven oscillators: delta = 440 * 2^(cv + val) / sr
    // (oscillator.rs:43-48,132) is an f64 exp2 and an f64 division per sample, stateless, and four
    // times the cost of the phase recurrence it feeds.  It becomes an OP_OSC_DELTA instruction of its
    // own writing the f64 delta to a pair of wires; the oscillator then reads delta like a CV-less
    // one reads its constant, and both can be time-split below.
    {
      std::vector<Pending> with_delta;
      for (Pending& p : code) {
        if (p.ins.op == OP_OSC && p.in_vw[0] >= 0 && spare > 0) {
          --spare;
          Pending d = p;
          d.ins.op = OP_OSC_DELTA;
          d.in_vw[1] = -1;
          for (int j = 0; j < 3; ++j) d.out_vw[j] = -1;
          for (int j = 0; j < 2; ++j) {
            vw.push_back(VWire());
            vw.back().last_use = 0;  // read by the oscillator below
            d.out_vw[j] = (int)vw.size() - 1;
            p.in_vw[2 + j] = d.out_vw[j];
          }
          p.in_vw[0] = -1;
          p.ins.n_ch = 1;
          with_delta.push_back(d);
        }
        with_delta.push_back(p);
      }
      code.swap(with_delta);
    }
    // SRK_OSC_PHASE_SPLIT=1 (experiment): an oscillator that would be time-split becomes OP_OSC_PHASE (the
    // recurrence, once) + OP_OSC_SHAPE (stateless, time-split below) joined by a wire pair carrying the f64 phase,
    // instead of n copies that each repeat the whole recurrence.
    {
      const char* ph_env = std::getenv("SRK_OSC_PHASE_SPLIT");
      std::vector<Pending> with_phase;
      for (Pending& p : code) {
        if (ph_env && ph_env[0] == '1' && p.ins.op == OP_OSC && p.in_vw[0] < 0 && spare > 0 &&
            (p.out_vw[0] >= 0 || p.out_vw[1] >= 0 || p.out_vw[2] >= 0)) {
          --spare;
          Pending ph = p;
          ph.ins.op = OP_OSC_PHASE;
          for (int j = 0; j < 3; ++j) ph.out_vw[j] = -1;
          p.ins.op = OP_OSC_SHAPE;
          for (int j = 0; j < 2; ++j) {
            vw.push_back(VWire());
            vw.back().last_use = 0;  // read by the shaper below
            ph.out_vw[j] = (int)vw.size() - 1;
            p.in_vw[j] = ph.out_vw[j];  // (the sync input stays with the phase instruction)
          }
          with_phase.push_back(ph);
        }
        with_phase.push_back(p);
      }
      code.swap(with_phase);
    }
    std::vector<int> copies(code.size(), 1);
    const char* sc_env = std::getenv("SRK_SHAPE_COST_SCALE");  // experiment knobs
    const int shape_scale = sc_env && *sc_env ? std::max(1, std::atoi(sc_env)) : 1;
    const char* mc_env = std::getenv("SRK_MAX_COPIES");
    const int env_copies = mc_env && *mc_env ? std::min(8, std::max(1, std::atoi(mc_env))) : 4;
    for (;;) {  // double the copies of the currently slowest splittable oscillator while warps last
      int best = -1, best_cost = 60;  // per-copy cost under which splitting further is pointless
      for (size_t i = 0; i < code.size(); ++i) {
        const Pending& p = code[i];
        const bool splittable = (p.ins.op == OP_OSC && p.in_vw[0] < 0) || p.ins.op == OP_OSC_DELTA || p.ins.op == OP_OSC_SHAPE;
        // stateless pieces may go to 8 copies (K / 8 is still a multiple of the sample group); an OP_OSC copy
        // repeats the recurrence, so 4 is where that stops paying
        const int max_copies = p.ins.op == OP_OSC ? 4 : env_copies;
        if (!splittable || copies[i] >= max_copies || spare < copies[i]) continue;
        const int c = osc_phase_cost(p) + shape_scale * osc_shape_cost(p) / copies[i];
        if (c > best_cost) { best = (int)i; best_cost = c; }
      }
      if (best < 0) break;
      spare -= copies[best];  // n -> 2n adds n warps
      copies[best] *= 2;
    }
    // Split CV-driven ladder filters: the coefficient block (filter.rs:61-68) is a pure function
    // of the CV sample, ~20 of the filter's ~75 instructions per sample and off the ladder's
    // dependency chain -- it moves to its own warp and feeds (f, p, q) through three wires.
    // Measured on B200 (profiles/r01o): 4.09 ms vs 4.13 ms per step at equal chunk length -- the
    // ladder is bound by its dependency chain, not by issue slots -- and the three extra wires
    // halve the chunk that fits in shared memory, so it is off unless SRK_MOOG_SPLIT=1.
    std::vector<int> coef_of(code.size(), 0);
    for (size_t i = 0; split_coef && i < code.size() && reserved_coef > 0; ++i)
      if (code[i].ins.op == OP_MOOG && code[i].in_vw[1] >= 0) { coef_of[i] = 1; --reserved_coef; }
    std::vector<Pending> split;
    for (size_t i = 0; i < code.size(); ++i)
      for (int c = 0; c < copies[i]; ++c) {
        if (coef_of[i]) {
          Pending k = code[i];
          k.ins.op = OP_MOOG_COEF;
          for (int j = 1; j < 4; ++j) k.in_vw[j] = -1;
          k.in_vw[0] = code[i].in_vw[1];
          for (int j = 0; j < 3; ++j) {
            vw.push_back(VWire());
            vw.back().last_use = 0;  // read by the ladder below
            k.out_vw[j] = (int)vw.size() - 1;
            code[i].in_vw[1 + j] = k.out_vw[j];
          }
          code[i].ins.flags |= F_MOOG_EXT_COEF;
          split.push_back(k);
        }
        Pending q = code[i];
        if (copies[i] > 1) q.ins.flags = (uint8_t)((q.ins.flags & F_OSC_NO_ANTIALIASING) | (copies[i] << 4) | c);
        split.push_back(q);
      }
    code.swap(split);
  }
  const int nc = (int)code.size();
  for (int i = 0; i < nc; ++i)
    for (int k = 0; k < 3; ++k)
      if (code[i].out_vw[k] >= 0) vw[code[i].out_vw[k]].def = i;
  for (auto& p : code) {
    p.cost = op_cost(p);
    prog.max_cost = std::max<uint32_t>(prog.max_cost, (uint32_t)p.cost);
  }

  std::vector<int> stage(nc, 0), warp(nc, 0);
  if (!pipelined) {
    // One warp, plan order, physical tiles shared by liveness (no in-place reuse inside one instr).
    std::vector<int> free_slots;
    int n_slots = 0;
    for (int i = 0; i < nc; ++i) {
      for (int k = 0; k < 3; ++k) {
        int w = code[i].out_vw[k];
        if (w < 0) continue;
        if (free_slots.empty()) vw[w].slot = n_slots++;
        else { vw[w].slot = free_slots.back(); free_slots.pop_back(); }
      }
      for (int k = 0; k < 4; ++k) {
        int w = code[i].in_vw[k];
        if (w >= 0 && vw[w].last_use == i && vw[w].slot >= 0) {
          // an instr may list the same wire twice (e.g. VCA audio == cv): free once
          bool dup = false;
          for (int j = 0; j < k; ++j) dup |= code[i].in_vw[j] == w;
          if (!dup) free_slots.push_back(vw[w].slot);
        }
      }
    }
    prog.wires.assign(n_slots, WireDesc{0, 0});
    for (int s = 0; s < n_slots; ++s) prog.wires[s].base = (uint16_t)s;
    prog.n_tiles = (uint32_t)n_slots;
    prog.n_warps = 1;
    prog.n_stages = 1;
  } else {
    // ASAP stages, then pull every non-sink instruction as late as its readers allow
    // (shorter rings), then one ring of tiles per wire.
    for (int i = 0; i < nc; ++i)
      for (int k = 0; k < 4; ++k) {
        int w = code[i].in_vw[k];
        if (w >= 0) stage[i] = std::max(stage[i], stage[vw[w].def] + 1);
      }
    for (int i = nc - 1; i >= 0; --i) {
      int latest = -1;
      for (int k = 0; k < 3; ++k) {
        int w = code[i].out_vw[k];
        if (w < 0) continue;
        for (int j = 0; j < nc; ++j)
          for (int q = 0; q < 4; ++q)
            if (code[j].in_vw[q] == w) latest = latest < 0 ? stage[j] - 1 : std::min(latest, stage[j] - 1);
      }
      if (latest >= 0) stage[i] = std::max(stage[i], latest);
    }
    int n_slots = 0, n_tiles = 0;
    for (int i = 0; i < nc; ++i)
      for (int k = 0; k < 3; ++k) {
        int w = code[i].out_vw[k];
        if (w < 0 || vw[w].slot >= 0) continue;  // (time-split copies write the same wire)
        int max_delta = 1;
        for (int j = 0; j < nc; ++j)
          for (int q = 0; q < 4; ++q)
            if (code[j].in_vw[q] == w) max_delta = std::max(max_delta, stage[j] - stage[i]);
        vw[w].slot = n_slots++;
        uint16_t depth = pow2_ceil(max_delta + 1);
        prog.wires.push_back(WireDesc{(uint16_t)n_tiles, (uint16_t)(depth - 1)});
        n_tiles += depth;
      }
    prog.n_tiles = (uint32_t)n_tiles;
    // Warps: hardware warp w issues from SM sub-partition w % 4.  Longest-processing-time
    // packing over the 4 sub-partitions, one instruction per warp while warps last.
    std::vector<int> order(nc);
    for (int i = 0; i < nc; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return code[a].cost > code[b].cost; });
    const int cap = std::min(max_warps, kMaxWarps);
    int load[4] = {0, 0, 0, 0}, count[4] = {0, 0, 0, 0};
    std::vector<int> warp_load(cap, 0);
    int n_warps = 1;
    // The slowest instruction (the pipeline's critical stage, normally the ladder filter's
    // dependent chain) gets sub-partition 0 to itself when the other three have warps enough
    // for everything else: co-resident warps steal its issue slots.
    const bool isolate = nc >= 2 && nc - 1 <= 3 * (cap / 4);
    bool first = true;
    for (int idx : order) {
      int best = -1;
      for (int s = 0; s < 4; ++s) {
        if (s + 4 * count[s] >= cap) continue;  // no free warp left on this sub-partition
        if (isolate && !first && s == 0) continue;
        if (best < 0 || load[s] < load[best]) best = s;
      }
      first = false;
      int w;
      if (best >= 0) {
        w = best + 4 * count[best]++;
        load[best] += code[idx].cost;
      } else {  // all warps taken: add to the least loaded warp
        w = (int)(std::min_element(warp_load.begin(), warp_load.end()) - warp_load.begin());
        load[w % 4] += code[idx].cost;
      }
      warp_load[w] += code[idx].cost;
      warp[idx] = w;
      n_warps = std::max(n_warps, w + 1);
    }
    prog.n_warps = (uint32_t)n_warps;
    int max_stage = 0;
    for (int i = 0; i < nc; ++i) {
      max_stage = std::max(max_stage, stage[i]);
      if (code[i].ins.op == OP_RING_STORE) prog.max_ring_store_stage = std::max<uint32_t>(prog.max_ring_store_stage, stage[i]);
    }
    prog.n_stages = (uint32_t)max_stage + 1;
    if (max_stage > 250) { err = "patch too deep to pipeline"; return SRK_ERR_LIMIT; }
  }

  // ---- emit, sorted by (warp, plan order) ------------------------------------
  std::vector<int> emit(nc);
  for (int i = 0; i < nc; ++i) emit[i] = i;
  std::stable_sort(emit.begin(), emit.end(), [&](int a, int b) { return warp[a] < warp[b]; });
  prog.warp_begin.assign(prog.n_warps + 1, 0);
  for (int i : emit) {
    Pending& p = code[i];
    for (int k = 0; k < 4; ++k) p.ins.in[k] = p.in_vw[k] >= 0 ? (int16_t)vw[p.in_vw[k]].slot : (int16_t)-1;
    for (int k = 0; k < 3; ++k) p.ins.out[k] = p.out_vw[k] >= 0 ? (int16_t)vw[p.out_vw[k]].slot : (int16_t)-1;
    p.ins.warp = (uint8_t)warp[i];
    p.ins.stage = (uint8_t)stage[i];
    prog.warp_begin[warp[i] + 1]++;
    prog.code.push_back(p.ins);
  }
  for (uint32_t w = 0; w < prog.n_warps; ++w) prog.warp_begin[w + 1] += prog.warp_begin[w];
  Pending end = blank();
  prog.code.push_back(end.ins);
  return SRK_OK;
}

}  // namespace srk
