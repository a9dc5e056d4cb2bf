// Patch compiler (host): planned module graph -> flat device program.
//
// Wire semantics follow block execution in the reference (src/synth.rs:97-101):
// a reader placed AFTER its source in the plan sees the source's samples of the
// same block (zero delay -> a plain wire); a reader placed BEFORE its source
// (only possible across a wire the cycle breaker removed, synth.rs:168-192) sees
// the source's previous block, i.e. exactly buffer_size samples of delay with
// zero initial history (synth.rs:32) -> a per-voice ring in HBM.
#include "program.hpp"

#include <algorithm>
#include <map>

#include "patch.hpp"

namespace srk {

int compile_program(const srk_patch& patch, Program& prog, std::string& err) {
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

  // ---- instruction emission ------------------------------------------------
  struct Pending { Instr ins; int in_vw[4]; int out_vw[3]; };
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
  auto alloc_state = [&](int words, std::initializer_list<uint32_t> init) {
    uint16_t first = (uint16_t)prog.state_init.size();
    std::vector<uint32_t> v(init);
    v.resize(words, 0u);
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
    Pending p = blank();
    for (size_t i = 0; i < mod->inputs.size() && i < 4; ++i) p.in_vw[i] = in_wire[m][i];
    for (int port = 0; port < mod->n_outputs(); ++port) {
      auto it = port_wire.find({m, port});
      if (it != port_wire.end()) p.out_vw[port] = it->second;
    }
    switch (mod->kind) {
      case SRK_KIND_OSCILLATOR:
        p.ins.op = OP_OSC;
        p.ins.state = alloc_state(kStateOsc, {0u, 0u, 1u});  // pos = 0.0, sync detector last = true
        p.ins.param = alloc_params(m, {SRK_OSC_VAL, -1, -2, SRK_OSC_ANTIALIASING});
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
        p.ins.param = alloc_params(m, {SRK_VCA_NEGATIVE});
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
      case SRK_KIND_OUTPUT: {
        if (mod != output) continue;  // only the first Output's bufs are ever read (ui.rs:84-96, main.rs:66)
        int prev = -2;
        for (size_t c = 0; c < mod->inputs.size(); ++c) {
          Pending q = blank();
          q.ins.op = OP_OUTPUT;
          q.ins.aux = (uint16_t)c;
          q.in_vw[0] = in_wire[m][c];
          if (c > 0 && q.in_vw[0] >= 0 && q.in_vw[0] == prev) q.ins.flags = F_OUT_SAME_AS_PREV;
          prev = q.in_vw[0];
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

  // ---- liveness + physical slot assignment (no in-place reuse inside one instr)
  for (size_t i = 0; i < code.size(); ++i) {
    for (int k = 0; k < 3; ++k)
      if (code[i].out_vw[k] >= 0) vw[code[i].out_vw[k]].def = (int)i;
    for (int k = 0; k < 4; ++k)
      if (code[i].in_vw[k] >= 0) vw[code[i].in_vw[k]].last_use = std::max(vw[code[i].in_vw[k]].last_use, (int)i);
  }
  std::vector<int> free_slots;
  int n_slots = 0;
  for (size_t i = 0; i < code.size(); ++i) {
    for (int k = 0; k < 3; ++k) {
      int w = code[i].out_vw[k];
      if (w < 0) continue;
      if (vw[w].last_use < 0) { code[i].out_vw[k] = -1; continue; }  // nobody reads it
      if (free_slots.empty()) vw[w].slot = n_slots++;
      else { vw[w].slot = free_slots.back(); free_slots.pop_back(); }
    }
    for (int k = 0; k < 4; ++k) {
      int w = code[i].in_vw[k];
      if (w >= 0 && vw[w].last_use == (int)i && vw[w].slot >= 0) {
        // an instr may list the same wire twice (e.g. VCA audio == cv): free once
        bool dup = false;
        for (int j = 0; j < k; ++j) dup |= code[i].in_vw[j] == w;
        if (!dup) free_slots.push_back(vw[w].slot);
      }
    }
  }
  prog.n_wires = (uint32_t)n_slots;
  for (auto& p : code) {
    for (int k = 0; k < 4; ++k) p.ins.in[k] = p.in_vw[k] >= 0 ? (int16_t)vw[p.in_vw[k]].slot : (int16_t)-1;
    for (int k = 0; k < 3; ++k) p.ins.out[k] = p.out_vw[k] >= 0 ? (int16_t)vw[p.out_vw[k]].slot : (int16_t)-1;
    prog.code.push_back(p.ins);
  }
  Pending end = blank();
  prog.code.push_back(end.ins);
  if (prog.code.size() > 4096 || prog.state_init.size() > 60000 || prog.param_src.size() > 60000) {
    err = "patch too large";
    return SRK_ERR_LIMIT;
  }
  return SRK_OK;
}

}  // namespace srk
