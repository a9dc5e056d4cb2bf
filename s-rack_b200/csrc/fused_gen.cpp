// Patch specialiser (host): one-warp program (program.cpp, plan order, wires by liveness) -> the CUDA C++
// translation unit of a fused voice kernel.  What is generated is only the wiring: one op object of
// fused_ops.cuh per instruction with its connectivity as template arguments, state / parameter word offsets as
// literals, and one call per instruction inside the sample-group body; wires become local arrays (registers).
// Parameters that are uniform over voices are read from the kernel arguments (constant bank operands), per-voice
// ones are loaded once into registers.
#include "fused.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>

namespace srk {

namespace {

std::string bits_f32(float x) {
  uint32_t b;
  std::memcpy(&b, &x, 4);
  char buf[48];
  std::snprintf(buf, sizeof buf, "__uint_as_float(0x%08xu)", b);
  return buf;
}

uint64_t fnv1a(const void* data, size_t n, uint64_t h) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
  return h;
}

}  // namespace

std::string fused_hash(const std::string& text, const std::string& salt) {
  uint64_t a = fnv1a(text.data(), text.size(), 0xcbf29ce484222325ull);
  a = fnv1a(salt.data(), salt.size(), a);
  uint64_t b = fnv1a(salt.data(), salt.size(), 0x84222325cbf29ce4ull);
  b = fnv1a(text.data(), text.size(), b ^ text.size());
  char buf[40];
  std::snprintf(buf, sizeof buf, "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
  return buf;
}

int fused_generate(const srk_patch& patch, const Program& prog, const FusedOptions& opt, FusedSpec& out, std::string& err) {
  int group = opt.group;
  const int min_blocks = opt.min_blocks;
  out = FusedSpec();
  if (prog.n_warps != 1) { err = "fused kernels are generated from the one-warp program"; return SRK_ERR_ARG; }
  const size_t P = prog.param_src.size();
  if (P > SRK_FUSED_MAX_UNIFORM) { err = "too many parameter words for a fused kernel"; return SRK_ERR_LIMIT; }
  if (prog.n_rings && prog.ring_len < (uint32_t)group) group = 1;  // a delayed sample must have been stored by an EARLIER group
  out.group = group;
  out.min_blocks = min_blocks;
  out.channels = (int)prog.channels;
  out.uniform.assign(P, 1);
  for (size_t w = 0; w < P; ++w) {
    const ParamSource& s = prog.param_src[w];
    const srk_module* m = patch.modules[s.module];
    const int pid = s.pid >= 0 ? s.pid : SRK_OSC_VAL;
    out.uniform[w] = m->param_pv[pid].empty() ? 1 : 0;
  }
  auto PW = [&](unsigned w) {
    char buf[48];
    if (out.uniform[w]) std::snprintf(buf, sizeof buf, "a.u[%u]", w);
    else std::snprintf(buf, sizeof buf, "c.ld_param(%uu)", w);
    return std::string(buf);
  };
  // ---- SSA: one local array per produced wire (the program's slots are shared by liveness in PLAN order, and the
  //      instructions are about to be reordered) --------------------------------------------------------------
  struct Node {
    Instr ins;
    int in[4], out[3];
    std::vector<int> deps;
    bool branchy = false;
    int rate = 0;
  };
  std::vector<Node> nodes;
  {
    std::vector<int> cur(prog.wires.size(), -1);  // slot -> SSA wire currently in it
    std::vector<int> producer;                    // SSA wire -> node
    std::map<int, int> ring_load_node;            // ring -> node of its RING_LOAD
    int n_ssa = 0;
    for (const Instr& ins : prog.code) {
      if (ins.op == OP_END || ins.op == OP_MIX) continue;
      Node nd;
      nd.ins = ins;
      for (int k = 0; k < 4; ++k) {
        nd.in[k] = ins.in[k] >= 0 ? cur[ins.in[k]] : -1;
        if (nd.in[k] >= 0) nd.deps.push_back(producer[nd.in[k]]);
      }
      for (int k = 0; k < 3; ++k) {
        nd.out[k] = -1;
        if (ins.out[k] >= 0) {
          nd.out[k] = n_ssa++;
          producer.push_back((int)nodes.size());
        }
      }
      for (int k = 0; k < 3; ++k)
        if (ins.out[k] >= 0) cur[ins.out[k]] = nd.out[k];
      if (ins.op == OP_RING_LOAD) ring_load_node[ins.aux] = (int)nodes.size();
      if (ins.op == OP_RING_STORE && ring_load_node.count(ins.aux)) nd.deps.push_back(ring_load_node[ins.aux]);  // the load reads the slots the store overwrites
      nodes.push_back(nd);
    }
    // Staged kernels: a CV-driven ladder filter becomes coefficients (MoogCoef: CV -> f, p, q on three wires) + ladder
    // (MoogCore), so that the stage cut can put the coefficient block on another warp than the ladder's dependent chain.
    if (opt.split_moog && opt.stages > 1) {
      const size_t n0 = nodes.size();
      for (size_t i = 0; i < n0; ++i) {
        if (nodes[i].ins.op != OP_MOOG || nodes[i].in[1] < 0 || (nodes[i].ins.flags & F_MOOG_EXT_COEF)) continue;
        Node coef;
        coef.ins = nodes[i].ins;
        coef.ins.op = OP_MOOG_COEF;
        coef.ins.flags = 0;
        coef.in[0] = nodes[i].in[1];
        coef.in[1] = coef.in[2] = coef.in[3] = -1;
        coef.deps.push_back(producer[coef.in[0]]);
        for (int k = 0; k < 3; ++k) {
          coef.out[k] = n_ssa++;
          producer.push_back((int)nodes.size());
        }
        coef.ins.in[0] = 0; coef.ins.in[1] = coef.ins.in[2] = coef.ins.in[3] = -1;  // (slot numbers are not used past this point;
        coef.ins.out[0] = coef.ins.out[1] = coef.ins.out[2] = 0;                    //  only "connected or not" is)
        Node& core = nodes[i];
        core.ins.flags |= F_MOOG_EXT_COEF;
        for (int k = 0; k < 3; ++k) { core.in[1 + k] = coef.out[k]; core.ins.in[1 + k] = 0; }
        core.deps.push_back((int)nodes.size());
        nodes.push_back(coef);
        out.split_moog = true;
      }
    }
    out.n_ssa = n_ssa;
  }
  // Oscillators with a constant delta: "audio rate" when every voice's delta is in [2^-200, 1/8) and large enough that
  // most 4-sample groups of a 32-voice warp touch a discontinuity (then the group test only costs).  The class is part
  // of the source, so a parameter edit that crosses the line selects another kernel.
  const int force_rate = [] { const char* e = std::getenv("SRK_FUSED_OSC_RATE"); return e && *e ? std::atoi(e) : -1; }();
  for (Node& nd : nodes) {
    const Instr& ins = nd.ins;
    if (ins.op == OP_OSC && ins.in[0] < 0 && ins.in[1] < 0) {
      const srk_module* m = patch.modules[prog.param_src[ins.param].module];
      const double sr = (double)m->osc_sample_rate;
      // delta = 440 * 2^val / sr is monotonic in val: the extremes of val give the extremes of delta
      float vmin = m->param[SRK_OSC_VAL], vmax = vmin;
      bool nan = vmin != vmin;
      const auto& pv = m->param_pv[SRK_OSC_VAL];
      if (!pv.empty()) {
        vmin = vmax = pv[0];
        for (float v : pv) { vmin = std::min(vmin, v); vmax = std::max(vmax, v); nan |= v != v; }
      }
      const double dmin = nan ? -1.0 : 440.0 * std::exp2((double)vmin) / sr, dmax = nan ? -1.0 : 440.0 * std::exp2((double)vmax) / sr;
      const bool all_small = dmin >= 0x1p-200 && dmax < 0.125;
      const double p_any = 1.0 - std::pow(std::max(0.0, 1.0 - (group + 1) * dmin), 32.0);  // some lane near a discontinuity
      nd.rate = all_small && p_any >= 0.7 ? 1 : 0;  // measured (profiles/r04d): below that the group test still pays
      if (force_rate >= 0) nd.rate = all_small ? force_rate : 0;
    }
    nd.branchy = (ins.op == OP_OSC && nd.rate == 0) || ins.op == OP_ADSR;
  }
  // Order: dependencies respected, instructions that still contain votes / branches (control-rate oscillators,
  // envelopes) as early as they can go, so that the branch-free ones end up next to each other: what follows the last
  // branch is one basic block, and the compiler overlaps the ladder filter's dependent chain with everything else in it.
  std::vector<int> order;
  {
    std::vector<char> done(nodes.size(), 0);
    const bool reorder = [] { const char* e = std::getenv("SRK_FUSED_REORDER"); return !(e && e[0] == '0'); }();
    while (order.size() < nodes.size()) {
      int pick = -1;
      for (int pass = 0; pass < 2 && pick < 0; ++pass)
        for (size_t i = 0; i < nodes.size() && pick < 0; ++i) {
          if (done[i]) continue;
          bool ready = true;
          for (int d : nodes[i].deps) ready &= done[d] != 0;
          if (!reorder) { pick = (int)i; break; }
          const uint8_t op = nodes[i].ins.op;
          const bool light = op == OP_MATH || op == OP_VCA || op == OP_MIXER || op == OP_MOOG_COEF;  // (nearly) stateless: next to its producers
          if (ready && (pass == 1 || nodes[i].branchy || op == OP_RING_LOAD || light)) pick = (int)i;
        }
      done[pick] = 1;
      order.push_back(pick);
    }
  }

  auto W = [](int id) { return id >= 0 ? "w" + std::to_string(id) : std::string("nullptr"); };
  auto outs_mask = [](const Instr& ins) { return (ins.out[0] >= 0 ? 1 : 0) | (ins.out[1] >= 0 ? 2 : 0) | (ins.out[2] >= 0 ? 4 : 0); };

  // channel -> distinct output wire (SSA)
  std::vector<int> chan_ssa(prog.channels, -1), distinct;
  for (const Node& nd : nodes) {
    if (nd.ins.op != OP_OUTPUT) continue;
    for (int j = 0; j < nd.ins.n_ch; ++j)
      if ((size_t)(nd.ins.aux + j) < chan_ssa.size()) chan_ssa[nd.ins.aux + j] = nd.in[j];
  }
  std::vector<int> chan_wire(prog.channels, -1);
  for (size_t ch = 0; ch < chan_ssa.size(); ++ch) {
    if (chan_ssa[ch] < 0) continue;
    size_t d = 0;
    while (d < distinct.size() && distinct[d] != chan_ssa[ch]) ++d;
    if (d == distinct.size()) distinct.push_back(chan_ssa[ch]);
    chan_wire[ch] = (int)d;
  }
  if (prog.channels == 0 || prog.channels > 64) { err = "fused kernels handle 1..64 output channels"; return SRK_ERR_LIMIT; }
  const int D = std::max<int>(1, (int)distinct.size());
  if (D > 6) { err = "more than 6 distinct output wires: not fused"; return SRK_ERR_LIMIT; }
  out.n_distinct = D;

  // ---- what every node contributes to the source ------------------------------------------------------------
  struct Piece { std::string decl, load, body, store, generic; };
  std::vector<Piece> piece(nodes.size());
  const char* T = "true";
  const char* F = "false";
  for (size_t idx = 0; idx < nodes.size(); ++idx) {
    const Node& nd = nodes[idx];
    const Instr& ins = nd.ins;
    const std::string m = "m" + std::to_string(idx);
    auto IN = [&](int k) { return W(nd.in[k]); };
    auto OUTW = [&](int k) { return W(nd.out[k]); };
    std::ostringstream decl, load, body, store;
    Piece& pc = piece[idx];
    switch (ins.op) {
      case OP_RING_LOAD:
        body << "        rings.load<U>(c, " << ins.aux << "u, " << OUTW(0) << ");\n";
        break;
      case OP_RING_STORE:
        body << "        rings.store<U>(c, " << ins.aux << "u, " << IN(0) << ");\n";
        break;
      case OP_OSC:
        decl << "      Osc<" << (ins.in[0] >= 0 ? T : F) << ", " << (ins.in[1] >= 0 ? T : F) << ", " << outs_mask(ins) << ", "
             << ((ins.flags & F_OSC_NO_ANTIALIASING) ? F : T) << ", " << nd.rate << "> " << m << ";\n";
        load << "      " << m << ".load(c, " << ins.state << "u, " << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2)
             << ", " << bits_f32(ins.imm) << ");\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << OUTW(0) << ", " << OUTW(1) << ", " << OUTW(2) << ");\n";
        store << "      " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_NOISE:
        decl << "      Noise " << m << ";\n";
        load << "      " << m << ".load(c, " << ins.state << "u, " << ins.aux << "u);\n";
        body << "        " << m << ".run<U, FAST>(" << OUTW(0) << ");\n";
        store << "      " << m << ".store(c, " << ins.state << "u);\n";
        pc.generic = " | " + m + ".needs_generic()";
        break;
      case OP_MOOG:
        if (ins.flags & F_MOOG_EXT_COEF) {  // coefficients arrive on wires in[1..3] from a MoogCoef
          decl << "      MoogCore<" << (ins.in[0] >= 0 ? T : F) << ", " << outs_mask(ins) << "> " << m << ";\n";
          load << "      " << m << ".load(c, " << ins.state << "u);\n";
          body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << IN(2) << ", " << IN(3) << ", " << OUTW(0) << ", " << OUTW(1)
               << ", " << OUTW(2) << ");\n";
          store << "      " << m << ".store(c, " << ins.state << "u);\n";
          break;
        }
        decl << "      Moog<" << (ins.in[0] >= 0 ? T : F) << ", " << (ins.in[1] >= 0 ? T : F) << ", " << outs_mask(ins) << "> " << m << ";\n";
        load << "      " << m << ".load(c, " << ins.state << "u, " << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2) << ");\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << OUTW(0) << ", " << OUTW(1) << ", " << OUTW(2) << ");\n";
        store << "      " << m << ".store(c, " << ins.state << "u);\n";
        if (ins.in[1] >= 0) pc.generic = " | " + m + ".needs_generic()";
        break;
      case OP_MOOG_COEF:
        decl << "      MoogCoef " << m << ";\n";
        load << "      " << m << ".load(c, " << ins.state << "u, " << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2) << ");\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << OUTW(0) << ", " << OUTW(1) << ", " << OUTW(2) << ");\n";
        store << "      " << m << ".store(c, " << ins.state << "u);\n";
        pc.generic = " | " + m + ".needs_generic()";
        break;
      case OP_ADSR:
        decl << "      Adsr<" << (ins.in[0] >= 0 ? T : F) << "> " << m << ";\n";
        load << "      " << m << ".load(c, " << ins.state << "u, " << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2)
             << ", " << PW(ins.param + 3) << ", " << bits_f32(ins.imm) << ");\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << OUTW(0) << ");\n";
        store << "      " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_VCA:
        decl << "      Vca<" << ((ins.in[0] >= 0 && ins.in[1] >= 0) ? T : F) << ", " << ((ins.flags & F_VCA_NEGATIVE) ? T : F) << "> " << m << ";\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << OUTW(0) << ");\n";
        break;
      case OP_MIXER: {
        int conn = 0;
        for (int k = 0; k < 4; ++k) conn |= ins.in[k] >= 0 ? 1 << k : 0;
        decl << "      Mixer<" << conn << "> " << m << ";\n";
        load << "      " << m << ".load(" << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2) << ", " << PW(ins.param + 3) << ");\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << IN(2) << ", " << IN(3) << ", " << OUTW(0) << ");\n";
        break;
      }
      case OP_MATH:
        decl << "      Math<" << (int)ins.flags << ", " << (ins.in[0] >= 0 ? T : F) << ", " << (ins.in[1] >= 0 ? T : F) << "> " << m << ";\n";
        load << "      " << m << ".load(" << PW(ins.param) << ");\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << OUTW(0) << ");\n";
        break;
      case OP_GRIDSEQ:
        decl << "      GridSeq " << m << ";\n";
        load << "      " << m << ".load(c, " << ins.state << "u, " << ins.aux << "u, " << (int)ins.n_ch << "u, " << bits_f32(ins.imm) << ");\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << OUTW(0) << ", " << OUTW(1) << ", " << OUTW(2) << ");\n";
        store << "      " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_PATSEQ:
        decl << "      PatSeq " << m << ";\n";
        load << "      " << m << ".load(c, " << ins.state << "u, " << ins.aux << "u, " << (int)ins.n_ch << "u, " << (int)ins.flags << "u);\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << OUTW(0) << ", " << OUTW(1) << ", " << OUTW(2) << ");\n";
        store << "      " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_SAMPLE:
        decl << "      Sample " << m << ";\n";
        load << "      " << m << ".load(c, " << ins.state << "u, " << ins.aux << "u);\n";
        body << "        " << m << ".run<U, FAST>(" << IN(0) << ", " << IN(1) << ", " << OUTW(0) << ");\n";
        store << "      " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_OUTPUT:
        break;  // every distinct wire goes to its tile once, in the last stage
      default:
        err = "instruction kind has no fused form";
        return SRK_ERR_UNSUPPORTED;
    }
    pc.decl = decl.str(); pc.load = load.str(); pc.body = body.str(); pc.store = store.str();
  }

  // ---- stages: the instruction order cut into S consecutive slices, one warp each ----------------------------
  // instructions per voice-sample of each op in a fused kernel (profiles/r04c per-op tables)
  auto cost_of = [&](const Node& nd) -> double {
    const Instr& ins = nd.ins;
    switch (ins.op) {
      case OP_OSC: {
        double c = ins.in[0] >= 0 ? 75.0 : 8.0;  // V/oct conversion: exp2 + division
        if (ins.in[1] >= 0) c += 6.0;
        if (ins.out[0] >= 0) c += 50.0;                                                          // sin
        const bool always = nd.rate == 1 || ins.in[0] >= 0 || ins.in[1] >= 0;
        if (!(ins.flags & F_OSC_NO_ANTIALIASING)) {
          if (ins.out[2] >= 0) c += always ? 26.0 : 14.0;                                       // saw + polyBLEP
          if (ins.out[1] >= 0) c += always ? 50.0 : 4.0;                                        // square + 2 polyBLEP
        } else {
          c += 4.0;
        }
        return c;
      }
      case OP_NOISE: return 17.0;
      case OP_MOOG: return (ins.flags & F_MOOG_EXT_COEF) ? 42.0 : ins.in[1] >= 0 ? 53.0 : 39.0;
      case OP_MOOG_COEF: return 15.0;
      case OP_ADSR: return 10.0;
      case OP_VCA: return 3.5;
      case OP_MIXER: return 6.0;
      case OP_MATH: return ins.flags == F_MATH_NONLIN ? 70.0 : 1.5;
      case OP_GRIDSEQ: case OP_PATSEQ: return 12.0;
      case OP_SAMPLE: return ins.in[1] >= 0 ? 45.0 : 15.0;
      case OP_RING_LOAD: case OP_RING_STORE: return 2.0;
      default: return 0.0;
    }
  };
  int S = std::max(1, std::min(opt.stages, 8));
  int tile = opt.tile == 16 ? 16 : 32;
  if (prog.n_rings && prog.ring_len < (uint32_t)tile) S = 1;  // a later stage's ring store must be a whole tile ahead of the load
  std::vector<int> work;  // positions in `order` that carry work
  for (size_t k = 0; k < order.size(); ++k)
    if (nodes[order[k]].ins.op != OP_OUTPUT) work.push_back((int)k);
  S = std::min<int>(S, std::max<size_t>(1, work.size()));
  std::vector<int> stage_of(nodes.size(), 0);
  out.max_stage_cost = 0.0;
  {
    // minimise the slowest stage: dynamic programme over the cut positions, for every stage count up to the one asked
    // for; more stages than it takes to get the slowest one down only add hand-overs (cfg4 @ 4096 voices: 7.6 ms with 5
    // or 6 stages, 8.7 ms with 8: profiles/r04g), so each extra stage has to buy 2 %
    const int n = (int)work.size();
    std::vector<double> pre(n + 1, 0.0);
    for (int k = 0; k < n; ++k) {
      // (a lone warp pays for every vote and branch with an instruction-fetch bubble: modules that still branch per
      // group count 1.6 times -- cfg2 @ 4096 voices: oscillator + envelope + oscillator in one stage 4.3 ms, the same
      // modules on two stages 2.9 ms, profiles/r04g, r04h)
      pre[k + 1] = pre[k] + cost_of(nodes[order[work[k]]]) * (nodes[order[work[k]]].branchy ? 1.6 : 1.0);
    }
    auto seg = [&](int a, int b) {  // cost of instructions a .. b-1 as one stage
      // (+ the tile hand-over; the last stage also owns the Output).  A lone warp retires an instruction every ~2.2
      // cycles whatever the module (dependent-instruction latency), so the stage with most instructions is the slowest:
      // the ladder filter's 80-cycle chain per sample is no floor of its own (cfg2 @ 4096 voices, profiles/r04g).
      return pre[b] - pre[a] + 4.0 + (b == n ? 8.0 : 0.0);
    };
    std::vector<std::vector<double>> best(S + 1, std::vector<double>(n + 1, 1e300));
    std::vector<std::vector<int>> cut(S + 1, std::vector<int>(n + 1, 0));
    best[0][0] = 0.0;
    for (int st = 1; st <= S; ++st)
      for (int b = st; b <= n; ++b)
        for (int a0 = st - 1; a0 < b; ++a0) {
          const double v = std::max(best[st - 1][a0], seg(a0, b));
          if (v < best[st][b]) { best[st][b] = v; cut[st][b] = a0; }
        }
    int pick = 1;
    for (int st = 2; st <= S && n >= st; ++st)
      if (best[st][n] * (1.0 + 0.02 * st) < best[pick][n] * (1.0 + 0.02 * pick)) pick = st;
    if (opt.exact_stages && n >= S) pick = S;
    S = n == 0 ? 1 : pick;
    out.max_stage_cost = n ? best[S][n] : 0.0;
    int b = n;
    for (int st = S; st >= 1; --st) {
      const int a0 = cut[st][b];
      for (int k = a0; k < b; ++k) stage_of[order[work[k]]] = st - 1;
      b = a0;
    }
  }
  for (size_t i = 0; i < nodes.size(); ++i)
    if (nodes[i].ins.op == OP_OUTPUT) stage_of[i] = S - 1;
  // wires that cross a stage boundary -> rings of tiles in shared memory, one tile more than the stages they span
  // (the producer may then run that many tiles ahead of its farthest consumer: every stage on a tile of its own)
  std::vector<int> prod_stage(out.n_ssa, 0), cross_id(out.n_ssa, -1), depth(out.n_ssa, 0), first_tile(out.n_ssa, 0);
  std::vector<std::vector<int>> cons_stages(out.n_ssa);
  for (size_t i = 0; i < nodes.size(); ++i)
    for (int k = 0; k < 3; ++k)
      if (nodes[i].out[k] >= 0) prod_stage[nodes[i].out[k]] = stage_of[i];
  int n_cross = 0;
  for (size_t i = 0; i < nodes.size(); ++i)
    for (int k = 0; k < 4; ++k) {
      const int w = nodes[i].in[k];
      if (w < 0 || stage_of[i] == prod_stage[w]) continue;
      if (stage_of[i] < prod_stage[w]) { err = "internal: stage order violates a dependency"; return SRK_ERR_LIMIT; }
      if (cross_id[w] < 0) cross_id[w] = n_cross++;
      if (std::find(cons_stages[w].begin(), cons_stages[w].end(), stage_of[i]) == cons_stages[w].end()) cons_stages[w].push_back(stage_of[i]);
    }
  std::map<int, int> ring_store_stage, ring_load_stage;
  for (size_t i = 0; i < nodes.size(); ++i) {
    if (nodes[i].ins.op == OP_RING_STORE) ring_store_stage[nodes[i].ins.aux] = stage_of[i];
    if (nodes[i].ins.op == OP_RING_LOAD) ring_load_stage[nodes[i].ins.aux] = stage_of[i];
  }
  int n_cross_tiles = 0;
  for (int w = 0; w < out.n_ssa; ++w) {
    if (cross_id[w] < 0) continue;
    int far = prod_stage[w];
    for (int cs : cons_stages[w]) far = std::max(far, cs);
    depth[w] = far - prod_stage[w] + 1;
    first_tile[w] = n_cross_tiles;
    n_cross_tiles += depth[w];
  }
  out.stages = S;
  out.tile = tile;
  out.n_cross = n_cross;
  out.n_cross_tiles = n_cross_tiles;
  const size_t tile_floats = (size_t)tile * 32;
  const size_t out_floats = 2 * D * tile_floats, cross_floats = (size_t)n_cross_tiles * tile_floats;
  out.smem_per_group = (out_floats + cross_floats + (S > 1 ? 32 : 0)) * sizeof(float);

  std::ostringstream src;
  src << "// generated by srack_b200 fused_gen.cpp -- the wiring of one patch; the DSP is fused_ops.cuh\n"
      << "#define SRK_TILE " << tile << "\n";
  // SRK_FUSED_DEFINE="NAME=VALUE,NAME2=VALUE2": macros ahead of the headers, part of the source and hence of the kernel
  // id.  (tests: SRK_SIN_TIE_BAND=268435456 sends every sine sample through the device's restatement of glibc's sin,
  // libm_glibc.cuh; experiments: SRK_SIN_COEF_SELECT=1)
  if (const char* e = std::getenv("SRK_FUSED_DEFINE")) {
    std::string item;
    for (const char* c = e;; ++c) {
      if (*c && *c != ',') { item += *c; continue; }
      const size_t eq = item.find('=');
      bool ok = !item.empty();
      for (char ch : item) ok &= std::isalnum((unsigned char)ch) || ch == '_' || ch == '=';
      if (ok) src << "#define " << item.substr(0, eq) << " " << (eq == std::string::npos ? "1" : item.substr(eq + 1)) << "\n";
      item.clear();
      if (!*c) break;
    }
  }
  src << "#include \"fused_ops.cuh\"\n"
      << "using namespace fz;\n"
      << "extern \"C\" __global__ void __launch_bounds__(" << std::max(kFusedMaxThreads, 32 * S) << ", "
      << std::max(1, min_blocks * kFusedMaxThreads / std::max(kFusedMaxThreads, 32 * S)) << ")\n"
      << "srk_fused_kernel(const __grid_constant__ SrkFusedArgs a, const __grid_constant__ SrkTensorMap tmap) {\n"
      << "  extern __shared__ __align__(1024) float srk_smem[];\n"
      << "  Ctx c;\n";
  if (S > 1) {
    // (every warp of the block reaches the barrier: all S warps of a group leave together below)
    src << "  for (u32 i = threadIdx.x; i < (blockDim.x >> 5) / " << S << "u; i += blockDim.x)\n"
        << "    for (u32 s = 0; s < " << S << "u; ++s)\n"
        << "      reinterpret_cast<volatile u32*>(srk_smem + (size_t)i * " << out.smem_per_group / 4 << "u + " << out_floats + cross_floats << "u)[s] = 0u;\n"
        << "  __syncthreads();\n";
  }
  src << "  if (!ctx_init(c, &a, " << S << "u)) return;\n"
      << "  float* const group_smem = srk_smem + (size_t)c.gib * " << out.smem_per_group / 4 << "u;\n"
      << "  const int chan_wire[" << prog.channels << "] = {";
  for (size_t ch = 0; ch < chan_wire.size(); ++ch) src << (ch ? ", " : "") << chan_wire[ch];
  src << "};\n"
      << "  const u32 N = a.n_samples;\n";
  if (S > 1) src << "  Pipe pipe;\n  pipe.init(c, group_smem, " << out_floats << "u, " << cross_floats << "u);\n  switch (c.stage) {\n";
  for (int st = 0; st < S; ++st) {
    const bool last = st == S - 1;
    std::ostringstream decl, load, body, store, generic, wires, waits;
    std::vector<int> used;  // SSA wires this stage touches
    auto use = [&](int w) { if (w >= 0 && std::find(used.begin(), used.end(), w) == used.end()) used.push_back(w); };
    bool has_ring = false;
    std::vector<int> wait_in;  // producer stages to wait for
    for (int idx : order) {
      if (stage_of[idx] != st) continue;
      const Node& nd = nodes[idx];
      for (int k = 0; k < 4; ++k) use(nd.in[k]);
      for (int k = 0; k < 3; ++k) use(nd.out[k]);
      has_ring |= nd.ins.op == OP_RING_LOAD || nd.ins.op == OP_RING_STORE;
    }
    // cross inputs first, then the ops, then cross outputs
    std::ostringstream gets, gets_first, gets_next, take_next;  // plain | prefetching: before the loop / next group / rotate
    std::vector<int> cross_in;
    for (int w : used)
      if (cross_id[w] >= 0 && prod_stage[w] != st) {
        const std::string slot = std::to_string(first_tile[w]) + "u + t % " + std::to_string(depth[w]) + "u";
        gets << "        pipe.get<U>(" << slot << ", row, w" << w << ");\n";
        gets_first << "        pipe.get<" << group << ">(" << slot << ", 0u, n" << w << ");\n";
        gets_next << "        pipe.get<U>(" << slot << ", row_next, n" << w << ");\n";
        take_next << "#pragma unroll\n        for (int j = 0; j < U; ++j) w" << w << "[j] = n" << w << "[j];\n";
        cross_in.push_back(w);
        if (std::find(wait_in.begin(), wait_in.end(), prod_stage[w]) == wait_in.end()) wait_in.push_back(prod_stage[w]);
      }
    for (int idx : order) {
      if (stage_of[idx] != st) continue;
      decl << piece[idx].decl; load << piece[idx].load; body << piece[idx].body; store << piece[idx].store; generic << piece[idx].generic;
      const Node& nd = nodes[idx];
      if (nd.ins.op == OP_RING_LOAD && ring_store_stage.count(nd.ins.aux) && ring_store_stage[nd.ins.aux] != st)
        waits << "        pipe.wait_ring(" << ring_store_stage[nd.ins.aux] << "u, t, a.B);\n";
    }
    for (int w : used)
      if (cross_id[w] >= 0 && prod_stage[w] == st) {
        body << "        pipe.put<U>(" << first_tile[w] << "u + t % " << depth[w] << "u, row, w" << w << ");\n";
        for (int cs : cons_stages[w])
          waits << "        pipe.wait_ge(" << cs << "u, (int)t + 1 - " << depth[w] << ");\n";
      }
    if (last)
      for (size_t d = 0; d < distinct.size(); ++d) body << "        out.put<U>(c, " << d << ", row, w" << distinct[d] << ");\n";
    if (has_ring) body << "        rings.advance<U>(c);\n";
    for (int ps : wait_in) waits << "        pipe.wait_ge(" << ps << "u, (int)t + 1);\n";
    if (!used.empty()) {
      wires << "        float ";
      for (size_t k = 0; k < used.size(); ++k) wires << (k ? ", " : "") << "w" << used[k] << "[U]";
      wires << ";\n";
    }
    // Prefetch (staged kernels): the cross-stage inputs of group g + 1 are loaded at the top of group g -- ptxas does not
    // move a load across the loop's back edge, and a stage that is one dependent chain (the ladder filter) otherwise
    // waits out the shared-memory latency at the head of every group.  The last group of a tile re-reads itself.
    const bool prefetch = opt.prefetch && S > 1 && group > 1 && !cross_in.empty();
    auto tick = [&](int U, bool fast) {
      std::ostringstream t;
      t << "      {\n        constexpr int U = " << U << ";\n        constexpr bool FAST = " << (fast ? "true" : "false") << ";\n" << wires.str();
      if (fast && prefetch)
        t << take_next.str() << "        const u32 row_next = min(row + " << group << "u, rows - " << group << "u);\n" << gets_next.str();
      else
        t << gets.str();
      t << body.str() << "      }\n";
      return t.str();
    };
    if (S > 1) src << "    case " << st << ": {\n";
    else src << "    {\n";
    src << decl.str() << load.str();
    if (has_ring) src << "      Rings rings;\n      rings.init(c);\n";
    if (last) src << "      Out<" << D << ", " << prog.channels << "> out;\n      out.init(c, group_smem);\n";
    src << "#pragma unroll 1\n"
        << "      for (u32 n0 = 0, t = 0; n0 < N; n0 += " << tile << "u, ++t) {\n"
        << "        const u32 rows = min(" << tile << "u, N - n0);\n";
    if (last) src << "        out.begin_tile(c);\n";
    if (S > 1) src << waits.str();
    src << "        u32 row = 0;\n";
    if (group > 1) {
      // the FAST body assumes what holds in all but a handful of tiles (no filter still on its all-zero coefficient
      // cache, noise counter a multiple of 4); a tile where it does not goes sample by sample through the generic body
      src << "        if (!(false" << generic.str() << ")) {\n";
      if (prefetch) {
        src << "        float ";
        for (size_t k = 0; k < cross_in.size(); ++k) src << (k ? ", " : "") << "n" << cross_in[k] << "[" << group << "]";
        src << ";\n        if (rows >= " << group << "u) {\n" << gets_first.str() << "        }\n";
      }
      src << "#pragma unroll 1\n"
          << "        for (; row + " << group << "u <= rows; row += " << group << "u)\n" << tick(group, true)
          << "        }\n";
    }
    src << "#pragma unroll 1\n"
        << "        for (; row < rows; ++row)\n" << tick(1, false);
    if (last) src << "        out.flush(c, &tmap, chan_wire, n0, rows);\n";
    if (S > 1) src << "        pipe.publish(" << st << "u, t + 1u);\n";
    src << "      }\n";
    if (last) src << "      out.finish(c);\n";
    src << store.str();
    if (S > 1) src << "    } break;\n";
    else src << "    }\n";
  }
  if (S > 1) src << "    default: break;\n  }\n";
  src << "}\n";
  out.source = src.str();
  return SRK_OK;
}

}  // namespace srk
