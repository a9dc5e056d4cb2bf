// Patch specialiser (host): one-warp program (program.cpp, plan order, wires by liveness) -> the CUDA C++
// translation unit of a fused voice kernel.  What is generated is only the wiring: one op object of
// fused_ops.cuh per instruction with its connectivity as template arguments, state / parameter word offsets as
// literals, and one call per instruction inside the sample-group body; wires become local arrays (registers).
// Parameters that are uniform over voices are read from the kernel arguments (constant bank operands), per-voice
// ones are loaded once into registers.
#include "fused.hpp"

#include <cstdio>
#include <cstring>
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

int fused_generate(const srk_patch& patch, const Program& prog, int group, int min_blocks, FusedSpec& out, std::string& err) {
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
  auto IN = [](const Instr& ins, int k) { return ins.in[k] >= 0 ? "w" + std::to_string(ins.in[k]) : std::string("nullptr"); };
  auto OUTW = [](const Instr& ins, int k) { return ins.out[k] >= 0 ? "w" + std::to_string(ins.out[k]) : std::string("nullptr"); };
  auto outs_mask = [](const Instr& ins) { return (ins.out[0] >= 0 ? 1 : 0) | (ins.out[1] >= 0 ? 2 : 0) | (ins.out[2] >= 0 ? 4 : 0); };

  // channel -> distinct output wire
  std::vector<int> chan_slot(prog.channels, -1), distinct;
  for (const Instr& ins : prog.code) {
    if (ins.op != OP_OUTPUT) continue;
    for (int j = 0; j < ins.n_ch; ++j)
      if ((size_t)(ins.aux + j) < chan_slot.size()) chan_slot[ins.aux + j] = ins.in[j];
  }
  std::vector<int> chan_wire(prog.channels, -1);
  for (size_t ch = 0; ch < chan_slot.size(); ++ch) {
    if (chan_slot[ch] < 0) continue;
    size_t d = 0;
    while (d < distinct.size() && distinct[d] != chan_slot[ch]) ++d;
    if (d == distinct.size()) distinct.push_back(chan_slot[ch]);
    chan_wire[ch] = (int)d;
  }
  if (prog.channels == 0 || prog.channels > 64) { err = "fused kernels handle 1..64 output channels"; return SRK_ERR_LIMIT; }
  const int D = std::max<int>(1, (int)distinct.size());
  if (D > 6) { err = "more than 6 distinct output wires: not fused"; return SRK_ERR_LIMIT; }
  out.n_distinct = D;
  out.smem_per_warp = (size_t)2 * D * SRK_FUSED_TILE * 32 * sizeof(float);

  std::ostringstream decl, load, body, store;
  int idx = 0;
  for (const Instr& ins : prog.code) {
    const std::string m = "m" + std::to_string(idx++);
    switch (ins.op) {
      case OP_END: case OP_MIX: break;
      case OP_RING_LOAD:
        body << "      rings.load<U>(c, " << ins.aux << "u, " << OUTW(ins, 0) << ");\n";
        break;
      case OP_RING_STORE:
        body << "      rings.store<U>(c, " << ins.aux << "u, " << IN(ins, 0) << ");\n";
        break;
      case OP_OSC:
        decl << "  Osc<" << (ins.in[0] >= 0 ? "true" : "false") << ", " << (ins.in[1] >= 0 ? "true" : "false") << ", " << outs_mask(ins)
             << ", " << ((ins.flags & F_OSC_NO_ANTIALIASING) ? "false" : "true") << "> " << m << ";\n";
        load << "  " << m << ".load(c, " << ins.state << "u, " << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2)
             << ", " << bits_f32(ins.imm) << ");\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << IN(ins, 1) << ", " << OUTW(ins, 0) << ", " << OUTW(ins, 1) << ", "
             << OUTW(ins, 2) << ");\n";
        store << "  " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_NOISE:
        decl << "  Noise " << m << ";\n";
        load << "  " << m << ".load(c, " << ins.state << "u, " << ins.aux << "u);\n";
        body << "      " << m << ".run<U>(" << OUTW(ins, 0) << ");\n";
        store << "  " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_MOOG:
        decl << "  Moog<" << (ins.in[0] >= 0 ? "true" : "false") << ", " << (ins.in[1] >= 0 ? "true" : "false") << ", " << outs_mask(ins)
             << "> " << m << ";\n";
        load << "  " << m << ".load(c, " << ins.state << "u, " << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2)
             << ");\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << IN(ins, 1) << ", " << OUTW(ins, 0) << ", " << OUTW(ins, 1) << ", "
             << OUTW(ins, 2) << ");\n";
        store << "  " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_ADSR:
        decl << "  Adsr<" << (ins.in[0] >= 0 ? "true" : "false") << "> " << m << ";\n";
        load << "  " << m << ".load(c, " << ins.state << "u, " << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2)
             << ", " << PW(ins.param + 3) << ", " << bits_f32(ins.imm) << ");\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << OUTW(ins, 0) << ");\n";
        store << "  " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_VCA:
        decl << "  Vca<" << ((ins.in[0] >= 0 && ins.in[1] >= 0) ? "true" : "false") << ", "
             << ((ins.flags & F_VCA_NEGATIVE) ? "true" : "false") << "> " << m << ";\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << IN(ins, 1) << ", " << OUTW(ins, 0) << ");\n";
        break;
      case OP_MIXER: {
        int conn = 0;
        for (int k = 0; k < 4; ++k) conn |= ins.in[k] >= 0 ? 1 << k : 0;
        decl << "  Mixer<" << conn << "> " << m << ";\n";
        load << "  " << m << ".load(" << PW(ins.param) << ", " << PW(ins.param + 1) << ", " << PW(ins.param + 2) << ", " << PW(ins.param + 3)
             << ");\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << IN(ins, 1) << ", " << IN(ins, 2) << ", " << IN(ins, 3) << ", "
             << OUTW(ins, 0) << ");\n";
        break;
      }
      case OP_MATH:
        decl << "  Math<" << (int)ins.flags << ", " << (ins.in[0] >= 0 ? "true" : "false") << ", " << (ins.in[1] >= 0 ? "true" : "false")
             << "> " << m << ";\n";
        load << "  " << m << ".load(" << PW(ins.param) << ");\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << IN(ins, 1) << ", " << OUTW(ins, 0) << ");\n";
        break;
      case OP_GRIDSEQ:
        decl << "  GridSeq " << m << ";\n";
        load << "  " << m << ".load(c, " << ins.state << "u, " << ins.aux << "u, " << (int)ins.n_ch << "u, " << bits_f32(ins.imm) << ");\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << IN(ins, 1) << ", " << OUTW(ins, 0) << ", " << OUTW(ins, 1) << ", "
             << OUTW(ins, 2) << ");\n";
        store << "  " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_PATSEQ:
        decl << "  PatSeq " << m << ";\n";
        load << "  " << m << ".load(c, " << ins.state << "u, " << ins.aux << "u, " << (int)ins.n_ch << "u, " << (int)ins.flags << "u);\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << IN(ins, 1) << ", " << OUTW(ins, 0) << ", " << OUTW(ins, 1) << ", "
             << OUTW(ins, 2) << ");\n";
        store << "  " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_SAMPLE:
        decl << "  Sample " << m << ";\n";
        load << "  " << m << ".load(c, " << ins.state << "u, " << ins.aux << "u);\n";
        body << "      " << m << ".run<U>(" << IN(ins, 0) << ", " << IN(ins, 1) << ", " << OUTW(ins, 0) << ");\n";
        store << "  " << m << ".store(c, " << ins.state << "u);\n";
        break;
      case OP_OUTPUT:
        break;  // handled below: every distinct wire goes to its tile once
      default:
        err = "instruction kind has no fused form";
        return SRK_ERR_UNSUPPORTED;
    }
  }
  for (size_t d = 0; d < distinct.size(); ++d)
    body << "      out.put<U>(c, " << d << ", row, w" << distinct[d] << ");\n";
  if (prog.n_rings) body << "      rings.advance<U>(c);\n";

  std::ostringstream wires;
  if (!prog.wires.empty()) {
    wires << "      float ";
    for (size_t s = 0; s < prog.wires.size(); ++s) wires << (s ? ", " : "") << "w" << s << "[U]";
    wires << ";\n";
  }
  auto tick = [&](int U) {
    std::ostringstream t;
    t << "    {\n      constexpr int U = " << U << ";\n" << wires.str() << body.str() << "    }\n";
    return t.str();
  };

  std::ostringstream src;
  src << "// generated by srack_b200 fused_gen.cpp -- the wiring of one patch; the DSP is fused_ops.cuh\n"
      << "#include \"fused_ops.cuh\"\n"
      << "using namespace fz;\n"
      << "extern \"C\" __global__ void __launch_bounds__(" << kFusedMaxThreads << ", " << min_blocks << ")\n"
      << "srk_fused_kernel(const __grid_constant__ SrkFusedArgs a, const __grid_constant__ SrkTensorMap tmap) {\n"
      << "  extern __shared__ __align__(1024) float srk_smem[];\n"
      << "  Ctx c;\n"
      << "  if (!ctx_init(c, &a)) return;\n"
      << decl.str() << load.str()
      << "  Rings rings;\n  rings.init(c);\n"
      << "  Out<" << D << ", " << prog.channels << "> out;\n"
      << "  out.init(c, srk_smem, threadIdx.x >> 5);\n"
      << "  const int chan_wire[" << prog.channels << "] = {";
  for (size_t ch = 0; ch < chan_wire.size(); ++ch) src << (ch ? ", " : "") << chan_wire[ch];
  src << "};\n"
      << "  const u32 N = a.n_samples;\n"
      << "#pragma unroll 1\n"
      << "  for (u32 n0 = 0; n0 < N; n0 += " << SRK_FUSED_TILE << "u) {\n"
      << "    const u32 rows = min(" << SRK_FUSED_TILE << "u, N - n0);\n"
      << "    out.begin_tile(c);\n"
      << "    u32 row = 0;\n";
  if (group > 1) {
    src << "#pragma unroll 1\n"
        << "    for (; row + " << group << "u <= rows; row += " << group << "u)\n" << tick(group);
  }
  src << "#pragma unroll 1\n"
      << "    for (; row < rows; ++row)\n" << tick(1)
      << "    out.flush(c, &tmap, chan_wire, n0, rows);\n"
      << "  }\n"
      << "  out.finish(c);\n"
      << store.str()
      << "}\n";
  out.source = src.str();
  return SRK_OK;
}

}  // namespace srk
