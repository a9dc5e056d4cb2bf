// Device side of the renderer: the voice kernel (one resident lane per voice, the
// whole sample loop in-kernel), the deterministic mixdown, and the Engine that owns
// HBM state for one patch.  sm_100a only; there is no CPU path.
//
// HBM layout (all structure-of-arrays, voice index fastest => every warp access is
// one contiguous 128-byte line):
//   state   u32 [S][V]      per-voice module state, loaded to shared memory at kernel
//                           start, stored back at kernel end
//   params  u32 [P][V]      per-voice parameters (uniform ones broadcast at upload)
//   rings   f32 [R][B][V]   history of delayed (feedback) wires, B = buffer_size
//   stems   f32 [C][N][V]   optional per-voice output
//   partial f32 [G][C][N]   per-group mix partials (G = voice groups of 32), reduced in
//                           a fixed order by mix_reduce_kernel => bit-reproducible mix
// One thread block = one group of 32 voices (lane = voice) x S warps.  Shared memory:
//   program blob (instructions, wire table, per-warp ranges; staged once) |
//   state [S][32] | params [P][32] | wire tiles [n_tiles][K][32]
// Pipelined schedule (S > 1): at iteration i a warp runs each of its instructions on
// chunk i - stage; one __syncthreads per iteration publishes the tiles (program.hpp).
#include "engine.hpp"

#include <cuda_runtime.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fused.hpp"
#include "voice_args.hpp"

namespace srk {

// mix[c][n] = sum over blocks, in block order (fixed tree => run-to-run identical bits)
__global__ void mix_reduce_kernel(const float* __restrict__ partial, float* __restrict__ mix, uint32_t n_blocks,
                                  size_t cn) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cn) return;
  float s = 0.0f;
  for (uint32_t b = 0; b < n_blocks; ++b) s = __fadd_rn(s, partial[(size_t)b * cn + i]);
  mix[i] = s;
}

__global__ void state_init_kernel(uint32_t* state, const uint32_t* init, uint32_t S, uint32_t V) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)S * V) return;
  state[i] = init[i / V];
}

// ----------------------------------------------------------------------------
// Engine
// ----------------------------------------------------------------------------
int compile_program(const srk_patch& patch, int max_warps, Program& prog, std::string& err);

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}

}  // namespace

static bool tune_allowed();

struct Engine {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // call start, kernel start, kernel end, call end
  bool timed = false;
  Program prog;
  uint64_t compiled_epoch = 0, uploaded_param_epoch = 0, compiled_table_epoch = 0, uploaded_wave_epoch = 0;
  int compiled_max_warps = 0;       // schedule the program was compiled for
  size_t compiled_voices = 0;       // ... and the voice count (schedule choice and chunk length follow it)
  int chunk = 0;                    // samples per chunk (K) for the compiled program
  std::vector<uint4> blob;          // device image of the program (see RenderArgs::blob)
  size_t V = 0, voice_offset = 0;
  uint64_t n_abs = 0;  // samples rendered since reset
  uint64_t state_epoch = 0;  // how often the voice state was (re)initialised: srk_state_epoch
  bool state_valid = false;
  DevBuf d_prog, d_state, d_state_init, d_params, d_rings, d_partial, d_stems, d_mix, d_waves;
  uint32_t* h_params = nullptr;  // pinned staging
  size_t h_params_bytes = 0;
  uint64_t launches = 0;
  int smem_optin = 0, smem_sm = 0, n_sm = 0;
  // geometry of the last launch
  int block_threads = 0, step = 0, n_warps = 0, n_stages = 0;
  size_t smem_bytes = 0;
  // fused (patch-specialised) kernel of the compiled program, when that is the schedule in use
  bool fused = false, last_fused = false;
  FusedSpec fspec;
  const FusedKernel* fkernel = nullptr;
  std::string fused_note;  // why the interpreter is used instead, if it is
  std::vector<uint32_t> uniform_words;  // parameter word w for voice 0 of the range (read by fused kernels for uniform words)
  // measured schedule choice (tune_schedule): pending until a render long enough to be worth it; the options of the
  // fused kernel that won are kept for the regenerations a parameter change can cause
  bool tune_pending = false, tuned = false, have_forced = false;
  FusedOptions forced;
  std::string tune_note;
  DevBuf d_tune_state, d_tune_rings, d_tune_stems, d_tune_partial, d_tune_prog;

  ~Engine() {
    if (device >= 0) {
      cudaSetDevice(device);
      for (auto* b : {&d_prog, &d_state, &d_state_init, &d_params, &d_rings, &d_partial, &d_stems, &d_mix, &d_waves, &d_tune_state,
                      &d_tune_rings, &d_tune_stems, &d_tune_partial, &d_tune_prog})
        b->release();
      if (h_params) cudaFreeHost(h_params);
      for (auto& e : ev)
        if (e) cudaEventDestroy(e);
      if (stream) cudaStreamDestroy(stream);
    }
  }
};

void engine_destroy(Engine* e) { delete e; }

#define SRK_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      patch->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);                      \
      return SRK_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

static int engine_open(srk_patch* patch) {
  if (patch->engine) return SRK_OK;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    patch->last_error = "no CUDA device: srack_b200 has no CPU path";
    return SRK_ERR_NO_DEVICE;
  }
  int dev = patch->device;
  if (dev < 0) SRK_CUDA(cudaGetDevice(&dev));
  if (dev >= count) { patch->last_error = "device ordinal out of range"; return SRK_ERR_ARG; }
  SRK_CUDA(cudaSetDevice(dev));
  std::unique_ptr<Engine, void (*)(Engine*)> eng(new Engine(), engine_destroy);
  eng->device = dev;
  SRK_CUDA(cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking));
  for (auto& ev : eng->ev) SRK_CUDA(cudaEventCreate(&ev));
  SRK_CUDA(cudaDeviceGetAttribute(&eng->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  SRK_CUDA(cudaDeviceGetAttribute(&eng->n_sm, cudaDevAttrMultiProcessorCount, dev));
  SRK_CUDA(cudaDeviceGetAttribute(&eng->smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  patch->engine = std::move(eng);
  return SRK_OK;
}

// Device image of a compiled program: instructions, wire table, per-warp ranges.
static size_t blob_table_offset(const Program& prog) {
  const size_t head = prog.code.size() * sizeof(Instr) + prog.wires.size() * sizeof(WireDesc) +
                      prog.warp_begin.size() * sizeof(uint16_t);
  return (head + 3) / 4 * 4;
}

static void build_blob(const Program& prog, std::vector<uint4>& blob) {
  const size_t bytes = blob_table_offset(prog) + prog.tables.size() * sizeof(int32_t);
  blob.assign((bytes + 15) / 16, uint4{0, 0, 0, 0});
  unsigned char* p = reinterpret_cast<unsigned char*>(blob.data());
  std::memcpy(p, prog.code.data(), prog.code.size() * sizeof(Instr));
  p += prog.code.size() * sizeof(Instr);
  if (!prog.wires.empty()) std::memcpy(p, prog.wires.data(), prog.wires.size() * sizeof(WireDesc));
  p += prog.wires.size() * sizeof(WireDesc);
  std::memcpy(p, prog.warp_begin.data(), prog.warp_begin.size() * sizeof(uint16_t));
  if (!prog.tables.empty())
    std::memcpy(reinterpret_cast<unsigned char*>(blob.data()) + blob_table_offset(prog), prog.tables.data(),
                prog.tables.size() * sizeof(int32_t));
}

// How many warps may share one 32-voice group for V voices: with few groups per SM the
// render is bound by the dependent-instruction latency of one voice, so the modules of a
// group are spread over warps as a pipeline; with many groups the SMs are already full
// and one warp per group (no barriers, smallest tiles) is the throughput shape.
static int choose_max_warps(const Engine& e, size_t V) {
  const int forced = env_int("SRK_WARPS", 0);
  if (forced > 0) return std::min(forced, (int)kMaxWarps);
  const size_t groups = (V + kVoicesPerGroup - 1) / kVoicesPerGroup;
  const size_t per_sm = (groups + std::max(e.n_sm, 1) - 1) / std::max(e.n_sm, 1);
  // Upper bound only; pipelined_pays() decides with the compiled program's cost model.
  return per_sm <= 8 ? kMaxWarps : 1;
}

// Pipelined groups finish in (groups per SM) x (slowest stage); a one-warp group in the sum of all
// its modules, however few the groups.  Measured on B200 (profiles/r01z_crossover_sweep.txt): cfg2
// pipelines up to ~5 groups per SM (16384 voices: 14.7 ms vs 24.6 ms; 24576: 31 ms vs 27 ms), cfg3 --
// whose CV-driven oscillator is one long stage -- only at 1-2 (16384 voices: 58 ms pipelined).
static bool pipelined_pays(const Engine& e, const Program& prog, size_t V) {
  const size_t groups = (V + kVoicesPerGroup - 1) / kVoicesPerGroup;
  const size_t per_sm = std::max<size_t>((groups + std::max(e.n_sm, 1) - 1) / std::max(e.n_sm, 1), 1);
  return env_int("SRK_WARPS", 0) > 0 || 2 * per_sm * prog.max_cost <= 3 * (size_t)prog.sum_cost;
}

static size_t smem_bytes_for(const Program& prog, size_t blob_vec, int K, int groups_per_block = 1) {
  return blob_vec * 16 + ((size_t)prog.state_init.size() + prog.param_src.size() + (size_t)prog.n_tiles * K) *
                             kVoicesPerGroup * sizeof(uint32_t) * groups_per_block;
}

// One-warp schedule: how many voice groups (warps) share a thread block.  All groups of the launch
// should be resident at once and spread evenly, and the warps of one SM should sit in as few blocks
// as possible -- the per-chunk barrier keeps a block's warps in the same op bodies, which is what
// makes the instruction caches work (profiles/r03c: 33 % "no instruction" stalls with 14 independent
// one-warp blocks per SM).
static int solo_groups_wanted(const Engine& e, size_t V) {
  const size_t groups = (V + kVoicesPerGroup - 1) / kVoicesPerGroup;
  const size_t n_sm = (size_t)std::max(e.n_sm, 1);
  const size_t per_sm = std::max<size_t>((groups + n_sm - 1) / n_sm, 1);
  int G = env_int("SRK_SOLO_GROUPS", 0);
  if (G <= 0) {
    const size_t blocks_per_sm = (per_sm + kMaxWarps - 1) / kMaxWarps;
    G = (int)((per_sm + blocks_per_sm - 1) / blocks_per_sm);
  }
  return std::max(1, std::min(G, (int)kMaxWarps));
}
static int choose_solo_groups(const Engine& e, const Program& prog, size_t blob_vec, int K, size_t V) {
  int G = solo_groups_wanted(e, V);
  while (G > 1 && smem_bytes_for(prog, blob_vec, K, G) > (size_t)e.smem_optin) --G;
  return G;
}

// Samples per chunk (a power of two) for this program and render length; 0 when it cannot fit.
// Pipelined groups want long chunks: every iteration costs a block barrier plus ~1.5k cycles of
// instruction-cache refill for the code outside the sample loops (profiles/r01i_k8), and only
// pays (stages - 1) chunks of fill/drain per render.
constexpr int kMaxChunk = 128;
static int choose_chunk(const Engine& e, const Program& prog, size_t blob_vec, size_t V) {
  const bool pipelined = prog.n_warps > 1;
  // shared memory one block may take so that every group of this launch is resident at once
  const size_t groups = (V + kVoicesPerGroup - 1) / kVoicesPerGroup;
  const size_t per_sm = std::max<size_t>((groups + std::max(e.n_sm, 1) - 1) / std::max(e.n_sm, 1), 1);
  const size_t smem_cap = pipelined ? std::min<size_t>(e.smem_optin, (size_t)(e.smem_sm / per_sm) - 1024)
                                    : (size_t)e.smem_optin;
  int K = env_int("SRK_STEP", pipelined ? kMaxChunk : 16);
  if (!pipelined && env_int("SRK_STEP", 0) <= 0) {
    // One-warp schedule: the longest chunk that still lets one block hold all the groups an SM gets.  Every
    // chunk costs each op a trip through the interpreter and its state through shared memory; measured
    // (profiles/r03m): cfg2 @ 32768 voices 30.0 ms at K = 16, 19.7 ms at K = 64; cfg3 @ 65536 42.7 -> 37.8 ms
    // at K = 32 -- but a chunk that forces the groups of an SM into a second wave loses more than it gains
    // (cfg2 @ 65536: 58 ms at K = 32 with 13 groups per block against 37 ms at K = 16 with 14).
    const int G = solo_groups_wanted(e, V);
    K = kMaxChunk;
    while (K > 16 && smem_bytes_for(prog, blob_vec, K, G) > (size_t)e.smem_optin) K /= 2;
  }
  if (K < 1) K = 1;
  if (K > kMaxChunk) K = kMaxChunk;
  while (K & (K - 1)) K &= K - 1;  // power of two
  if (prog.n_rings) {
    // a delayed wire's sample n - B must have been stored in an EARLIER iteration than the one
    // that loads sample n: K <= B when one warp runs everything in order, K * (stage + 2) <= B
    // when the store runs `stage` iterations behind the load.
    const size_t B = std::max<uint32_t>(prog.ring_len, 1);
    const size_t lim = pipelined ? B / (prog.max_ring_store_stage + 2) : B;
    while (K > 1 && (size_t)K > lim) K /= 2;
    if ((size_t)K > lim) return 0;
  }
  while (K > 1 && smem_bytes_for(prog, blob_vec, K) > smem_cap) K /= 2;
  if (smem_bytes_for(prog, blob_vec, K) > smem_cap) return 0;
  if (pipelined && K < (per_sm > 2 ? 32 : 8)) return 0;  // short chunks: the barrier interval costs more than it buys
  return K;
}

// The chunk actually used for a render of n_samples: a pipelined program pays (stages - 1) chunks of
// fill and drain, kept under ~1/8 of the render.
static int chunk_for_length(const Program& prog, int K, size_t n_samples) {
  if (prog.n_warps > 1)
    while (K > 8 && (size_t)K * (prog.n_stages - 1) * 8 > std::max<size_t>(n_samples, 1)) K /= 2;
  return K;
}

// SRK_FUSED: 0 = never use the fused (patch-specialised) kernel, 1 = fail when it cannot be used, unset = use it
// whenever it can be generated and compiled, the interpreter kernels otherwise.
static int fused_mode() { return env_int("SRK_FUSED", -1); }

// Build options of the fused kernel for a launch of n_voices.  With few voice groups per SM a group is cut into
// stages (one warp each, a tile apart): the render is then bound by dependent-instruction latency and more warps per
// SM are what hides it (cfg2 @ 4096 voices: 6.7 ms as one warp per group, profiles/r04c).
static FusedOptions fused_options(const Engine& e, size_t n_voices, const FusedOptions* forced = nullptr) {
  if (forced) return *forced;  // (a measured choice, tune_schedule)
  FusedOptions o;
  o.group = env_int("SRK_FUSED_GROUP", 4);
  o.min_blocks = env_int("SRK_FUSED_MINB", 4);
  const size_t groups = (n_voices + kVoicesPerGroup - 1) / kVoicesPerGroup;
  const size_t n_sm = (size_t)std::max(e.n_sm, 1);
  const size_t per_sm = std::max<size_t>((groups + n_sm - 1) / n_sm, 1);
  // up to 16 warps per SM (128 registers each); with more than 4 groups per SM one warp per group is faster
  // (cfg4 @ 32768 voices: 25.4 ms against 30.1 ms as two stages, profiles/r04g)
  o.stages = per_sm <= 4 ? (int)std::min<size_t>(8, 16 / per_sm) : 1;
  if (env_int("SRK_FUSED_STAGES", 0) > 0) { o.stages = env_int("SRK_FUSED_STAGES", 1); o.exact_stages = true; }
  o.tile = env_int("SRK_FUSED_TILE_ROWS", 32);
  o.split_moog = env_int("SRK_FUSED_SPLIT_MOOG", 1) != 0;
  o.prefetch = env_int("SRK_FUSED_PREFETCH", 0) != 0;
  return o;
}
// ... and the variant that fits: every group an SM gets must be resident at once (the launch is one wave)
static int fused_generate_fitting(const srk_patch& patch, const Engine& e, const Program& prog, size_t n_voices, FusedSpec& spec, std::string& why,
                                  const FusedOptions* forced = nullptr) {
  FusedOptions o = fused_options(e, n_voices, forced);
  const size_t groups = (n_voices + kVoicesPerGroup - 1) / kVoicesPerGroup;
  const size_t n_sm = (size_t)std::max(e.n_sm, 1);
  const size_t per_sm = std::max<size_t>((groups + n_sm - 1) / n_sm, 1);
  const size_t budget = (size_t)e.smem_sm - 1024 * std::min<size_t>(per_sm, 32);  // (1 KB per block is the system's)
  for (;;) {
    int rc = fused_generate(patch, prog, o, spec, why);
    if (rc != SRK_OK) return rc;
    if (env_int("SRK_DEBUG", 0))
      std::fprintf(stderr, "[srk] fused: stages %d tile %d cross wires %d smem/group %zu x %zu groups/SM (budget %zu)\n", spec.stages, spec.tile,
                   spec.n_cross, spec.smem_per_group, per_sm, budget);
    if (spec.stages == 1 || spec.smem_per_group * per_sm <= budget) return SRK_OK;
    if (forced) return SRK_ERR_LIMIT;  // (a measured candidate either fits as asked for or is no candidate)
    if (spec.split_moog) o.split_moog = false;  // the coefficient wires first, then shorter tiles, then fewer stages
    else if (o.tile == 32) o.tile = 16;
    else { o.stages = spec.stages - 1; o.exact_stages = false; o.tile = 32; }
  }
}

// Compiles the planned patch for n_voices: pipelined program, one-warp program or one-warp program + fused
// kernel source.  `e` supplies the device limits (a probe without a device assumes sm_100).
static int schedule_interpreter(const srk_patch& patch, const Engine& e, size_t n_voices, Program& prog, std::vector<uint4>& blob, int& K,
                                std::string& err) {
  int rc = compile_program(patch, choose_max_warps(e, n_voices), prog, err);
  if (rc != SRK_OK) return rc;
  build_blob(prog, blob);
  K = prog.n_warps > 1 && !pipelined_pays(e, prog, n_voices) ? 0 : choose_chunk(e, prog, blob.size(), n_voices);
  if (K == 0 && prog.n_warps > 1) {  // does not fit or does not pay as a pipeline: one warp, plan order
    rc = compile_program(patch, 1, prog, err);
    if (rc != SRK_OK) return rc;
    build_blob(prog, blob);
    K = choose_chunk(e, prog, blob.size(), n_voices);
  }
  if (K == 0) { err = "patch needs more shared memory than one block can have"; return SRK_ERR_LIMIT; }
  return SRK_OK;
}

static int schedule_program(const srk_patch& patch, const Engine& e, size_t n_voices, Program& prog, std::vector<uint4>& blob,
                            int& K, bool& fused, FusedSpec& spec, std::string& note, std::string& err) {
  fused = false;
  note.clear();
  const int mode = fused_mode();
  int rc = SRK_OK;
  if (mode != 0) {  // the fused kernel is generated from the one-warp program
    rc = compile_program(patch, 1, prog, err);
    if (rc != SRK_OK) return rc;
    std::string why;
    if (fused_generate_fitting(patch, e, prog, n_voices, spec, why) == SRK_OK) {
      // Few voice groups per SM and a slowest stage that is one long module (a CV-driven oscillator: exp2, division
      // and sin in one warp): the interpreter's pipeline splits such a module over several warps and wins
      // (cfg3 @ 4096 voices: 4.2 ms against 14.7 ms, profiles/r04g).
      const size_t groups = (n_voices + kVoicesPerGroup - 1) / kVoicesPerGroup;
      const size_t per_sm = std::max<size_t>((groups + std::max(e.n_sm, 1) - 1) / std::max(e.n_sm, 1), 1);
      // cycles per sample, both ways.  A lone fused warp retires an instruction every ~2.6 cycles (measured at 4096
      // voices: cfg2 1.95, cfg4 3.3, cfg3 2.9 -- f64-heavy stages are the slow ones).  The interpreter's pipeline runs at
      // its slowest stage or, with more instructions than warps, at its 16 warps' share of everything (measured, cycles
      // per sample: cfg2 132, cfg4 442, cfg3 172, cfg3b 262, cfg1 125 against (max, sum) of its cost model (115, 437),
      // (115, 770), (57, 352), (80, 464), (47, 133)).  What this decides: patches whose weight is one CV-driven or sine
      // oscillator go to the interpreter, which splits such a module over several warps (cfg3: 4.2 against 14.7 ms).
      if (mode != 1 && per_sm <= 2) {
        Program pp;
        std::vector<uint4> pb;
        int pk = 0;
        std::string perr;
        if (schedule_interpreter(patch, e, n_voices, pp, pb, pk, perr) == SRK_OK && pp.n_warps > 1) {
          const double t_fused = 2.6 * spec.max_stage_cost, t_interp = std::max(1.15 * pp.max_cost, 0.45 * pp.sum_cost) * (double)per_sm;
          if (env_int("SRK_DEBUG", 0))
            std::fprintf(stderr, "[srk] schedule: fused %d stages, slowest %.0f instr -> %.0f cycles/sample; interpreter pipeline %.0f (max %u sum %u warps %u)\n",
                         spec.stages, spec.max_stage_cost, t_fused, t_interp, pp.max_cost, pp.sum_cost, pp.n_warps);
          if (t_interp < t_fused) {
            prog = std::move(pp); blob = std::move(pb); K = pk;
            note = "interpreter pipeline estimated faster than the fused stages";
            return SRK_OK;
          }
        }
      }
      build_blob(prog, blob);
      fused = true;
      K = spec.tile;
      return SRK_OK;
    }
    note = why;
  }
  return schedule_interpreter(patch, e, n_voices, prog, blob, K, err);
}

static int build_param_table(srk_patch* patch, Engine& e) {
  const size_t P = e.prog.param_src.size(), V = e.V;
  const size_t bytes = std::max<size_t>(P * V, 1) * sizeof(uint32_t);
  if (bytes > e.h_params_bytes) {
    if (e.h_params) cudaFreeHost(e.h_params);
    e.h_params = nullptr;
    SRK_CUDA(cudaMallocHost(&e.h_params, bytes));
    e.h_params_bytes = bytes;
  }
  for (size_t w = 0; w < P; ++w) {
    const ParamSource& src = e.prog.param_src[w];
    const srk_module* m = patch->modules[src.module];
    uint32_t* row = e.h_params + w * V;
    if (src.pid >= 0) {
      const auto& pv = m->param_pv[src.pid];
      if (!pv.empty()) {
        if (pv.size() < e.voice_offset + V) {
          patch->last_error = "per-voice parameter array shorter than voice_offset + n_voices";
          return SRK_ERR_SIZE;
        }
        std::memcpy(row, pv.data() + e.voice_offset, V * sizeof(float));
      } else {
        uint32_t bits;
        std::memcpy(&bits, &m->param[src.pid], 4);
        std::fill(row, row + V, bits);
      }
    } else {
      // Oscillator delta for an unconnected CV input, in f64 on the host exactly as the
      // reference computes it per sample: 440 * 2^val / sample_rate (oscillator.rs:46,132).
      const auto& pv = m->param_pv[SRK_OSC_VAL];
      const double sr = (double)m->osc_sample_rate;
      auto word = [&](float val) {
        double d = 440.0 * std::exp2((double)val) / sr;
        uint64_t bits;
        std::memcpy(&bits, &d, 8);
        return src.pid == -1 ? (uint32_t)bits : (uint32_t)(bits >> 32);
      };
      if (!pv.empty()) {
        if (pv.size() < e.voice_offset + V) {
          patch->last_error = "per-voice parameter array shorter than voice_offset + n_voices";
          return SRK_ERR_SIZE;
        }
        for (size_t v = 0; v < V; ++v) row[v] = word(pv[e.voice_offset + v]);
      } else {
        std::fill(row, row + V, word(m->param[SRK_OSC_VAL]));
      }
    }
  }
  e.uniform_words.assign(P, 0u);
  for (size_t w = 0; w < P && V; ++w) e.uniform_words[w] = e.h_params[w * V];  // what a fused kernel reads for a uniform word
  SRK_CUDA(e.d_params.ensure(bytes));
  SRK_CUDA(cudaMemcpyAsync(e.d_params.p, e.h_params, P * V * sizeof(uint32_t), cudaMemcpyHostToDevice, e.stream));
  return SRK_OK;
}

static int reset_state(srk_patch* patch, Engine& e) {
  const size_t S = e.prog.state_init.size();
  if (S && e.V) {
    const size_t n = S * e.V;
    state_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e.stream>>>(
        (uint32_t*)e.d_state.p, (const uint32_t*)e.d_state_init.p, (uint32_t)S, (uint32_t)e.V);
    ++e.launches;
    SRK_CUDA(cudaGetLastError());
  }
  if (e.prog.n_rings && e.V)
    SRK_CUDA(cudaMemsetAsync(e.d_rings.p, 0, (size_t)e.prog.n_rings * e.prog.ring_len * e.V * sizeof(float), e.stream));
  e.n_abs = 0;
  e.state_valid = true;
  ++e.state_epoch;
  return SRK_OK;
}

// The voice count the schedule is chosen for: this launch's plus the voices other patches render on the same device at the
// same time (srk_set_co_resident_voices): what matters to every choice below is how many voice groups an SM holds.
static size_t sched_voices(const srk_patch& patch, size_t n_voices) { return n_voices + patch.co_resident_voices; }

// (Re)compile + (re)allocate for the current plan and voice range.
static int engine_prepare(srk_patch* patch, size_t n_voices, size_t voice_offset) {
  Engine& e = *patch->engine;
  bool fresh = false;
  const size_t sv = sched_voices(*patch, n_voices);
  const int want_warps = choose_max_warps(e, sv);
  const bool rewired = e.compiled_epoch != patch->wiring_epoch || e.compiled_max_warps != want_warps ||
                       e.compiled_voices != sv;  // (a new voice count resets the voice state anyway)
  if (rewired || e.compiled_table_epoch != patch->table_epoch) {
    // (a sequence-table edit alone rebuilds the program image but keeps the voice state: the state
    // layout depends on the wiring only)
    std::string err;
    int rc;
    if (!rewired && e.tuned && e.fused && e.have_forced) {
      // a table edit under a measured choice: the same launch shape around the new program image
      std::string why;
      const int compiler = e.fspec.compiler;  // (the measured choice includes which NVRTC builds the kernel)
      rc = compile_program(*patch, 1, e.prog, err);
      if (rc == SRK_OK && fused_generate_fitting(*patch, e, e.prog, sv, e.fspec, why, &e.forced) != SRK_OK) { rc = SRK_ERR_LIMIT; err = why; }
      e.fspec.compiler = compiler;
      if (rc == SRK_OK) { build_blob(e.prog, e.blob); e.chunk = e.fspec.tile; }
    } else if (!rewired && e.tuned && !e.fused) {
      rc = schedule_interpreter(*patch, e, sv, e.prog, e.blob, e.chunk, err);
    } else {
      rc = schedule_program(*patch, e, sv, e.prog, e.blob, e.chunk, e.fused, e.fspec, e.fused_note, err);
    }
    if (rc != SRK_OK) { patch->last_error = err; return rc; }
    e.fkernel = nullptr;
    if (e.fused) {
      std::string why;
      if (fused_kernel(e.fspec, &e.fkernel, why) != SRK_OK) {
        // no NVRTC here, or the generated source did not compile: the interpreter kernels take over
        if (fused_mode() == 1) { patch->last_error = why; return SRK_ERR_UNSUPPORTED; }
        rc = schedule_interpreter(*patch, e, sv, e.prog, e.blob, e.chunk, err);
        if (rc != SRK_OK) { patch->last_error = err; return rc; }
        e.fused = false;
        e.fused_note = why;
      }
    }
    SRK_CUDA(e.d_prog.ensure(e.blob.size() * sizeof(uint4)));
    SRK_CUDA(cudaMemcpyAsync(e.d_prog.p, e.blob.data(), e.blob.size() * sizeof(uint4), cudaMemcpyHostToDevice, e.stream));
    SRK_CUDA(e.d_state_init.ensure(std::max<size_t>(e.prog.state_init.size(), 1) * sizeof(uint32_t)));
    if (!e.prog.state_init.empty())
      SRK_CUDA(cudaMemcpyAsync(e.d_state_init.p, e.prog.state_init.data(), e.prog.state_init.size() * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, e.stream));
    // the host vectors must outlive the async copies
    SRK_CUDA(cudaStreamSynchronize(e.stream));
    if (rewired) {
      e.tune_pending = tune_allowed();
      e.tuned = false;
      e.have_forced = false;
      e.tune_note.clear();
    }
    e.compiled_epoch = patch->wiring_epoch;
    e.compiled_max_warps = want_warps;
    e.compiled_voices = sv;
    e.compiled_table_epoch = patch->table_epoch;
    fresh = rewired;
    e.uploaded_wave_epoch = 0;  // WaveDesc offsets follow the plan: re-concatenate
  }
  if (e.uploaded_wave_epoch != patch->wave_epoch) {
    // Sample tables (WaveBox.samples), back to back in the order compile_program laid them out
    SRK_CUDA(e.d_waves.ensure(std::max<size_t>(e.prog.wave_total, 1) * sizeof(float)));
    size_t off = 0;
    for (int mi : e.prog.wave_modules) {
      const std::vector<float>& w = patch->modules[mi]->wave;
      if (!w.empty())
        SRK_CUDA(cudaMemcpyAsync((float*)e.d_waves.p + off, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, e.stream));
      off += w.size();
    }
    SRK_CUDA(cudaStreamSynchronize(e.stream));  // pageable host vectors may change after we return
    e.uploaded_wave_epoch = patch->wave_epoch;
  }
  if (fresh || e.V != n_voices || e.voice_offset != voice_offset || !e.state_valid) {
    e.V = n_voices;
    e.voice_offset = voice_offset;
    SRK_CUDA(e.d_state.ensure(std::max<size_t>(e.prog.state_init.size() * e.V, 1) * sizeof(uint32_t)));
    SRK_CUDA(e.d_rings.ensure(std::max<size_t>((size_t)e.prog.n_rings * e.prog.ring_len * e.V, 1) * sizeof(float)));
    int rc = reset_state(patch, e);
    if (rc != SRK_OK) return rc;
    e.uploaded_param_epoch = 0;
  }
  if (e.uploaded_param_epoch != patch->param_epoch) {
    if (e.fused) {
      // which parameter words are uniform over voices is part of a fused kernel's source: a parameter that became
      // per-voice (or uniform again) selects another kernel
      FusedSpec spec;
      std::string why;
      spec.compiler = e.fspec.compiler;
      const int rc_gen = fused_generate_fitting(*patch, e, e.prog, sv, spec, why, e.have_forced ? &e.forced : nullptr);
      spec.compiler = e.fspec.compiler;
      if (rc_gen == SRK_OK && spec.source != e.fspec.source) {
        const FusedKernel* k = nullptr;
        if (fused_kernel(spec, &k, why) == SRK_OK) {
          e.fspec = std::move(spec);
          e.fkernel = k;
        } else {
          patch->last_error = why;  // (the state layout is the same, but the interpreter's program may be a pipelined one: replan)
          return SRK_ERR_UNSUPPORTED;
        }
      }
    }
    int rc = build_param_table(patch, e);
    if (rc != SRK_OK) return rc;
    SRK_CUDA(cudaStreamSynchronize(e.stream));  // staging buffer is reused by the next upload
    e.uploaded_param_epoch = patch->param_epoch;
  }
  return SRK_OK;
}

// One way to run the compiled patch for a launch shape: the interpreter's program (pipelined or one warp per group), or
// the one-warp program plus the fused kernel generated from it.
struct Schedule {
  Program prog;
  std::vector<uint4> blob;
  int chunk = 0;
  bool fused = false;
  FusedSpec fspec;
  FusedOptions fopt;  // what the fused kernel was generated with, when that was forced
  bool forced = false;
  const FusedKernel* fkernel = nullptr;
  std::string id;     // "fused:<key>" or "interpreter:<warps>x<chunk>"
};
struct SchedView {
  const Program* prog;
  size_t blob_vec;
  int chunk;
  bool fused;
  const FusedSpec* fspec;
  const FusedKernel* fkernel;
};
static SchedView view_of(const Engine& e) { return SchedView{&e.prog, e.blob.size(), e.chunk, e.fused, &e.fspec, e.fkernel}; }
static SchedView view_of(const Schedule& s) { return SchedView{&s.prog, s.blob.size(), s.chunk, s.fused, &s.fspec, s.fkernel}; }

struct LaunchIO {
  uint32_t* state;
  float* rings;
  float* stems;    // device, or null
  float* partial;  // device [G][C][N], or null (no mix)
  size_t n_voices, voice_offset, n_samples;
  uint64_t n_abs;
};
struct LaunchShape { int K = 0, G = 1, T = 0; size_t smem = 0; };

// One voice-kernel launch of a schedule -- the engine's current one or a tuning candidate -- on `work`.
static int launch_voice_kernel(srk_patch* patch, Engine& e, const SchedView& v, const void* d_prog, const LaunchIO& io, cudaStream_t work,
                               LaunchShape& shape) {
  const Program& prog = *v.prog;
  const size_t C = prog.channels;
  const size_t n_voices = io.n_voices, voice_offset = io.voice_offset, n_samples = io.n_samples;
  const unsigned n_groups = (unsigned)((n_voices + kVoicesPerGroup - 1) / kVoicesPerGroup);
  int K = 0, G = 1, T = 0;
  size_t smem = 0;
  if (v.fused) {
    // ---- fused kernel: S warps (stages) per voice group, everything in registers; single-stage groups may share a block
    const int S = std::max(1, v.fspec->stages);
    const int wpb = S > 1 ? 1 : std::max(1, std::min(env_int("SRK_FUSED_WPB", 1), kFusedMaxThreads / 32));  // groups per block
    SrkFusedArgs args{};
    args.state = (unsigned*)io.state;
    args.params = (const unsigned*)e.d_params.p;
    args.rings = io.rings;
    args.stems = io.stems;
    args.partial = io.partial;
    args.waves = (const float*)e.d_waves.p;
    args.tables = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(d_prog) + blob_table_offset(prog));
    args.V = (unsigned)n_voices;
    args.voice_offset = (unsigned)voice_offset;
    args.n_samples = (unsigned)n_samples;
    args.C = (unsigned)C;
    args.B = std::max<uint32_t>(prog.ring_len, 1);
    args.ring_phase = (unsigned)(io.n_abs % args.B);
    args.n_abs = (unsigned)io.n_abs;
    args.seed_lo = (unsigned)patch->seed;
    args.seed_hi = (unsigned)(patch->seed >> 32);
    SrkTensorMap tmap{};
    args.use_tma = 0;
    if (io.stems && n_voices % 4 == 0 && env_int("SRK_FUSED_TMA", 1)) {
      std::string why;
      if (fused_stems_map(&tmap, io.stems, C, n_samples, n_voices, (unsigned)v.fspec->tile, why) == SRK_OK) args.use_tma = 1;
    }
    for (size_t w = 0; w < e.uniform_words.size() && w < SRK_FUSED_MAX_UNIFORM; ++w) args.u[w] = e.uniform_words[w];
    smem = v.fspec->smem_per_group * wpb;
    FusedKernel* fk = const_cast<FusedKernel*>(v.fkernel);
    if (smem > fk->max_smem_set) {
      SRK_CUDA(cudaFuncSetAttribute((const void*)fk->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SRK_CUDA(cudaFuncSetAttribute((const void*)fk->kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      fk->max_smem_set = smem;
    }
    void* kargs[] = {&args, &tmap};
    const unsigned grid = (n_groups + wpb - 1) / wpb;
    SRK_CUDA(cudaLaunchKernel((const void*)fk->kernel, dim3(grid), dim3(32u * S * wpb), kargs, smem, work));
    K = v.fspec->tile;
    G = wpb;
    T = 32 * S;
  } else {
  K = chunk_for_length(prog, v.chunk, n_samples);  // <= the chunk that fitted
  T = (int)prog.n_warps * 32;
  G = prog.n_warps == 1 ? choose_solo_groups(e, prog, v.blob_vec, K, sched_voices(*patch, n_voices)) : 1;
  smem = smem_bytes_for(prog, v.blob_vec, K, G);
  const unsigned grid = (n_groups + G - 1) / G;

  RenderArgs a{};
  a.blob = (const uint4*)d_prog;
  a.state = io.state;
  a.params = (const uint32_t*)e.d_params.p;
  a.rings = io.rings;
  a.stems = io.stems;
  a.partial = io.partial;
  a.waves = (const float*)e.d_waves.p;
  a.blob_vec = (uint32_t)v.blob_vec;
  a.table_off = (uint32_t)blob_table_offset(prog);
  a.n_instr = (uint32_t)prog.code.size();
  a.n_wires = (uint32_t)prog.wires.size();
  a.n_warps = prog.n_warps;
  a.n_stages = prog.n_stages;
  a.n_tiles = prog.n_tiles;
  a.V = (uint32_t)n_voices;
  a.voice_offset = (uint32_t)voice_offset;
  a.n_samples = (uint32_t)n_samples;
  a.S = (uint32_t)prog.state_init.size();
  a.P = (uint32_t)prog.param_src.size();
  a.C = (uint32_t)C;
  a.B = std::max<uint32_t>(prog.ring_len, 1);
  a.K = (uint32_t)K;
  a.log2K = 0;
  while ((1u << a.log2K) < a.K) ++a.log2K;
  a.ring_phase = (uint32_t)(io.n_abs % a.B);
  a.n_abs = (uint32_t)io.n_abs;
  a.seed_lo = (uint32_t)patch->seed;
  a.seed_hi = (uint32_t)(patch->seed >> 32);
  a.solo_op_barrier = G > 1 && env_int("SRK_SOLO_OP_BARRIER", 0) ? 1u : 0u;

  bool beyond_baseline = false;  // sequencers / sample player: the larger one-warp image
  for (const Instr& ins : prog.code) beyond_baseline |= ins.op == OP_GRIDSEQ || ins.op == OP_PATSEQ || ins.op == OP_SAMPLE;
  SRK_CUDA(prog.n_warps > 1 ? launch_voices_pipelined(a, grid, (unsigned)T, smem, work)
           : beyond_baseline ? launch_voices_solo_full(a, grid, 32u * G, smem, work)
                             : launch_voices_solo(a, grid, 32u * G, smem, work));
  }
  SRK_CUDA(cudaGetLastError());
  shape.K = K; shape.G = G; shape.T = T; shape.smem = smem;
  return SRK_OK;
}

// ----------------------------------------------------------------------------
// Measured schedule choice.  schedule_program() picks a launch shape from a cost model (instructions per sample, voice
// groups per SM); which shape is actually fastest also depends on things the model does not see -- how well ptxas
// interleaves a stage's straight-line groups, instruction-cache footprint, the stems traffic.  So the first render of
// a schedule that is long enough to be worth it runs the plausible alternatives for a few thousand samples each on a
// scratch copy of the voice state, timed with CUDA events on the render stream, and keeps the fastest.  Every
// candidate computes the same bits (tests/test_gpu_parity.py::test_measured_schedule_choice_keeps_the_bits), so the
// choice is invisible in the output.  The decision is cached next to the cubins (kernel_cache/<key>.tune).
// Off with SRK_TUNE=0 or whenever a schedule knob is forced through the environment.
// ----------------------------------------------------------------------------
constexpr size_t kTuneMinSamples = 16384;  // shorter renders keep the cost model's choice (tuning would dominate them)

static bool tune_allowed() {
  if (!env_int("SRK_TUNE", 1)) return false;
  for (const char* k : {"SRK_FUSED", "SRK_WARPS", "SRK_STEP", "SRK_FUSED_STAGES", "SRK_FUSED_GROUP", "SRK_FUSED_TILE_ROWS", "SRK_SOLO_GROUPS",
                        "SRK_FUSED_MINB", "SRK_FUSED_WPB", "SRK_SOLO_OP_BARRIER", "SRK_FUSED_SPLIT_MOOG", "SRK_FUSED_PREFETCH"}) {
    const char* v = std::getenv(k);
    if (v && *v) return false;
  }
  return true;
}

static std::string interp_id(const Program& prog, int chunk) {
  return "interpreter:" + std::to_string(prog.n_warps) + "x" + std::to_string(chunk);
}

// The alternatives worth measuring next to the schedule `cur` the cost model picked for `sv` scheduling voices.
// cur itself is out[0].  Needs no device.
static void schedule_candidates(const srk_patch& patch, const Engine& lim, size_t sv, Schedule cur, std::vector<Schedule>& out) {
  const size_t groups = (sv + kVoicesPerGroup - 1) / kVoicesPerGroup;
  const size_t n_sm = (size_t)std::max(lim.n_sm, 1);
  const size_t per_sm = std::max<size_t>((groups + n_sm - 1) / n_sm, 1);
  const size_t budget = (size_t)lim.smem_sm - 1024 * std::min<size_t>(per_sm, 32);
  cur.id = cur.fused ? "fused:" + fused_key(cur.fspec) : interp_id(cur.prog, cur.chunk);
  const bool cur_fused = cur.fused;
  const int S0 = cur.fused ? cur.fspec.stages : 0, g0 = cur.fused ? cur.fspec.group : 4, t0 = cur.fused ? cur.fspec.tile : 32;
  out.clear();
  out.push_back(std::move(cur));
  auto add_fused = [&](int stages, int group, int tile, bool split) {
    Program one;
    std::string err, why;
    if (compile_program(patch, 1, one, err) != SRK_OK) return;
    FusedOptions o;
    o.group = group; o.min_blocks = 4; o.stages = stages; o.exact_stages = true; o.tile = tile; o.split_moog = split;
    FusedSpec spec;
    if (fused_generate(patch, one, o, spec, why) != SRK_OK || spec.stages != stages || spec.group != group) return;
    if (spec.stages > 1 && spec.smem_per_group * per_sm > budget) return;
    for (const Schedule& c : out)
      if (c.fused && c.fspec.source == spec.source) return;
    Schedule s;
    s.prog = std::move(one);
    build_blob(s.prog, s.blob);
    s.chunk = spec.tile;
    s.fused = true;
    s.fopt = o;
    s.forced = true;
    s.id = "fused:" + fused_key(spec);
    s.fspec = std::move(spec);
    out.push_back(std::move(s));
  };
  if (cur_fused) {
    const bool sp0 = out[0].fspec.split_moog;
    add_fused(S0, g0 == 8 ? 4 : 8, t0, sp0);                 // samples per straight-line group: 4 or 8
    if (S0 > 1 && S0 < 8) { add_fused(S0 + 1, 4, t0, sp0); add_fused(S0 + 1, 8, t0, sp0); }  // one more pipeline stage
    if (S0 > 1) {                                            // the ladder filters' coefficients inside / outside their stage
      add_fused(S0, 4, t0, !sp0); add_fused(S0, 8, t0, !sp0);
      if (S0 < 8) { add_fused(S0 + 1, 4, t0, !sp0); add_fused(S0 + 1, 8, t0, !sp0); }
    }
    if (per_sm <= 2) {                                       // the interpreter's pipeline (splits single modules over warps)
      Schedule s;
      std::string err;
      if (schedule_interpreter(patch, lim, sv, s.prog, s.blob, s.chunk, err) == SRK_OK && s.prog.n_warps > 1) {
        s.id = interp_id(s.prog, s.chunk);
        out.push_back(std::move(s));
      }
    }
  } else {
    Program one;
    std::string err, why;
    FusedSpec spec;
    if (fused_mode() != 0 && compile_program(patch, 1, one, err) == SRK_OK && fused_generate_fitting(patch, lim, one, sv, spec, why) == SRK_OK) {
      add_fused(spec.stages, 4, spec.tile, spec.split_moog);
      add_fused(spec.stages, 8, spec.tile, spec.split_moog);
    }
  }
  // every fused candidate once more as the other NVRTC version builds it, when there is one (fused_rt.cpp: neither is
  // better everywhere)
  if (fused_compilers() > 1) {
    const size_t n = out.size();
    for (size_t i = 0; i < n; ++i) {
      if (!out[i].fused) continue;
      Schedule s = out[i];
      s.fspec.compiler = 1;
      s.fkernel = nullptr;
      if (!s.forced) {  // the cost model's own shape, spelled out so that a later regeneration keeps it
        s.fopt.group = s.fspec.group; s.fopt.min_blocks = s.fspec.min_blocks; s.fopt.stages = s.fspec.stages;
        s.fopt.exact_stages = true; s.fopt.tile = s.fspec.tile; s.fopt.split_moog = s.fspec.split_moog;
        s.forced = true;
      }
      s.id = "fused:" + fused_key(s.fspec);
      out.push_back(std::move(s));
    }
  }
}

static void adopt_schedule(Engine& e, Schedule&& s) {
  e.prog = std::move(s.prog);
  e.blob = std::move(s.blob);
  e.chunk = s.chunk;
  e.fused = s.fused;
  e.fspec = std::move(s.fspec);
  e.fkernel = s.fkernel;
  e.have_forced = s.fused && s.forced;
  e.forced = s.fopt;
}

static int tune_schedule(srk_patch* patch, Engine& e, size_t n_voices, size_t voice_offset, bool want_stems, bool want_mix) {
  e.tune_pending = false;
  const size_t sv = sched_voices(*patch, n_voices);
  Schedule cur;
  cur.prog = e.prog; cur.blob = e.blob; cur.chunk = e.chunk; cur.fused = e.fused; cur.fspec = e.fspec; cur.fkernel = e.fkernel;
  std::vector<Schedule> cand;
  schedule_candidates(*patch, e, sv, std::move(cur), cand);
  // candidates must share the state and parameter layout of the current program (they do: both follow the plan)
  for (size_t i = cand.size(); i-- > 1;) {
    const Program& a = cand[0].prog;
    const Program& b = cand[i].prog;
    bool same = a.state_init == b.state_init && a.param_src.size() == b.param_src.size() && a.n_rings == b.n_rings && a.ring_len == b.ring_len;
    for (size_t w = 0; same && w < a.param_src.size(); ++w)
      same = a.param_src[w].module == b.param_src[w].module && a.param_src[w].pid == b.param_src[w].pid;
    if (!same) cand.erase(cand.begin() + (long)i);
  }
  for (size_t i = cand.size(); i-- > 1;) {  // load (compile) the fused kernels; one that does not build is no candidate
    std::string why;
    if (cand[i].fused && fused_kernel(cand[i].fspec, &cand[i].fkernel, why) != SRK_OK) cand.erase(cand.begin() + (long)i);
  }
  if (cand.size() < 2) { e.tuned = true; e.tune_note = "no alternative schedule"; return SRK_OK; }

  // ---- a decision made earlier for exactly these candidates on this kind of launch?
  std::string all;
  for (const Schedule& c : cand) { all += c.id; all += '\n'; }
  all += "V=" + std::to_string(sv) + " stems=" + std::to_string((int)want_stems) + " mix=" + std::to_string((int)want_mix) + " sm=" + std::to_string(e.n_sm);
  const std::string name = fused_hash(all, "tune-v1") + ".tune";
  const std::string path = fused_cache_dir() + "/" + name;
  size_t best = cand.size();
  std::string found_in;
  const char* nocache = std::getenv("SRK_KERNEL_CACHE_OFF");
  if (!(nocache && nocache[0] == '1')) {
    // this machine's own measurements first, then the decisions shipped with the library (tuned/: measured on B200 for
    // the BASELINE launches, so that a fresh checkout runs the kernels the committed profiles describe)
    for (const std::string& file : {path, fused_tuned_dir() + "/" + name}) {
      FILE* f = std::fopen(file.c_str(), "r");
      if (!f) continue;
      char line[256] = {0};
      if (std::fgets(line, sizeof line, f)) {
        std::string id(line);
        while (!id.empty() && (id.back() == '\n' || id.back() == '\r')) id.pop_back();
        for (size_t i = 0; i < cand.size(); ++i)
          if (cand[i].id == id) { best = i; found_in = file; }
      }
      std::fclose(f);
      if (best != cand.size()) break;
    }
  }
  std::string report;
  const int pick = env_int("SRK_TUNE_PICK", -1);  // tests: take candidate #pick instead of measuring
  if (pick >= 0) {
    best = std::min<size_t>((size_t)pick, cand.size() - 1);
    e.tune_note = "picked " + cand[best].id + " (SRK_TUNE_PICK) of " + std::to_string(cand.size()) + " candidates";
    if (best != 0) {
      adopt_schedule(e, std::move(cand[best]));
      SRK_CUDA(e.d_prog.ensure(e.blob.size() * sizeof(uint4)));
      SRK_CUDA(cudaMemcpyAsync(e.d_prog.p, e.blob.data(), e.blob.size() * sizeof(uint4), cudaMemcpyHostToDevice, e.stream));
      SRK_CUDA(cudaStreamSynchronize(e.stream));
    }
    e.tuned = true;
    return SRK_OK;
  }
  if (best == cand.size()) {
    // ---- measure: T(2K) - T(K) per candidate (launch overhead, state load/store and pipeline fill cancel)
    const Program& p0 = cand[0].prog;
    const size_t C = p0.channels;
    size_t K1 = 2048;
    if (want_stems)
      while (K1 > 256 && C * 2 * K1 * n_voices * sizeof(float) > (256u << 20)) K1 /= 2;
    if (p0.n_rings) K1 = std::max<size_t>(K1, std::min<size_t>(4 * p0.ring_len, 8192));  // past the first wrap of a feedback ring
    const size_t K2 = 2 * K1;
    const unsigned n_groups = (unsigned)((n_voices + kVoicesPerGroup - 1) / kVoicesPerGroup);
    const size_t state_bytes = std::max<size_t>(p0.state_init.size() * n_voices, 1) * sizeof(uint32_t);
    const size_t ring_bytes = std::max<size_t>((size_t)p0.n_rings * p0.ring_len * n_voices, 1) * sizeof(float);
    SRK_CUDA(e.d_tune_state.ensure(state_bytes));
    SRK_CUDA(e.d_tune_rings.ensure(ring_bytes));
    if (want_stems) SRK_CUDA(e.d_tune_stems.ensure(C * K2 * n_voices * sizeof(float)));
    if (want_mix) SRK_CUDA(e.d_tune_partial.ensure((size_t)n_groups * C * K2 * sizeof(float)));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    SRK_CUDA(cudaEventCreate(&ev0));
    SRK_CUDA(cudaEventCreate(&ev1));
    double best_ms = 0.0;
    int rc = SRK_OK;
    for (size_t i = 0; i < cand.size() && rc == SRK_OK; ++i) {
      const Schedule& c = cand[i];
      if (e.d_tune_prog.ensure(c.blob.size() * sizeof(uint4)) != cudaSuccess ||
          cudaMemcpyAsync(e.d_tune_prog.p, c.blob.data(), c.blob.size() * sizeof(uint4), cudaMemcpyHostToDevice, e.stream) != cudaSuccess) { rc = SRK_ERR_CUDA; break; }
      // (the first launch also pays for loading the kernel image; then three (K1, K2) pairs, the fastest of each length:
      //  a single pair left candidates 2 % apart in an order that changed from run to run)
      constexpr int kRuns = 7;
      float t[kRuns] = {0};
      const size_t len[kRuns] = {K1, K1, K2, K1, K2, K1, K2};
      for (int k = 0; k < kRuns && rc == SRK_OK; ++k) {
        cudaMemcpyAsync(e.d_tune_state.p, e.d_state.p, state_bytes, cudaMemcpyDeviceToDevice, e.stream);
        cudaMemcpyAsync(e.d_tune_rings.p, e.d_rings.p, ring_bytes, cudaMemcpyDeviceToDevice, e.stream);
        LaunchIO io{};
        io.state = (uint32_t*)e.d_tune_state.p;
        io.rings = (float*)e.d_tune_rings.p;
        io.stems = want_stems ? (float*)e.d_tune_stems.p : nullptr;
        io.partial = want_mix ? (float*)e.d_tune_partial.p : nullptr;
        io.n_voices = n_voices;
        io.voice_offset = voice_offset;
        io.n_samples = len[k];
        io.n_abs = e.n_abs;
        LaunchShape shape;
        cudaEventRecord(ev0, e.stream);
        rc = launch_voice_kernel(patch, e, view_of(c), e.d_tune_prog.p, io, e.stream, shape);
        cudaEventRecord(ev1, e.stream);
        if (rc == SRK_OK && cudaEventSynchronize(ev1) != cudaSuccess) rc = SRK_ERR_CUDA;
        if (rc == SRK_OK) cudaEventElapsedTime(&t[k], ev0, ev1);
        ++e.launches;
      }
      if (rc != SRK_OK && i > 0) {  // an alternative that cannot be launched on this device is no candidate
        cudaGetLastError();
        cudaStreamSynchronize(e.stream);
        report += std::string(i ? ", " : "") + c.id + " failed";
        rc = SRK_OK;
        continue;
      }
      // ms per K1 samples in steady state
      const double slope = (double)std::min(t[2], std::min(t[4], t[6])) - (double)std::min(t[1], std::min(t[3], t[5]));
      char buf[160];
      std::snprintf(buf, sizeof buf, "%s%s %.4f", i ? ", " : "", c.id.c_str(), slope);
      report += buf;
      // an alternative has to win by more than the measurement resolves (1 %) to displace an earlier candidate
      if (rc == SRK_OK && slope > 0.0 && (best == cand.size() || slope < 0.99 * best_ms)) { best = i; best_ms = slope; }
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    for (DevBuf* b : {&e.d_tune_state, &e.d_tune_rings, &e.d_tune_stems, &e.d_tune_partial, &e.d_tune_prog}) b->release();
    if (rc != SRK_OK) return rc;
    if (best == cand.size()) best = 0;
    report = "measured ms per " + std::to_string(K1) + " samples: " + report;
    if (!(nocache && nocache[0] == '1')) {
      mkdir(fused_cache_dir().c_str(), 0755);
      const std::string tmp = path + ".tmp." + std::to_string((long)getpid());
      if (FILE* f = std::fopen(tmp.c_str(), "w")) {
        std::fprintf(f, "%s\n%s\n%s\n", cand[best].id.c_str(), report.c_str(), all.c_str());
        std::fclose(f);
        if (std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
      }
    }
  } else {
    report = "decision from " + found_in;
  }
  e.tune_note = "chose " + cand[best].id + "; " + report;
  if (env_int("SRK_DEBUG", 0)) std::fprintf(stderr, "[srk] tune: %s\n", e.tune_note.c_str());
  if (best != 0) {
    adopt_schedule(e, std::move(cand[best]));
    SRK_CUDA(e.d_prog.ensure(e.blob.size() * sizeof(uint4)));
    SRK_CUDA(cudaMemcpyAsync(e.d_prog.p, e.blob.data(), e.blob.size() * sizeof(uint4), cudaMemcpyHostToDevice, e.stream));
    SRK_CUDA(cudaStreamSynchronize(e.stream));
  }
  e.tuned = true;
  return SRK_OK;
}

int engine_render(srk_patch* patch, size_t n_voices, size_t voice_offset, size_t n_samples, unsigned flags,
                  float* stems, float* mix, void* user_stream, bool use_user_stream) {
  if (!patch->planned) { patch->last_error = "srk_plan() has not been called since the last wiring change"; return SRK_ERR_NOT_PLANNED; }
  if (n_voices > (0xFFFFFFFFull / kMaxChunk) || n_samples > 0xFFFFFFFFull) { patch->last_error = "n_voices exceeds 2^25-1 or n_samples 2^32-1"; return SRK_ERR_ARG; }
  if ((flags & SRK_RENDER_ASYNC) && !(flags & SRK_RENDER_DEVICE_OUT)) { patch->last_error = "SRK_RENDER_ASYNC needs SRK_RENDER_DEVICE_OUT"; return SRK_ERR_ARG; }
  int rc = engine_open(patch);
  if (rc != SRK_OK) return rc;
  Engine& e = *patch->engine;
  SRK_CUDA(cudaSetDevice(e.device));
  cudaStream_t work = e.stream;
  const bool foreign = use_user_stream;
  cudaStream_t caller = static_cast<cudaStream_t>(user_stream);
  if (foreign) {
    // order our stream after everything the caller has enqueued so far
    SRK_CUDA(cudaEventRecord(e.ev[0], caller));
    SRK_CUDA(cudaStreamWaitEvent(work, e.ev[0], 0));
  }
  SRK_CUDA(cudaEventRecord(e.ev[0], work));
  rc = engine_prepare(patch, n_voices, voice_offset);
  if (rc != SRK_OK) return rc;
  if (n_voices == 0 || n_samples == 0) {
    // no voices: the mix of nothing is silence (a rank that got no voices still contributes zeros to the NCCL sum)
    if (mix && n_samples) {
      const size_t bytes = (size_t)e.prog.channels * n_samples * sizeof(float);
      if (flags & SRK_RENDER_DEVICE_OUT) {
        SRK_CUDA(cudaMemsetAsync(mix, 0, bytes, work));
        if (foreign) { SRK_CUDA(cudaEventRecord(e.ev[3], work)); SRK_CUDA(cudaStreamWaitEvent(caller, e.ev[3], 0)); }
        if (!(flags & SRK_RENDER_ASYNC)) SRK_CUDA(cudaStreamSynchronize(work));
      } else {
        std::memset(mix, 0, bytes);
      }
    }
    e.timed = false;
    return SRK_OK;
  }

  if (e.tune_pending && n_samples >= kTuneMinSamples) {
    rc = tune_schedule(patch, e, n_voices, voice_offset, stems != nullptr, mix != nullptr);
    if (rc != SRK_OK) return rc;
    SRK_CUDA(cudaEventRecord(e.ev[0], work));  // the measurement is not part of this call's device time
  }
  const Program& prog = e.prog;
  const size_t C = prog.channels;
  const unsigned n_groups = (unsigned)((n_voices + kVoicesPerGroup - 1) / kVoicesPerGroup);
  const bool device_out = flags & SRK_RENDER_DEVICE_OUT;

  float* d_stems = nullptr;
  float* d_mix = nullptr;
  if (stems) {
    if (device_out) d_stems = stems;
    else { SRK_CUDA(e.d_stems.ensure(C * n_samples * n_voices * sizeof(float))); d_stems = (float*)e.d_stems.p; }
  }
  if (mix) {
    if (device_out) d_mix = mix;
    else { SRK_CUDA(e.d_mix.ensure(C * n_samples * sizeof(float))); d_mix = (float*)e.d_mix.p; }
    SRK_CUDA(e.d_partial.ensure((size_t)n_groups * C * n_samples * sizeof(float)));
  }

  LaunchIO io{};
  io.state = (uint32_t*)e.d_state.p;
  io.rings = (float*)e.d_rings.p;
  io.stems = d_stems;
  io.partial = mix ? (float*)e.d_partial.p : nullptr;
  io.n_voices = n_voices;
  io.voice_offset = voice_offset;
  io.n_samples = n_samples;
  io.n_abs = e.n_abs;
  LaunchShape shape;
  SRK_CUDA(cudaEventRecord(e.ev[1], work));
  rc = launch_voice_kernel(patch, e, view_of(e), e.d_prog.p, io, work, shape);
  if (rc != SRK_OK) return rc;
  const int K = shape.K, G = shape.G, T = shape.T;
  const size_t smem = shape.smem;
  SRK_CUDA(cudaGetLastError());
  ++e.launches;
  SRK_CUDA(cudaEventRecord(e.ev[2], work));
  if (mix) {
    const size_t cn = C * n_samples;
    mix_reduce_kernel<<<(unsigned)((cn + 255) / 256), 256, 0, work>>>((const float*)e.d_partial.p, d_mix, n_groups, cn);
    SRK_CUDA(cudaGetLastError());
    ++e.launches;
  }
  if (!device_out) {
    if (stems) SRK_CUDA(cudaMemcpyAsync(stems, d_stems, C * n_samples * n_voices * sizeof(float), cudaMemcpyDeviceToHost, work));
    if (mix) SRK_CUDA(cudaMemcpyAsync(mix, d_mix, C * n_samples * sizeof(float), cudaMemcpyDeviceToHost, work));
  }
  SRK_CUDA(cudaEventRecord(e.ev[3], work));
  e.timed = true;
  e.n_abs += n_samples;
  // WaveBox.new is consumed by the first calc() after a load (sample.rs:212-216): the program image of
  // the next render carries is_new = 0 again
  for (int mi : prog.wave_modules)
    if (patch->modules[mi]->wave_new) { patch->modules[mi]->wave_new = false; ++patch->table_epoch; }
  e.block_threads = e.fused ? T * G : prog.n_warps == 1 ? 32 * G : T;
  e.last_fused = e.fused;
  e.step = K;
  e.n_warps = e.fused ? e.fspec.stages : (int)prog.n_warps;
  e.n_stages = e.fused ? e.fspec.stages : (int)prog.n_stages;
  e.smem_bytes = smem;
  if (foreign) SRK_CUDA(cudaStreamWaitEvent(caller, e.ev[3], 0));  // caller's stream sees the results
  if (!(flags & SRK_RENDER_ASYNC)) SRK_CUDA(cudaStreamSynchronize(work));
  return SRK_OK;
}

int engine_sync(srk_patch* patch) {
  if (!patch->engine) return SRK_OK;
  SRK_CUDA(cudaSetDevice(patch->engine->device));
  SRK_CUDA(cudaStreamSynchronize(patch->engine->stream));
  return SRK_OK;
}

int engine_reset(srk_patch* patch) {
  if (!patch->engine) return SRK_OK;
  Engine& e = *patch->engine;
  SRK_CUDA(cudaSetDevice(e.device));
  if (e.compiled_epoch != patch->wiring_epoch || !e.V) { e.state_valid = false; return SRK_OK; }
  return reset_state(patch, e);
}

void engine_invalidate_state(srk_patch* patch) {
  if (patch->engine) patch->engine->state_valid = false;
}

int engine_last_ms(srk_patch* patch, float* kernel_ms, float* total_ms) {
  if (!patch->engine || !patch->engine->timed) { patch->last_error = "no timed render yet"; return SRK_ERR_ARG; }
  Engine& e = *patch->engine;
  SRK_CUDA(cudaSetDevice(e.device));
  SRK_CUDA(cudaEventSynchronize(e.ev[3]));
  if (kernel_ms) SRK_CUDA(cudaEventElapsedTime(kernel_ms, e.ev[1], e.ev[2]));
  if (total_ms) SRK_CUDA(cudaEventElapsedTime(total_ms, e.ev[0], e.ev[3]));
  return SRK_OK;
}

uint64_t engine_launches(const srk_patch* patch) { return patch->engine ? patch->engine->launches : 0; }
uint64_t engine_state_epoch(const srk_patch* patch) { return patch->engine ? patch->engine->state_epoch : 0; }

// Compiles the planned patch the way a render of n_voices would (no device needed: the sm_100
// limits are assumed when the patch has no engine yet).
static int probe_program(srk_patch* patch, size_t n_voices, Program& prog, std::vector<uint4>& blob, int& K, int* solo_groups = nullptr,
                         bool* fused = nullptr, FusedSpec* spec = nullptr) {
  if (!patch->planned) { patch->last_error = "not planned"; return SRK_ERR_NOT_PLANNED; }
  Engine probe;
  probe.smem_optin = patch->engine ? patch->engine->smem_optin : 227 * 1024;
  probe.n_sm = patch->engine ? patch->engine->n_sm : 148;
  probe.smem_sm = patch->engine ? patch->engine->smem_sm : 228 * 1024;
  std::string err, note;
  bool is_fused = false;
  FusedSpec local;
  n_voices = sched_voices(*patch, n_voices);
  int rc = schedule_program(*patch, probe, n_voices, prog, blob, K, is_fused, spec ? *spec : local, note, err);
  if (rc != SRK_OK) { patch->last_error = err; return rc; }
  if (is_fused && !fused) {  // the caller wants the interpreter's program (srk_get_program)
    is_fused = false;
    rc = schedule_interpreter(*patch, probe, n_voices, prog, blob, K, err);
    if (rc != SRK_OK) { patch->last_error = err; return rc; }
  }
  if (fused) *fused = is_fused;
  if (solo_groups) *solo_groups = prog.n_warps == 1 && !is_fused ? choose_solo_groups(probe, prog, blob.size(), K, n_voices) : 1;
  return SRK_OK;
}

// The engine holds the schedule (possibly a measured choice) a render of n_voices uses right now.
static bool engine_has_schedule_for(const srk_patch* patch, size_t n_voices) {
  const Engine* e = patch->engine.get();
  return e && patch->planned && e->compiled_epoch == patch->wiring_epoch && e->compiled_table_epoch == patch->table_epoch &&
         e->compiled_voices == sched_voices(*patch, n_voices) && (!e->fused || e->fkernel);
}

int engine_program_info(srk_patch* patch, size_t n_voices, srk_program_info* out) {
  Program prog;
  std::vector<uint4> blob;
  int K = 0, G = 1;
  bool fused = false;
  FusedSpec spec;
  if (engine_has_schedule_for(patch, n_voices)) {
    const Engine& e = *patch->engine;
    prog = e.prog; blob = e.blob; K = e.chunk; fused = e.fused; spec = e.fspec;
    G = prog.n_warps == 1 && !fused ? choose_solo_groups(e, prog, blob.size(), K, sched_voices(*patch, n_voices)) : 1;
  } else {
    int rc = probe_program(patch, n_voices, prog, blob, K, &G, &fused, &spec);
    if (rc != SRK_OK) return rc;
  }
  std::memset(out, 0, sizeof *out);
  out->n_instr = (uint32_t)prog.code.size();
  out->step_samples = (uint32_t)K;
  out->n_wires = (uint32_t)prog.wires.size();
  out->state_words = (uint32_t)prog.state_init.size();
  out->param_words = (uint32_t)prog.param_src.size();
  out->n_rings = prog.n_rings;
  out->n_warps = prog.n_warps;
  out->n_stages = prog.n_stages;
  out->n_tiles = prog.n_tiles;
  if (fused) {
    const int wpb = spec.stages > 1 ? 1 : std::max(1, std::min(env_int("SRK_FUSED_WPB", 1), kFusedMaxThreads / 32));
    out->fused = 1;
    out->fused_group = (uint32_t)spec.group;
    out->n_warps = out->n_stages = (uint32_t)spec.stages;  // warps per voice group = pipeline stages
    out->n_tiles = (uint32_t)(spec.n_cross_tiles + 2 * spec.n_distinct);
    out->block_threads = 32u * spec.stages * wpb;
    out->smem_bytes = (uint32_t)(spec.smem_per_group * wpb);
    out->groups_per_block = (uint32_t)wpb;
    const Engine* e = patch->engine.get();
    if (e && e->fused && e->fkernel && e->fspec.source == spec.source) {  // the kernel is loaded: what ptxas made of it
      out->fused_regs = (uint32_t)e->fkernel->regs;
      out->fused_local_bytes = (uint32_t)e->fkernel->local_bytes;
    }
  } else {
    out->block_threads = prog.n_warps * 32 * G;  // one-warp schedule: G voice groups (warps) per block
    out->smem_bytes = (uint32_t)smem_bytes_for(prog, blob.size(), K, G);
    out->groups_per_block = (uint32_t)G;
  }
  return SRK_OK;
}

// The generated source of the fused kernel a render of n_voices would use (empty when it would not use one).
int engine_fused_source(srk_patch* patch, size_t n_voices, std::string& source) {
  Program prog;
  std::vector<uint4> blob;
  int K = 0;
  bool fused = false;
  FusedSpec spec;
  int rc = probe_program(patch, n_voices, prog, blob, K, nullptr, &fused, &spec);
  if (rc != SRK_OK) return rc;
  source = fused ? spec.source : std::string();
  return SRK_OK;
}

// Compiles that kernel into the on-disk cache (NVRTC; needs no GPU).  *compiled: 1 compiled now, 0 already cached
// or no fused kernel for this launch shape.
int engine_precompile(srk_patch* patch, size_t n_voices, int* compiled) {
  Program prog;
  std::vector<uint4> blob;
  int K = 0;
  bool fused = false;
  FusedSpec spec;
  if (compiled) *compiled = 0;
  int rc = probe_program(patch, n_voices, prog, blob, K, nullptr, &fused, &spec);
  if (rc != SRK_OK) return rc;
  // the schedule the cost model picks and the alternatives a long render would measure against it (tune_schedule)
  Engine probe;
  probe.smem_optin = patch->engine ? patch->engine->smem_optin : 227 * 1024;
  probe.n_sm = patch->engine ? patch->engine->n_sm : 148;
  probe.smem_sm = patch->engine ? patch->engine->smem_sm : 228 * 1024;
  Schedule cur;
  cur.prog = std::move(prog); cur.blob = std::move(blob); cur.chunk = K; cur.fused = fused; cur.fspec = std::move(spec);
  std::vector<Schedule> cand;
  if (tune_allowed()) schedule_candidates(*patch, probe, sched_voices(*patch, n_voices), std::move(cur), cand);
  else cand.push_back(std::move(cur));
  for (const Schedule& c : cand) {
    if (!c.fused) continue;
    std::vector<char> cubin;
    std::string key, err;
    bool from_disk = false;
    rc = fused_cubin(c.fspec, cubin, key, &from_disk, nullptr, err);
    if (rc != SRK_OK) { patch->last_error = err; return rc; }
    if (compiled && !from_disk) ++*compiled;
  }
  return SRK_OK;
}

int engine_tune_report(srk_patch* patch, std::string& report) {
  report = patch->engine ? patch->engine->tune_note : std::string();
  return SRK_OK;
}

// Which kernel image a render of n_voices runs: "fused:<hash of the generated source, the op headers and the compiler
// options>" or "interpreter:<hash of dsp.cuh + voice_kernel.cuh at build time>:<pipelined|solo|solo_full>".  Profiles are
// stamped with it (profiles/traffic.json) so that a number captured on another kernel is never quoted for this one.
#ifndef SRK_SOURCE_HASH
#define SRK_SOURCE_HASH "unknown"
#endif
int engine_kernel_id(srk_patch* patch, size_t n_voices, std::string& id) {
  Program prog;
  std::vector<uint4> blob;
  int K = 0;
  bool fused = false;
  FusedSpec spec;
  if (engine_has_schedule_for(patch, n_voices)) {  // what the engine launches right now (a measured choice included)
    const Engine& e = *patch->engine;
    prog = e.prog; fused = e.fused; spec = e.fspec;
  } else {
    int rc = probe_program(patch, n_voices, prog, blob, K, nullptr, &fused, &spec);
    if (rc != SRK_OK) return rc;
  }
  if (fused) { id = "fused:" + fused_key(spec); return SRK_OK; }
  bool beyond_baseline = false;
  for (const Instr& ins : prog.code) beyond_baseline |= ins.op == OP_GRIDSEQ || ins.op == OP_PATSEQ || ins.op == OP_SAMPLE;
  id = std::string("interpreter:") + SRK_SOURCE_HASH + (prog.n_warps > 1 ? ":pipelined" : beyond_baseline ? ":solo_full" : ":solo");
  return SRK_OK;
}

// ---- per-voice state export / import: the device-side analogue of the DSP state every reference module serializes
// (ui.rs:98-134; `#[derive(Serialize)]` on oscillator phase, filter memory, envelope stage ...).  The blob is the raw
// state words [S][V], the feedback rings [R][B][V] and the absolute sample index, behind a header that names the layout.
namespace {
struct StateHeader {
  char magic[8];  // "SRKSTATE"
  uint32_t version, S, R, B;
  uint64_t V, voice_offset, n_abs, layout;
};
constexpr uint32_t kStateVersion = 1;

// the state layout follows from the plan alone: module kinds in plan order and the wires the cycle breaker cut
uint64_t state_layout_hash(const srk_patch& patch, const Program& prog) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](uint64_t v) { for (int k = 0; k < 8; ++k) { h ^= (v >> (8 * k)) & 0xff; h *= 1099511628211ull; } };
  auto plan_index = [&](const srk_module* m) -> uint64_t { for (size_t i = 0; i < patch.plan.size(); ++i) if (patch.plan[i] == m) return (uint64_t)i; return ~0ull; };
  for (const srk_module* m : patch.plan) mix((uint64_t)m->kind);
  for (const auto& c : patch.cuts) { mix(plan_index(c.first)); mix(plan_index(c.second)); }
  mix(prog.state_init.size()); mix(prog.n_rings); mix(prog.ring_len);
  return h;
}
}  // namespace

int engine_state_export(srk_patch* patch, const void** blob, size_t* n_bytes) {
  if (!patch->engine || !patch->engine->state_valid || !patch->engine->V || patch->engine->compiled_epoch != patch->wiring_epoch) {
    patch->last_error = "no voice state to export: nothing has been rendered since the last wiring change";
    return SRK_ERR_ARG;
  }
  Engine& e = *patch->engine;
  SRK_CUDA(cudaSetDevice(e.device));
  SRK_CUDA(cudaStreamSynchronize(e.stream));
  const size_t state_bytes = e.prog.state_init.size() * e.V * sizeof(uint32_t);
  const size_t ring_bytes = (size_t)e.prog.n_rings * e.prog.ring_len * e.V * sizeof(float);
  patch->saved_state.resize(sizeof(StateHeader) + state_bytes + ring_bytes);
  StateHeader h{};
  std::memcpy(h.magic, "SRKSTATE", 8);
  h.version = kStateVersion;
  h.S = (uint32_t)e.prog.state_init.size();
  h.R = e.prog.n_rings;
  h.B = e.prog.ring_len;
  h.V = e.V;
  h.voice_offset = e.voice_offset;
  h.n_abs = e.n_abs;
  h.layout = state_layout_hash(*patch, e.prog);
  unsigned char* out = patch->saved_state.data();
  std::memcpy(out, &h, sizeof h);
  if (state_bytes) SRK_CUDA(cudaMemcpy(out + sizeof h, e.d_state.p, state_bytes, cudaMemcpyDeviceToHost));
  if (ring_bytes) SRK_CUDA(cudaMemcpy(out + sizeof h + state_bytes, e.d_rings.p, ring_bytes, cudaMemcpyDeviceToHost));
  *blob = out;
  *n_bytes = patch->saved_state.size();
  return SRK_OK;
}

int engine_state_import(srk_patch* patch, const void* blob, size_t n_bytes) {
  if (!patch->planned) { patch->last_error = "srk_plan() has not been called since the last wiring change"; return SRK_ERR_NOT_PLANNED; }
  StateHeader h;
  if (n_bytes < sizeof h) { patch->last_error = "state blob shorter than its header"; return SRK_ERR_ARG; }
  std::memcpy(&h, blob, sizeof h);
  if (std::memcmp(h.magic, "SRKSTATE", 8) != 0 || h.version != kStateVersion) { patch->last_error = "not a state blob of this version"; return SRK_ERR_ARG; }
  if (h.V == 0 || h.V > (0xFFFFFFFFull / kMaxChunk)) { patch->last_error = "state blob: bad voice count"; return SRK_ERR_ARG; }
  const size_t state_bytes = (size_t)h.S * h.V * sizeof(uint32_t);
  const size_t ring_bytes = (size_t)h.R * h.B * h.V * sizeof(float);
  if (n_bytes != sizeof h + state_bytes + ring_bytes) { patch->last_error = "state blob: size does not match its header"; return SRK_ERR_SIZE; }
  int rc = engine_open(patch);
  if (rc != SRK_OK) return rc;
  Engine& e = *patch->engine;
  SRK_CUDA(cudaSetDevice(e.device));
  rc = engine_prepare(patch, (size_t)h.V, (size_t)h.voice_offset);  // compiles, allocates, X::new() state
  if (rc != SRK_OK) return rc;
  if (h.S != e.prog.state_init.size() || h.R != e.prog.n_rings || h.B != e.prog.ring_len || h.layout != state_layout_hash(*patch, e.prog)) {
    patch->last_error = "state blob was exported from a different patch (module kinds in plan order, cut wires or buffer_size differ)";
    return SRK_ERR_ARG;
  }
  const unsigned char* in = static_cast<const unsigned char*>(blob) + sizeof h;
  if (state_bytes) SRK_CUDA(cudaMemcpyAsync(e.d_state.p, in, state_bytes, cudaMemcpyHostToDevice, e.stream));
  if (ring_bytes) SRK_CUDA(cudaMemcpyAsync(e.d_rings.p, in + state_bytes, ring_bytes, cudaMemcpyHostToDevice, e.stream));
  SRK_CUDA(cudaStreamSynchronize(e.stream));  // the caller's blob may go away
  e.n_abs = h.n_abs;
  e.state_valid = true;
  ++e.state_epoch;
  // the imported play positions supersede a pending "rewind at the next render" of freshly loaded Sample tables
  for (srk_module* m : patch->modules)
    if (m->wave_new) { m->wave_new = false; ++patch->table_epoch; }
  return SRK_OK;
}

int engine_program_dump(srk_patch* patch, size_t n_voices, srk_instr_info* instrs, size_t instr_cap, size_t* n_instr,
                        srk_wire_info* wires, size_t wire_cap, size_t* n_wires) {
  Program prog;
  std::vector<uint4> blob;
  int K = 0;
  int rc = probe_program(patch, n_voices, prog, blob, K);
  if (rc != SRK_OK) return rc;
  if (n_instr) *n_instr = prog.code.size();
  if (n_wires) *n_wires = prog.wires.size();
  for (size_t i = 0; instrs && i < prog.code.size() && i < instr_cap; ++i) {
    const Instr& c = prog.code[i];
    srk_instr_info& o = instrs[i];
    o.op = c.op; o.flags = c.flags; o.warp = c.warp; o.stage = c.stage;
    for (int k = 0; k < 4; ++k) o.in[k] = c.in[k];
    for (int k = 0; k < 3; ++k) o.out[k] = c.out[k];
    o.n_ch = c.n_ch;
    o.state = c.state; o.param = c.param; o.aux = c.aux;
  }
  for (size_t i = 0; wires && i < prog.wires.size() && i < wire_cap; ++i) {
    wires[i].first_tile = prog.wires[i].base;
    wires[i].n_tiles = (uint16_t)(prog.wires[i].mask + 1);
  }
  return SRK_OK;
}

}  // namespace srk
