// Device side of the renderer: the voice kernel (one resident thread per voice,
// the whole sample loop in-kernel), the deterministic mixdown, and the Engine that
// owns HBM state for one patch.  sm_100a only; there is no CPU path.
//
// HBM layout (all structure-of-arrays, voice index fastest => every warp access is
// one contiguous 128-byte line):
//   state   u32 [S][V]      per-voice module state, loaded to shared memory at kernel
//                           start, stored back at kernel end
//   params  u32 [P][V]      per-voice parameters (uniform ones broadcast at upload)
//   rings   f32 [R][B][V]   history of delayed (feedback) wires, B = buffer_size
//   stems   f32 [C][N][V]   optional per-voice output
//   partial f32 [G][C][N]   per-block mix partials (G = grid size), reduced in a
//                           fixed order by mix_reduce_kernel => bit-reproducible mix
// Shared memory per block (T threads, K samples per inner step):
//   program (32 B / instr, staged once) | state [S][T] | params [P][T] |
//   wires [W][K][T] | reduction scratch [T]
#include "engine.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dsp.cuh"

namespace srk {

struct RenderArgs {
  const Instr* prog;
  uint32_t* state;
  const uint32_t* params;
  float* rings;
  float* stems;
  float* partial;
  uint32_t n_instr;
  uint32_t V;             // voices rendered by this launch
  uint32_t voice_offset;  // global index of voice 0 (noise key)
  uint32_t n_samples;
  uint32_t S, P, W, C, B;
  uint32_t step;          // samples per inner step, <= K and <= B
  uint32_t ring_phase;    // absolute sample index of sample 0, mod B
  uint32_t seed_lo, seed_hi;
};

template <int K>
__global__ void __launch_bounds__(256) render_voices_kernel(const RenderArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  Instr* prog = reinterpret_cast<Instr*>(smem_raw);
  uint32_t* st = reinterpret_cast<uint32_t*>(prog + a.n_instr);
  uint32_t* pr = st + (size_t)a.S * T;
  float* wires = reinterpret_cast<float*>(pr + (size_t)a.P * T);
  float* red = wires + (size_t)a.W * K * T;

  // stage the patch program (port/wire table) once per block
  for (uint32_t i = tid; i < a.n_instr * 2; i += T)
    reinterpret_cast<uint4*>(prog)[i] = reinterpret_cast<const uint4*>(a.prog)[i];
  const uint32_t v_raw = blockIdx.x * T + tid;
  const bool active = v_raw < a.V;
  const uint32_t v = active ? v_raw : a.V - 1;  // idle lanes shadow the last voice, never store
  for (uint32_t w = 0; w < a.S; ++w) st[w * T + tid] = a.state[(size_t)w * a.V + v];
  for (uint32_t w = 0; w < a.P; ++w) pr[w * T + tid] = a.params[(size_t)w * a.V + v];
  __syncthreads();

  const dsp::Lane L{st + tid, pr + tid, wires + tid, T};
  float mix_prev = 0.0f;
  for (uint32_t n0 = 0; n0 < a.n_samples; n0 += a.step) {
    const int kk = (int)min(a.step, a.n_samples - n0);
    for (uint32_t pc = 0;; ++pc) {
      const Instr ins = prog[pc];
      if (ins.op == OP_END) break;
      switch (ins.op) {
        case OP_OSC: dsp::op_osc<K>(ins, L, kk); break;
        case OP_MOOG: dsp::op_moog<K>(ins, L, kk); break;
        case OP_ADSR: dsp::op_adsr<K>(ins, L, kk); break;
        case OP_VCA: dsp::op_vca<K>(ins, L, kk); break;
        case OP_MIXER: dsp::op_mixer<K>(ins, L, kk); break;
        case OP_MATH: dsp::op_math<K>(ins, L, kk); break;
        case OP_NOISE: dsp::op_noise<K>(ins, L, kk, a.voice_offset + v, a.seed_lo, a.seed_hi); break;
        case OP_RING_LOAD: {
          const float* ring = a.rings + (size_t)ins.aux * a.B * a.V;
          float* out = dsp::wire<K>(L, ins.out[0]);
          uint32_t idx = (a.ring_phase + n0) % a.B;
          for (int k = 0; k < kk; ++k) {
            out[k * T] = ring[(size_t)idx * a.V + v];
            idx = idx + 1 == a.B ? 0 : idx + 1;
          }
          break;
        }
        case OP_RING_STORE: {
          float* ring = a.rings + (size_t)ins.aux * a.B * a.V;
          const float* in = dsp::wire<K>(L, ins.in[0]);
          uint32_t idx = (a.ring_phase + n0) % a.B;
          for (int k = 0; k < kk; ++k) {
            if (active) ring[(size_t)idx * a.V + v] = in[k * T];
            idx = idx + 1 == a.B ? 0 : idx + 1;
          }
          break;
        }
        case OP_OUTPUT: {  // OutputModule::calc, src/synth/output.rs:46-60, + mixdown
          const uint32_t c = ins.aux;
          const float* src = dsp::wire<K>(L, ins.in[0]);
          if (a.stems && active) {
            float* dst = a.stems + ((size_t)c * a.n_samples + n0) * a.V + v;
            for (int k = 0; k < kk; ++k) dst[(size_t)k * a.V] = src ? src[k * T] : 0.0f;
          }
          if (a.partial) {
            float* dst = a.partial + ((size_t)blockIdx.x * a.C + c) * a.n_samples + n0;
            if (!src) {
              if (tid < kk) dst[tid] = 0.0f;
            } else if (ins.flags & F_OUT_SAME_AS_PREV) {
              if (tid < kk) dst[tid] = mix_prev;
            } else {
              __syncthreads();  // every column of the wire is final
              // thread (k, seg) adds K columns of sample row k; the rotated start column
              // keeps the 32 lanes on 32 different banks
              const int k = tid & (K - 1), seg = tid / K;
              const float* row = wires + ((size_t)ins.in[0] * K + k) * T + seg * K;
              const uint32_t col0 = blockIdx.x * T + seg * K;
              float acc = 0.0f;
#pragma unroll
              for (int j = 0; j < K; ++j) {
                const int jj = (j + k) & (K - 1);
                const float x = row[jj];
                acc = dsp::fadd(acc, col0 + jj < a.V ? x : 0.0f);
              }
              red[tid] = acc;
              __syncthreads();
              if (tid < K) {
                float sum = 0.0f;
                for (int s = 0; s < T / K; ++s) sum = dsp::fadd(sum, red[s * K + tid]);
                mix_prev = sum;
                if (tid < kk) dst[tid] = sum;
              }
              __syncthreads();  // scratch and wire may be rewritten from here on
            }
          }
          break;
        }
        default: break;
      }
    }
  }
  if (active)
    for (uint32_t w = 0; w < a.S; ++w) a.state[(size_t)w * a.V + v] = st[w * T + tid];
}

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
int compile_program(const srk_patch& patch, Program& prog, std::string& err);

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

struct Engine {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // call start, kernel start, kernel end, call end
  bool timed = false;
  Program prog;
  uint64_t compiled_epoch = 0, uploaded_param_epoch = 0;
  size_t V = 0, voice_offset = 0;
  uint64_t n_abs = 0;  // samples rendered since reset
  bool state_valid = false;
  DevBuf d_prog, d_state, d_state_init, d_params, d_rings, d_partial, d_stems, d_mix;
  uint32_t* h_params = nullptr;  // pinned staging
  size_t h_params_bytes = 0;
  uint64_t launches = 0;
  int smem_optin = 0, n_sm = 0;
  // geometry of the last launch
  int block_threads = 0, step = 0;
  size_t smem_bytes = 0;

  ~Engine() {
    if (device >= 0) {
      cudaSetDevice(device);
      for (auto* b : {&d_prog, &d_state, &d_state_init, &d_params, &d_rings, &d_partial, &d_stems, &d_mix}) b->release();
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
  patch->engine = std::move(eng);
  return SRK_OK;
}

// Threads per block / samples per step for V voices of this program.
static void choose_geometry(const Engine& e, const Program& prog, size_t V, int& T, int& K, size_t& smem) {
  (void)V;
  T = env_int("SRK_BLOCK_THREADS", 32);
  K = env_int("SRK_STEP", 16);
  if (K != 8 && K != 16 && K != 32) K = 16;
  if (T < 32) T = 32;
  if (T > 256) T = 256;
  T = (T + 31) / 32 * 32;
  if (T < K) T = K;
  auto bytes = [&](int t, int k) {
    return prog.code.size() * sizeof(Instr) +
           ((size_t)prog.state_init.size() + prog.param_src.size() + (size_t)prog.n_wires * k + 1) * t * sizeof(uint32_t);
  };
  while (bytes(T, K) > (size_t)e.smem_optin && K > 8) K /= 2;
  while (bytes(T, K) > (size_t)e.smem_optin && T > 32) T -= 32;
  if (T < K) T = K;
  smem = bytes(T, K);
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
  return SRK_OK;
}

// (Re)compile + (re)allocate for the current plan and voice range.
static int engine_prepare(srk_patch* patch, size_t n_voices, size_t voice_offset) {
  Engine& e = *patch->engine;
  bool fresh = false;
  if (e.compiled_epoch != patch->wiring_epoch) {
    std::string err;
    int rc = compile_program(*patch, e.prog, err);
    if (rc != SRK_OK) { patch->last_error = err; return rc; }
    SRK_CUDA(e.d_prog.ensure(e.prog.code.size() * sizeof(Instr)));
    SRK_CUDA(cudaMemcpyAsync(e.d_prog.p, e.prog.code.data(), e.prog.code.size() * sizeof(Instr),
                             cudaMemcpyHostToDevice, e.stream));
    SRK_CUDA(e.d_state_init.ensure(std::max<size_t>(e.prog.state_init.size(), 1) * sizeof(uint32_t)));
    if (!e.prog.state_init.empty())
      SRK_CUDA(cudaMemcpyAsync(e.d_state_init.p, e.prog.state_init.data(), e.prog.state_init.size() * sizeof(uint32_t),
                               cudaMemcpyHostToDevice, e.stream));
    // the host vectors must outlive the async copies
    SRK_CUDA(cudaStreamSynchronize(e.stream));
    e.compiled_epoch = patch->wiring_epoch;
    fresh = true;
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
    int rc = build_param_table(patch, e);
    if (rc != SRK_OK) return rc;
    SRK_CUDA(cudaStreamSynchronize(e.stream));  // staging buffer is reused by the next upload
    e.uploaded_param_epoch = patch->param_epoch;
  }
  return SRK_OK;
}

template <int K>
static cudaError_t launch_voices(const RenderArgs& a, unsigned grid, int T, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(render_voices_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  render_voices_kernel<K><<<grid, T, smem, s>>>(a);
  return cudaGetLastError();
}

int engine_render(srk_patch* patch, size_t n_voices, size_t voice_offset, size_t n_samples, unsigned flags,
                  float* stems, float* mix, void* user_stream, bool use_user_stream) {
  if (!patch->planned) { patch->last_error = "srk_plan() has not been called since the last wiring change"; return SRK_ERR_NOT_PLANNED; }
  if (n_voices > 0xFFFFFFFFull || n_samples > 0xFFFFFFFFull) { patch->last_error = "n_voices / n_samples exceed 2^32-1"; return SRK_ERR_ARG; }
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
  if (n_voices == 0 || n_samples == 0) { e.timed = false; return SRK_OK; }

  const Program& prog = e.prog;
  const size_t C = prog.channels;
  int T, K;
  size_t smem;
  choose_geometry(e, prog, n_voices, T, K, smem);
  if (smem > (size_t)e.smem_optin) { patch->last_error = "patch needs more shared memory than one block can have"; return SRK_ERR_LIMIT; }
  const unsigned grid = (unsigned)((n_voices + T - 1) / T);
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
    SRK_CUDA(e.d_partial.ensure((size_t)grid * C * n_samples * sizeof(float)));
  }

  RenderArgs a{};
  a.prog = (const Instr*)e.d_prog.p;
  a.state = (uint32_t*)e.d_state.p;
  a.params = (const uint32_t*)e.d_params.p;
  a.rings = (float*)e.d_rings.p;
  a.stems = d_stems;
  a.partial = mix ? (float*)e.d_partial.p : nullptr;
  a.n_instr = (uint32_t)prog.code.size();
  a.V = (uint32_t)n_voices;
  a.voice_offset = (uint32_t)voice_offset;
  a.n_samples = (uint32_t)n_samples;
  a.S = (uint32_t)prog.state_init.size();
  a.P = (uint32_t)prog.param_src.size();
  a.W = prog.n_wires;
  a.C = (uint32_t)C;
  a.B = std::max<uint32_t>(prog.ring_len, 1);
  a.step = (uint32_t)std::min<size_t>(K, prog.n_rings ? a.B : (size_t)K);
  a.ring_phase = (uint32_t)(e.n_abs % a.B);
  a.seed_lo = (uint32_t)patch->seed;
  a.seed_hi = (uint32_t)(patch->seed >> 32);

  SRK_CUDA(cudaEventRecord(e.ev[1], work));
  cudaError_t le = K == 8 ? launch_voices<8>(a, grid, T, smem, work)
                   : K == 16 ? launch_voices<16>(a, grid, T, smem, work)
                             : launch_voices<32>(a, grid, T, smem, work);
  SRK_CUDA(le);
  ++e.launches;
  SRK_CUDA(cudaEventRecord(e.ev[2], work));
  if (mix) {
    const size_t cn = C * n_samples;
    mix_reduce_kernel<<<(unsigned)((cn + 255) / 256), 256, 0, work>>>((const float*)e.d_partial.p, d_mix, grid, cn);
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
  e.block_threads = T;
  e.step = (int)a.step;
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

int engine_program_info(srk_patch* patch, size_t n_voices, srk_program_info* out) {
  if (!patch->planned) { patch->last_error = "not planned"; return SRK_ERR_NOT_PLANNED; }
  Program prog;
  std::string err;
  int rc = compile_program(*patch, prog, err);
  if (rc != SRK_OK) { patch->last_error = err; return rc; }
  Engine probe;  // geometry without a device: assume the sm_100 opt-in limit
  probe.smem_optin = patch->engine ? patch->engine->smem_optin : 227 * 1024;
  int T, K;
  size_t smem;
  choose_geometry(probe, prog, n_voices, T, K, smem);
  out->n_instr = (uint32_t)prog.code.size();
  out->step_samples = (uint32_t)std::min<size_t>(K, prog.n_rings ? std::max<uint32_t>(prog.ring_len, 1) : (size_t)K);
  out->block_threads = (uint32_t)T;
  out->smem_bytes = (uint32_t)smem;
  out->n_wires = prog.n_wires;
  out->state_words = (uint32_t)prog.state_init.size();
  out->param_words = (uint32_t)prog.param_src.size();
  out->n_rings = prog.n_rings;
  return SRK_OK;
}

}  // namespace srk
