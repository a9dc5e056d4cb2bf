// Fused (patch-specialised) voice kernels: generation (fused_gen.cpp) and the runtime that turns the
// generated source into a loaded kernel (fused_rt.cpp: in-memory cache -> cubin cache on disk -> NVRTC).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "fused_args.h"
#include "patch.hpp"
#include "program.hpp"

namespace srk {

constexpr int kFusedMaxThreads = 128;  // __launch_bounds__ of every fused kernel; blocks of 1..4 warps are launched

struct FusedSpec {
  std::string source;            // the generated translation unit (#includes "fused_ops.cuh")
  std::vector<uint8_t> uniform;  // per parameter word: 1 = the kernel reads args.u[w], 0 = params[w][voice]
  int n_distinct = 0;            // distinct wires feeding the Output module (one shared-memory tile pair each)
  int channels = 0;
  int n_ssa = 0;                 // wires (one local array each)
  int group = 4;                 // samples per straight-line group
  int min_blocks = 4;            // second __launch_bounds__ argument (register cap = 65536 / (128 * min_blocks))
  int stages = 1;                // warps per voice group: consecutive slices of the patch, one tile apart
  bool split_moog = false;       // the ladder filters' coefficient blocks are ops of their own
  int tile = 32;                 // samples per output / cross-stage tile
  double max_stage_cost = 0.0;   // cost model: instructions per voice-sample of the slowest stage
  int n_cross = 0;               // wires that cross a stage boundary
  int n_cross_tiles = 0;         // ... and the tiles of their rings (stages spanned + 1 each)
  size_t smem_per_group = 0;
  int compiler = 0;              // which NVRTC builds it (fused_rt.cpp: 0 = the toolkit's, 1 = another version, if present)
};

struct FusedOptions {
  int group = 4;       // samples per straight-line group
  int min_blocks = 4;
  int stages = 1;      // at most; the generator picks the count whose slowest stage is cheapest
  bool exact_stages = false;  // ... unless told to use exactly that many (experiments, tests)
  int tile = 32;
  bool split_moog = true;  // staged kernels: the ladder filter's coefficient block as an op of its own (another stage)
  bool prefetch = false;   // staged kernels: a stage loads its cross-stage inputs one sample group ahead (measured: slower)
};

// `prog` must be the one-warp program of the planned patch (compile_program(patch, 1, ...)).
int fused_generate(const srk_patch& patch, const Program& prog, const FusedOptions& opt, FusedSpec& out, std::string& err);

std::string fused_hash(const std::string& text, const std::string& salt);
// cache key of a generated kernel: hash of its source, the op headers and the compiler options
std::string fused_key(const FusedSpec& spec);
int fused_compilers();                           // NVRTC versions available: 0, 1 or 2
std::string fused_compiler_name(int compiler);   // "nvrtc 12.9"

struct FusedKernel {
  void* library = nullptr;  // cudaLibrary_t
  void* kernel = nullptr;   // cudaKernel_t (accepted by cudaLaunchKernel / cudaFuncSetAttribute as is)
  std::string key;
  int regs = 0;
  size_t local_bytes = 0;   // spills show up here
  size_t max_smem_set = 0;
  bool from_disk = false;
  double compile_ms = 0.0;
};

// cubin for `spec` (cache key included): from the disk cache, else compiled with NVRTC for sm_100a and stored.
// Needs no GPU.  SRK_OK or SRK_ERR_UNSUPPORTED (no NVRTC on this machine) / SRK_ERR_LIMIT (compile error: log in err).
int fused_cubin(const FusedSpec& spec, std::vector<char>& cubin, std::string& key, bool* from_disk, double* compile_ms, std::string& err);

// The loaded kernel for `spec` on the current device (cached per process by key).
int fused_kernel(const FusedSpec& spec, const FusedKernel** out, std::string& err);

// cuTensorMapEncodeTiled for the f32 [C][N][V] stems tensor with a {32, 32, 1} box.
int fused_stems_map(SrkTensorMap* map, float* stems, uint64_t C, uint64_t N, uint64_t V, unsigned tile_rows, std::string& err);

std::string fused_cache_dir();
std::string fused_tuned_dir();  // read-only: schedule decisions shipped with the library (tuned/ next to it, or $SRK_TUNED_DIR)

}  // namespace srk
