// Runtime of the fused voice kernels: generated source -> cubin (disk cache, else NVRTC for sm_100a) -> loaded
// kernel (cudaLibraryLoadData).  NVRTC is dlopen()ed so that the library itself loads on machines without it;
// when it is missing the engine keeps using the interpreter kernels (voice_kernel.cuh), still on the GPU.
#include "fused.hpp"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>

#include "fused_embed.inc"  // kFusedOpsSrc, kFusedArgsSrc, kFusedLibmSrc: the headers as strings (Makefile)

namespace srk {

namespace {

const char* const kNvrtcOptions[] = {
    "--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "--ftz=false", "--prec-div=true", "--prec-sqrt=true",
    "-lineinfo",
};
constexpr int kNumNvrtcOptions = sizeof(kNvrtcOptions) / sizeof(kNvrtcOptions[0]);

struct Nvrtc {
  void* handle = nullptr;
  decltype(&nvrtcCreateProgram) create = nullptr;
  decltype(&nvrtcCompileProgram) compile = nullptr;
  decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
  decltype(&nvrtcGetCUBIN) cubin = nullptr;
  decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
  decltype(&nvrtcGetProgramLog) log = nullptr;
  decltype(&nvrtcDestroyProgram) destroy = nullptr;
  decltype(&nvrtcVersion) version = nullptr;
  std::string why;
  bool ok() const { return handle != nullptr; }
};

void bind(Nvrtc& r, const std::vector<std::string>& names, const char* missing) {
  for (const std::string& name : names) {
    r.handle = dlopen(name.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (r.handle) break;
  }
  if (!r.handle) { r.why = missing; return; }
#define SRK_SYM(field, name)                                                   \
  r.field = reinterpret_cast<decltype(r.field)>(dlsym(r.handle, #name));       \
  if (!r.field) { r.why = "libnvrtc lacks " #name; r.handle = nullptr; return; }
  SRK_SYM(create, nvrtcCreateProgram)
  SRK_SYM(compile, nvrtcCompileProgram)
  SRK_SYM(cubin_size, nvrtcGetCUBINSize)
  SRK_SYM(cubin, nvrtcGetCUBIN)
  SRK_SYM(log_size, nvrtcGetProgramLogSize)
  SRK_SYM(log, nvrtcGetProgramLog)
  SRK_SYM(destroy, nvrtcDestroyProgram)
  SRK_SYM(version, nvrtcVersion)
#undef SRK_SYM
}

std::string version_of(const Nvrtc& rt) {
  int major = 0, minor = 0;
  if (!rt.ok() || rt.version(&major, &minor) != NVRTC_SUCCESS) return "";
  return std::to_string(major) + "." + std::to_string(minor);
}

// Compiler 0: the toolkit's library by PATH before any bare name.  A bare "libnvrtc.so.12" resolves to whatever copy the
// process has already mapped -- `import torch` maps the 12.8 one of its own wheel -- and the same kernel id would then
// name two different cubins (seen: 128 registers without and with spills).  The version is part of the id (salt()).
// Compiler 1, when there is one: ANOTHER version of the library -- SRK_NVRTC_ALT, else whatever the bare soname gives
// if that differs from compiler 0.  Measured on B200 (profiles/r06h_tune_all*.txt): 12.8 schedules the staged kernels
// better (cfg2 @ 4096 2.50 against 2.59 ms, cfg4 @ 16384 12.9 against 15.1), 12.9 the one-warp sine kernels (gated sine
// @ 32768 7.2 against 7.9), so the compiler is one more dimension of the measured schedule choice (engine.cu).
const Nvrtc& nvrtc(int which) {
  static Nvrtc first = [] {
    Nvrtc r;
    std::vector<std::string> names;
    if (const char* e = std::getenv("SRK_NVRTC_LIB")) names.push_back(e);
    for (const char* s : {"/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so"})
      names.push_back(s);
    bind(r, names, "libnvrtc.so.12 not found (set SRK_NVRTC_LIB)");
    return r;
  }();
  if (which == 0) return first;
  static Nvrtc second = [] {
    Nvrtc r;
    const char* off = std::getenv("SRK_NVRTC_ALT");
    if (!first.ok() || (off && off[0] == '0' && !off[1])) { r.why = "no alternative compiler"; return r; }
    std::vector<std::string> names;
    if (off && *off) names.push_back(off);
    names.push_back("libnvrtc.so.12");
    bind(r, names, "no alternative compiler");
    if (r.ok() && (r.handle == first.handle || version_of(r) == version_of(first) || version_of(r).empty())) {
      r.handle = nullptr;
      r.why = "no alternative compiler";
    }
    return r;
  }();
  return second;
}

std::string salt(int compiler) {
  std::string s(kFusedOpsSrc);
  s += kFusedArgsSrc;
  s += kFusedLibmSrc;
  for (int i = 0; i < kNumNvrtcOptions; ++i) { s += kNvrtcOptions[i]; s += ' '; }
  s += "nvrtc " + version_of(nvrtc(compiler));  // (no NVRTC: nothing gets compiled under the id anyway)
  return s;
}

bool read_file(const std::string& path, std::vector<char>& data) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  data.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return !data.empty();
}

void write_file_atomic(const std::string& path, const std::vector<char>& data) {
  const std::string tmp = path + ".tmp." + std::to_string((long)getpid());
  {
    std::ofstream f(tmp, std::ios::binary);
    if (!f) return;
    f.write(data.data(), (std::streamsize)data.size());
    if (!f) { std::remove(tmp.c_str()); return; }
  }
  if (std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}

}  // namespace

std::string fused_cache_dir() {
  if (const char* e = std::getenv("SRK_KERNEL_CACHE")) return e;
  Dl_info info;
  std::string dir = ".";
  if (dladdr(reinterpret_cast<const void*>(&fused_cache_dir), &info) && info.dli_fname) {
    dir = info.dli_fname;
    const size_t slash = dir.rfind('/');
    dir = slash == std::string::npos ? "." : dir.substr(0, slash);
  }
  return dir + "/kernel_cache";
}

std::string fused_key(const FusedSpec& spec) {
  static const std::string kSalt[2] = {salt(0), salt(1)};
  return fused_hash(spec.source, kSalt[spec.compiler ? 1 : 0]);
}

int fused_compilers() { return nvrtc(1).ok() ? 2 : (nvrtc(0).ok() ? 1 : 0); }

std::string fused_compiler_name(int compiler) { return "nvrtc " + version_of(nvrtc(compiler ? 1 : 0)); }

std::string fused_tuned_dir() {
  if (const char* e = std::getenv("SRK_TUNED_DIR")) return e;
  Dl_info info;
  std::string dir = ".";
  if (dladdr(reinterpret_cast<const void*>(&fused_tuned_dir), &info) && info.dli_fname) {
    dir = info.dli_fname;
    const size_t slash = dir.rfind('/');
    dir = slash == std::string::npos ? "." : dir.substr(0, slash);
  }
  return dir + "/tuned";
}

int fused_cubin(const FusedSpec& spec, std::vector<char>& cubin, std::string& key, bool* from_disk, double* compile_ms, std::string& err) {
  key = fused_key(spec);
  const std::string dir = fused_cache_dir();
  const std::string path = dir + "/" + key + ".cubin";
  if (from_disk) *from_disk = false;
  if (compile_ms) *compile_ms = 0.0;
  const char* nocache = std::getenv("SRK_KERNEL_CACHE_OFF");
  if (!(nocache && nocache[0] == '1') && read_file(path, cubin)) {
    if (from_disk) *from_disk = true;
    return SRK_OK;
  }
  const Nvrtc& rt = nvrtc(spec.compiler ? 1 : 0);
  if (!rt.ok()) { err = "fused kernels unavailable: " + rt.why; return SRK_ERR_UNSUPPORTED; }
  const auto t0 = std::chrono::steady_clock::now();
  nvrtcProgram prog = nullptr;
  const char* headers[] = {kFusedOpsSrc, kFusedArgsSrc, kFusedLibmSrc};
  const char* names[] = {"fused_ops.cuh", "fused_args.h", "libm_glibc.cuh"};
  if (rt.create(&prog, spec.source.c_str(), "srk_fused.cu", 3, headers, names) != NVRTC_SUCCESS) {
    err = "nvrtcCreateProgram failed";
    return SRK_ERR_LIMIT;
  }
  const nvrtcResult rc = rt.compile(prog, kNumNvrtcOptions, kNvrtcOptions);
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    rt.log_size(prog, &n);
    std::string log(n, '\0');
    if (n) rt.log(prog, &log[0]);
    rt.destroy(&prog);
    err = "NVRTC could not compile the fused kernel:\n" + log;
    if (const char* dump = std::getenv("SRK_FUSED_DUMP")) {
      std::ofstream f(std::string(dump) + "/" + key + ".failed.cu");
      f << spec.source;
    }
    return SRK_ERR_LIMIT;
  }
  size_t n = 0;
  rt.cubin_size(prog, &n);
  cubin.resize(n);
  rt.cubin(prog, cubin.data());
  rt.destroy(&prog);
  if (compile_ms) *compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  mkdir(dir.c_str(), 0755);
  write_file_atomic(path, cubin);
  if (const char* dump = std::getenv("SRK_FUSED_DUMP")) {
    std::ofstream f(std::string(dump) + "/" + key + ".cu");
    f << spec.source;
  }
  return SRK_OK;
}

int fused_kernel(const FusedSpec& spec, const FusedKernel** out, std::string& err) {
  static std::mutex mu;
  static std::map<std::pair<int, std::string>, std::unique_ptr<FusedKernel>> loaded;  // (device, key)
  std::vector<char> cubin;
  std::string key;
  bool from_disk = false;
  double ms = 0.0;
  key = fused_key(spec);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { err = "no current CUDA device"; return SRK_ERR_CUDA; }
  std::lock_guard<std::mutex> lock(mu);
  auto it = loaded.find({dev, key});
  if (it != loaded.end()) { *out = it->second.get(); return SRK_OK; }
  int rc = fused_cubin(spec, cubin, key, &from_disk, &ms, err);
  if (rc != SRK_OK) return rc;
  auto k = std::make_unique<FusedKernel>();
  cudaLibrary_t lib = nullptr;
  cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (e != cudaSuccess) { err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); return SRK_ERR_CUDA; }
  cudaKernel_t fn = nullptr;
  e = cudaLibraryGetKernel(&fn, lib, "srk_fused_kernel");
  if (e != cudaSuccess) { cudaLibraryUnload(lib); err = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e); return SRK_ERR_CUDA; }
  cudaFuncAttributes attr{};
  if (cudaFuncGetAttributes(&attr, (const void*)fn) == cudaSuccess) {
    k->regs = attr.numRegs;
    k->local_bytes = attr.localSizeBytes;
  }
  k->library = lib;
  k->kernel = fn;
  k->key = key;
  k->from_disk = from_disk;
  k->compile_ms = ms;
  *out = k.get();
  loaded[{dev, key}] = std::move(k);
  return SRK_OK;
}

int fused_stems_map(SrkTensorMap* map, float* stems, uint64_t C, uint64_t N, uint64_t V, unsigned tile_rows, std::string& err) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (!encode) { err = "cuTensorMapEncodeTiled is not available from this driver"; return SRK_ERR_UNSUPPORTED; }
  static_assert(sizeof(SrkTensorMap) == sizeof(CUtensorMap), "tensor map size");
  const cuuint64_t dims[3] = {V, N, C};
  const cuuint64_t strides[2] = {V * sizeof(float), N * V * sizeof(float)};
  const cuuint32_t box[3] = {32, tile_rows, 1};
  const cuuint32_t elem[3] = {1, 1, 1};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, stems, dims, strides, box, elem,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return SRK_ERR_CUDA; }
  return SRK_OK;
}

}  // namespace srk
