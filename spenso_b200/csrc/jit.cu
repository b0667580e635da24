// see jit.hpp
#include "jit.hpp"

#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>

#include "tensor.hpp"

namespace spn {

namespace {
// ---- JIT: CUDA source -> cubin (NVRTC, sm_100a) -> cudaKernel_t -------------------------------

uint64_t fnv1a_impl(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : s) {
    h ^= c;
    h *= 1099511628211ull;
  }
  return h;
}

std::string cache_dir() {
  if (const char* e = getenv("SPN_JIT_CACHE")) return e;
  Dl_info info;
  if (dladdr((void*)&fnv1a_impl, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t s = p.find_last_of('/');
    return (s == std::string::npos ? std::string(".") : p.substr(0, s)) + "/_jit_cache";
  }
  return "/tmp/spn_jit_cache";
}

const char* kJitArch = "--gpu-architecture=sm_100a";

// NVRTC is bound at run time by absolute path: a host process may already hold another libnvrtc.so.12
// (PyTorch bundles the CUDA 12.8 one, whose ptxas rejects the 256-bit loads of PTX 8.8), and a link-time
// dependency would silently resolve to that copy.
struct Nvrtc {
  void* h = nullptr;
  int major = 0, minor = 0;
  nvrtcResult (*Version)(int*, int*) = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  bool has_256bit_loads() const { return major > 12 || (major == 12 && minor >= 9); }
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  if (n.h) return n;
  std::vector<std::string> cand;
  if (const char* e = getenv("SPN_NVRTC")) cand.push_back(e);
  cand.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
  cand.push_back("/usr/local/cuda/lib64/libnvrtc.so");
  cand.push_back("libnvrtc.so.12");
  cand.push_back("libnvrtc.so");
  for (auto& c : cand) {
    n.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (n.h) break;
  }
  if (!n.h) throw Error(SPN_ERR_JIT, std::string("cannot load NVRTC: ") + dlerror());
#define SPN_NVRTC_SYM(field, name)                                              \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.h, name));              \
  if (!n.field) throw Error(SPN_ERR_JIT, std::string("NVRTC symbol missing: ") + name)
  SPN_NVRTC_SYM(Version, "nvrtcVersion");
  SPN_NVRTC_SYM(CreateProgram, "nvrtcCreateProgram");
  SPN_NVRTC_SYM(CompileProgram, "nvrtcCompileProgram");
  SPN_NVRTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
  SPN_NVRTC_SYM(GetProgramLog, "nvrtcGetProgramLog");
  SPN_NVRTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
  SPN_NVRTC_SYM(GetCUBIN, "nvrtcGetCUBIN");
  SPN_NVRTC_SYM(DestroyProgram, "nvrtcDestroyProgram");
#undef SPN_NVRTC_SYM
  n.Version(&n.major, &n.minor);
  return n;
}

void jit_compile_impl(JitKernel& k) {
  Nvrtc& rt = nvrtc();
  std::string key = k.source + "|" + kJitArch + "|" + std::to_string(rt.major) + "." + std::to_string(rt.minor);
  char hex[32];
  std::snprintf(hex, sizeof hex, "%016llx", (unsigned long long)fnv1a_impl(key));
  const std::string dir = cache_dir();
  const std::string path = dir + "/" + hex + ".cubin";
  const bool use_cache = !getenv("SPN_JIT_NO_CACHE");
  if (use_cache) {
    std::ifstream f(path, std::ios::binary);
    if (f) {
      k.cubin.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
      if (!k.cubin.empty()) return;
    }
  }
  nvrtcProgram prog;
  if (rt.CreateProgram(&prog, k.source.c_str(), (k.name + ".cu").c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS)
    throw Error(SPN_ERR_JIT, "nvrtcCreateProgram failed");
  const char* opts[] = {kJitArch, "-lineinfo", "--std=c++17", "--fmad=true", "-default-device"};
  nvrtcResult r = rt.CompileProgram(prog, 5, opts);
  if (r != NVRTC_SUCCESS) {
    size_t n = 0;
    rt.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    rt.GetProgramLog(prog, &log[0]);
    rt.DestroyProgram(&prog);
    throw Error(SPN_ERR_JIT, "nvrtc: " + log.substr(0, 2000));
  }
  size_t n = 0;
  rt.GetCUBINSize(prog, &n);
  k.cubin.resize(n);
  rt.GetCUBIN(prog, k.cubin.data());
  rt.DestroyProgram(&prog);
  if (use_cache) {
    mkdir(dir.c_str(), 0755);
    std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    std::ofstream f(tmp, std::ios::binary);
    if (f) {
      f.write(k.cubin.data(), (std::streamsize)k.cubin.size());
      f.close();
      rename(tmp.c_str(), path.c_str());
      std::ofstream s(dir + "/" + hex + ".cu");
      if (s) s << k.source;
    }
  }
}

void jit_load_impl(JitKernel& k, int block) {
  if (k.kernel) return;
  cuda_check(cudaLibraryLoadData(&k.lib, k.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0),
             "cudaLibraryLoadData (fused network kernel)");
  cuda_check(cudaLibraryGetKernel(&k.kernel, k.lib, k.name.c_str()), "cudaLibraryGetKernel");
  cudaFuncAttributes attr{};
  if (cudaFuncGetAttributes(&attr, (const void*)k.kernel) == cudaSuccess) k.num_regs = attr.numRegs;
  if (k.smem_bytes >= 48 * 1024) {
    int dev = 0;
    cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
    cuda_check(cudaKernelSetAttributeForDevice(k.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem_bytes, dev),
               "cudaKernelSetAttributeForDevice(MaxDynamicSharedMemorySize)");
  }
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)k.kernel, block, k.smem_bytes) != cudaSuccess || nb <= 0) {
    cudaGetLastError();
    nb = k.num_regs > 0 ? std::max(1, 65536 / (k.num_regs * block)) : 4;
    if (k.smem_bytes) nb = std::max(1, std::min(nb, (int)(200 * 1024 / k.smem_bytes)));
  }
  k.blocks_per_sm = nb;
}

}  // namespace

uint64_t fnv1a(const std::string& s) { return fnv1a_impl(s); }
void jit_compile(JitKernel& k) { jit_compile_impl(k); }
void jit_load(JitKernel& k, int block) { jit_load_impl(k, block); }
bool jit_has_256bit_loads() { return nvrtc().has_256bit_loads(); }

}  // namespace spn
