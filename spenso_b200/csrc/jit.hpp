// JIT plumbing shared by the fused network compiler (net.cu) and the specialised pairwise kernels
// (pairwise_jit.cu): generated CUDA C -> cubin for sm_100a with NVRTC (-lineinfo), an on-disk cubin cache next to
// the library (spenso_b200/_jit_cache, pre-filled by build()), loading through cudaLibraryLoadData.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace spn {

struct JitKernel {
  std::string source, name;
  std::vector<char> cubin;
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  int blocks_per_sm = 0;
  int block = 0;          // threads per CTA this kernel was generated for (0: the caller's default)
  bool flat = false;      // launched with one point per thread (hardware-scheduled CTAs) instead of one persistent wave
  int num_regs = 0;
  size_t smem_bytes = 0;  // dynamic shared memory (prefetch ring)
  ~JitKernel() {
    if (lib) cudaLibraryUnload(lib);
  }
};

uint64_t fnv1a(const std::string& s);
// compile k.source (entry point k.name) to k.cubin, through the cubin cache; throws Error(SPN_ERR_JIT)
void jit_compile(JitKernel& k);
// load the cubin on the current device and query registers / occupancy for `block` threads per CTA
void jit_load(JitKernel& k, int block);
// the NVRTC in use understands the 256-bit ld.global.nc.v4.f64 of PTX 8.8 (CUDA >= 12.9)
bool jit_has_256bit_loads();

}  // namespace spn
