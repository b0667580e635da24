// Probe: what the FP64 (non-tensor) pipe of a B200 SM sustains, by instruction mix, chains per thread and resident warps.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe_probe fp64_pipe_probe.cu
// Prints warp-level FP64 instructions per second and the fraction of 148 SMs x 64 lanes/clk x the measured SM clock.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int ILP, int MIX>   // MIX 0: DFMA only; 1: DMUL + DADD alternating; 2: DFMA with one integer op between;
                              // 3: DFMA whose three operands are three different live registers (the shape of real code)
__global__ void __launch_bounds__(256) k(double* out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  unsigned n = threadIdx.x;
  double y[ILP], z[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { y[i] = 1.0 + a * (i + 1) * 1e-9 + threadIdx.x * 1e-12; z[i] = b * (i + 1); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MIX == 0) x[i] = fma(x[i], a, b);
        if (MIX == 1) x[i] = (r & 1) ? __dmul_rn(x[i], a) : __dadd_rn(x[i], b);
        if (MIX == 2) { x[i] = fma(x[i], a, b); n = n * 1664525u + 1013904223u; }
        if (MIX == 3) x[i] = fma(x[i], y[(i + r) % ILP], z[(i + 2 * r + 1) % ILP]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (MIX == 3)
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += y[i] + z[i];
  if (s == 12345.678 || n == 7u) out[0] = s;   // keep the chains alive
}

template <int ILP, int MIX>
static void run(const char* what, int ctas_per_sm, int threads, int sms, double clk_hz) {
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP, MIX><<<sms * ctas_per_sm, threads>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e0);
  k<ILP, MIX><<<sms * ctas_per_sm, threads>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double lane_instr = (double)sms * ctas_per_sm * threads * iters * 8.0 * ILP;
  const double rate = lane_instr / (ms * 1e-3);
  printf("%-22s ILP %2d  %2d warps/SM  %8.3f ms  %.3e lane-instr/s  = %.3f of %d SMs x 64 x %.0f MHz\n", what, ILP,
         ctas_per_sm * threads / 32, ms, rate, rate / (sms * 64.0 * clk_hz), sms, clk_hz / 1e6);
  cudaFree(out);
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  for (int w : {1, 2, 4}) {   // CTAs of 256 threads per SM: 8 / 16 / 32 warps
    run<1, 0>("DFMA", w, 256, sms, clk);
    run<4, 0>("DFMA", w, 256, sms, clk);
    run<8, 0>("DFMA", w, 256, sms, clk);
    run<16, 0>("DFMA", w, 256, sms, clk);
    run<8, 1>("DMUL/DADD", w, 256, sms, clk);
    run<8, 2>("DFMA + IMAD", w, 256, sms, clk);
    run<8, 3>("DFMA 3 live operands", w, 256, sms, clk);
    run<16, 3>("DFMA 3 live operands", w, 256, sms, clk);
  }
  return 0;
}
