// Do DMMA (tensor sub-pipe) and DFMA (FP64 ALU) share one datapath on B200?  Time DMMA-only, DFMA-only and an interleaved mix.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ND, int NF>   // per inner step: ND DMMAs + NF DFMAs (independent accumulators)
__global__ void k(double* out, int iters) {
  double acc[16][2], f[32];
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  for (int i = 0; i < 32; ++i) f[i] = threadIdx.x * 1e-3 + i;
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < ND; ++i) dmma884(acc[i][0], acc[i][1], a, b);
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0.0;
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  for (int i = 0; i < 32; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ND, int NF>
void run(const char* name, double* out) {
  const int iters = 2000, blocks = 148, threads = 512;
  k<ND, NF><<<blocks, threads>>>(out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<ND, NF><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = blocks * threads / 32.0;
  const double fl_d = warps * iters * 4.0 * ND * 512.0, fl_f = warps * iters * 4.0 * NF * 64.0;   // DMMA 8x8x4 = 256 FMA = 512 flop per warp; DFMA 32 lanes x 2
  printf("%-22s %.3f ms  DMMA %.1f TF/s + DFMA %.1f TF/s = %.1f TF/s\n", name, ms, fl_d / ms * 1e-9, fl_f / ms * 1e-9, (fl_d + fl_f) / ms * 1e-9);
}
int main() {
  double* out; cudaMalloc(&out, 8 * 148 * 512);
  run<16, 0>("DMMA only", out);
  run<0, 32>("DFMA only", out);
  run<16, 8>("16 DMMA + 8 DFMA", out);
  run<16, 32>("16 DMMA + 32 DFMA", out);
  run<8, 32>("8 DMMA + 32 DFMA", out);
  run<16, 16>("16 DMMA + 16 DFMA", out);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
