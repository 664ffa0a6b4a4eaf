// dependent-chain latencies on sm_100a: DFMA, DADD, DMUL, SHFL(64-bit)+DADD, F2F, MUFU.RSQ(f32), FFMA, LDS
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0, int iters) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = x0;
  __syncthreads();
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001, z = 1e-9;
  long long t0, t1;
  // DFMA
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, y, z);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DADD
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = x + z;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // DMUL
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = x * y;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // SHFL64 + DADD
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x += __shfl_xor_sync(0xffffffffu, x, 1 << (j & 3));
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // F2F roundtrip (f64->f32->f64) + DADD
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = (double)((float)x) + z;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // MUFU.RSQ f32 chain
  float f = (float)x + 2.0f;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f = rsqrtf(f) + 1.5f;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // FFMA
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f = fmaf(f, 1.0000001f, 1e-9f);
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // LDS dependent (pointer chase through an index)
  int idx = threadIdx.x & 63;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) idx = ((int)sm[idx] + idx) & 63;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  // FSEL on double (select) + DADD
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = ((threadIdx.x >> (j & 3)) & 1) ? x + z : x - z;
  }
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + f + idx;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024 * 64); cudaMalloc(&cyc, 8 * 16);
  const char* names[] = {"DFMA", "DADD", "DMUL", "SHFL64+DADD", "F2F.rt+DADD", "MUFU.RSQ+FADD", "FFMA", "LDS+cvt+iadd", "DADD+sel"};
  for (int warps : {1, 8, 16}) {
    k<<<1, 32 * warps>>>(out, cyc, 1.0, 64);
    cudaDeviceSynchronize();
    k<<<1, 32 * warps>>>(out, cyc, 1.0, 64);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 8 * 16, cudaMemcpyDeviceToHost);
    printf("warps/SM=%d:", warps);
    for (int i = 0; i < 9; ++i) printf("  %s %.1f", names[i], h[i] / (64.0 * 16));
    printf("  (cycles per dependent op)\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
