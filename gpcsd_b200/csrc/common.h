// Shared host-side helpers of libgpcsd_b200.so: error reporting across the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gpcsd_b200.h"

int gp_fail(const char* msg);                 // records msg, returns 1
int gp_fail_cuda(cudaError_t e, const char* what, int line);
int gp_num_sms();
void gp_set_sm_reserve(int n);   // thread-local: SMs left free when this thread sizes full-GPU grids

#define GP_CUDA(expr)                                                  \
  do {                                                                 \
    cudaError_t _e = (expr);                                           \
    if (_e != cudaSuccess) return gp_fail_cuda(_e, #expr, __LINE__);   \
  } while (0)

// Per-DEVICE one-shot guard for cudaFuncSetAttribute calls (function attributes are per device; a process may drive more
// than one): flags is a zero-initialised static array of GP_MAX_DEVICES ints owned by the call site.
#define GP_MAX_DEVICES 64
static inline bool gp_first_use_on_device(int* flags) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= GP_MAX_DEVICES) return true;
  if (flags[dev]) return false;
  flags[dev] = 1;
  return true;
}
