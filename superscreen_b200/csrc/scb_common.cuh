// Shared helpers for libsc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/scb.h"

namespace scb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SCB_CHECK_ARG(cond, msg)                       \
  do {                                                 \
    if (!(cond)) {                                     \
      scb::set_error("%s: %s", __func__, msg);         \
      return SCB_ERR_INVALID;                          \
    }                                                  \
  } while (0)

#define SCB_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      scb::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return SCB_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define SCB_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    scb::count_launch();                                                            \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      scb::set_error("%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return SCB_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr double kOneOver4Pi = 0.07957747154594767;  // 1/(4*pi)

// r2^(-3/2) in fp64.  One MUFU seed + Newton steps inside rsqrt(); cubed.
__device__ __forceinline__ double inv_r3(double r2) {
  double inv = rsqrt(r2);
  return inv * inv * inv;
}

}  // namespace scb
