// Shared helpers for the rave_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rave_b200.h"

namespace rave {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define RAVE_CHECK_ARG(cond, ...)                \
  do {                                           \
    if (!(cond)) {                               \
      rave::set_error(__VA_ARGS__);              \
      return 1;                                  \
    }                                            \
  } while (0)

#define RAVE_CHECK_LAUNCH(name)                                              \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      rave::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return 2;                                                              \
    }                                                                        \
    rave::count_launch();                                                    \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// activation(dim) of rave/blocks.py: LeakyReLU(slope) / Snake(alpha)
__device__ __forceinline__ float act_apply(float x, int act, float slope, float alpha) {
  if (act == RAVE_ACT_LEAKY) return x > 0.f ? x : x * slope;
  if (act == RAVE_ACT_SNAKE) {
    float s = sinf(alpha * x);
    return x + s * s / (alpha + 1e-9f);
  }
  return x;
}
__device__ __forceinline__ float act_grad(float x, int act, float slope, float alpha) {
  if (act == RAVE_ACT_LEAKY) return x > 0.f ? 1.f : slope;
  if (act == RAVE_ACT_SNAKE) {
    // d/dx [x + sin^2(a x)/(a+eps)] = 1 + a sin(2 a x)/(a+eps)
    return 1.f + alpha * sinf(2.f * alpha * x) / (alpha + 1e-9f);
  }
  return 1.f;
}

}  // namespace rave
