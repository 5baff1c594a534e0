// Shared helpers for the sparf_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/sparf_b200.h"

namespace sparf {

// thread-local error text behind sparf_last_error()
void set_error(const char* fmt, ...);

#define SPARF_CHECK_CUDA(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ::sparf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SPARF_ERR_CUDA;                                                                     \
    }                                                                                            \
  } while (0)

// every kernel launch of the library goes through this macro: it also feeds sparf_launch_count()
extern unsigned long long g_launch_count;
#define SPARF_CHECK_LAUNCH(name)                                                                 \
  do {                                                                                           \
    ++::sparf::g_launch_count;                                                                   \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess) {                                                                     \
      ::sparf::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SPARF_ERR_CUDA;                                                                     \
    }                                                                                            \
  } while (0)

#define SPARF_REQUIRE(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::sparf::set_error(__VA_ARGS__);           \
      return SPARF_ERR_INVALID;                  \
    }                                            \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// number of SMs of the current device (cached)
int num_sms();

// ---- arithmetic that must round exactly like the reference's separate fp32 torch ops (no FMA
// contraction): the 2^9*pi positional-encoding band amplifies a 1-ulp difference in x by ~1e3.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// torch.nn.functional.softplus (beta=1, threshold=20) and its derivative
__device__ __forceinline__ float softplus_f(float z) { return z > 20.f ? z : log1pf(expf(z)); }
__device__ __forceinline__ float softplus_grad_f(float z) {
  if (z > 20.f) return 1.f;
  float e = expf(z);
  return e / (e + 1.f);
}
__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + expf(-z)); }

// BARF coarse-to-fine weight of frequency band j (frequency_nerf.py:248-253), fp32 op-for-op:
//   alpha = (progress - start) / (end - start) * L ; w = (1 - cos(pi * clamp(alpha - j, 0, 1))) / 2
__device__ __forceinline__ float c2f_weight(float progress, float start, float inv_den /*unused*/, float den,
                                            int L, int j) {
  float alpha = mul_rn(__fdiv_rn(__fsub_rn(progress, start), den), (float)L);
  float x = fminf(fmaxf(__fsub_rn(alpha, (float)j), 0.f), 1.f);
  float c = cosf(mul_rn(x, 3.14159274101257324f));
  return __fdiv_rn(__fsub_rn(1.f, c), 2.f);
}

struct C2F {
  int enabled;
  float start, den;
  const float* progress;
};

__device__ __forceinline__ float band_weight(const C2F& c, int L, int j) {
  if (!c.enabled) return 1.f;
  return c2f_weight(*c.progress, c.start, 0.f, c.den, L, j);
}

// frequency of band j: 2^j * float(pi) (exact scaling of the fp32 constant), frequency_nerf.py:52
__device__ __forceinline__ float band_freq(int j) { return ldexpf(3.14159274101257324f, j); }

}  // namespace sparf
