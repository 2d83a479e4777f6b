// Shared host/device helpers for the d2r_b200 library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/d2r_b200.h"

namespace d2r {

// ---- error reporting (thread-local message; the C ABI never throws) -------------------
char* last_error_buf();
int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define D2R_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::d2r::set_error(D2R_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define D2R_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::d2r::set_error(D2R_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(D2R_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return D2R_OK;
}

int num_sms();   // SM count of the current device (cached)

// ---- dtype-generic element access -----------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 8 consecutive elements <-> 8 floats (16B vector for bf16, 2x16B for fp32); pointers must be
// aligned to 16 bytes.
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == D2R_ACT_RELU) return fmaxf(v, 0.f);
  if (act == D2R_ACT_TANH) return tanhf(v);
  return v;
}

inline size_t dtype_size(int dt) { return dt == D2R_BF16 ? 2 : 4; }

// dispatch on a runtime dtype enum to a template functor taking the element type
#define D2R_DISPATCH_DTYPE(dt, T, ...)                                          \
  do {                                                                          \
    if ((dt) == D2R_BF16) {                                                     \
      using T = __nv_bfloat16;                                                  \
      __VA_ARGS__;                                                              \
    } else if ((dt) == D2R_F32) {                                               \
      using T = float;                                                          \
      __VA_ARGS__;                                                              \
    } else {                                                                    \
      return ::d2r::set_error(D2R_ERR_ARG, "bad dtype %d", (int)(dt));          \
    }                                                                           \
  } while (0)

}  // namespace d2r
