// Shared helpers for the nrb200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/nrb200.h"

namespace nrb {

void set_error(const char* fmt, ...);

#define NRB_CUDA_CHECK(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::nrb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NRB_E_CUDA;                                                                       \
    }                                                                                          \
  } while (0)

#define NRB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::nrb::set_error(__VA_ARGS__);    \
      return NRB_E_INVALID;             \
    }                                   \
  } while (0)

static inline cudaStream_t as_stream(nrb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count_cached();
void note_launch(int n = 1);  // cumulative kernel-launch counter (nrb_kernel_launches)

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}

// 128-bit streaming load through the read-only path, not allocated in L1:
// table rows are touched once per impression and never reused by the same SM.
__device__ __forceinline__ uint4 ldg_stream_128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Unpack one 16-byte vector into fp32 lanes.  EPV = elements per vector.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int EPV = 4;
  __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z);
    f[3] = __uint_as_float(v.w);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int EPV = 8;
  __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
    f[0] = bf16_lo(v.x);
    f[1] = bf16_hi(v.x);
    f[2] = bf16_lo(v.y);
    f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z);
    f[5] = bf16_hi(v.z);
    f[6] = bf16_lo(v.w);
    f[7] = bf16_hi(v.w);
  }
};

}  // namespace nrb
