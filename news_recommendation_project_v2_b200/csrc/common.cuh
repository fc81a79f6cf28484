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

// ---- all-gather pushes carried by the GEMM kernels (push_rows.cu, gemm_tc.cu) ----------------------------------
// One segment = a block of finished rows that goes to EVERY rank's copy of a table through one NVLink multicast
// (NVLS) store per 16 bytes.  `dst` is the multicast address of the first destination row.
struct PushSeg {
  const char* src;
  char* dst;
  int64_t n_rows;
  int64_t src_stride_bytes;
  int64_t dst_stride_bytes;
  int vecs_per_row;  // 16-byte OUTPUT vectors per row
  int f32_to_bf16;   // 1: src rows are fp32, destination rows bf16
};
constexpr int kMaxPushSegs = 3;
struct PushJob {
  PushSeg seg[kMaxPushSegs];
  int n_seg;
};
// pending segments attached by nrb_push_attach: the next tcgen05 GEMM launches of this host thread each take a share
void take_push_share(PushJob* job);

__device__ __forceinline__ void multimem_st_v4(char* mc_addr, const uint4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// One warp's share (part `part` of `n_parts`) of a push job: 128-bit loads, `UN` vectors per lane in flight, one
// multicast store per vector.  Runs in the spare control warp of the persistent GEMM CTAs (and in the standalone
// flush kernel), so the transfer overlaps the tensor-core work of the same kernel.
__device__ __forceinline__ void push_job_warp(const PushJob& job, int64_t part, int64_t n_parts, int lane) {
  constexpr int UN = 4;
  for (int sidx = 0; sidx < job.n_seg; ++sidx) {
    const PushSeg& sg = job.seg[sidx];
    const int64_t total = sg.n_rows * sg.vecs_per_row;
    for (int64_t base = part * (32 * UN); base < total; base += n_parts * (32 * UN)) {
      uint4 buf[UN];
      char* dst[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int64_t idx = base + u * 32 + lane;
        dst[u] = nullptr;
        if (idx < total) {
          const int64_t r = idx / sg.vecs_per_row;
          const int v = (int)(idx - r * sg.vecs_per_row);
          dst[u] = sg.dst + r * sg.dst_stride_bytes + (size_t)v * 16;
          if (sg.f32_to_bf16) {
            const char* sp = sg.src + r * sg.src_stride_bytes + (size_t)v * 32;
            const uint4 a = ldg_stream_128(sp), b = ldg_stream_128(sp + 16);
            buf[u] = make_uint4(pack_bf16x2(__uint_as_float(a.x), __uint_as_float(a.y)),
                                pack_bf16x2(__uint_as_float(a.z), __uint_as_float(a.w)),
                                pack_bf16x2(__uint_as_float(b.x), __uint_as_float(b.y)),
                                pack_bf16x2(__uint_as_float(b.z), __uint_as_float(b.w)));
          } else {
            buf[u] = ldg_stream_128(sg.src + r * sg.src_stride_bytes + (size_t)v * 16);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u)
        if (dst[u] != nullptr) multimem_st_v4(dst[u], buf[u]);
    }
  }
  __threadfence_system();
}

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
