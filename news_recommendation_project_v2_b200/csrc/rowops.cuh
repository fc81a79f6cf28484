// Row-wise arithmetic of the HBM-bound passes of stage A (LayerNorm, masked mean + L2 norm), written ONCE and used
//   * by the stand-alone kernels (dense.cu layer_norm_vec_kernel, latent.cu pool_items_kernel) and
//   * by the RIDER warps of the persistent tcgen05 GEMM kernels (gemm_tc.cu): the two control warps that are idle
//     after set-up (TMEM allocator, all-gather carrier) run the LayerNorm / pooling pass of a NEIGHBOURING sub-chunk
//     while the MMA / epilogue warps of the same CTA work on the GEMM -- the tensor-bound kernels leave 55-85 % of the
//     HBM bandwidth unused, the passes are pure HBM streams, so they disappear from the critical path.
// Every floating-point operation is spelled with an explicit rounding intrinsic (no contraction left to the compiler)
// so that both users produce the same bits: chunking, packing and riding stay invisible in the results.
//
// reference: torch.nn.LayerNorm (latent_attention.py:10-19), masked mean + F.normalize (latent_attention.py:165-170)
#pragma once

#include "common.cuh"

namespace nrb {

// ---- LayerNorm of one row held in registers (one warp per row) -------------------------------------------------
// NVMAX 16-byte vectors per lane, `nv` of them valid (warp-uniform; nv == NVMAX in the stand-alone kernel).
template <typename TIN, int NVMAX>
struct LnRow {
  static constexpr int EPV = Vec16<TIN>::EPV;
  float v[NVMAX][EPV];

  __device__ __forceinline__ void load(const char* base, int lane, int nv) {
#pragma unroll
    for (int i = 0; i < NVMAX; ++i)
      if (i < nv) {
        const uint4 u = *reinterpret_cast<const uint4*>(base + (size_t)(lane + 32 * i) * 16);
        Vec16<TIN>::unpack(u, v[i]);
      }
  }

  // two-pass mean / biased variance in fp32 like ATen; y row in bf16 or fp32, optional fp32 copy of the raw row
  __device__ __forceinline__ void finish(int lane, int nv, int dim, float eps, const float* gamma, const float* beta,
                                         void* y_row, int y_dtype, float* copy_row) const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NVMAX; ++i)
      if (i < nv) {
#pragma unroll
        for (int k = 0; k < EPV; ++k) s = __fadd_rn(s, v[i][k]);
      }
    const float mean = __fdiv_rn(warp_sum(s), (float)dim);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NVMAX; ++i)
      if (i < nv) {
#pragma unroll
        for (int k = 0; k < EPV; ++k) {
          const float d = __fsub_rn(v[i][k], mean);
          q = __fmaf_rn(d, d, q);
        }
      }
    const float rstd = rsqrtf(__fadd_rn(__fdiv_rn(warp_sum(q), (float)dim), eps));
#pragma unroll
    for (int i = 0; i < NVMAX; ++i)
      if (i < nv) {
        const int e0 = (lane + 32 * i) * EPV;
        float o[EPV];
#pragma unroll
        for (int k = 0; k < EPV; k += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(gamma + e0 + k);
          const float4 b4 = *reinterpret_cast<const float4*>(beta + e0 + k);
          o[k] = __fmaf_rn(__fmul_rn(__fsub_rn(v[i][k], mean), rstd), g4.x, b4.x);
          o[k + 1] = __fmaf_rn(__fmul_rn(__fsub_rn(v[i][k + 1], mean), rstd), g4.y, b4.y);
          o[k + 2] = __fmaf_rn(__fmul_rn(__fsub_rn(v[i][k + 2], mean), rstd), g4.z, b4.z);
          o[k + 3] = __fmaf_rn(__fmul_rn(__fsub_rn(v[i][k + 3], mean), rstd), g4.w, b4.w);
        }
        if (y_dtype == NRB_BF16) {
          __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(y_row) + e0;
          if (EPV == 8) {
            *reinterpret_cast<uint4*>(yo) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                       pack_bf16x2(o[EPV - 4], o[EPV - 3]), pack_bf16x2(o[EPV - 2], o[EPV - 1]));
          } else {
            *reinterpret_cast<uint2*>(yo) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
          }
        } else {
          float* yo = reinterpret_cast<float*>(y_row) + e0;
#pragma unroll
          for (int k = 0; k < EPV; k += 4) *reinterpret_cast<float4*>(yo + k) = make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]);
        }
        if (copy_row != nullptr) {
#pragma unroll
          for (int k = 0; k < EPV; k += 4)
            *reinterpret_cast<float4*>(copy_row + e0 + k) = make_float4(v[i][k], v[i][k + 1], v[i][k + 2], v[i][k + 3]);
        }
      }
  }
};

// ---- masked mean + L2 normalisation of one item (latent_attention.py:165-170) ------------------------------------
// The stand-alone kernel gives one CTA of 256 threads to an item: thread t owns the float4 columns t, t+256, ...; the
// squared norm is reduced per warp (butterfly) and then over the 8 warps in order.  A rider warp reproduces exactly
// that association: lane l plays the threads l, l+32, ..., l+224 one after the other.
__device__ __forceinline__ void pool_add(float4& a, const float4& t) {
  a.x = __fadd_rn(a.x, t.x);
  a.y = __fadd_rn(a.y, t.y);
  a.z = __fadd_rn(a.z, t.z);
  a.w = __fadd_rn(a.w, t.w);
}
__device__ __forceinline__ void pool_mean(float4& a, float cnt) {  // 0/0 -> NaN for an all-masked item, like the reference
  a.x = __fdiv_rn(a.x, cnt);
  a.y = __fdiv_rn(a.y, cnt);
  a.z = __fdiv_rn(a.z, cnt);
  a.w = __fdiv_rn(a.w, cnt);
}
__device__ __forceinline__ float pool_sq4(const float4& a) {
  float t = __fmul_rn(a.x, a.x);
  t = __fmaf_rn(a.y, a.y, t);
  t = __fmaf_rn(a.z, a.z, t);
  return __fmaf_rn(a.w, a.w, t);
}
__device__ __forceinline__ float pool_norm(float tot) { return fmaxf(sqrtf(tot), 1e-12f); }  // F.normalize eps
__device__ __forceinline__ float4 pool_unit(const float4& a, float nrm) {
  return make_float4(__fdiv_rn(a.x, nrm), __fdiv_rn(a.y, nrm), __fdiv_rn(a.z, nrm), __fdiv_rn(a.w, nrm));
}

// ---- rider jobs -----------------------------------------------------------------------------------------------------
enum { NRB_RIDE_NONE = 0, NRB_RIDE_LN = 1, NRB_RIDE_POOL = 2 };
struct RiderJob {
  int kind;
  // LayerNorm: y[r] = LN(x[row_map ? row_map[r] : r]) in bf16, r < min(rows, *rows_dev)
  const void* x;
  int x_dtype;
  int64_t ldx;
  const int32_t* row_map;
  const float* gamma;
  const float* beta;
  void* y;
  int64_t ldy;
  int64_t rows;
  const int* rows_dev;
  int dim;
  float eps;
  // pooling: out[i] = normalize(mean(h[item_off[i] .. item_off[i+1])))
  const float* h;
  int64_t ldh;
  const int32_t* item_off;
  int64_t items;
  float* out;
};
// shapes a rider warp can take: rows of 256 / 512 / 768 / 1024 elements (a ring of rows has to fit the registers)
static inline bool rider_dim_ok(int dim) { return dim == 256 || dim == 512 || dim == 768 || dim == 1024; }
static inline bool rider_ln_ok(int x_dtype, int dim) { return (x_dtype == NRB_F32 || x_dtype == NRB_BF16) && rider_dim_ok(dim); }
static inline bool rider_pool_ok(int dim) { return rider_dim_ok(dim); }

// A rider warp is alone with the memory latency (~2 us under the load of the GEMM's own traffic): it keeps a RING of R
// raw rows in registers -- the row consumed now was requested R rows ago -- and the row-map entries one ring cycle
// further ahead, so that no instruction waits for a load that was issued less than R rows of work earlier.

// LayerNorm rider: warp `part` of `n_parts` takes rows part, part + n_parts, ...; bf16 output.
template <typename TIN, int NV, int R>
__device__ __forceinline__ void rider_ln_ring(const RiderJob& j, int64_t part, int64_t n_parts, int lane) {
  const int64_t n = j.rows_dev != nullptr ? min(j.rows, (int64_t)*j.rows_dev) : j.rows;
  if (part >= n) return;
  const int64_t K = (n - part + n_parts - 1) / n_parts;  // my rows: part + k * n_parts, k < K
  const TIN* xin = reinterpret_cast<const TIN*>(j.x);
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(j.y);
  uint4 raw[R][NV];
  int32_t src[R];
  auto src_of = [&](int64_t k) -> int32_t {
    const int64_t r = part + k * n_parts;
    return j.row_map != nullptr ? j.row_map[r] : (int32_t)r;
  };
  auto request = [&](uint4 (&dst)[NV], int32_t s) {
    const char* base = reinterpret_cast<const char*>(xin + (int64_t)s * j.ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i] = *reinterpret_cast<const uint4*>(base + (size_t)(lane + 32 * i) * 16);
  };
#pragma unroll
  for (int s = 0; s < R; ++s) src[s] = s < K ? src_of(s) : 0;
#pragma unroll
  for (int s = 0; s < R; ++s)
    if (s < K) request(raw[s], src[s]);
#pragma unroll
  for (int s = 0; s < R; ++s) src[s] = R + s < K ? src_of(R + s) : 0;
  for (int64_t k0 = 0; k0 < K; k0 += R) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int64_t k = k0 + s;
      if (k < K) {  // warp-uniform
        LnRow<TIN, NV> row;
#pragma unroll
        for (int i = 0; i < NV; ++i) Vec16<TIN>::unpack(raw[s][i], row.v[i]);
        if (k + R < K) request(raw[s], src[s]);
        if (k + 2 * R < K) src[s] = src_of(k + 2 * R);
        row.finish(lane, NV, j.dim, j.eps, j.gamma, j.beta, y + (part + k * n_parts) * j.ldy, NRB_BF16, nullptr);
      }
    }
  }
}
template <typename TIN>
__device__ __forceinline__ void rider_ln_typed(const RiderJob& j, int64_t part, int64_t n_parts, int lane) {
  constexpr int Q = (int)sizeof(TIN) == 4 ? 2 : 1;  // 16-byte vectors per lane per 256 elements
  constexpr int RB = (int)sizeof(TIN) == 4 ? 4 : 6; // ring depth at <= 768 elements (96 / 72 raw registers)
  switch (j.dim) {
    case 256: rider_ln_ring<TIN, 1 * Q, RB>(j, part, n_parts, lane); break;
    case 512: rider_ln_ring<TIN, 2 * Q, RB>(j, part, n_parts, lane); break;
    case 768: rider_ln_ring<TIN, 3 * Q, RB>(j, part, n_parts, lane); break;
    case 1024: rider_ln_ring<TIN, 4 * Q, (int)sizeof(TIN) == 4 ? 3 : 5>(j, part, n_parts, lane); break;
    default: break;
  }
}

// pooling rider: one warp per item; the item's rows stream through a ring of R rows
template <int NVP, int R>
__device__ __forceinline__ void rider_pool_ring(const RiderJob& j, int64_t part, int64_t n_parts, int lane) {
  for (int64_t i = part; i < j.items; i += n_parts) {
    const int64_t r0 = j.item_off[i], r1 = j.item_off[i + 1];
    const int64_t K = r1 - r0;
    const float cnt = (float)K;
    float4 acc[NVP];
#pragma unroll
    for (int w = 0; w < NVP; ++w) acc[w] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint4 raw[R][NVP];
    auto request = [&](uint4 (&dst)[NVP], int64_t r) {
      const float* p = j.h + r * j.ldh;
#pragma unroll
      for (int w = 0; w < NVP; ++w) dst[w] = ldg_stream_128(p + (size_t)(lane + 32 * w) * 4);
    };
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (s < K) request(raw[s], r0 + s);
    for (int64_t k0 = 0; k0 < K; k0 += R) {
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int64_t k = k0 + s;
        if (k < K) {  // warp-uniform
          float4 t[NVP];
#pragma unroll
          for (int w = 0; w < NVP; ++w)
            t[w] = make_float4(__uint_as_float(raw[s][w].x), __uint_as_float(raw[s][w].y), __uint_as_float(raw[s][w].z),
                               __uint_as_float(raw[s][w].w));
          if (k + R < K) request(raw[s], r0 + k + R);
#pragma unroll
          for (int w = 0; w < NVP; ++w) pool_add(acc[w], t[w]);
        }
      }
    }
    // lane l plays the threads l + 32 w of the stand-alone kernel's CTA (w >= NVP: no columns, they contribute +0)
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < NVP; ++w) {
      pool_mean(acc[w], cnt);
      tot = __fadd_rn(tot, warp_sum(__fadd_rn(0.f, pool_sq4(acc[w]))));
    }
    const float nrm = pool_norm(tot);
#pragma unroll
    for (int w = 0; w < NVP; ++w)
      *reinterpret_cast<float4*>(j.out + i * (int64_t)j.dim + (size_t)(lane + 32 * w) * 4) = pool_unit(acc[w], nrm);
  }
}
__device__ __forceinline__ void rider_pool(const RiderJob& j, int64_t part, int64_t n_parts, int lane) {
  switch (j.dim) {
    case 256: rider_pool_ring<2, 4>(j, part, n_parts, lane); break;
    case 512: rider_pool_ring<4, 4>(j, part, n_parts, lane); break;
    case 768: rider_pool_ring<6, 4>(j, part, n_parts, lane); break;
    case 1024: rider_pool_ring<8, 3>(j, part, n_parts, lane); break;
    default: break;
  }
}

__device__ __forceinline__ void rider_warp(const RiderJob& j, int64_t part, int64_t n_parts, int lane) {
  if (j.kind == NRB_RIDE_LN) {
    if (j.x_dtype == NRB_F32)
      rider_ln_typed<float>(j, part, n_parts, lane);
    else
      rider_ln_typed<__nv_bfloat16>(j, part, n_parts, lane);
  } else if (j.kind == NRB_RIDE_POOL) {
    rider_pool(j, part, n_parts, lane);
  }
}

// host side (gemm_tc.cu): the next rider-capable tcgen05 GEMM launch of this host thread carries the job
void set_next_rider(const RiderJob& job);
bool riders_enabled();

}  // namespace nrb
