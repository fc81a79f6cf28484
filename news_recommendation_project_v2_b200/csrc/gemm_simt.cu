// fp32 "exact" dense row kernel: y = epilogue(a[M,K] @ w[N,K]^T + bias) with FFMA accumulation.
//
// This is the reference's own arithmetic (config.py:39 TORCH_DTYPE = float32; torch's fp32
// matmul does not use TF32 by default), kept as the parity path for
// modeling_utils.py:218-222 and latent_attention.py:65-74,33-37.  The throughput path is the
// tcgen05 kernel in gemm_tc.cu; this one exists so that rank parity against the fp32 reference
// does not depend on bf16 rounding.
//
// Classic register-tiled SGEMM: 128x128 CTA tile, BK = 16, 256 threads, 8x8 outputs per thread,
// operands staged k-major in shared memory with register prefetch of the next k-block.
#include "common.cuh"

#include <algorithm>

namespace nrb {

constexpr int SBM = 128, SBN = 128, SBK = 16, STHREADS = 256;

struct SimtParams {
  const float* a;
  int64_t lda;
  const float* w;
  int64_t ldw;
  const float* bias;
  const void* res;
  int res_dtype;
  const int32_t* res_map;
  int64_t ldres;
  void* y;
  int64_t ldy;
  int64_t M;
  const int* m_dev;  // optional device-side row count (absolute, before the y0 split)
  int64_t row_base;  // first absolute row of this launch
  int N, K;
};

__device__ __forceinline__ float gelu_erf_f32(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752440f)); }

template <int EPI, bool OUT_BF16>
__global__ void __launch_bounds__(STHREADS)
gemm_simt_kernel(const SimtParams p) {
  __shared__ float As[2][SBK][SBM + 4];
  __shared__ float Bs[2][SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15;   // column group: columns tx*4..+3 and 64+tx*4..+3
  const int ty = tid >> 4;   // row group:    rows    ty*4..+3 and 64+ty*4..+3
  const int64_t m0 = (int64_t)blockIdx.y * SBM;
  const int n0 = blockIdx.x * SBN;
  const int64_t Meff = p.m_dev != nullptr ? min(p.M, (int64_t)*p.m_dev - p.row_base) : p.M;
  if (m0 >= Meff) return;

  // global -> register staging: each thread moves 2 float4 of A and 2 of W per k-block
  const int lr = tid >> 2;        // 0..63 (row within half tile)
  const int lk = (tid & 3) * 4;   // k offset 0,4,8,12
  float4 ra[2], rb[2];

  auto load_tiles = [&](int kb) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t row = m0 + lr + h * 64;
      ra[h] = row < Meff ? *reinterpret_cast<const float4*>(p.a + row * p.lda + kb * SBK + lk)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      const int col = n0 + lr + h * 64;
      rb[h] = col < p.N ? *reinterpret_cast<const float4*>(p.w + (int64_t)col * p.ldw + kb * SBK + lk)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      As[buf][lk + 0][r] = ra[h].x;
      As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z;
      As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x;
      Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z;
      Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int kblocks = p.K / SBK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < kblocks; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < kblocks) load_tiles(kb + 1);
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < kblocks) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: each thread owns 2x2 blocks of 4 rows x 4 consecutive columns ----------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= Meff) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + jh * 64 + tx * 4;
      if (col >= p.N) continue;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (p.bias != nullptr) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += p.bias[col + q];
      }
      if (EPI == NRB_EPI_RELU) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
      } else if (EPI == NRB_EPI_EXP) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = expf(v[q]);
      } else if (EPI == NRB_EPI_RESIDUAL) {
        const int64_t arow = p.row_base + row;  // row maps are indexed by the absolute row
        const int64_t rrow = p.res_map != nullptr ? (int64_t)p.res_map[arow] : row;
        if (p.res_dtype == NRB_F32) {
          const float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + rrow * p.ldres + col);
          v[0] += r.x;
          v[1] += r.y;
          v[2] += r.z;
          v[3] += r.w;
        } else {
          const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.res) + rrow * p.ldres + col);
          v[0] += bf16_lo(r.x);
          v[1] += bf16_hi(r.x);
          v[2] += bf16_lo(r.y);
          v[3] += bf16_hi(r.y);
        }
      }
      if (EPI == NRB_EPI_GEGLU) {
        // W rows interleaved in pairs (a0,a1,g0,g1,...): same layout as the tensor-core kernel
        const float o0 = v[0] * gelu_erf_f32(v[2]);
        const float o1 = v[1] * gelu_erf_f32(v[3]);
        const int oc = col >> 1;
        if (OUT_BF16) {
          *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.y) + row * p.ldy + oc) = pack_bf16x2(o0, o1);
        } else {
          *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.y) + row * p.ldy + oc) = make_float2(o0, o1);
        }
      } else if (OUT_BF16) {
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.y) + row * p.ldy + col) =
            make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + row * p.ldy + col) =
            make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

template <bool OUT_BF16>
static int dispatch_simt(int epi, const SimtParams& p, dim3 grid, cudaStream_t st) {
  switch (epi) {
#define NRB_CASE(E)                                                    \
  case E:                                                              \
    gemm_simt_kernel<E, OUT_BF16><<<grid, STHREADS, 0, st>>>(p); note_launch();       \
    break;
    NRB_CASE(NRB_EPI_NONE)
    NRB_CASE(NRB_EPI_RELU)
    NRB_CASE(NRB_EPI_EXP)
    NRB_CASE(NRB_EPI_RESIDUAL)
    NRB_CASE(NRB_EPI_GEGLU)
#undef NRB_CASE
    default:
      set_error("nrb_linear(fp32): unsupported epilogue %d", epi);
      return NRB_E_INVALID;
  }
  NRB_CUDA_CHECK(cudaGetLastError());
  return NRB_OK;
}

int gemm_f32_simt(int epi, int out_dtype, const void* a, int64_t lda, const void* w, int64_t ldw,
                  const float* bias, const void* res, int64_t ldres, void* y, int64_t ldy, int64_t M,
                  const int* m_dev, int N, int K, cudaStream_t st, int res_dtype, const int32_t* res_map) {
  NRB_REQUIRE(M > 0 && N > 0 && K > 0, "nrb_linear: empty problem");
  NRB_REQUIRE(K % SBK == 0, "nrb_linear(fp32): K must be a multiple of 16 (got %d)", K);
  NRB_REQUIRE(N % 4 == 0, "nrb_linear(fp32): N must be a multiple of 4 (got %d)", N);
  NRB_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w) & 15) == 0,
              "nrb_linear(fp32): operands must be 16-byte aligned");
  NRB_REQUIRE(epi != NRB_EPI_RESIDUAL || res != nullptr, "nrb_linear: residual epilogue needs res");
  NRB_REQUIRE(out_dtype == NRB_F32 || out_dtype == NRB_BF16, "nrb_linear: bad out_dtype");
  NRB_REQUIRE(ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "nrb_linear: y alignment");
  SimtParams p;
  p.a = (const float*)a;
  p.lda = lda;
  p.w = (const float*)w;
  p.ldw = ldw;
  p.bias = bias;
  p.res = res;
  p.res_dtype = res_dtype;
  p.res_map = res_map;
  p.ldres = ldres;
  p.y = y;
  p.ldy = ldy;
  p.M = M;
  p.m_dev = m_dev;
  p.row_base = 0;
  p.N = N;
  p.K = K;
  const int64_t my = (M + SBM - 1) / SBM;
  NRB_REQUIRE(my <= 65535 * 64ll, "nrb_linear(fp32): M too large for one launch");
  // grid.y is limited to 65535: split very tall problems
  const int64_t max_y = 65535;
  for (int64_t y0 = 0; y0 < my; y0 += max_y) {
    SimtParams q = p;
    const int64_t rows0 = y0 * SBM;
    q.a = p.a + rows0 * lda;
    if (p.res && p.res_map == nullptr)
      q.res = (const char*)p.res + (size_t)rows0 * ldres * (res_dtype == NRB_F32 ? 4 : 2);
    q.y = out_dtype == NRB_BF16 ? (void*)((__nv_bfloat16*)p.y + rows0 * ldy) : (void*)((float*)p.y + rows0 * ldy);
    q.M = std::min<int64_t>(M - rows0, max_y * SBM);
    q.row_base = rows0;
    dim3 grid((N + SBN - 1) / SBN, (unsigned)std::min<int64_t>(my - y0, max_y));
    int rc = out_dtype == NRB_BF16 ? dispatch_simt<true>(epi, q, grid, st) : dispatch_simt<false>(epi, q, grid, st);
    if (rc != NRB_OK) return rc;
  }
  return NRB_OK;
}

}  // namespace nrb
