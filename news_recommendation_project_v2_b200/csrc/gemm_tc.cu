// Dense row kernel on the 5th-gen tensor cores:  y = epilogue(a[M,K] @ w[N,K]^T + bias).
//
// This is the torch.nn.Linear contraction of the reference (latent_attention.py:65-74,33-37;
// modeling_utils.py:218-222), bf16 operands, fp32 accumulation.
//
// B200 design (one CTA per SM, persistent over output tiles, warp specialised):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2-D tiles of A (128x64) and W (BNx64) into a
//              kStages-deep ring of 128B-swizzled shared-memory buffers, mbarrier complete_tx.
//   warp 1   : MMA issuer    -- one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//              (M=128, N=BN, K=16) reading the smem ring through UMMA descriptors; accumulators live
//              in TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps the
//              mainloop of tile i+1; tcgen05.commit releases smem stages / publishes accumulators.
//   warps 2-5: epilogue      -- tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> fused
//              bias / ReLU / exp / residual / GEGLU -> vectorised global stores.
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <mutex>

namespace nrb {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // double-buffered accumulator
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  int64_t M;             // row capacity (TMA extent); effective rows = min(M, *m_dev) when m_dev != NULL
  const int* m_dev;      // optional device-side row count (varlen token packing, no host sync)
  int N, K;
  const float* bias;
  const float* res;
  int64_t ldres;
  void* y;
  int64_t ldy;
};

__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752440f)); }

template <int EPI, bool OUT_BF16>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], const GemmParams& p, int64_t row, int col0) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + j);
      v[j] += b.x;
      v[j + 1] += b.y;
      v[j + 2] += b.z;
      v[j + 3] += b.w;
    }
  }
  if (EPI == NRB_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (EPI == NRB_EPI_EXP) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = expf(v[j]);
  } else if (EPI == NRB_EPI_RESIDUAL) {
    const float* r = p.res + row * p.ldres + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(r + j);
      v[j] += b.x;
      v[j + 1] += b.y;
      v[j + 2] += b.z;
      v[j + 3] += b.w;
    }
  }
  if (EPI == NRB_EPI_GEGLU) {
    // W rows interleaved (a0,g0,a1,g1,...): 32 accumulator columns -> 16 outputs
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = v[2 * j] * gelu_erf(v[2 * j + 1]);
    const int oc = col0 >> 1;
    if (OUT_BF16) {
      __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + row * p.ldy + oc;
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        uint4 u = make_uint4(pack_bf16x2(o[j], o[j + 1]), pack_bf16x2(o[j + 2], o[j + 3]),
                             pack_bf16x2(o[j + 4], o[j + 5]), pack_bf16x2(o[j + 6], o[j + 7]));
        *reinterpret_cast<uint4*>(y + j) = u;
      }
    } else {
      float* y = reinterpret_cast<float*>(p.y) + row * p.ldy + oc;
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(y + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
    return;
  }
  if (OUT_BF16) {
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(p.y) + row * p.ldy + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 u = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                           pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
      *reinterpret_cast<uint4*>(y + j) = u;
    }
  } else {
    float* y = reinterpret_cast<float*>(p.y) + row * p.ldy + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(y + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
}

template <int BN, int EPI, bool OUT_BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atom
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                    // [kStages] TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;     // [2] MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int64_t M = p.m_dev != nullptr ? min(p.M, (int64_t)*p.m_dev) : p.M;
  const int m_tiles = (int)((M + kBM - 1) / kBM);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int64_t total_tiles = (int64_t)m_tiles * n_tiles;
  const int k_blocks = p.K / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_a);
    ptx::prefetch_tensormap(&map_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = (int)(tile / n_tiles);
        const int n_blk = (int)(tile % n_tiles);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          ptx::tma_load_2d(smem_a + stage * Cfg::kABytes, &map_a, &full_bar[stage], kb * kBK, m_blk * kBM);
          ptx::tma_load_2d(smem_b + stage * Cfg::kBBytes, &map_w, &full_bar[stage], kb * kBK, n_blk * BN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * kUmmaK * 2);
            const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 2);
            ptx::mma_bf16_ss(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::mma_commit(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = (int)(tile / n_tiles);
      const int n_blk = (int)(tile % n_tiles);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const int64_t row = (int64_t)m_blk * kBM + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        const int col0 = n_blk * BN + c;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr + (uint32_t)c, r);
        ptx::tmem_ld_wait();
        if (row < M) epilogue_chunk<EPI, OUT_BF16>(r, p, row, col0);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host side ---------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 64 cols], 128B swizzle.
int make_tmap_bf16(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return NRB_E_CUDA;
  }
  NRB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
  NRB_REQUIRE((ld * 2) % 16 == 0, "TMA operand leading dimension must be a multiple of 8 elements");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)(ld * 2)};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
    return NRB_E_CUDA;
  }
  return NRB_OK;
}

template <int BN, int EPI, bool OUT_BF16>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mw, const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tc_kernel<BN, EPI, OUT_BF16>;
  static bool attr_set = false;
  if (!attr_set) {
    NRB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int64_t tiles = ((p.M + kBM - 1) / kBM) * (int64_t)((p.N + BN - 1) / BN);
  const int grid = (int)std::min<int64_t>(tiles, sm_count_cached());
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(ma, mw, p); note_launch();
  NRB_CUDA_CHECK(cudaGetLastError());
  return NRB_OK;
}

template <int BN, bool OUT_BF16>
static int dispatch_epi(int epi, const CUtensorMap& ma, const CUtensorMap& mw, const GemmParams& p,
                        cudaStream_t st) {
  switch (epi) {
    case NRB_EPI_NONE:
      return launch_gemm<BN, NRB_EPI_NONE, OUT_BF16>(ma, mw, p, st);
    case NRB_EPI_RELU:
      return launch_gemm<BN, NRB_EPI_RELU, OUT_BF16>(ma, mw, p, st);
    case NRB_EPI_EXP:
      return launch_gemm<BN, NRB_EPI_EXP, OUT_BF16>(ma, mw, p, st);
    case NRB_EPI_RESIDUAL:
      return launch_gemm<BN, NRB_EPI_RESIDUAL, OUT_BF16>(ma, mw, p, st);
    case NRB_EPI_GEGLU:
      return launch_gemm<BN, NRB_EPI_GEGLU, OUT_BF16>(ma, mw, p, st);
    default:
      set_error("nrb_linear(bf16): unsupported epilogue %d", epi);
      return NRB_E_INVALID;
  }
}

// y = epi(a @ w^T + bias); a [M,K] bf16, w [N,K] bf16.
int gemm_bf16_tc(int epi, int out_dtype, const void* a, int64_t lda, const void* w, int64_t ldw, const float* bias,
                 const float* res, int64_t ldres, void* y, int64_t ldy, int64_t M, const int* m_dev, int N, int K,
                 cudaStream_t st) {
  NRB_REQUIRE(M > 0 && N > 0 && K > 0, "nrb_linear: empty problem");
  NRB_REQUIRE(K % kBK == 0, "nrb_linear(bf16): K must be a multiple of 64 (got %d)", K);
  NRB_REQUIRE(N % 32 == 0, "nrb_linear(bf16): N must be a multiple of 32 (got %d)", N);
  NRB_REQUIRE(out_dtype == NRB_F32 || out_dtype == NRB_BF16, "nrb_linear: bad out_dtype");
  NRB_REQUIRE(epi != NRB_EPI_RESIDUAL || res != nullptr, "nrb_linear: residual epilogue needs res");
  const int64_t ycols_align = out_dtype == NRB_BF16 ? 8 : 4;
  NRB_REQUIRE(ldy % ycols_align == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
              "nrb_linear: y must be 16-byte aligned with a 16-byte multiple row pitch");
  NRB_REQUIRE(res == nullptr || (ldres % 4 == 0 && (reinterpret_cast<uintptr_t>(res) & 15) == 0),
              "nrb_linear: res must be 16-byte aligned");
  NRB_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "nrb_linear: bias alignment");
  const bool small_n = N <= 128;
  CUtensorMap ma, mw;
  int rc = make_tmap_bf16(&ma, a, M, K, lda, kBM);
  if (rc != NRB_OK) return rc;
  rc = make_tmap_bf16(&mw, w, N, K, ldw, small_n ? 128 : 256);
  if (rc != NRB_OK) return rc;
  GemmParams p;
  p.M = M;
  p.m_dev = m_dev;
  p.N = N;
  p.K = K;
  p.bias = bias;
  p.res = res;
  p.ldres = ldres;
  p.y = y;
  p.ldy = ldy;
  if (small_n) {
    return out_dtype == NRB_BF16 ? dispatch_epi<128, true>(epi, ma, mw, p, st)
                                 : dispatch_epi<128, false>(epi, ma, mw, p, st);
  }
  return out_dtype == NRB_BF16 ? dispatch_epi<256, true>(epi, ma, mw, p, st)
                               : dispatch_epi<256, false>(epi, ma, mw, p, st);
}

}  // namespace nrb
