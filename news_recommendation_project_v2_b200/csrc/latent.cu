// Stage A: latent-attention pooling head (reference latent_attention.py:77-171).
//
//   nrb_latent_fold    : one-time weight preparation.  K,V = to_kv(LN_ctx(latents)) do not depend on
//                        the input (the reference recomputes them for every batch row, :161,:67), so they
//                        are projected once and folded through to_q / to_out into two per-head matrices
//                        A [heads*Lp, d] and B [d, heads*Lp]:  logits = LN(x) A^T, attn = softmax(logits) B^T.
//   nrb_latent_forward : varlen token packing (padded tokens are never computed) -> LN -> GEMM(A) ->
//                        per-head softmax -> GEMM(B)+residual -> LN -> GEMM(FF1)+bias+GEGLU ->
//                        GEMM(FF2)+bias+residual -> masked mean -> L2 normalise.
// All contractions go through nrb::linear (tcgen05 for bf16 weights, FFMA for fp32 weights).
#include "dense.cuh"
#include "rowops.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace nrb {

// ---- varlen packing ----------------------------------------------------------------------
// counts[i] = number of non-zero mask entries of item i (one warp per item)
__global__ void __launch_bounds__(256)
count_valid_kernel(const int32_t* mask, int64_t items, int seq, int32_t* counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp_id; i < items; i += n_warps) {
    int c = 0;
    for (int s = lane; s < seq; s += 32) c += mask[i * seq + s] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFullMask, c, o);
    if (lane == 0) counts[i] = c;
  }
}

// exclusive scan of counts -> item_off[items+1]; total -> *m_dev.  Single CTA (items <= ~1e6).
__global__ void __launch_bounds__(1024)
scan_items_kernel(const int32_t* counts, int64_t items, int32_t* item_off, int32_t* m_dev) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < items; base += 1024) {
    const int64_t i = base + tid;
    const int32_t v = i < items ? counts[i] : 0;
    int32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(kFullMask, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(kFullMask, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int32_t before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl - v;
    if (i < items) item_off[i] = before;
    __syncthreads();
    if (tid == 1023) carry = before + v;
    __syncthreads();
  }
  if (tid == 0) {
    item_off[items] = carry;
    *m_dev = carry;
  }
}

// row_map[item_off[i] + rank] = i*seq + s for every valid token (one warp per item)
__global__ void __launch_bounds__(256)
fill_row_map_kernel(const int32_t* mask, int64_t items, int seq, const int32_t* item_off, int32_t* row_map) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp_id; i < items; i += n_warps) {
    int32_t pos = item_off[i];
    for (int s0 = 0; s0 < seq; s0 += 32) {
      const int s = s0 + lane;
      const bool v = s < seq && mask[i * seq + s] != 0;
      const unsigned b = __ballot_sync(kFullMask, v);
      if (v) row_map[pos + __popc(b & ((1u << lane) - 1))] = (int32_t)(i * seq + s);
      pos += __popc(b);
    }
  }
}

// int32 item offsets -> int64 CSR offsets (the layout nrb_score_rank takes)
__global__ void widen_offsets_kernel(const int32_t* in, int64_t n, int64_t* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}

// pooled[i] = normalize(mean over the item's packed rows)  (latent_attention.py:165-170)
// one CTA per item; thread t owns float4 columns t, t+256, ...
template <int MAXV>
__global__ void __launch_bounds__(256)
pool_items_kernel(const float* h, int64_t ldh, const int32_t* item_off, int seq_if_dense, int64_t items, int dim,
                  float* out) {
  __shared__ float red[8];
  const int tid = threadIdx.x;
  const int nvec = dim / 4;
  for (int64_t i = blockIdx.x; i < items; i += gridDim.x) {
    const int64_t r0 = item_off != nullptr ? item_off[i] : i * seq_if_dense;
    const int64_t r1 = item_off != nullptr ? item_off[i + 1] : (i + 1) * seq_if_dense;
    const float cnt = (float)(r1 - r0);
    float4 acc[MAXV];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int v = tid + k * 256;
      acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < nvec) {
        for (int64_t r = r0; r < r1; ++r) pool_add(acc[k], *reinterpret_cast<const float4*>(h + r * ldh + (size_t)v * 4));
        pool_mean(acc[k], cnt);  // 0/0 -> NaN for an all-masked item, like the reference
        ss = __fadd_rn(ss, pool_sq4(acc[k]));
      }
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot = __fadd_rn(tot, red[w]);
    const float nrm = pool_norm(tot);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
      const int v = tid + k * 256;
      if (v < nvec) *reinterpret_cast<float4*>(out + i * (int64_t)dim + (size_t)v * 4) = pool_unit(acc[k], nrm);
    }
    __syncthreads();
  }
}

// padded latent count: a power of two >= 32 so that softmax groups tile the 256-column MMA tiles
static int pad_latents(int L) {
  int p = 32;
  while (p < L) p <<= 1;
  return p;
}

struct FoldWs {
  float *cn, *kv, *wqT, *a32, *b32;
  size_t bytes;
};
static FoldWs fold_ws(void* base, int dim, int heads, int dim_head, int L) {
  const int inner = heads * dim_head;
  const int Lp = pad_latents(L);
  Workspace ws(base, (size_t)-1);
  FoldWs f;
  f.cn = (float*)ws.take((size_t)L * dim * 4);
  f.kv = (float*)ws.take((size_t)L * 2 * inner * 4);
  f.wqT = (float*)ws.take((size_t)dim * inner * 4);
  f.a32 = (float*)ws.take((size_t)heads * Lp * dim * 4);
  f.b32 = (float*)ws.take((size_t)dim * heads * Lp * 4);
  f.bytes = ws.used + 256;
  return f;
}

// Workspace of the forward pass.  Up to two SUB-CHUNKS are in flight (set 0 / set 1): their GEMMs alternate, and the
// HBM-bound passes of one (LayerNorm 1 / 2, masked mean) ride in the idle control warps of the other's GEMM kernels.
struct PackBufs {
  int32_t *counts, *item_off, *m_dev, *row_map;
};
struct SubSet {
  PackBufs pk[2];  // the packing of the NEXT sub-chunk of a set is built while the current one is still in use
  void *xn, *p, *hn, *g;
  float *logits, *h1;
};
struct FwdWs {
  SubSet set[2];
  float* h2;  // [cap0 + cap1, dim]: set 1's rows follow set 0's
  size_t bytes;
};
// fused softmax epilogue (logits never leave TMEM): bf16, whole 256-column tiles, a softmax row within one cluster
static bool fused_softmax(const nrb_latent_weights* w) {
  return w->precision == NRB_BF16 && (w->heads * w->latents_padded) % 256 == 0 && w->latents_padded <= 1024;
}

static FwdWs fwd_ws(void* base, const nrb_latent_weights* w, int64_t cap0, int64_t items0, int64_t cap1,
                    int64_t items1) {
  const size_t es = dtype_size(w->precision);
  const int64_t hl = (int64_t)w->heads * w->latents_padded;
  Workspace ws(base, (size_t)-1);
  FwdWs f;
  for (int k = 0; k < 2; ++k) {
    const int64_t cap = k == 0 ? cap0 : cap1, items = k == 0 ? items0 : items1;
    SubSet& q = f.set[k];
    for (int b = 0; b < 2; ++b) {
      q.pk[b].counts = (int32_t*)ws.take((size_t)items * 4);
      q.pk[b].item_off = (int32_t*)ws.take((size_t)(items + 1) * 4);
      q.pk[b].m_dev = (int32_t*)ws.take(16);
      q.pk[b].row_map = (int32_t*)ws.take((size_t)cap * 4);
    }
    q.xn = ws.take((size_t)cap * w->dim * es);
    q.logits = fused_softmax(w) ? nullptr : (float*)ws.take((size_t)cap * hl * 4);
    q.p = ws.take((size_t)cap * hl * es);
    q.h1 = (float*)ws.take((size_t)cap * w->dim * 4);
    q.hn = ws.take((size_t)cap * w->dim * es);
    q.g = ws.take((size_t)cap * 4 * w->dim * es);
  }
  f.h2 = (float*)ws.take((size_t)(cap0 + cap1) * w->dim * 4);
  f.bytes = ws.used + 256;
  return f;
}

// sub-chunks shorter than this are not worth splitting (wave quantisation of the GEMMs, launch gaps)
static int64_t min_sub_tokens() {  // read per call: tests and A/B runs flip it at run time
  const char* e = getenv("NRB200_RIDER_MIN_TOKENS");
  const long long n = e != nullptr ? atoll(e) : 0;
  return (int64_t)(n > 0 ? n : 65536);
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_mask_to_csr(const int32_t* mask, int64_t batch, int seq, int32_t* idx_out, int64_t* off_out,
                               int32_t* workspace, nrb_stream_t stream) {
  NRB_REQUIRE(batch >= 0 && seq > 0, "nrb_mask_to_csr: bad sizes");
  NRB_REQUIRE(batch * (int64_t)seq < (int64_t)1 << 31, "nrb_mask_to_csr: batch * seq must fit int32");
  NRB_REQUIRE(off_out != nullptr, "nrb_mask_to_csr: null pointer");
  cudaStream_t st = as_stream(stream);
  if (batch == 0) {
    NRB_CUDA_CHECK(cudaMemsetAsync(off_out, 0, 8, st));
    return NRB_OK;
  }
  NRB_REQUIRE(mask && idx_out && workspace, "nrb_mask_to_csr: null pointer");
  int32_t* counts = workspace;                // [batch]
  int32_t* item_off = workspace + batch;      // [batch + 1]
  int32_t* total = workspace + 2 * batch + 1;  // [1]
  const int g1 = (int)std::min<int64_t>((batch + 7) / 8, (int64_t)sm_count_cached() * 16);
  count_valid_kernel<<<g1, 256, 0, st>>>(mask, batch, seq, counts); note_launch();
  scan_items_kernel<<<1, 1024, 0, st>>>(counts, batch, item_off, total); note_launch();
  fill_row_map_kernel<<<g1, 256, 0, st>>>(mask, batch, seq, item_off, idx_out); note_launch();
  const int g2 = (int)std::min<int64_t>((batch + 1 + 255) / 256, (int64_t)sm_count_cached() * 4);
  widen_offsets_kernel<<<g2, 256, 0, st>>>(item_off, batch + 1, off_out); note_launch();
  NRB_CUDA_CHECK(cudaGetLastError());
  return NRB_OK;
}

extern "C" size_t nrb_latent_fold_workspace_bytes(int dim, int heads, int dim_head, int num_latents) {
  return fold_ws(nullptr, dim, heads, dim_head, num_latents).bytes;
}

extern "C" int nrb_latent_fold(int precision, int dim, int heads, int dim_head, int num_latents,
                               const float* latents, const float* ln_ctx_w, const float* ln_ctx_b,
                               const float* w_q, const float* w_kv, const float* w_out, void* a_out, void* b_out,
                               void* workspace, size_t workspace_bytes, nrb_stream_t stream) {
  NRB_REQUIRE(precision == NRB_F32 || precision == NRB_BF16, "nrb_latent_fold: bad precision");
  NRB_REQUIRE(dim > 0 && dim % 16 == 0, "nrb_latent_fold: dim must be a multiple of 16");
  NRB_REQUIRE(heads > 0 && dim_head > 0 && dim_head % 16 == 0, "nrb_latent_fold: dim_head must be a multiple of 16");
  NRB_REQUIRE(num_latents > 0 && num_latents % 4 == 0, "nrb_latent_fold: num_latents must be a multiple of 4");
  NRB_REQUIRE(latents && ln_ctx_w && ln_ctx_b && w_q && w_kv && w_out && a_out && b_out && workspace,
              "nrb_latent_fold: null pointer");
  FoldWs f = fold_ws(workspace, dim, heads, dim_head, num_latents);
  if (workspace_bytes < f.bytes) {
    set_error("nrb_latent_fold: workspace too small (%zu < %zu)", workspace_bytes, f.bytes);
    return NRB_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const int L = num_latents, Lp = pad_latents(L), inner = heads * dim_head, hl = heads * Lp;
  int rc;
  // context = LN_ctx(latents); kv = to_kv(context)            (latent_attention.py:18-19, 67)
  if ((rc = layer_norm_rows(latents, NRB_F32, dim, nullptr, ln_ctx_w, ln_ctx_b, f.cn, NRB_F32, dim, nullptr, 0, L,
                            nullptr, dim, st)) != NRB_OK)
    return rc;
  if ((rc = gemm_f32_simt(NRB_EPI_NONE, NRB_F32, f.cn, dim, w_kv, dim, nullptr, nullptr, 0, f.kv, 2 * inner, L,
                          nullptr, 2 * inner, dim, st)) != NRB_OK)
    return rc;
  // fold the SDPA scale dim_head^-0.5 (:69-72) into K
  if ((rc = scale_cols_f32(f.kv, 2 * inner, L, inner, 1.0f / sqrtf((float)dim_head), st)) != NRB_OK) return rc;
  if ((rc = transpose_f32(w_q, inner, dim, f.wqT, st)) != NRB_OK) return rc;
  NRB_CUDA_CHECK(cudaMemsetAsync(f.a32, 0, (size_t)hl * dim * 4, st));
  NRB_CUDA_CHECK(cudaMemsetAsync(f.b32, 0, (size_t)dim * hl * 4, st));
  for (int h = 0; h < heads; ++h) {
    // A_h[l, i] = sum_c K[l, h*dh+c] * Wq[h*dh+c, i]
    if ((rc = gemm_f32_simt(NRB_EPI_NONE, NRB_F32, f.kv + h * dim_head, 2 * inner, f.wqT + h * dim_head, inner,
                            nullptr, nullptr, 0, f.a32 + (size_t)h * Lp * dim, dim, L, nullptr, dim, dim_head,
                            st)) != NRB_OK)
      return rc;
    // B[i, h*Lp + l] = sum_c Wout[i, h*dh+c] * V[l, h*dh+c]
    if ((rc = gemm_f32_simt(NRB_EPI_NONE, NRB_F32, w_out + h * dim_head, inner, f.kv + inner + h * dim_head,
                            2 * inner, nullptr, nullptr, 0, f.b32 + (size_t)h * Lp, hl, dim, nullptr, L, dim_head,
                            st)) != NRB_OK)
      return rc;
  }
  if ((rc = convert_rows(f.a32, NRB_F32, dim, a_out, precision, dim, hl, dim, st)) != NRB_OK) return rc;
  if ((rc = convert_rows(f.b32, NRB_F32, hl, b_out, precision, hl, dim, hl, st)) != NRB_OK) return rc;
  return NRB_OK;
}

// ---- the chain over sub-chunks -------------------------------------------------------------------------------------
// One sub-chunk = a block of token rows that goes through
//     L1: xn = LN1(x)  ->  S: P = softmax_h(xn A^T)  ->  V: h1 = P B^T + x  ->  L2: hn = LN2(h1)
//     ->  G: g = GEGLU(hn W1^T + b1)  ->  F: h2 = g W2^T + b2 + h1  ->  PL: pooled = normalize(masked mean(h2))
// (latent_attention.py:154-170).  S / V / G / F are tcgen05 GEMMs (tensor bound), L1 / L2 / PL are HBM streams.
struct Sub {
  int64_t items = 0;     // item mode: items of this sub-chunk
  int64_t rows_cap = 0;  // token rows (capacity; the effective count is *m_dev when tokens are packed on device)
  const void* xc = nullptr;
  const int32_t* mask = nullptr;  // item mode with packing: [items, seq]
  SubSet* set = nullptr;
  PackBufs* pk = nullptr;
  const int32_t* row_map = nullptr;
  const int* m_dev = nullptr;
  float* h2 = nullptr;
  float* pooled = nullptr;  // item mode: this sub-chunk's rows of pooled_out
};

struct RiderGuard {  // nothing stale may ride in a later, unrelated GEMM launch
  ~RiderGuard() {
    RiderJob none = {};
    set_next_rider(none);
  }
};

struct Chain {
  const nrb_latent_weights* w;
  int x_dtype, seq;
  bool pack, pool, ride;
  cudaStream_t st;
  int sms;

  int do_pack(Sub& s) const {
    if (!pack) return NRB_OK;
    const int g1 = (int)std::min<int64_t>((s.items + 7) / 8, (int64_t)sms * 16);
    count_valid_kernel<<<g1, 256, 0, st>>>(s.mask, s.items, seq, s.pk->counts); note_launch();
    scan_items_kernel<<<1, 1024, 0, st>>>(s.pk->counts, s.items, s.pk->item_off, s.pk->m_dev); note_launch();
    fill_row_map_kernel<<<g1, 256, 0, st>>>(s.mask, s.items, seq, s.pk->item_off, s.pk->row_map); note_launch();
    NRB_CUDA_CHECK(cudaGetLastError());
    s.row_map = s.pk->row_map;
    s.m_dev = s.pk->m_dev;
    return NRB_OK;
  }
  RiderJob ln_job(const void* x, int xdt, const int32_t* row_map, const float* g, const float* b, void* y,
                  const Sub& s) const {
    RiderJob j = {};
    j.kind = NRB_RIDE_LN;
    j.x = x;
    j.x_dtype = xdt;
    j.ldx = w->dim;
    j.row_map = row_map;
    j.gamma = g;
    j.beta = b;
    j.y = y;
    j.ldy = w->dim;
    j.rows = s.rows_cap;
    j.rows_dev = s.m_dev;
    j.dim = w->dim;
    j.eps = 1e-5f;
    return j;
  }
  // xn = LN1(x), gathered through the row map when tokens are packed      latent_attention.py:16
  RiderJob l1(const Sub& s) const { return ln_job(s.xc, x_dtype, s.row_map, w->ln1_w, w->ln1_b, s.set->xn, s); }
  // hn = LN2(h1)                                                           :16 (second PreNorm)
  RiderJob l2(const Sub& s) const { return ln_job(s.set->h1, NRB_F32, nullptr, w->ln2_w, w->ln2_b, s.set->hn, s); }
  // pooled = normalize(masked mean(h2))                                    :165-170
  RiderJob pl(const Sub& s) const {
    RiderJob j = {};
    j.kind = NRB_RIDE_POOL;
    j.h = s.h2;
    j.ldh = w->dim;
    j.item_off = s.pk->item_off;
    j.items = s.items;
    j.out = s.pooled;
    j.dim = w->dim;
    return j;
  }
  int alone(const RiderJob& j) const {
    if (j.kind == NRB_RIDE_LN)
      return layer_norm_rows(j.x, j.x_dtype, j.ldx, j.row_map, j.gamma, j.beta, j.y, w->precision, j.ldy, nullptr, 0,
                             j.rows, j.rows_dev, j.dim, st, j.eps);
    const int gp = (int)std::min<int64_t>(j.items, (int64_t)sms * 16);
    if (gp > 0) {
      pool_items_kernel<4><<<gp, 256, 0, st>>>(j.h, j.ldh, j.item_off, 0, j.items, j.dim, j.out); note_launch();
      NRB_CUDA_CHECK(cudaGetLastError());
    }
    return NRB_OK;
  }
  // hand the pass to the GEMM launch that FOLLOWS (`hosted`), or run it as a kernel of its own
  int pass(const RiderJob& j, bool hosted) const {
    const bool fits = j.kind == NRB_RIDE_LN ? rider_ln_ok(j.x_dtype, j.dim) : rider_pool_ok(j.dim);
    if (ride && hosted && fits) {
      set_next_rider(j);
      return NRB_OK;
    }
    return alone(j);
  }
  // P = softmax_h(xn A^T)  (SDPA scale folded into A)                       :65-72
  int S(const Sub& s) const {
    const int d = w->dim, P = w->precision, hl = w->heads * w->latents_padded;
    if (fused_softmax(w))  // logits never leave TMEM; row statistics exchanged across the cluster through DSMEM
      return linear(P, NRB_EPI_SOFTMAX, P, s.set->xn, d, w->a, d, nullptr, nullptr, 0, s.set->p, hl, s.rows_cap,
                    s.m_dev, hl, d, st, w->latents_padded, w->num_latents);
    int rc = linear(P, NRB_EPI_NONE, NRB_F32, s.set->xn, d, w->a, d, nullptr, nullptr, 0, s.set->logits, hl,
                    s.rows_cap, s.m_dev, hl, d, st);
    if (rc != NRB_OK) return rc;
    return softmax_groups(s.set->logits, hl, s.set->p, P, hl, s.rows_cap, s.m_dev, w->heads, w->latents_padded,
                          w->num_latents, st);
  }
  // h1 = P B^T + x   (the residual operand is the RAW input row, read in place through the row map)   :74, :162
  int V(const Sub& s) const {
    const int d = w->dim, hl = w->heads * w->latents_padded;
    return linear(w->precision, NRB_EPI_RESIDUAL, NRB_F32, s.set->p, hl, w->b, hl, nullptr, s.xc, d, s.set->h1, d,
                  s.rows_cap, s.m_dev, d, hl, st, 0, 0, x_dtype, s.row_map);
  }
  // g = GEGLU(hn W1^T + b1)                                                 :33-35, 24-27
  int G(const Sub& s) const {
    const int d = w->dim, P = w->precision;
    return linear(P, NRB_EPI_GEGLU, P, s.set->hn, d, w->w_ff1, d, w->b_ff1, nullptr, 0, s.set->g, 4 * d, s.rows_cap,
                  s.m_dev, 8 * d, d, st);
  }
  // h2 = g W2^T + b2 + h1                                                   :36, :163
  int F(const Sub& s) const {
    const int d = w->dim;
    return linear(w->precision, NRB_EPI_RESIDUAL, NRB_F32, s.set->g, 4 * d, w->w_ff2, 4 * d, w->b_ff2, s.set->h1, d,
                  s.h2, d, s.rows_cap, s.m_dev, d, 4 * d, st);
  }

  // Launch order for a pair (A, B) of sub-chunks; the pass named in brackets rides in that GEMM's idle warps:
  //   S_A [PL of the previous A]   S_B [PL of the previous B]   V_A [L1 of the next A]   V_B [L2_A]
  //   G_A [L2_B]                   G_B [L1 of the next B]       F_A                      F_B
  // (first pair: S_A carries L1_B; last pair: F_B carries PL_A; only L1 of the very first and PL of the very last
  //  sub-chunk are kernels of their own)
  // Every pass reads only results of EARLIER launches and is consumed by LATER ones, so kernel boundaries are the
  // only synchronisation.  With one sub-chunk (or riders off) every pass is a kernel of its own, in program order.
  int run(Sub* subs, int n) const {
#define NRB_TRY(expr)                 \
  do {                                \
    const int _rc = (expr);           \
    if (_rc != NRB_OK) return _rc;    \
  } while (0)
    RiderGuard guard;
    for (int i = 0; i < n; i += ride ? 2 : 1) {
      // without riders the sub-chunks share ONE workspace set and run strictly one after the other
      Sub& A = subs[i];
      Sub* B = ride && i + 1 < n ? &subs[i + 1] : nullptr;
      Sub* nA = ride && i + 2 < n ? &subs[i + 2] : nullptr;
      Sub* nB = ride && i + 3 < n ? &subs[i + 3] : nullptr;
      Sub* pA = ride && i >= 2 ? &subs[i - 2] : nullptr;
      Sub* pB = ride && i >= 2 ? &subs[i - 1] : nullptr;
      if (i == 0 || !ride) {  // first pair: nothing earlier to ride in
        NRB_TRY(do_pack(A));
        NRB_TRY(alone(l1(A)));
        if (B) NRB_TRY(do_pack(*B));
      }
      if (pool && pA)
        NRB_TRY(pass(pl(*pA), true));
      else if (i == 0 && B)
        NRB_TRY(pass(l1(*B), true));
      NRB_TRY(S(A));
      if (pool && pB) NRB_TRY(pass(pl(*pB), B != nullptr));
      if (B) NRB_TRY(S(*B));
      if (nA) {
        NRB_TRY(do_pack(*nA));
        NRB_TRY(pass(l1(*nA), true));
      }
      NRB_TRY(V(A));
      NRB_TRY(pass(l2(A), B != nullptr));
      if (B) {
        NRB_TRY(V(*B));
        NRB_TRY(pass(l2(*B), true));
      }
      NRB_TRY(G(A));
      if (B) {
        if (nB) {
          NRB_TRY(do_pack(*nB));
          NRB_TRY(pass(l1(*nB), true));
        }
        NRB_TRY(G(*B));
      }
      NRB_TRY(F(A));
      const bool last = pool && !nA;
      if (B) {
        if (last) NRB_TRY(pass(pl(A), true));
        NRB_TRY(F(*B));
      }
      if (last) NRB_TRY(alone(pl(B ? *B : A)));  // nothing left to ride in
    }
#undef NRB_TRY
    return NRB_OK;
  }
};

static bool chain_rides(const nrb_latent_weights* w) {
  return riders_enabled() && w->precision == NRB_BF16 && fused_softmax(w);
}

static int64_t chunk_items_for(int64_t max_tokens, int seq) { return std::max<int64_t>(1, max_tokens / seq); }

extern "C" size_t nrb_latent_forward_workspace_bytes(const nrb_latent_weights* w, int64_t max_tokens) {
  if (w == nullptr || max_tokens <= 0) return 0;
  // capacity in items is unknown here (depends on seq): bound it by the token capacity (seq >= 1); one sub-chunk of
  // max_tokens or two of half of it
  const int64_t half = (max_tokens + 1) / 2;
  return std::max(fwd_ws(nullptr, w, max_tokens, max_tokens, 0, 0).bytes, fwd_ws(nullptr, w, half, half, half, half).bytes);
}

extern "C" int nrb_latent_forward(const nrb_latent_weights* w, const void* x, int x_dtype, int64_t batch, int seq,
                                  const int32_t* token_mask, float* pooled_out, float* unpooled_out,
                                  void* workspace, size_t workspace_bytes, int64_t max_tokens,
                                  int64_t* n_tokens_host, nrb_stream_t stream) {
  NRB_REQUIRE(w != nullptr, "nrb_latent_forward: null weights");
  NRB_REQUIRE(w->precision == NRB_F32 || w->precision == NRB_BF16, "nrb_latent_forward: bad precision");
  NRB_REQUIRE(x_dtype == NRB_F32 || x_dtype == NRB_BF16, "nrb_latent_forward: bad x dtype");
  NRB_REQUIRE(batch >= 0 && seq > 0, "nrb_latent_forward: bad batch/seq");
  NRB_REQUIRE((pooled_out != nullptr) != (unpooled_out != nullptr),
              "nrb_latent_forward: exactly one of pooled_out / unpooled_out must be given");
  NRB_REQUIRE(pooled_out == nullptr || token_mask != nullptr, "nrb_latent_forward: pooling needs token_mask");
  NRB_REQUIRE(w->dim % 64 == 0, "nrb_latent_forward: dim must be a multiple of 64 (got %d)", w->dim);
  NRB_REQUIRE(w->latents_padded >= 32 && (w->latents_padded & (w->latents_padded - 1)) == 0 &&
                  w->latents_padded >= w->num_latents,
              "nrb_latent_forward: latents_padded must be a power of two >= max(32, num_latents)");
  NRB_REQUIRE(max_tokens >= seq, "nrb_latent_forward: max_tokens (%lld) must be >= seq (%d)", (long long)max_tokens,
              seq);
  NRB_REQUIRE(w->dim <= 4096, "nrb_latent_forward: dim > 4096 unsupported");
  if (n_tokens_host) *n_tokens_host = -1;
  if (batch == 0) return NRB_OK;
  NRB_REQUIRE(x && workspace, "nrb_latent_forward: null pointer");
  // sub-chunks: one of up to max_tokens token slots, or -- when the call is large enough for the HBM-bound passes to
  // ride in a neighbour's GEMMs -- an even number of up to max_tokens / 2 slots each, two in flight
  const bool rides = chain_rides(w);
  const int64_t ib_full = std::min<int64_t>(batch, chunk_items_for(max_tokens, seq));
  const int64_t ib_half = chunk_items_for(max_tokens / 2, seq);
  const bool dual = rides && max_tokens / 2 >= seq && batch >= 2 && batch * (int64_t)seq >= 2 * min_sub_tokens() &&
                    ib_half * seq >= std::min<int64_t>(min_sub_tokens(), batch * (int64_t)seq / 2);
  int64_t ib = ib_full;
  if (dual) {
    int64_t n_sub = (batch + ib_half - 1) / ib_half;
    n_sub += n_sub & 1;  // pairs
    ib = (batch + n_sub - 1) / n_sub;
  }
  const int64_t cap = ib * seq;
  FwdWs f = dual ? fwd_ws(workspace, w, cap, ib, cap, ib) : fwd_ws(workspace, w, cap, ib, 0, 0);
  if (workspace_bytes < f.bytes) {
    set_error("nrb_latent_forward: workspace too small (%zu < %zu)", workspace_bytes, f.bytes);
    return NRB_E_WORKSPACE;
  }
  const int d = w->dim;
  const size_t xs = dtype_size(x_dtype);
  const bool packed = pooled_out != nullptr;
  const int64_t n_sub = (batch + ib - 1) / ib;
  NRB_REQUIRE(n_sub <= (1 << 20), "nrb_latent_forward: too many sub-chunks (raise max_tokens)");
  std::vector<Sub> subs((size_t)n_sub);
  for (int64_t k = 0; k < n_sub; ++k) {
    Sub& s = subs[(size_t)k];
    const int64_t i0 = k * ib;
    s.items = std::min<int64_t>(ib, batch - i0);
    s.rows_cap = s.items * seq;
    s.xc = (const char*)x + (size_t)i0 * seq * d * xs;
    s.set = &f.set[dual ? (k & 1) : 0];
    if (packed) {
      s.mask = token_mask + i0 * seq;
      s.pk = &s.set->pk[dual ? ((k >> 1) & 1) : 0];
      s.h2 = f.h2 + (dual && (k & 1) ? (size_t)cap * d : 0);
      s.pooled = pooled_out + i0 * d;
    } else {
      s.h2 = unpooled_out + (size_t)i0 * seq * d;
    }
  }
  Chain c{w, x_dtype, seq, packed, packed, rides && dual, as_stream(stream), sm_count_cached()};
  return c.run(subs.data(), (int)n_sub);
}

// Varlen entry: tokens already packed [n_tokens, dim] with CSR item offsets (device int32 [batch+1]); this is the
// layout a packed token store hands over (the reference pads per batch: data_utils.py:878-933, 753-781).
extern "C" int nrb_latent_forward_packed(const nrb_latent_weights* w, const void* x_packed, int x_dtype,
                                         int64_t n_tokens, const int32_t* item_off, int64_t batch,
                                         float* pooled_out, void* workspace, size_t workspace_bytes,
                                         nrb_stream_t stream) {
  NRB_REQUIRE(w != nullptr, "nrb_latent_forward_packed: null weights");
  NRB_REQUIRE(w->precision == NRB_F32 || w->precision == NRB_BF16, "nrb_latent_forward_packed: bad precision");
  NRB_REQUIRE(x_dtype == NRB_F32 || x_dtype == NRB_BF16, "nrb_latent_forward_packed: bad x dtype");
  NRB_REQUIRE(n_tokens >= 0 && batch >= 0, "nrb_latent_forward_packed: bad sizes");
  NRB_REQUIRE(w->dim % 64 == 0 && w->dim <= 4096, "nrb_latent_forward_packed: unsupported dim %d", w->dim);
  if (batch == 0) return NRB_OK;
  if (n_tokens == 0) {
    // every item is empty: the reference's masked mean is 0/0 = NaN (latent_attention.py:165-168)
    NRB_REQUIRE(item_off && pooled_out, "nrb_latent_forward_packed: null pointer");
    const int gp0 = (int)std::min<int64_t>(batch, (int64_t)sm_count_cached() * 16);
    pool_items_kernel<4><<<gp0, 256, 0, as_stream(stream)>>>(nullptr, w->dim, item_off, 0, batch, w->dim, pooled_out);
    note_launch();
    NRB_CUDA_CHECK(cudaGetLastError());
    return NRB_OK;
  }
  NRB_REQUIRE(x_packed && item_off && pooled_out && workspace, "nrb_latent_forward_packed: null pointer");
  // rows are independent until the pooling: two sub-chunks of rows (split anywhere, not at item boundaries) let the
  // LayerNorm passes ride; the pooling runs over all rows afterwards
  const bool dual = chain_rides(w) && n_tokens >= 2 * min_sub_tokens();
  const int64_t cap0 = dual ? ((n_tokens / 2 + 127) / 128) * 128 : n_tokens;
  const int64_t cap1 = n_tokens - cap0;
  FwdWs f = fwd_ws(workspace, w, cap0, 0, cap1, 0);
  if (workspace_bytes < f.bytes) {
    set_error("nrb_latent_forward_packed: workspace too small (%zu < %zu)", workspace_bytes, f.bytes);
    return NRB_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  int rc;
  Sub subs[2];
  subs[0].rows_cap = cap0;
  subs[0].xc = x_packed;
  subs[0].set = &f.set[0];
  subs[0].h2 = f.h2;
  subs[1].rows_cap = cap1;
  subs[1].xc = (const char*)x_packed + (size_t)cap0 * w->dim * dtype_size(x_dtype);
  subs[1].set = &f.set[1];
  subs[1].h2 = f.h2 + (size_t)cap0 * w->dim;
  Chain c{w, x_dtype, 0, false, false, dual, st, sm_count_cached()};
  if ((rc = c.run(subs, dual ? 2 : 1)) != NRB_OK) return rc;
  const int gp = (int)std::min<int64_t>(batch, (int64_t)sm_count_cached() * 16);
  pool_items_kernel<4><<<gp, 256, 0, st>>>(f.h2, w->dim, item_off, 0, batch, w->dim, pooled_out); note_launch();
  NRB_CUDA_CHECK(cudaGetLastError());
  return NRB_OK;
}
