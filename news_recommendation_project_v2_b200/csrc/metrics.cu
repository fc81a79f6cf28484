// MIND metrics on the device: per-impression AUC / MRR / nDCG@5 / nDCG@10 from dense ranks + labels.
//
// Replaces evaluation.py:34-98 (score_row in a 4-process pool, 2.2 ms per impression):
//   y_score = 1 / rank                                   (evaluation.py:41-47)
//   auc     = sklearn roc_auc_score(y_true, y_score)     = Mann-Whitney U with ties counted 1/2
//   order   = np.argsort(y_score)[::-1]                  (evaluation.py:14,28)
//   mrr     = sum_k y[order[k]] / (k+1) / sum(y)         (evaluation.py:26-31)
//   ndcg@k  = dcg@k(order) / dcg@k(argsort(y_true)[::-1]), gains 2^y - 1, discounts log2(pos + 2)
// Since y_score is a decreasing function of the rank, position(j) = #{k : rank_k < rank_j} +
// #{k : rank_k == rank_j and k > j}: the reversed stable argsort puts the LATER of two tied candidates
// first.  (numpy's default argsort is not stable above 16 elements, so the reference's own order among
// exactly tied scores is platform dependent; ties need bit-identical scores.)
//
// One warp per impression, ranks / labels staged in shared memory, float64 accumulation like numpy.
#include "common.cuh"

#include <algorithm>

namespace nrb {

constexpr int kMetWarps = 8;
constexpr int kMetCap = 512;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

__global__ void __launch_bounds__(kMetWarps * 32)
mind_metrics_kernel(const int32_t* ranks, const int8_t* labels, const int64_t* offsets, int64_t n_imp,
                    double* per_imp, double* sums) {
  __shared__ int32_t s_rank[kMetWarps][kMetCap];
  __shared__ int8_t s_lab[kMetWarps][kMetCap];
  __shared__ double s_part[kMetWarps][5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kMetWarps;
  double acc[5] = {0, 0, 0, 0, 0};
  for (int64_t imp = (int64_t)blockIdx.x * kMetWarps + warp; imp < n_imp; imp += stride) {
    const int64_t c0 = offsets[imp];
    const int n = (int)(offsets[imp + 1] - c0);
    const bool in_smem = n <= kMetCap;
    if (in_smem) {
      for (int k = lane; k < n; k += 32) {
        s_rank[warp][k] = ranks[c0 + k];
        s_lab[warp][k] = labels[c0 + k];
      }
      __syncwarp();
    }
    const int32_t* rk = in_smem ? s_rank[warp] : ranks + c0;
    const int8_t* lb = in_smem ? s_lab[warp] : labels + c0;
    double n_pos = 0, n_neg = 0, auc_num = 0, rr = 0, dcg5 = 0, dcg10 = 0, idcg5 = 0, idcg10 = 0, ysum = 0;
    bool bad = false;
    for (int j = lane; j < n; j += 32) {
      const int rj = rk[j];
      const int yj = lb[j];
      bad |= (rj <= 0);  // 0 = NaN sentinel of nrb_dense_rank
      int pos = 0, ipos = 0, less_neg = 0, eq_neg = 0;
      for (int k = 0; k < n; ++k) {
        const int rk_ = rk[k];
        const int yk = lb[k];
        pos += (rk_ < rj) | ((rk_ == rj) & (k > j));
        ipos += (yk > yj) | ((yk == yj) & (k > j));
        if (yj > 0 && yk <= 0) {  // (positive j, negative k) pairs
          less_neg += (rj < rk_);
          eq_neg += (rj == rk_);
        }
      }
      const double gain = exp2((double)yj) - 1.0;
      if (yj > 0) {
        n_pos += 1;
        auc_num += (double)less_neg + 0.5 * (double)eq_neg;
      } else {
        n_neg += 1;
      }
      ysum += (double)yj;
      rr += (double)yj / (double)(pos + 1);
      const double disc = 1.0 / log2((double)pos + 2.0);
      const double idisc = 1.0 / log2((double)ipos + 2.0);
      if (pos < 5) dcg5 += gain * disc;
      if (pos < 10) dcg10 += gain * disc;
      if (ipos < 5) idcg5 += gain * idisc;
      if (ipos < 10) idcg10 += gain * idisc;
    }
    n_pos = warp_sum_f64(n_pos);
    n_neg = warp_sum_f64(n_neg);
    auc_num = warp_sum_f64(auc_num);
    rr = warp_sum_f64(rr);
    ysum = warp_sum_f64(ysum);
    dcg5 = warp_sum_f64(dcg5);
    dcg10 = warp_sum_f64(dcg10);
    idcg5 = warp_sum_f64(idcg5);
    idcg10 = warp_sum_f64(idcg10);
    bad = __any_sync(kFullMask, bad);
    const bool valid = !bad && n_pos > 0 && n_neg > 0;  // roc_auc_score raises with a single class
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const double auc = valid ? auc_num / (n_pos * n_neg) : qnan;
    const double mrr = valid ? rr / ysum : qnan;
    const double nd5 = valid ? dcg5 / idcg5 : qnan;
    const double nd10 = valid ? dcg10 / idcg10 : qnan;
    if (lane == 0) {
      if (per_imp != nullptr) {
        per_imp[imp * 4 + 0] = auc;
        per_imp[imp * 4 + 1] = mrr;
        per_imp[imp * 4 + 2] = nd5;
        per_imp[imp * 4 + 3] = nd10;
      }
      if (valid) {
        acc[0] += auc;
        acc[1] += mrr;
        acc[2] += nd5;
        acc[3] += nd10;
        acc[4] += 1.0;
      }
    }
    __syncwarp();
  }
  if (lane == 0)
    for (int q = 0; q < 5; ++q) s_part[warp][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0;
    for (int w = 0; w < kMetWarps; ++w) t += s_part[w][threadIdx.x];
    if (t != 0) atomicAdd(&sums[threadIdx.x], t);
  }
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_mind_metrics(const int32_t* ranks, const int8_t* labels, const int64_t* offsets, int64_t n_imp,
                                double* per_imp_out, double* sums, nrb_stream_t stream) {
  NRB_REQUIRE(n_imp >= 0, "nrb_mind_metrics: n_imp < 0");
  if (n_imp == 0) return NRB_OK;
  NRB_REQUIRE(ranks && labels && offsets && sums, "nrb_mind_metrics: null pointer");
  const int64_t want = (n_imp + kMetWarps - 1) / kMetWarps;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count_cached() * 16);
  mind_metrics_kernel<<<grid, kMetWarps * 32, 0, as_stream(stream)>>>(ranks, labels, offsets, n_imp, per_imp_out,
                                                                        sums);
  note_launch();
  NRB_CUDA_CHECK(cudaGetLastError());
  return NRB_OK;
}
