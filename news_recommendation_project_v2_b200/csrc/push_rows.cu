// Row-sharded table all-gather over NVLink peer memory (BASELINE configs[4]).
//
// Each rank transforms its own shard of the table; this kernel then PUSHES the freshly produced rows
// into every peer's copy of the full table with plain 128-bit stores on peer-mapped (symmetric) memory:
// NVSwitch gives every GPU full bandwidth to every peer, so one pass over the local chunk feeds all
// `world` destinations (the row is read once from local HBM / L2 and written `world` times).  Launched on
// the compute stream right behind the transform of the chunk, it overlaps with the next chunk's GEMMs.
// fp32 sources can be converted to bf16 on the way (the transform's residual stream is fp32).
#include "common.cuh"

#include <algorithm>

namespace nrb {

constexpr int kMaxPeers = 16;

struct PushParams {
  const char* src;
  int64_t src_stride_bytes;
  int64_t n_rows;
  int dim;
  int src_f32_to_bf16;  // 1: src is fp32, destinations are bf16
  int world;
  int64_t dst_row_offset;
  int64_t dst_stride_bytes;
  char* dst[kMaxPeers];
};

__global__ void __launch_bounds__(256)
push_rows_kernel(const PushParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int out_vecs = p.src_f32_to_bf16 ? p.dim / 8 : 0;  // 16-byte bf16 output vectors per row
  for (int64_t r = warp_id; r < p.n_rows; r += n_warps) {
    const char* s = p.src + r * p.src_stride_bytes;
    const int64_t doff = (p.dst_row_offset + r) * p.dst_stride_bytes;
    if (p.src_f32_to_bf16) {
      for (int v = lane; v < out_vecs; v += 32) {
        const float4 a = *reinterpret_cast<const float4*>(s + (size_t)v * 32);
        const float4 b = *reinterpret_cast<const float4*>(s + (size_t)v * 32 + 16);
        const uint4 o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y),
                                   pack_bf16x2(b.z, b.w));
        for (int g = 0; g < p.world; ++g) *reinterpret_cast<uint4*>(p.dst[g] + doff + (size_t)v * 16) = o;
      }
    } else {
      const int nv = p.dim;  // here `dim` counts 16-byte vectors per row
      for (int v = lane; v < nv; v += 32) {
        const uint4 o = *reinterpret_cast<const uint4*>(s + (size_t)v * 16);
        for (int g = 0; g < p.world; ++g) *reinterpret_cast<uint4*>(p.dst[g] + doff + (size_t)v * 16) = o;
      }
    }
  }
}

// ---- multicast pushes that ride inside the GEMM kernels ----------------------------------------------------------
// Pending segments of this host thread.  Every tcgen05 GEMM launch takes rows [done, done + share) of each segment,
// share = ceil(rows / spread): after `spread` launches everything is on its way; nrb_push_flush sends the rest.
struct PendingPush {
  PushSeg seg[kMaxPushSegs];
  int64_t done[kMaxPushSegs];
  int64_t share[kMaxPushSegs];
  int n_seg = 0;
  int device = -1;  // only GEMM launches on the device the segments live on may carry them
};
static thread_local PendingPush g_pending;

static void cut_share(PushJob* job, bool everything) {
  job->n_seg = 0;
  PendingPush& q = g_pending;
  if (q.n_seg == 0) return;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev != q.device) return;  // another device's launch: not its rows
  bool left = false;
  for (int i = 0; i < q.n_seg; ++i) {
    const int64_t remain = q.seg[i].n_rows - q.done[i];
    if (remain <= 0) continue;
    const int64_t take = everything ? remain : std::min<int64_t>(remain, q.share[i]);
    PushSeg sg = q.seg[i];
    sg.src += q.done[i] * sg.src_stride_bytes;
    sg.dst += q.done[i] * sg.dst_stride_bytes;
    sg.n_rows = take;
    job->seg[job->n_seg++] = sg;
    q.done[i] += take;
    left = left || q.done[i] < q.seg[i].n_rows;
  }
  if (!left) q.n_seg = 0;
}

void take_push_share(PushJob* job) { cut_share(job, false); }

__global__ void __launch_bounds__(128)
push_multicast_kernel(const PushJob job) {
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  push_job_warp(job, warp_id, n_warps, threadIdx.x & 31);
}

}  // namespace nrb

using namespace nrb;

extern "C" int nrb_push_attach(const nrb_push_seg* segs, int n_segs, int spread) {
  NRB_REQUIRE(n_segs >= 0 && n_segs <= kMaxPushSegs, "nrb_push_attach: at most %d segments", kMaxPushSegs);
  NRB_REQUIRE(spread >= 1, "nrb_push_attach: spread must be >= 1");
  NRB_REQUIRE(g_pending.n_seg == 0, "nrb_push_attach: earlier segments are still pending (call nrb_push_flush)");
  NRB_REQUIRE(n_segs == 0 || segs != nullptr, "nrb_push_attach: null pointer");
  for (int i = 0; i < n_segs; ++i) {
    const nrb_push_seg& u = segs[i];
    NRB_REQUIRE(u.src && u.mc_dst && u.n_rows >= 0 && u.dim > 0, "nrb_push_attach: bad segment %d", i);
    NRB_REQUIRE((u.src_dtype == NRB_F32 || u.src_dtype == NRB_BF16) && (u.dst_dtype == NRB_F32 || u.dst_dtype == NRB_BF16),
                "nrb_push_attach: bad dtype");
    NRB_REQUIRE(u.src_dtype == u.dst_dtype || (u.src_dtype == NRB_F32 && u.dst_dtype == NRB_BF16),
                "nrb_push_attach: only same-dtype or fp32 -> bf16 pushes");
    const int ses = u.src_dtype == NRB_F32 ? 4 : 2, des = u.dst_dtype == NRB_F32 ? 4 : 2;
    NRB_REQUIRE((u.dim * des) % 16 == 0 && (u.src_stride * ses) % 16 == 0 && (u.dst_stride * des) % 16 == 0 &&
                    (reinterpret_cast<uintptr_t>(u.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(u.mc_dst) & 15) == 0,
                "nrb_push_attach: rows must be 16-byte aligned multiples");
    PushSeg& sg = g_pending.seg[i];
    sg.src = (const char*)u.src;
    sg.dst = (char*)u.mc_dst + u.dst_row_offset * u.dst_stride * des;
    sg.n_rows = u.n_rows;
    sg.src_stride_bytes = u.src_stride * ses;
    sg.dst_stride_bytes = u.dst_stride * des;
    sg.vecs_per_row = u.dim * des / 16;
    sg.f32_to_bf16 = (u.src_dtype == NRB_F32 && u.dst_dtype == NRB_BF16) ? 1 : 0;
    g_pending.done[i] = 0;
    g_pending.share[i] = (u.n_rows + spread - 1) / spread;
  }
  g_pending.n_seg = n_segs;
  NRB_CUDA_CHECK(cudaGetDevice(&g_pending.device));
  return NRB_OK;
}

extern "C" void nrb_push_cancel(void) { g_pending.n_seg = 0; }

extern "C" int nrb_push_flush(nrb_stream_t stream) {
  PushJob job;
  cut_share(&job, true);
  if (job.n_seg == 0) return NRB_OK;
  int64_t vecs = 0;
  for (int i = 0; i < job.n_seg; ++i) vecs += job.seg[i].n_rows * job.seg[i].vecs_per_row;
  const int64_t want = (vecs + 128 * 4 - 1) / (128 * 4);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count_cached() * 8));
  push_multicast_kernel<<<grid, 128, 0, as_stream(stream)>>>(job);
  note_launch();
  NRB_CUDA_CHECK(cudaGetLastError());
  return NRB_OK;
}

extern "C" int nrb_push_rows(const void* src, int src_dtype, int64_t src_stride, int64_t n_rows, int dim,
                             void* const* dst_ptrs_host, int world, int dst_dtype, int64_t dst_row_offset,
                             int64_t dst_stride, nrb_stream_t stream) {
  NRB_REQUIRE(src && dst_ptrs_host, "nrb_push_rows: null pointer");
  NRB_REQUIRE(world >= 1 && world <= kMaxPeers, "nrb_push_rows: world must be in [1, %d]", kMaxPeers);
  NRB_REQUIRE(n_rows >= 0 && dim > 0, "nrb_push_rows: bad sizes");
  NRB_REQUIRE((src_dtype == NRB_F32 || src_dtype == NRB_BF16) && (dst_dtype == NRB_F32 || dst_dtype == NRB_BF16),
              "nrb_push_rows: bad dtype");
  NRB_REQUIRE(src_dtype == dst_dtype || (src_dtype == NRB_F32 && dst_dtype == NRB_BF16),
              "nrb_push_rows: only same-dtype or fp32 -> bf16 pushes");
  if (n_rows == 0) return NRB_OK;
  const int ses = src_dtype == NRB_F32 ? 4 : 2, des = dst_dtype == NRB_F32 ? 4 : 2;
  NRB_REQUIRE((dim * des) % 16 == 0 && (src_stride * ses) % 16 == 0 && (dst_stride * des) % 16 == 0,
              "nrb_push_rows: rows must be 16-byte multiples");
  PushParams p;
  p.src = (const char*)src;
  p.src_stride_bytes = src_stride * ses;
  p.n_rows = n_rows;
  p.src_f32_to_bf16 = (src_dtype == NRB_F32 && dst_dtype == NRB_BF16) ? 1 : 0;
  p.dim = p.src_f32_to_bf16 ? dim : dim * des / 16;
  p.world = world;
  p.dst_row_offset = dst_row_offset;
  p.dst_stride_bytes = dst_stride * des;
  for (int g = 0; g < kMaxPeers; ++g) p.dst[g] = g < world ? (char*)dst_ptrs_host[g] : nullptr;
  for (int g = 0; g < world; ++g) NRB_REQUIRE(p.dst[g] != nullptr, "nrb_push_rows: null destination %d", g);
  const int64_t want = (n_rows + 7) / 8;
  const int grid = (int)std::min<int64_t>(want, (int64_t)sm_count_cached() * 8);
  push_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  note_launch();
  NRB_CUDA_CHECK(cudaGetLastError());
  return NRB_OK;
}

extern "C" int nrb_push_bytes(const void* src, int64_t n_bytes, void* const* dst_ptrs_host, int world,
                              int64_t dst_byte_offset, nrb_stream_t stream) {
  NRB_REQUIRE(src && dst_ptrs_host, "nrb_push_bytes: null pointer");
  NRB_REQUIRE(world >= 1 && world <= kMaxPeers, "nrb_push_bytes: world must be in [1, %d]", kMaxPeers);
  NRB_REQUIRE(n_bytes >= 0 && dst_byte_offset >= 0, "nrb_push_bytes: bad sizes");
  if (n_bytes == 0) return NRB_OK;
  for (int g = 0; g < world; ++g) {
    NRB_REQUIRE(dst_ptrs_host[g] != nullptr, "nrb_push_bytes: null destination %d", g);
    char* dst = (char*)dst_ptrs_host[g] + dst_byte_offset;
    if ((const void*)dst == src) continue;  // already in place in the local table
    NRB_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)n_bytes, cudaMemcpyDefault, as_stream(stream)));
  }
  return NRB_OK;
}
