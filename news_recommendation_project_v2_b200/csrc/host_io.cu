// Host-side staging helper for the packed token file (the data format in front of stage A).
//
// nrb_host_copy: memcpy of one contiguous byte range (page cache -> pinned staging buffer) split over worker
// threads.  The reference reads one pickled blob per item through sqlite and pads on the CPU
// (data_utils.py:878-933); here a chunk of items is ONE byte range of the mapped file, and moving it into the
// DMA-able buffer is the only host work left on the path -- at PCIe Gen5 rates a single memcpy thread is the
// bottleneck, and Python threads share a GIL.
#include "common.cuh"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

using namespace nrb;

extern "C" int nrb_host_copy(void* dst_host, const void* src_host, int64_t n_bytes, int n_threads) {
  NRB_REQUIRE(n_bytes >= 0 && n_threads >= 1, "nrb_host_copy: bad sizes");
  if (n_bytes == 0) return NRB_OK;
  NRB_REQUIRE(dst_host && src_host, "nrb_host_copy: null pointer");
  const int64_t kMinSlice = 1 << 20;  // below 1 MiB per thread the spawn costs more than it buys
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_bytes / kMinSlice));
  if (nt == 1) {
    std::memcpy(dst_host, src_host, (size_t)n_bytes);
    return NRB_OK;
  }
  // 4 KiB-aligned slices: every thread walks whole pages of the mapping
  const int64_t slice = ((n_bytes + nt - 1) / nt + 4095) / 4096 * 4096;
  std::vector<std::thread> workers;
  workers.reserve(nt);
  for (int t = 0; t < nt; ++t) {
    const int64_t a = (int64_t)t * slice, b = std::min<int64_t>(n_bytes, a + slice);
    if (a >= b) break;
    workers.emplace_back([=] { std::memcpy((char*)dst_host + a, (const char*)src_host + a, (size_t)(b - a)); });
  }
  for (auto& w : workers) w.join();
  return NRB_OK;
}
