// Library-level entry points of the nrb200 C ABI: version, errors, device checks.
#include "common.cuh"

#include <atomic>
#include <cstring>

namespace nrb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count_cached() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace nrb

using namespace nrb;

extern "C" const char* nrb_version(void) { return "nrb200 0.1.0 (sm_100a)"; }

extern "C" const char* nrb_last_error(void) { return g_err; }

extern "C" int nrb_check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device visible (%s); nrb200 has no CPU fallback", cudaGetErrorString(e));
    return NRB_E_CUDA;
  }
  NRB_REQUIRE(device >= 0 && device < n, "device %d out of range (count %d)", device, n);
  int major = 0, minor = 0;
  NRB_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  NRB_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    set_error("device %d is sm_%d%d; nrb200 is built for sm_100a (B200) only", device, major, minor);
    return NRB_E_ARCH;
  }
  return NRB_OK;
}

extern "C" long long nrb_kernel_launches(void) { return nrb::launches(); }

extern "C" int nrb_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  return n;
}
