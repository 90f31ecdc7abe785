// host_common.cu -- library identification, error reporting and launch counting (C ABI).
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <string>

#include "host_common.h"
#include "lasr_b200.h"

namespace lasr {

static std::mutex g_err_mu;
static std::string g_err;
static std::atomic<uint64_t> g_launches{0};

int lasr_fail(const std::string& msg) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = msg;
  return 1;
}

int lasr_check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return lasr_fail(std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}

}  // namespace lasr

extern "C" int lasr_b200_abi_version(void) { return 8; }   // 8: + losses_forward_total / head_forward_total; 7: + lbs_pose_forward / backward; 6: + head_backward_packed / render_backward_packed; 5: + lasr_b200_raster_forward_bg (4: + chamfer, texture atlas, staging)

extern "C" const char* lasr_b200_last_error(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(lasr::g_err_mu);
  copy = lasr::g_err;
  return copy.c_str();
}

extern "C" uint64_t lasr_b200_launch_count(void) { return lasr::g_launches.load(); }
