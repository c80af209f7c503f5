// Version / error-string / measurement entry points of the C ABI.
#include <atomic>
#include <mutex>
#include <vector>

#include "host_common.h"

namespace tcl {

static std::atomic<int64_t> g_launches{0};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
struct EvPair { cudaEvent_t a, b; int dev; };  // events belong to the device they were created on
static std::vector<EvPair> g_pool[TCL_K_COUNT];
static size_t g_used[TCL_K_COUNT];
static constexpr size_t kMaxPairs = 1 << 14;

void prof_begin(int id, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_used[id] >= kMaxPairs) return;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  if (g_used[id] < g_pool[id].size() && g_pool[id][g_used[id]].dev != dev) {  // pooled pair of another device: replace it
    cudaEventDestroy(g_pool[id][g_used[id]].a);
    cudaEventDestroy(g_pool[id][g_used[id]].b);
    EvPair e;
    e.dev = dev;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) { g_pool[id].resize(g_used[id]); return; }
    g_pool[id][g_used[id]] = e;
  }
  if (g_used[id] == g_pool[id].size()) {
    EvPair e;
    e.dev = dev;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    g_pool[id].push_back(e);
  }
  cudaEventRecord(g_pool[id][g_used[id]].a, st);
}
void prof_end(int id, cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_used[id] >= g_pool[id].size()) return;
  cudaEventRecord(g_pool[id][g_used[id]].b, st);
  g_used[id]++;
}

}  // namespace tcl

using namespace tcl;

extern "C" int tcl_version(void) { return TCL_ABI_VERSION; }
extern "C" const char* tcl_last_error_string(void) { return last_error_buf(); }
extern "C" int64_t tcl_launch_count(void) { return g_launches.load(); }

extern "C" int tcl_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < TCL_K_COUNT; ++i) g_used[i] = 0;
  g_prof_on.store(on != 0);
  return TCL_OK;
}

extern "C" int tcl_profile_read(int kernel_id, double* total_ms, int64_t* launches) {
  TCL_REQUIRE(kernel_id >= 0 && kernel_id < TCL_K_COUNT && total_ms && launches, TCL_ERR_BAD_ARG, "profile_read: bad argument");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double tot = 0.0;
  for (size_t i = 0; i < g_used[kernel_id]; ++i) {
    TCL_CHECK_CUDA(cudaEventSynchronize(g_pool[kernel_id][i].b));
    float ms = 0.f;
    TCL_CHECK_CUDA(cudaEventElapsedTime(&ms, g_pool[kernel_id][i].a, g_pool[kernel_id][i].b));
    tot += ms;
  }
  *total_ms = tot;
  *launches = static_cast<int64_t>(g_used[kernel_id]);
  return TCL_OK;
}

extern "C" int tcl_copy_rows(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes,
                             int64_t width_bytes, int64_t rows, void* stream) {
  TCL_REQUIRE(dst && src, TCL_ERR_BAD_ARG, "copy_rows: null pointer");
  TCL_REQUIRE(rows >= 0 && width_bytes >= 1 && dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes,
              TCL_ERR_BAD_SHAPE, "copy_rows: rows %lld, width %lld, pitches %lld / %lld", (long long)rows,
              (long long)width_bytes, (long long)dst_pitch_bytes, (long long)src_pitch_bytes);
  if (rows == 0) return TCL_OK;
  TCL_CHECK_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(dst_pitch_bytes), src, static_cast<size_t>(src_pitch_bytes),
                                   static_cast<size_t>(width_bytes), static_cast<size_t>(rows), cudaMemcpyDeviceToDevice,
                                   static_cast<cudaStream_t>(stream)));
  return TCL_OK;
}
