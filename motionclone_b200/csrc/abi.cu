// Library-wide state of the C ABI: last-error string and launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "mc_common.cuh"

namespace mc {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky-less launch error so the caller's next launch is judged on its own
    return MC_E_CUDA;
  }
  return MC_OK;
}

}  // namespace mc

extern "C" int mc_abi_version(void) { return MC_ABI_VERSION; }
extern "C" const char* mc_last_error(void) { return mc::g_err; }
extern "C" uint64_t mc_launch_count(void) { return mc::g_launches.load(std::memory_order_relaxed); }
extern "C" void mc_reset_launch_count(void) { mc::g_launches.store(0, std::memory_order_relaxed); }
extern "C" void mc_add_launch_count(uint64_t n) { mc::g_launches.fetch_add(n, std::memory_order_relaxed); }
