// libuvx core: error reporting, launch accounting.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "uvx_common.cuh"

namespace uvx {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UVX_PDL");
    // on by default since round 2 (UVX_PDL=0 turns it off): with the W boxes of a GEMM's first ring round requested before
    // griddepcontrol.wait the programmatic edges are worth 0.06-0.15 ms per prefill step (profiles/r2_ab_bench_v3..v5.txt);
    // round 1, without the early loads, measured them slower than plain graph edges
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return UVX_ERR_CUDA;
  }
  return UVX_OK;
}

}  // namespace uvx

extern "C" int uvx_abi_version(void) { return UVX_ABI_VERSION; }
extern "C" const char* uvx_last_error(void) { return uvx::g_err; }
extern "C" int64_t uvx_launch_count(void) { return uvx::g_launches.load(std::memory_order_relaxed); }
