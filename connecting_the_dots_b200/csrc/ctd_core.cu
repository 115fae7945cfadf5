// Library-wide state of libctd_b200: error messages, launch accounting, options.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "ctd_common.cuh"

namespace ctd {

int g_force_generic = 0;  // tests: route every op through its generic kernel
static std::atomic<uint64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky launch error so the next call starts clean
    return fail(CTD_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  }
  return CTD_OK;
}

}  // namespace ctd

CTD_API const char* ctd_last_error(void) { return ctd::err_buf(); }
CTD_API const char* ctd_version(void) { return "ctd_b200 0.1 (sm_100a)"; }
CTD_API uint64_t ctd_launch_count(void) { return ctd::g_launches.load(std::memory_order_relaxed); }

CTD_API int ctd_set_option(const char* name, int value) {
  if (name && !strcmp(name, "force_generic")) {
    ctd::g_force_generic = value;
    return CTD_OK;
  }
  return ctd::fail(CTD_ERR_INVALID, "ctd_set_option: unknown option '%s'", name ? name : "(null)");
}
