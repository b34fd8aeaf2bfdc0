// flowdec_b200 — C-ABI plumbing shared by all kernels: last-error string, launch checks.
#include "fd_common.cuh"

#include <cstdarg>
#include <cstdio>

namespace fd {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

}  // namespace fd

extern "C" const char* fd_last_error(void) { return fd::g_last_error; }

extern "C" int fd_abi_version(void) { return 1; }
