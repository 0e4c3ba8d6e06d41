// cabi.cu -- error plumbing and bookkeeping shared by every C-ABI entry point.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace sb2 {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_error (int code, const char *fmt, ...)
{
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (g_err, sizeof (g_err), fmt, ap);
  va_end (ap);
  return code;
}

int check_cuda (cudaError_t e, const char *what)
{
  if (e == cudaSuccess) return SB2_OK;
  snprintf (g_err, sizeof (g_err), "%s: %s (%s)", what, cudaGetErrorString (e), cudaGetErrorName (e));
  return SB2_ERR_CUDA;
}

void count_launch (unsigned n) { g_launches.fetch_add (n, std::memory_order_relaxed); }

}  // namespace sb2

extern "C" const char *sb2_last_error (void) { return sb2::g_err; }
extern "C" int sb2_version (void) { return 1; }
extern "C" unsigned long long sb2_launch_count (void)
{
  return sb2::g_launches.load (std::memory_order_relaxed);
}
