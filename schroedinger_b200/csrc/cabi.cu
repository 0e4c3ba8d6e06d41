// cabi.cu -- error plumbing and bookkeeping shared by every C-ABI entry point.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

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

// ---- optional per-launch device timing (used by bench.py for the roofline) ----
// Events come from a pool that survives sb2_profile_reset, so a measured loop pays two event
// records per launch and no event creation.
struct ProfRec { char tag[48]; cudaEvent_t e0, e1; double bytes; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_events;
static std::atomic<int> g_prof_on{0};

bool profiling () { return g_prof_on.load (std::memory_order_relaxed) != 0; }

static cudaEvent_t prof_event ()
{
  if (!g_prof_events.empty ()) {
    cudaEvent_t e = g_prof_events.back ();
    g_prof_events.pop_back ();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate (&e);
  return e;
}

int prof_begin (const char *tag, double bytes, cudaStream_t st)
{
  if (!profiling ()) return -1;
  ProfRec r;
  snprintf (r.tag, sizeof (r.tag), "%s", tag);
  r.bytes = bytes;
  std::lock_guard<std::mutex> lk (g_prof_mu);
  r.e0 = prof_event ();
  r.e1 = prof_event ();
  cudaEventRecord (r.e0, st);
  g_prof.push_back (r);
  return (int) g_prof.size () - 1;
}

void prof_end (int id, cudaStream_t st)
{
  if (id < 0) return;
  std::lock_guard<std::mutex> lk (g_prof_mu);
  if (id < (int) g_prof.size ()) cudaEventRecord (g_prof[id].e1, st);
}

}  // namespace sb2

extern "C" const char *sb2_last_error (void) { return sb2::g_err; }
extern "C" int sb2_version (void) { return 1; }
extern "C" unsigned long long sb2_launch_count (void)
{
  return sb2::g_launches.load (std::memory_order_relaxed);
}

// Per-launch CUDA-event timing on the launching stream.  Enable, run, synchronise,
// then read the records back.
extern "C" void sb2_profile_enable (int on)
{
  sb2::g_prof_on.store (on, std::memory_order_relaxed);
}
extern "C" void sb2_profile_reset (void)
{
  std::lock_guard<std::mutex> lk (sb2::g_prof_mu);
  for (auto &r : sb2::g_prof) { sb2::g_prof_events.push_back (r.e0); sb2::g_prof_events.push_back (r.e1); }
  sb2::g_prof.clear ();
}
extern "C" int sb2_profile_count (void)
{
  std::lock_guard<std::mutex> lk (sb2::g_prof_mu);
  return (int) sb2::g_prof.size ();
}
extern "C" int sb2_profile_get (int i, char *tag, int tag_len, float *ms, double *bytes)
{
  std::lock_guard<std::mutex> lk (sb2::g_prof_mu);
  if (i < 0 || i >= (int) sb2::g_prof.size ()) return SB2_ERR_ARG;
  auto &r = sb2::g_prof[i];
  if (tag && tag_len > 0) snprintf (tag, tag_len, "%s", r.tag);
  if (bytes) *bytes = r.bytes;
  float t = 0.f;
  cudaError_t e = cudaEventElapsedTime (&t, r.e0, r.e1);
  if (ms) *ms = t;
  return e == cudaSuccess ? SB2_OK : SB2_ERR_CUDA;
}
