// Library-wide C ABI plumbing: version, thread-local error text, launch counter.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace c3d {

static thread_local char t_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

// ---- optional per-kernel device timing (CUDA events on the launching stream) ----
namespace {
struct Rec { const char* name; cudaEvent_t a, b; };
std::mutex g_mu;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
bool g_prof_on = false;
std::string g_filter;
cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace

KernelTimer::KernelTimer(const char* name, cudaStream_t stream) : name_(nullptr), stream_(stream) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_prof_on || (!g_filter.empty() && g_filter != name)) return;
  name_ = name;
  a_ = get_event(); b_ = get_event();
  cudaEventRecord(a_, stream_);
}
KernelTimer::~KernelTimer() {
  if (!name_) return;
  cudaEventRecord(b_, stream_);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back(Rec{name_, a_, b_});
}

}  // namespace c3d

extern "C" int c3d_profile_enable(const char* kernel_name) {
  std::lock_guard<std::mutex> lk(c3d::g_mu);
  if (kernel_name == nullptr) { c3d::g_prof_on = false; return C3D_OK; }
  c3d::g_prof_on = true;
  c3d::g_filter = kernel_name;  // "" = every kernel
  return C3D_OK;
}

extern "C" int c3d_profile_read(const char* kernel_name, double* total_ms, long long* count) {
  std::lock_guard<std::mutex> lk(c3d::g_mu);
  double tot = 0; long long n = 0;
  for (auto& r : c3d::g_recs) {
    if (kernel_name && kernel_name[0] && strcmp(kernel_name, r.name) != 0) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { tot += ms; ++n; }
  }
  if (total_ms) *total_ms = tot;
  if (count) *count = n;
  return C3D_OK;
}

extern "C" int c3d_profile_names(char* buf, int buf_len) {
  std::lock_guard<std::mutex> lk(c3d::g_mu);
  std::string out;
  std::vector<const char*> seen;
  for (auto& r : c3d::g_recs) {
    bool dup = false;
    for (auto* s : seen) if (strcmp(s, r.name) == 0) { dup = true; break; }
    if (!dup) { seen.push_back(r.name); if (!out.empty()) out += ","; out += r.name; }
  }
  if (!buf || buf_len <= 0) return C3D_INVALID_ARGUMENT;
  snprintf(buf, buf_len, "%s", out.c_str());
  return C3D_OK;
}

extern "C" int c3d_profile_timeline(char* buf, int buf_len) {
  // "name,start_us,end_us\n" per recorded launch, relative to the first record.
  std::lock_guard<std::mutex> lk(c3d::g_mu);
  if (!buf || buf_len <= 0) return C3D_INVALID_ARGUMENT;
  std::string out;
  if (!c3d::g_recs.empty()) {
    cudaEvent_t base = c3d::g_recs[0].a;
    for (auto& r : c3d::g_recs) {
      if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
      float t0 = 0, t1 = 0;
      if (cudaEventElapsedTime(&t0, base, r.a) != cudaSuccess) continue;
      if (cudaEventElapsedTime(&t1, base, r.b) != cudaSuccess) continue;
      char line[160];
      snprintf(line, sizeof(line), "%s,%.2f,%.2f\n", r.name, t0 * 1e3f, t1 * 1e3f);
      out += line;
    }
  }
  snprintf(buf, buf_len, "%s", out.c_str());
  return C3D_OK;
}

extern "C" int c3d_profile_reset(void) {
  std::lock_guard<std::mutex> lk(c3d::g_mu);
  for (auto& r : c3d::g_recs) { c3d::g_pool.push_back(r.a); c3d::g_pool.push_back(r.b); }
  c3d::g_recs.clear();
  return C3D_OK;
}

extern "C" int c3d_version(void) { return 100; }
extern "C" const char* c3d_last_error(void) { return c3d::t_error; }
extern "C" long long c3d_launch_count(void) { return c3d::g_launches.load(); }
