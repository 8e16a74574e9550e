// Library-wide C ABI plumbing: version, thread-local error text, launch counter.
#include <stdarg.h>
#include <string.h>

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

}  // namespace c3d

extern "C" int c3d_version(void) { return 100; }
extern "C" const char* c3d_last_error(void) { return c3d::t_error; }
extern "C" long long c3d_launch_count(void) { return c3d::g_launches.load(); }
