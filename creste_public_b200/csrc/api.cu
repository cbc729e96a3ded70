// api.cu -- version / error plumbing of the C ABI (include/creste_b200.h).
#include "common.cuh"

namespace creste {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }
int num_sms() {
  static int cached = -1;
  if (cached >= 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  cached = n;
  return n;
}
}  // namespace creste

extern "C" int creste_version(void) { return 100; }
extern "C" const char* creste_last_error(void) { return creste::g_err; }
extern "C" int creste_num_sms(void) { return creste::num_sms(); }
extern "C" unsigned long long creste_launch_count(void) { return creste::g_launches; }
