// Library-wide plumbing of libpmgt_b200.so: thread-local error string, ABI
// version, cached device properties.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace pmgt {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

static int g_pdl = -1;

bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("PMGT_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl == 1;
}

int set_alternate(int enabled);

int set_pdl(int enabled) {
  const int prev = pdl_enabled() ? 1 : 0;
  g_pdl = enabled ? 1 : 0;
  return prev;
}

static int g_alternate = -1;
static int g_order = 0;

int next_tile_order() {
  if (g_alternate < 0) {
    const char* e = getenv("PMGT_ALTERNATE");
    g_alternate = (e && e[0] == '0') ? 0 : 1;
  }
  if (!g_alternate) return 0;
  g_order ^= 1;
  return g_order;
}

int set_alternate(int enabled) {
  const int prev = g_alternate < 0 ? 1 : g_alternate;
  g_alternate = enabled ? 1 : 0;
  g_order = 0;
  return prev;
}

}  // namespace pmgt

extern "C" {
int pmgt_abi_version(void) { return PMGT_B200_ABI_VERSION; }
const char* pmgt_last_error(void) { return pmgt::g_err; }
int pmgt_set_pdl(int enabled) { return pmgt::set_pdl(enabled); }
int pmgt_set_alternate_order(int enabled) { return pmgt::set_alternate(enabled); }
}
