// C-ABI plumbing shared by every entry point: thread-local error string, device attribute cache.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ldt_b200.h"

namespace ldt {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
  if (e == cudaSuccess) return LDT_OK;
  set_last_error("CUDA error %d (%s) at %s:%d in `%s`", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
  return LDT_ERR_CUDA;
}

void set_pdl(int on);

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < LDT_MAX_DEVICES ? dev : LDT_MAX_DEVICES - 1;
}

int num_sms() {
  static PerDevice<int> cached;   // zero-initialised: 0 = not queried on this device yet
  int& c = cached.get();
  if (c == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, current_device()) == cudaSuccess && n > 0)
      c = n;
    else
      c = 148;  // B200
  }
  return c;
}

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("LDT_PDL");
    g_pdl = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_pdl != 0;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }

}  // namespace ldt

extern "C" int ldt_set_pdl(int enable) {
  ldt::set_pdl(enable);
  return LDT_OK;
}
extern "C" int ldt_abi_version(void) { return LDT_ABI_VERSION; }
extern "C" const char* ldt_last_error_string(void) { return ldt::g_err; }
extern "C" int ldt_device_sm_count(void) { return ldt::num_sms(); }
