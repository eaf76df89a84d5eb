// Thread-local error text behind the C-ABI (mtrl_last_error).
#include <stdarg.h>

#include "common.cuh"
#include "mtrl_b200.h"

static thread_local char g_err[1024] = "";

void mtrl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mtrl_last_error(void) { return g_err; }
extern "C" int mtrl_abi_version(void) { return 6; }
