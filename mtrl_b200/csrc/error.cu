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

// The host -> device transfer of one batch as ONE call: the update takes five host arrays (observations, actions,
// next_observations, dones, rewards: mtrl/types.py:30-35, handed over by `self.update(data)`, base.py:221); a host that copies
// them one framework call at a time spends more CPU time dispatching the copies (20-35 us each from Python) than PCIe spends
// moving the 4.7 MB.  Pinned sources are copied asynchronously on `stream` (they must stay untouched until the stream has passed
// the copies); pageable ones return once staged, as cudaMemcpyAsync defines it.
extern "C" int mtrl_memcpy_h2d_batch(int n, void* const* dst, const void* const* src, const long long* bytes, void* stream) {
  MTRL_REQUIRE(n >= 0 && (n == 0 || (dst && src && bytes)), "mtrl_memcpy_h2d_batch: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < n; ++i) {
    MTRL_REQUIRE(dst[i] && src[i] && bytes[i] >= 0, "mtrl_memcpy_h2d_batch: bad entry %d", i);
    if (bytes[i]) MTRL_CUDA_CHECK(cudaMemcpyAsync(dst[i], src[i], static_cast<size_t>(bytes[i]), cudaMemcpyHostToDevice, st));
  }
  return MTRL_OK;
}
