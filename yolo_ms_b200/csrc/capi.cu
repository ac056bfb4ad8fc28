// Error reporting and bookkeeping shared by every entry point of libyms_b200.so.
#include "common.cuh"

#include <stdarg.h>

namespace yms {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
long long* g_prof_buf = nullptr;

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace yms

extern "C" int yms_abi_version(void) { return YMS_ABI_VERSION; }
extern "C" const char* yms_last_error(void) { return yms::g_err; }
extern "C" long long yms_launch_count(void) { return yms::g_launches.load(); }

/* Debug hook (not part of the reference-facing ABI): device buffer [148][16] int64 that -DYMS_PROF builds of
 * the conv kernels fill with per-role cycle counters; returns 1 in profiling builds, 0 otherwise. */
extern "C" int yms_debug_set_prof(void* dev_buf) {
    yms::g_prof_buf = reinterpret_cast<long long*>(dev_buf);
#ifdef YMS_PROF
    return 1;
#else
    return 0;
#endif
}
