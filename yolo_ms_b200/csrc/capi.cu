// Error reporting and bookkeeping shared by every entry point of libyms_b200.so.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

namespace yms {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
long long* g_prof_buf = nullptr;
DebugOptions g_opt = {0, 0, 0, 0, 8, 256, 0};

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

/* Debug hook (not part of the reference-facing ABI): experiment switches by name; returns 0, or YMS_E_ARG for an unknown name.
 * Affects plans created / kernels launched AFTER the call. */
extern "C" int yms_debug_set_option(const char* name, int value) {
    using yms::g_opt;
    if (!name) return yms::fail(YMS_E_ARG, "option: null name");
    if (!strcmp(name, "pdl_off")) g_opt.pdl_off = value;
    else if (!strcmp(name, "conv_half")) g_opt.conv_half = value;
    else if (!strcmp(name, "stem_gather")) g_opt.stem_gather = value;
    else if (!strcmp(name, "nms_groups")) g_opt.nms_groups = value;
    else if (!strcmp(name, "nms_mask_tiles")) g_opt.nms_mask_tiles = value;
    else if (!strcmp(name, "nms_poll_ns")) g_opt.nms_poll_ns = value;
    else if (!strcmp(name, "nms_sort_bitonic")) g_opt.nms_sort_bitonic = value;
    else return yms::fail(YMS_E_ARG, "option: unknown name '%s'", name);
    return 0;
}

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
