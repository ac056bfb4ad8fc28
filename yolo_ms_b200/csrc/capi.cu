// Error reporting and bookkeeping shared by every entry point of libyms_b200.so.
#include "common.cuh"

#include <stdarg.h>

namespace yms {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

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
