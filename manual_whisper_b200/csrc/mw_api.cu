// Library-wide C ABI plumbing: version, thread-local error string, launch counter.
#include "mw_common.cuh"
#include <stdarg.h>

namespace mw {
static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};
thread_local bool t_capturing = false;
thread_local uint64_t t_captured = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace mw

extern "C" int mw_abi_version(void) { return MW_ABI_VERSION; }
extern "C" const char* mw_last_error(void) { return mw::g_err; }
extern "C" uint64_t mw_launch_count(void) { return mw::g_launches.load(); }
extern "C" int mw_storage_dtype(void) {
#ifdef MW_STORAGE_BF16
    return 1;
#else
    return 0;
#endif
}
