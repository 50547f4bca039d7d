// Error reporting shared by every translation unit of libpivp.so (thread-local message, no exceptions cross the ABI).
#include "common.cuh"
#include <stdarg.h>

namespace pivp {
static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
void note_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace pivp

extern "C" {
const char* pivp_last_error(void) { return pivp::g_err; }
int pivp_abi_version(void) { return 1; }
long pivp_launch_count(void) { return (long)__atomic_load_n(&pivp::g_launches, __ATOMIC_RELAXED); }
int pivp_device_sync_check(void) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { pivp::set_error("device: %s", cudaGetErrorString(e)); return PIVP_ECUDA; }
    return PIVP_OK;
}
}
