// Error reporting shared by every translation unit of libpivp.so (thread-local message, no exceptions cross the ABI).
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace pivp {
static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
void note_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
static int g_pdl = -1;         // -1: read PIVP_PDL on first use (default on)
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("PIVP_PDL");
        g_pdl = e ? (atoi(e) != 0) : 1;
    }
    return g_pdl != 0;
}
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
/* Programmatic dependent launch on (1) / off (0) for every kernel launched from now on (captured graphs keep what they were captured with). */
int pivp_set_pdl(int on) { pivp::g_pdl = on ? 1 : 0; return PIVP_OK; }
int pivp_device_sync_check(void) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { pivp::set_error("device: %s", cudaGetErrorString(e)); return PIVP_ECUDA; }
    return PIVP_OK;
}
}
