// Library-wide plumbing: ABI version, thread-local error string, stats vector reset.
#include <stdarg.h>

#include "common.cuh"

namespace sg {
namespace {
thread_local char g_error[512] = "";
}
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
namespace {
__global__ void stats_reset_kernel(double* stats) {
    const int i = threadIdx.x;
    if (i < SG_STATS_LEN) stats[i] = (i == 3) ? 1e300 : (i == 4) ? -1e300 : 0.0;
}
}  // namespace
}  // namespace sg

extern "C" {

int sg_abi_version(void) { return SG_ABI_VERSION; }

const char* sg_last_error(void) { return sg::g_error; }

int sg_device_sm_count(void) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return sms;
}

int sg_stats_reset(double* stats, void* stream) {
    SG_CHECK_ARG(stats != nullptr, "stats_reset: NULL");
    sg::stats_reset_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(stats);
    return sg::launch_status("stats_reset_kernel");
}

}  // extern "C"
