// Process-wide state of the C ABI: error string, launch counter, device query.
#include <cstdarg>

#include "common.cuh"

namespace se3 {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<long long> g_tc_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

}  // namespace se3

extern "C" const char* se3_last_error(void) { return se3::g_err; }
extern "C" int se3_version(void) { return 100; }
extern "C" int64_t se3_launch_count(void) { return (int64_t)se3::g_launches.load(); }
extern "C" int64_t se3_tc_launch_count(void) { return (int64_t)se3::g_tc_launches.load(); }
