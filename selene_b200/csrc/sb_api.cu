// selene_b200 — C ABI plumbing: error string, version, launch counter, device info.
#include "sb_common.cuh"

#include <mutex>

unsigned long long g_sb_launches = 0;

static thread_local char t_sb_error[1024] = "";

void sb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_sb_error, sizeof(t_sb_error), fmt, ap);
    va_end(ap);
}

int sb_sm_count(int device) {
    static int cache[64];
    static std::once_flag once;
    std::call_once(once, [] { for (int i = 0; i < 64; ++i) cache[i] = 0; });
    if (device < 0 || device >= 64) return 148;
    if (cache[device] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cache[device] = v;
    }
    return cache[device];
}

extern "C" int sb_version(void) { return 100; }

extern "C" int sb_last_error(char* buf, int n) {
    if (!buf || n <= 0) return SB_ERR_ARG;
    strncpy(buf, t_sb_error, (size_t)n - 1);
    buf[n - 1] = 0;
    return SB_OK;
}

extern "C" unsigned long long sb_launch_count(void) { return g_sb_launches; }

extern "C" int sb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    cudaDeviceProp prop;
    SB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return SB_OK;
}
