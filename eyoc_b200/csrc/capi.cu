// C-ABI plumbing: version + thread-local error string.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void eyoc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

unsigned long long g_eyoc_launches = 0;

extern "C" int eyoc_version(void) { return 100; }

extern "C" unsigned long long eyoc_launch_count(void) { return g_eyoc_launches; }

extern "C" const char* eyoc_last_error(void) { return g_err; }

extern "C" int eyoc_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    cudaDeviceProp p;
    EYOC_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return EYOC_OK;
}
