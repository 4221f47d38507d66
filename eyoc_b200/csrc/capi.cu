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

// ------------------------------------------------------------------------------------------------ index compositions
// out[i, :] = src[idx[i], :]: the row gathers of find_corr / random_sample / match_pair (scripts/test_kitti.py:36-42,
// 69-73, scripts/SC2_PCR/SC2_PCR.py:290-305: `F[inds]`, `xyz[inds]`) for a whole block of pairs in one launch.
// c floats per row (c % 4 == 0: 16-byte pieces, c / 4 lanes per row; else scalar), idx int64, rows independent.
namespace {
__global__ void gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, long long m, int c,
                                   float* __restrict__ out) {
    if ((c & 3) == 0) {
        const int c4 = c >> 2;
        const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= m * c4) return;
        const long long i = e / c4;
        const int j = (int)(e - i * c4);
        reinterpret_cast<float4*>(out)[e] = __ldg(reinterpret_cast<const float4*>(src + (size_t)idx[i] * c) + j);
    } else {
        const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= m * c) return;
        const long long i = e / c;
        out[e] = __ldg(src + (size_t)idx[i] * c + (e - i * c));
    }
}
}  // namespace

extern "C" int eyoc_gather_rows(const float* src, const int64_t* idx, int64_t m, int c, float* out, cudaStream_t stream) {
    EYOC_CHECK_ARG(src && idx && out && m >= 0 && c >= 1, "eyoc_gather_rows: bad argument");
    if (m == 0) return EYOC_OK;
    const long long n = (c & 3) == 0 ? (long long)m * (c >> 2) : (long long)m * c;
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, idx, m, c, out);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
