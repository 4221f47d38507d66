// Shared helpers for the eyoc_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define EYOC_OK 0
#define EYOC_ERR_ARG (-1)
#define EYOC_ERR_CUDA (-2)
#define EYOC_ERR_WORKSPACE (-3)
#define EYOC_ERR_DEGENERATE (-4)

void eyoc_set_error(const char* fmt, ...);

#define EYOC_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            eyoc_set_error(__VA_ARGS__);          \
            return EYOC_ERR_ARG;                  \
        }                                         \
    } while (0)

#define EYOC_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            eyoc_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return EYOC_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

// every kernel launch of the library passes through here: error check + launch accounting
extern unsigned long long g_eyoc_launches;
#define EYOC_LAUNCH_CHECK()            \
    do {                               \
        ++g_eyoc_launches;             \
        EYOC_CUDA(cudaGetLastError()); \
    } while (0)

static inline size_t eyoc_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-owned workspace.
struct WsCarver {
    char* base;
    size_t off;
    size_t cap;
    bool ok;
    WsCarver(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes), ok(true) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = eyoc_align(n * sizeof(T));
        T* r = (T*)(base + off);
        off += bytes;
        if (off > cap && base != nullptr) ok = false;
        return r;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// torch.norm(p_i - p_j) on CPU == sequential-FMA sum of squares, then correctly rounded sqrt
// (pinned empirically in tests/test_oracle_arith.py).  Intrinsics keep nvcc from re-contracting.
__device__ __forceinline__ float dist3_fma(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    float s = __fmul_rn(dx, dx);
    s = __fmaf_rn(dy, dy, s);
    s = __fmaf_rn(dz, dz, s);
    return __fsqrt_rn(s);
}
// ((a-b)**2).sum(-1)**0.5 on the reference == separately rounded squares, (x2+y2)+z2, sqrt.
__device__ __forceinline__ float dist3_sum(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return __fsqrt_rn(s);
}
