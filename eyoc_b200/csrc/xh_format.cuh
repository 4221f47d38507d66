// The split-half activation row format (include/eyoc_b200.h): per 32-channel chunk 128 bytes = 32 fp16 hi | 32 fp16 lo',
// x = hi + lo' / 2048.  Shared by the tensor-core convolution (sparse_conv_h.cu) and the stem convolution (coordmap.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace {

constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

__device__ __forceinline__ void xh_split(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(__fmul_rn(__fsub_rn(x, __half2float(hi)), LO_SCALE));
}
__device__ __forceinline__ float xh_join(__half hi, __half lo) { return __fmaf_rn(__half2float(lo), LO_INV, __half2float(hi)); }

// 8 consecutive channels of a split-half row: 16 bytes of hi, 16 bytes of lo' 64 bytes further
__device__ __forceinline__ void xh_load8(const uint8_t* row, int col, float* y) {
    const uint8_t* p = row + (col >> 5) * 128 + (col & 31) * 2;
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(p + 64));
    const __half2* hh = reinterpret_cast<const __half2*>(&h);
    const __half2* ll = reinterpret_cast<const __half2*>(&l);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        y[2 * i] = xh_join(__low2half(hh[i]), __low2half(ll[i]));
        y[2 * i + 1] = xh_join(__high2half(hh[i]), __high2half(ll[i]));
    }
}
__device__ __forceinline__ void xh_store8(uint8_t* row, int col, const float* y) {
    uint8_t* p = row + (col >> 5) * 128 + (col & 31) * 2;
    uint4 h, l;
    __half2* hh = reinterpret_cast<__half2*>(&h);
    __half2* ll = reinterpret_cast<__half2*>(&l);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half h0, l0, h1, l1;
        xh_split(y[2 * i], h0, l0);
        xh_split(y[2 * i + 1], h1, l1);
        hh[i] = __halves2half2(h0, h1);
        ll[i] = __halves2half2(l0, l1);
    }
    *reinterpret_cast<uint4*>(p) = h;
    *reinterpret_cast<uint4*>(p + 64) = l;
}

}  // namespace
