// Probe of the TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4) on sm_100a: one CTA gathers 256 rows of 128 bytes by
// index into a SWIZZLE_128B operand stage, `reps` times, and dumps the stage - correctness (swizzle, zero fill of negative
// row indices) and the issue / landing rate of 64 gather4 per stage.  Development aid for the sparse-convolution producer.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/eyoc_b200.h"

using namespace tcp;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}

__global__ void __launch_bounds__(128, 1)
gather4_probe_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx, int n_idx, int reps, uint4* __restrict__ out,
                     long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const uint32_t st = smem_u32(smem), b = smem_u32(&bar);
    if (threadIdx.x == 0) { mbar_init(b, 1); mbar_init_fence(); }
    __syncthreads();
    long long t0 = 0;
    for (int r = 0; r < reps; ++r) {
        if (threadIdx.x < 32) {
            if (r == 0 && threadIdx.x == 0) t0 = clock64();
            if (threadIdx.x == 0) mbar_expect_tx(b, 256 * 128);
            __syncwarp();
            const int* ix = idx + ((size_t)blockIdx.x * reps + r) % (n_idx / 256) * 256;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int q = threadIdx.x * 2 + g;                         // group of 4 rows: tile rows 4 q .. 4 q + 3
                const int4 v = *reinterpret_cast<const int4*>(ix + 4 * q);
                tma_gather4(st + q * 512, &map, b, 0, v.x, v.y, v.z, v.w);
            }
        }
        mbar_wait(b, r & 1);
        __syncthreads();
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < 256 * 8; i += 128) out[i] = reinterpret_cast<const uint4*>(smem)[i];
}

}  // namespace

// X [n_rows, 64] fp16 (128-byte rows); idx [n_idx] int32 (multiple of 256; -1 = no row); out 32 KB: the stage of the last
// repetition of CTA 0 as it lies in shared memory; cycles [ctas].  box_rows: the tensor map's box height to try (1 or 4).
extern "C" int eyoc_debug_gather4_probe(const void* X, int64_t n_rows, const int32_t* idx, int n_idx, int ctas, int reps, int box_rows,
                                        void* out, long long* cycles, cudaStream_t stream) {
    EYOC_CHECK_ARG(X && idx && out && cycles && n_idx >= 256 && n_idx % 256 == 0 && ctas >= 1 && reps >= 1, "eyoc_debug_gather4_probe: bad argument");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { eyoc_set_error("cuTensorMapEncodeTiled not available"); return EYOC_ERR_CUDA; }
    CUtensorMap map;
    const cuuint64_t dims[2] = {64, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { eyoc_set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return EYOC_ERR_CUDA; }
    EYOC_CUDA(cudaFuncSetAttribute(gather4_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 + 1024));
    gather4_probe_kernel<<<ctas, 128, 256 * 128 + 1024, stream>>>(map, idx, n_idx, reps, (uint4*)out, cycles);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}
