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

// NW issuing warps (each takes 64 / NW of a stage's 64 gather4, one per lane), NS stages in flight
template <int NW, int NS>
__global__ void __launch_bounds__(NW * 32, 1)
gather4_probe_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx, int n_idx, int reps, uint4* __restrict__ out,
                     long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[NS];
    const uint32_t st0 = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) mbar_init(smem_u32(&bars[i]), NW);
        mbar_init_fence();
    }
    __syncthreads();
    const long long t0 = clock64();
    constexpr int PER = 64 / NW;                                   // gather4 per warp and stage
    for (int r = 0; r < reps; ++r) {
        const int s = r % NS;
        const uint32_t b = smem_u32(&bars[s]), st = st0 + s * 256 * 128;
        if (r >= NS) mbar_wait(b, ((r / NS) - 1) & 1);             // the stage's previous fill has landed (nothing consumes it here)
        if (lane == 0) mbar_expect_tx(b, PER * 512);
        __syncwarp();
        const int* ix = idx + ((size_t)blockIdx.x * reps + r) % (n_idx / 256) * 256;
        for (int g = lane; g < PER; g += 32) {
            const int q = warp * PER + g;
            const int4 v = *reinterpret_cast<const int4*>(ix + 4 * q);
            tma_gather4(st + q * 512, &map, b, 0, v.x, v.y, v.z, v.w);
        }
    }
    for (int r = reps > NS ? reps - NS : 0; r < reps; ++r) mbar_wait(smem_u32(&bars[r % NS]), (r / NS) & 1);
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    if (blockIdx.x == 0) {
        const uint4* last = reinterpret_cast<const uint4*>(smem + ((reps - 1) % NS) * 256 * 128);
        for (int i = threadIdx.x; i < 256 * 8; i += NW * 32) out[i] = last[i];
    }
}

template <int NW, int NS>
int launch_probe(const CUtensorMap& map, const int* idx, int n_idx, int ctas, int reps, void* out, long long* cycles, cudaStream_t stream) {
    const int smem = NS * 256 * 128 + 1024;
    EYOC_CUDA(cudaFuncSetAttribute(gather4_probe_kernel<NW, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gather4_probe_kernel<NW, NS><<<ctas, NW * 32, smem, stream>>>(map, idx, n_idx, reps, (uint4*)out, cycles);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

}  // namespace

// X [n_rows, 64] fp16 (128-byte rows); idx [n_idx] int32 (multiple of 256; -1 = no row); out 32 KB: the stage of the last
// repetition of CTA 0 as it lies in shared memory; cycles [ctas].  `config`: issuing warps / stages in flight (see the switch below).
extern "C" int eyoc_debug_gather4_probe(const void* X, int64_t n_rows, const int32_t* idx, int n_idx, int ctas, int reps, int box_rows,
                                        void* out, long long* cycles, cudaStream_t stream) {
    EYOC_CHECK_ARG(X && idx && out && cycles && n_idx >= 256 && n_idx % 256 == 0 && ctas >= 1 && reps >= 1, "eyoc_debug_gather4_probe: bad argument");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { eyoc_set_error("cuTensorMapEncodeTiled not available"); return EYOC_ERR_CUDA; }
    CUtensorMap map;
    const cuuint64_t dims[2] = {64, (cuuint64_t)n_rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 1};        // ONE row: the instruction fetches four of them (a 4-row box is an illegal instruction)
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { eyoc_set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return EYOC_ERR_CUDA; }
    switch (box_rows) {        // (second use of the argument: the issue configuration)
        case 1: return launch_probe<1, 1>(map, idx, n_idx, ctas, reps, out, cycles, stream);
        case 2: return launch_probe<2, 6>(map, idx, n_idx, ctas, reps, out, cycles, stream);
        case 4: return launch_probe<4, 6>(map, idx, n_idx, ctas, reps, out, cycles, stream);
        case 8: return launch_probe<8, 6>(map, idx, n_idx, ctas, reps, out, cycles, stream);
        case 16: return launch_probe<16, 6>(map, idx, n_idx, ctas, reps, out, cycles, stream);
        case 116: return launch_probe<16, 1>(map, idx, n_idx, ctas, reps, out, cycles, stream);
        case 102: return launch_probe<1, 6>(map, idx, n_idx, ctas, reps, out, cycles, stream);
        default: eyoc_set_error("eyoc_debug_gather4_probe: unknown configuration %d", box_rows); return EYOC_ERR_ARG;
    }
}
