// Micro-benchmark: issue rate / latency of tcgen05.mma.kind::tf32 (M=128, K=8) from shared memory as a function of N,
// of whether consecutive MMAs accumulate into the same TMEM tile, and of how many MMAs sit between two commits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long spins = 0;
    while (!done) {
        if (++spins > (1ull << 24)) __trap();
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

// mode bit0: rotate the accumulator tile between consecutive MMAs; per_commit: MMAs between commits (and waits)
template <int N>
__global__ void bench(int iters, int per_commit, int rotate, int wait_each, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tbase;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) ((volatile float*)(raw + (base - smem_u32(raw))))[i] = 0.f;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase;
    if (threadIdx.x == 0) {
        const uint32_t a = base, b = base + 16384;
        const int ntile = 512 / N;
        uint32_t phase = 0;
        long long t0 = clock64();
        int cnt = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tb + (rotate ? (uint32_t)((i % ntile) * N) : 0u);
            umma_tf32(d, make_desc(a + (i & 3) * 32), make_desc(b + (i & 3) * 32), IDESC, 1u);
            if (++cnt == per_commit) {
                cnt = 0;
                if (wait_each) { umma_commit(smem_u32(&bar)); mbar_wait(smem_u32(&bar), phase); phase ^= 1u; }
                else umma_commit(smem_u32(&bar2));          // un-awaited commit (arrivals on a spare barrier)
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int N>
void run(int per_commit, int rotate, int wait_each, int grid) {
    long long* d;
    cudaMalloc(&d, grid * 8);
    const int iters = 4096;
    const size_t smem = 1024 + 16384 + N * 128;
    cudaFuncSetAttribute(bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<N><<<grid, 128, smem>>>(iters, per_commit, rotate, wait_each, d);
    bench<N><<<grid, 128, smem>>>(iters, per_commit, rotate, wait_each, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%3d per_commit=%4d rotate=%d wait_each=%d grid=%3d : %.1f cycles/MMA (floor %d)  %s\n", N, per_commit, rotate, wait_each,
           grid, (double)mx / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        for (int rot : {0, 1}) {
            run<32>(4096, rot, 0, grid);
            run<64>(4096, rot, 0, grid);
            run<128>(4096, rot, 0, grid);
            run<256>(4096, rot, 0, grid);
        }
        run<64>(12, 0, 0, grid);
        run<64>(12, 1, 0, grid);
        run<64>(12, 0, 1, grid);
        run<64>(1, 0, 1, grid);
        run<128>(12, 0, 1, grid);
        run<256>(12, 0, 1, grid);
        run<256>(1, 0, 1, grid);
    }
    return 0;
}
