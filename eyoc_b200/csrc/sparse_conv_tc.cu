// Sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same contract and epilogue as csrc/sparse_conv.cu (replaces ME.MinkowskiConvolution[Transpose] + BN + ReLU +
// residual + ME.cat + L2 norm of model/resunet.py:142-193), different data path:
//   * one CTA owns up to 8 accumulator tiles of 128 output rows x N = C_out channels, all resident in TMEM
//     (512 columns x 128 lanes x fp32), so a W[k] slab staged in shared memory is reused by up to 1024 rows;
//   * producer groups of 4 warps gather the input rows of one (tile, kernel offset, 32-channel chunk) straight into
//     the SWIZZLE_128B K-major UMMA layout (8 lanes cover one 128-byte row chunk: coalesced loads, conflict-free
//     st.shared.v4); the gather of a group's NEXT item is in flight while it splits and stores the current one;
//   * weight slabs are pre-split and pre-swizzled once (eyoc_conv_split_weights) so that one TMA bulk copy
//     (cp.async.bulk) per (offset, chunk) drops the exact shared-memory image, completion on an mbarrier;
//   * 1 thread issues tcgen05.mma.kind::tf32; fp32-level accuracy comes from the 3-term split
//     x = hi + lo (both rounded to tf32):  A_hi W_hi + A_hi W_lo + A_lo W_hi  accumulated in fp32 in TMEM;
//   * smem ring (A) / ring (W) hand-shaken with mbarriers, slots released by tcgen05.commit;
//   * the producer warps become the epilogue: tcgen05.ld 32 lanes x 16 columns, fused affine / residual / ReLU /
//     L2 norm, one output row per thread.
// Kernel offsets (and whole W slabs) with no neighbour in a tile are skipped by both sides from a shared bit mask.
#include "common.cuh"
#include "../../include/eyoc_b200.h"

namespace {

constexpr int UM = 128;        // rows per accumulator tile (UMMA M)
constexpr int KC = 32;         // channels per chunk = one 128-byte swizzle-atom row of fp32
constexpr int MAXACC = 8;
constexpr int A_BYTES = UM * KC * 4;   // 16 KB per (hi | lo) tile

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned long long spins = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1ull << 26)) __trap();      // watchdog: never hang the device
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D), completion counted on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// x = hi + lo with hi = x rounded to tf32 (round-to-nearest, ties away = cvt.rna.tf32.f32 for finite x: add half an
// ulp of the 10-bit mantissa to the magnitude, clear the 13 low bits), lo = tf32(x - hi) the same way.  Two integer
// ops per rounding instead of the four-instruction sequence ptxas emits for cvt.rna (Inf/NaN guard).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);          // start address (16-byte units), bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

struct TcArgs {
    const float* in0; int c0;
    const float* in1; int c1;
    const int32_t* nbr;       // [K, n_out]; column = output row, or tile position when nbr_tiled
    const int32_t* row_perm;  // [n_out] tile position -> output row, or null
    const float* wt_img;      // [K][cin/32][hi|lo][cout][32] swizzled shared-memory images (eyoc_conv_split_weights)
    const float* scale;
    const float* shift;
    const float* residual;
    float* out;
    int K, n_out, relu, l2norm, nacc, nbr_tiled;
};

constexpr int MAX_ITEMS = 27 * 12 * MAXACC;      // (kernel offset, 32-channel chunk, tile) work items per CTA

// Work item: bits [0,5) kernel offset, [5,9) chunk index, [9,12) tile, bit 12 = first item of its (offset, chunk),
// i.e. the MMA side must switch to the next weight slab.
__device__ __forceinline__ int item_k(uint32_t it) { return it & 31; }
__device__ __forceinline__ int item_c(uint32_t it) { return (it >> 5) & 15; }
__device__ __forceinline__ int item_t(uint32_t it) { return (it >> 9) & 7; }
__device__ __forceinline__ bool item_first(uint32_t it) { return (it >> 12) & 1; }

// Thread map: NPG producer groups of 128 threads (group g owns A stage g), one weight-loader warp (TMA bulk copies,
// one elected thread), one MMA-issuer warp (one elected thread).
template <int N, int NPG, int NSW>
__global__ void __launch_bounds__(NPG * 128 + 64, 1)
sparse_conv_tc_kernel(TcArgs a) {
    constexpr int W_BYTES = N * KC * 4;
    constexpr int NPT = NPG * 128;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);
    constexpr int TMEM_COLS = (MAXACC * N > 512) ? 512 : MAXACC * N;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem0;                                   // NPG x (hi | lo)
    const uint32_t sW = sA + NPG * 2 * A_BYTES;                  // NSW x (hi | lo)
    __shared__ uint64_t bars[2 * NPG + 2 * NSW + 1];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t valid[MAXACC];
    __shared__ uint32_t tmask[32];
    __shared__ int pair_off[27 * 12 + 1];
    __shared__ uint16_t items[MAX_ITEMS];
    __shared__ int nitems_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cin = a.c0 + a.c1;
    const int nch = cin / KC;
    const int nacc = a.nacc;
    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[NPG]);
    const uint32_t w_full = smem_u32(&bars[2 * NPG]), w_empty = smem_u32(&bars[2 * NPG + NSW]);
    const uint32_t done_bar = smem_u32(&bars[2 * NPG + 2 * NSW]);
    const int wload_warp = NPT / 32, mma_warp = NPT / 32 + 1;

    if (tid == 0) {
        for (int i = 0; i < NPG; ++i) { mbar_init(a_full + 8 * i, 128); mbar_init(a_empty + 8 * i, 1); }
        for (int i = 0; i < NSW; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < MAXACC) valid[tid] = 0;
    if (warp == mma_warp) tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    // ---- which kernel offsets have a neighbour in each tile
    const int tile0 = blockIdx.x * nacc;
    if (tid < NPT) {
        const int r = tid & 127;
        for (int t = tid >> 7; t < nacc; t += NPG) {
            const int rr = (tile0 + t) * UM + r;
            int col = -1;
            if (rr < a.n_out) col = (a.row_perm && !a.nbr_tiled) ? a.row_perm[rr] : rr;
            uint32_t m = 0;
            if (a.nbr == nullptr) {
                m = col >= 0 ? 1u : 0u;
            } else if (col >= 0) {
                for (int k = 0; k < a.K; ++k) m |= (uint32_t)(__ldg(a.nbr + (size_t)k * a.n_out + col) >= 0) << k;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
            if (lane == 0 && m) atomicOr(&valid[t], m);
        }
    }
    __syncthreads();
    // ---- the CTA's ordered work list: (offset, chunk) outer, tiles inner
    if (tid < 32) {
        uint32_t m = 0;
        if (tid < a.K)
            for (int t = 0; t < nacc; ++t) m |= ((valid[t] >> tid) & 1u) << t;
        tmask[tid] = m;
    }
    __syncthreads();
    if (warp == 0) {
        const int npairs = a.K * nch;
        int carry = 0;
        for (int p0 = 0; p0 < npairs; p0 += 32) {
            const int p = p0 + lane;
            const int c = p < npairs ? __popc(tmask[p / nch]) : 0;
            int x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (p < npairs) pair_off[p] = carry + x - c;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) nitems_s = carry;
    }
    __syncthreads();
    for (int p = tid; p < a.K * nch; p += blockDim.x) {
        const int k = p / nch, ci = p % nch;
        uint32_t m = tmask[k];
        int o = pair_off[p];
        bool first = true;
        while (m) {
            const int t = __ffs(m) - 1;
            m &= m - 1;
            items[o++] = (uint16_t)(k | (ci << 5) | (t << 9) | ((first ? 1 : 0) << 12));
            first = false;
        }
    }
    __syncthreads();
    const int nitems = nitems_s;

    if (tid < NPT) {
        // =========================================================== producers: gather + tf32 split -> smem
        const int g = tid >> 7, w4 = warp & 3;
        const int c = lane & 7, rsub = lane >> 3;
        const uint32_t dh = sA + g * 2 * A_BYTES;
        // row r = w4*32 + q*4 + rsub, 16-byte chunk c of its 128-byte line lands at r*128 + ((c ^ (r & 7)) << 4)
        const uint32_t st0 = dh + (uint32_t)(w4 * 32 + rsub) * 128u + (uint32_t)((c ^ rsub) << 4);        // q even
        const uint32_t st1 = dh + (uint32_t)(w4 * 32 + rsub) * 128u + (uint32_t)((c ^ rsub ^ 4) << 4);    // q odd
        // lane l of the warp fetches the neighbour index of row w4*32 + l (coalesced); rows are handed out by shuffle
        auto load_idx = [&](uint32_t it) -> int {
            const int rr = (tile0 + item_t(it)) * UM + w4 * 32 + lane;
            int v = -1;
            if (rr < a.n_out) {
                const int col = (a.row_perm && !a.nbr_tiled) ? __ldg(a.row_perm + rr) : rr;
                v = a.nbr ? __ldg(a.nbr + (size_t)item_k(it) * a.n_out + col) : (a.row_perm ? __ldg(a.row_perm + rr) : rr);
            }
            return v;
        };
        auto gather = [&](float4 (&x)[8], uint32_t it, int idx) {
            const int cc = item_c(it) * KC;
            const float* src = cc < a.c0 ? a.in0 : a.in1;
            const int cs = cc < a.c0 ? a.c0 : a.c1;
            const int co = (cc < a.c0 ? cc : cc - a.c0) + c * 4;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int v = __shfl_sync(0xffffffffu, idx, q * 4 + rsub);
                x[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v >= 0) x[q] = __ldg(reinterpret_cast<const float4*>(src + (size_t)v * cs + co));
            }
        };
        auto split_store = [&](const float4 (&x)[8]) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                split_tf32(x[q].x, h0, l0); split_tf32(x[q].y, h1, l1);
                split_tf32(x[q].z, h2, l2); split_tf32(x[q].w, h3, l3);
                const uint32_t addr = ((q & 1) ? st1 : st0) + (uint32_t)q * 512u;
                sts128(addr, h0, h1, h2, h3);
                sts128(addr + A_BYTES, l0, l1, l2, l3);
            }
        };
        // software pipeline per group: indices two items ahead, gathered rows one item ahead
        float4 xa[8], xb[8];
        int i = g;
        int idx_a = i < nitems ? load_idx(items[i]) : -1;
        int idx_b = i + NPG < nitems ? load_idx(items[i + NPG]) : -1;
        if (i < nitems) gather(xa, items[i], idx_a);
        uint32_t n = 0;
        while (i < nitems) {
            // ---- even step: xa is current, xb receives the next item
            if (i + NPG < nitems) gather(xb, items[i + NPG], idx_b);
            idx_a = i + 2 * NPG < nitems ? load_idx(items[i + 2 * NPG]) : -1;
            mbar_wait(a_empty + 8 * g, (n & 1u) ^ 1u);
            split_store(xa);
            fence_proxy_async();
            mbar_arrive(a_full + 8 * g);
            i += NPG; ++n;
            if (i >= nitems) break;
            // ---- odd step: xb is current, xa receives the next item
            if (i + NPG < nitems) gather(xa, items[i + NPG], idx_a);
            idx_b = i + 2 * NPG < nitems ? load_idx(items[i + 2 * NPG]) : -1;
            mbar_wait(a_empty + 8 * g, (n & 1u) ^ 1u);
            split_store(xb);
            fence_proxy_async();
            mbar_arrive(a_full + 8 * g);
            i += NPG; ++n;
        }
        // =========================================================== epilogue: TMEM -> registers -> global
        mbar_wait(done_bar, 0);
        tc_fence_after();
        const int rloc = w4 * 32 + lane;
        for (int t = g; t < nacc; t += NPG) {
            const int rr = (tile0 + t) * UM + rloc;
            const int row = rr < a.n_out ? (a.row_perm ? a.row_perm[rr] : rr) : -1;
            const bool started = valid[t] != 0;
            const uint32_t taddr = tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(t * N);
            float ss = 0.f;
            float keep[N == 32 ? 32 : 1];
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t v[16];
                if (started) tmem_ld16(taddr + c0, v);       // warp-uniform branch (sync.aligned)
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0u;
                }
                float y[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    y[j] = __uint_as_float(v[j]);
                    const int col = c0 + j;
                    if (a.scale) y[j] = __fmaf_rn(y[j], __ldg(a.scale + col), a.shift ? __ldg(a.shift + col) : 0.f);
                    else if (a.shift) y[j] += __ldg(a.shift + col);
                }
                if (a.residual && row >= 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(a.residual + (size_t)row * N + c0 + j));
                        y[j] += q.x; y[j + 1] += q.y; y[j + 2] += q.z; y[j + 3] += q.w;
                    }
                }
                if (a.relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
                }
                if (N == 32 && a.l2norm) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) { keep[(N == 32 ? c0 : 0) + (N == 32 ? j : 0)] = y[j]; ss = __fmaf_rn(y[j], y[j], ss); }
                } else if (row >= 0) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(a.out + (size_t)row * N + c0 + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
                }
            }
            if (N == 32 && a.l2norm && row >= 0) {
                const float nrm = sqrtf(ss);
#pragma unroll
                for (int j = 0; j < (N == 32 ? 32 : 0); j += 4)
                    *reinterpret_cast<float4*>(a.out + (size_t)row * N + j) =
                        make_float4(__fdiv_rn(keep[j], nrm), __fdiv_rn(keep[j + 1], nrm), __fdiv_rn(keep[j + 2], nrm), __fdiv_rn(keep[j + 3], nrm));
            }
        }
        tc_fence_before();
    } else if (warp == wload_warp) {
        // =========================================================== weight slabs: one TMA bulk copy each
        if (lane == 0) {
            uint32_t w_it = 0;
            for (int i = 0; i < nitems; ++i) {
                const uint32_t it = items[i];
                if (!item_first(it)) continue;
                const uint32_t ws = w_it % NSW;
                mbar_wait(w_empty + 8 * ws, ((w_it / NSW) & 1u) ^ 1u);
                mbar_expect_tx(w_full + 8 * ws, 2 * W_BYTES);
                bulk_g2s(sW + ws * 2 * W_BYTES, a.wt_img + ((size_t)item_k(it) * nch + item_c(it)) * (2 * N * KC), 2 * W_BYTES,
                         w_full + 8 * ws);
                ++w_it;
            }
        }
        __syncwarp();
    } else {
        // =========================================================== MMA issuer (one thread)
        uint32_t w_it = 0, started = 0;
        uint32_t wh = 0, wl = 0;
        for (int i = 0; i < nitems; ++i) {
            const uint32_t it = items[i];
            if (item_first(it)) {
                if (w_it > 0 && lane == 0) umma_commit(w_empty + 8 * ((w_it - 1) % NSW));   // previous slab fully consumed
                const uint32_t ws = w_it % NSW;
                mbar_wait(w_full + 8 * ws, (w_it / NSW) & 1u);
                wh = sW + ws * 2 * W_BYTES;
                wl = wh + W_BYTES;
                ++w_it;
            }
            const int g = i % NPG, t = item_t(it);
            mbar_wait(a_full + 8 * g, (uint32_t)(i / NPG) & 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t ah = sA + g * 2 * A_BYTES, al = ah + A_BYTES;
                const uint32_t d = tmem_base + (uint32_t)(t * N);
                uint32_t acc = (started >> t) & 1u;
#pragma unroll
                for (int j = 0; j < KC / 8; ++j) {
                    const uint64_t dah = make_desc(ah + j * 32), dal = make_desc(al + j * 32);
                    const uint64_t dwh = make_desc(wh + j * 32), dwl = make_desc(wl + j * 32);
                    umma_tf32(d, dah, dwh, IDESC, acc);
                    umma_tf32(d, dah, dwl, IDESC, 1u);
                    umma_tf32(d, dal, dwh, IDESC, 1u);
                    acc = 1u;
                }
                umma_commit(a_empty + 8 * g);
            }
            started |= 1u << t;
            __syncwarp();
        }
        if (lane == 0) umma_commit(done_bar);
        __syncwarp();
    }
    __syncthreads();
    if (warp == mma_warp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// weight [K, cin, cout] -> shared-memory images [K][cin/32][hi|lo][cout rows][32 channels], each row's eight 16-byte
// chunks XOR-swizzled with (row & 7) exactly as the SWIZZLE_128B K-major UMMA descriptor expects them.
__global__ void split_weights_kernel(const float* __restrict__ w, int K, int cin, int cout, float* __restrict__ img) {
    const size_t total = (size_t)K * cin * cout;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % cin);
    const int n = (int)((i / cin) % cout);
    const int k = (int)(i / ((size_t)cin * cout));
    const float x = w[((size_t)k * cin + c) * cout + n];
    const float h = __uint_as_float(to_tf32(x));
    const float l = __uint_as_float(to_tf32(x - h));
    const int ci = c / KC, cl = c % KC;
    const size_t slab = ((size_t)k * (cin / KC) + ci) * (2 * (size_t)cout * KC);
    const int off = n * KC + ((((cl >> 2) ^ (n & 7)) << 2) | (cl & 3));
    img[slab + off] = h;
    img[slab + (size_t)cout * KC + off] = l;
}

template <int N, int NPG, int NSW>
int launch_tc(const TcArgs& a, cudaStream_t stream) {
    const size_t smem = 1024 + (size_t)NPG * 2 * A_BYTES + (size_t)NSW * 2 * N * KC * 4;
    EYOC_CUDA(cudaFuncSetAttribute(sparse_conv_tc_kernel<N, NPG, NSW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (a.n_out + UM - 1) / UM;
    const int grid = (tiles + a.nacc - 1) / a.nacc;
    sparse_conv_tc_kernel<N, NPG, NSW><<<grid, NPG * 128 + 64, smem, stream>>>(a);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

}  // namespace

extern "C" int eyoc_conv_split_weights(const float* weight, int K, int cin, int cout, float* wt_img, cudaStream_t stream) {
    EYOC_CHECK_ARG(weight && wt_img && K >= 1 && cin >= 1 && cout >= 1, "eyoc_conv_split_weights: bad argument");
    EYOC_CHECK_ARG(cin % KC == 0, "eyoc_conv_split_weights: cin must be a multiple of 32");
    const size_t total = (size_t)K * cin * cout;
    split_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(weight, K, cin, cout, wt_img);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_sparse_conv_tc_supported(int c0, int c1, int cout, int K, int l2norm) {
    const int cin = c0 + c1;
    if (cin % KC || c0 % KC || K < 1 || K > 27 || cin / KC > 12) return 0;
    if (!(cout == 32 || cout == 64 || cout == 128 || cout == 256)) return 0;
    if (l2norm && cout != 32) return 0;
    return 1;
}

extern "C" int eyoc_sparse_conv_tc(const float* in0, int c0, const float* in1, int c1, const int32_t* nbr, int K, int64_t n_out,
                                   const int32_t* row_perm, int nbr_tiled, const float* wt_img, const float* scale,
                                   const float* shift, const float* residual, int relu, int l2norm, float* out, int cout,
                                   cudaStream_t stream) {
    EYOC_CHECK_ARG(in0 && wt_img && out, "eyoc_sparse_conv_tc: null argument");
    EYOC_CHECK_ARG((in1 != nullptr) == (c1 > 0), "eyoc_sparse_conv_tc: in1 and c1 must be given together");
    EYOC_CHECK_ARG(nbr || K == 1, "eyoc_sparse_conv_tc: a neighbour table is required when K > 1");
    EYOC_CHECK_ARG(eyoc_sparse_conv_tc_supported(c0, c1, cout, K, l2norm), "eyoc_sparse_conv_tc: unsupported shape c0=%d c1=%d cout=%d K=%d", c0, c1, cout, K);
    EYOC_CHECK_ARG(n_out >= 0 && n_out < (1ll << 31), "eyoc_sparse_conv_tc: bad n_out");
    EYOC_CHECK_ARG(!nbr_tiled || row_perm, "eyoc_sparse_conv_tc: a tiled neighbour table needs row_perm");
    if (n_out == 0) return EYOC_OK;
    const int maxacc = cout <= 64 ? 8 : (cout == 128 ? 4 : 2);
    const int64_t tiles = (n_out + UM - 1) / UM;
    int nacc = (int)(tiles / (2 * 148));               // keep >= 2 CTAs per SM's worth of work before widening
    nacc = nacc < 1 ? 1 : (nacc > maxacc ? maxacc : nacc);
    TcArgs a{in0, c0, in1, c1, nbr, row_perm, wt_img, scale, shift, residual, out, K, (int)n_out, relu, l2norm, nacc, nbr_tiled};
    switch (cout) {
        case 32: return launch_tc<32, 4, 4>(a, stream);
        case 64: return launch_tc<64, 4, 4>(a, stream);
        case 128: return launch_tc<128, 4, 2>(a, stream);
        default: return launch_tc<256, 2, 2>(a, stream);
    }
}
