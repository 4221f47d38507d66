// Sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Same contract and epilogue as csrc/sparse_conv.cu (replaces ME.MinkowskiConvolution[Transpose] + BN + ReLU +
// residual + ME.cat + L2 norm of model/resunet.py:142-193), different data path.
//
// Measured on B200 (tools/ubench/mma_rate.cu): one tcgen05.mma kind::tf32 costs ~142 cycles whatever its N (32..256)
// at M = 128, K = 8 - the instruction count, not the FLOP count, bounds this kernel.  So the GEMM is issued
// TRANSPOSED: the gathered input rows are the N side (256 rows per instruction), the weights the M side:
//     D[channel, row] += W[k][channel, c] * X[row, c]
//   * C_out <= 64: the tf32 hi and lo parts of W are stacked along M (lanes 0..63 = W_hi, 64..127 = W_lo), so TWO
//     instructions per 8 input channels per 256 rows give the full 4-term product  (W_hi + W_lo)(X_hi + X_lo);
//   * C_out = 128 (and each 128-channel half of C_out = 256): three instructions W_hi X_hi + W_lo X_hi + W_hi X_lo.
//   Accumulators: 2 tiles x 256 fp32 columns of TMEM per CTA.
//   * producer groups of 4 warps gather the rows of one (tile, kernel offset, 32-channel chunk) straight into the
//     SWIZZLE_128B K-major UMMA layout (8 lanes cover one 128-byte row chunk: coalesced loads, conflict-free
//     st.shared.v4), splitting x = hi + lo with integer rounding; the gather of a group's NEXT item is in flight
//     while it splits and stores the current one;
//   * weight slabs are pre-split and pre-swizzled once (eyoc_conv_split_weights) so that one TMA bulk copy
//     (cp.async.bulk) per (offset, chunk) drops the exact shared-memory image, completion on an mbarrier;
//   * smem rings hand-shaken with mbarriers, slots released by tcgen05.commit; one thread issues the MMAs;
//   * epilogue: tcgen05.ld (lane = channel) -> transpose through shared memory (hi + lo lanes summed) -> fused
//     affine / residual / ReLU / L2 norm -> fully coalesced row-major stores.
// Kernel offsets (and whole W slabs) with no neighbour in a tile are skipped by both sides from a shared bit mask;
// the tile order of csrc/coordmap.cu (eyoc_tile_order) makes the surviving (tile, offset) items dense.
#include "common.cuh"
#include "../../include/eyoc_b200.h"

namespace {

constexpr int TR = 256;        // output rows per accumulator tile (UMMA N)
constexpr int KC = 32;         // channels per chunk = one 128-byte swizzle-atom row of fp32
constexpr int NTILE = 2;       // accumulator tiles per CTA: 2 x 256 TMEM columns
constexpr int NPG = 4;         // producer warpgroups (16 warps x 16 rows = one 256-row item)
constexpr int XH_BYTES = TR * KC * 4;          // 32 KB: 256 rows x 128 B (hi); lo follows
constexpr int XS_BYTES = 2 * XH_BYTES;         // 64 KB per stage
constexpr int NPT = NPG * 128;

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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {      // non-blocking
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
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
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// x = hi + lo with hi = x rounded to tf32 (round-to-nearest, ties away = cvt.rna.tf32.f32 for finite x: add half an
// ulp of the 10-bit mantissa to the magnitude, clear the 13 low bits), lo = x - hi rounded the same way (its low bits
// are left in place: tf32 operands are read from the upper 19 bits).  Integer ops instead of the four-instruction
// sequence ptxas emits per cvt.rna (Inf/NaN guard).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
#ifdef EYOC_TRUNC_TEST
    hi = __float_as_uint(x);                                       // raw fp32: the tensor core truncates to tf32
    lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u)) + 0x1000u;
#else
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;      // the tensor core ignores the 13 low bits: no mask
#endif
}
// x_lo of the operand split x = trunc_tf32(x) + x_lo: exact difference, rounded to tf32 (half an ulp added to the
// magnitude; the 13 low bits are left in place - the tensor core does not read them)
__device__ __forceinline__ uint32_t lo_tf32(float x) {
    return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u)) + 0x1000u;
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
    const float* wt_img;      // per (k, chunk, part): [W_hi rows | W_lo rows] x 32 channels, swizzled smem image
    const float* scale;
    const float* shift;
    const float* residual;
    float* out;
    int K, n_out, cout, relu, l2norm, nbr_tiled;
};

// Debug / measurement only (tools/conv_ablate.py): bit 0 = skip the MMAs, bit 1 = skip the gather copies and the x_lo
// pass, bit 2 = skip the weight-slab copies.  Results are garbage when non-zero.
__device__ int g_ablate = 0;
__device__ long long g_times[1024][6];      // bit 3 of g_ablate: per-CTA phase timestamps of the first 1024 CTAs

constexpr int MAX_ITEMS = 27 * 12 * NTILE;      // (kernel offset, 32-channel chunk, tile) work items per CTA

// Work item: bits [0,5) kernel offset, [5,9) chunk index, bit 9 tile, bit 12 = first item of its (offset, chunk),
// i.e. the MMA side must switch to the next weight slab.
__device__ __forceinline__ int item_k(uint32_t it) { return it & 31; }
__device__ __forceinline__ int item_c(uint32_t it) { return (it >> 5) & 15; }
__device__ __forceinline__ int item_t(uint32_t it) { return (it >> 9) & 1; }
__device__ __forceinline__ bool item_first(uint32_t it) { return (it >> 12) & 1; }

__device__ __forceinline__ void bar_sync_producers() { asm volatile("bar.sync 1, %0;" ::"n"(NPT) : "memory"); }

// WIDE = false: C_out <= 64, W_hi / W_lo stacked along M (64 lanes each).  WIDE = true: 128 output channels per CTA
// (blockIdx.y selects the half when C_out = 256), W_hi and W_lo are separate 128-row operands.
// Thread map: 4 producer groups of 128 threads, one weight-loader warp (TMA bulk copies, one elected thread), one
// MMA-issuer warp (one elected thread).
template <bool WIDE, int NSW, int NXS>
__global__ void __launch_bounds__(NPT + 64, 1)
sparse_conv_tc_kernel(TcArgs a) {
    constexpr int W_BYTES = (WIDE ? 256 : 128) * KC * 4;      // slab: 32 KB (hi 128 rows | lo 128 rows) or 16 KB (64 | 64)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_off = ((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw);      // 0: declared aligned
    const uint32_t sX = smem_u32(smem_raw) + smem_off;           // NXS x (hi | lo); reused by the epilogue transpose
    const uint32_t sW = sX + NXS * XS_BYTES;                     // NSW slabs
    float* const sOut = reinterpret_cast<float*>(smem_raw + smem_off);
    __shared__ uint64_t bars[2 * NXS + 2 * NSW + 1];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t valid[NTILE];
    __shared__ uint32_t tmask[32];
    __shared__ uint16_t pair_off[27 * 12 + 1];
    __shared__ uint16_t items[MAX_ITEMS];
    __shared__ int nitems_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t_start = clock64();
    const int cin = a.c0 + a.c1;
    const int nch = cin / KC;
    const int nparts = WIDE ? a.cout / 128 : 1;
    const int part = WIDE ? blockIdx.y : 0;
    const int cpart = WIDE ? 128 : a.cout;                       // output channels this CTA produces
    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[NXS]);
    const uint32_t w_full = smem_u32(&bars[2 * NXS]), w_empty = smem_u32(&bars[2 * NXS + NSW]);
    const uint32_t done_bar = smem_u32(&bars[2 * NXS + 2 * NSW]);
    const int wload_warp = NPT / 32, mma_warp = NPT / 32 + 1;

    if (tid == 0) {
        for (int i = 0; i < NXS; ++i) { mbar_init(a_full + 8 * i, NPT); mbar_init(a_empty + 8 * i, 1); }
        for (int i = 0; i < NSW; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
        mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < NTILE) valid[tid] = 0;
    if (warp == mma_warp) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    // ---- which kernel offsets have a neighbour in each tile
    const int tile0 = blockIdx.x * NTILE;
    if (tid < NPT) {
        const int t = tid >> 8, r = tid & 255;
        const int rr = (tile0 + t) * TR + r;
        int col = -1;
        if (rr < a.n_out) col = (a.row_perm && !a.nbr_tiled) ? a.row_perm[rr] : rr;
        uint32_t m = 0;
        if (a.nbr == nullptr) {
            m = col >= 0 ? 1u : 0u;
        } else if (col >= 0) {
            int v[27];
#pragma unroll
            for (int k = 0; k < 27; ++k) v[k] = k < a.K ? __ldg(a.nbr + (size_t)k * a.n_out + col) : -1;      // 27 loads in flight
#pragma unroll
            for (int k = 0; k < 27; ++k) m |= (uint32_t)(v[k] >= 0) << k;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
        if (lane == 0 && m) atomicOr(&valid[t], m);
    }
    __syncthreads();
    // ---- the CTA's ordered work list: (offset, chunk) outer, tiles inner
    if (tid < 32) {
        uint32_t m = 0;
        if (tid < a.K)
            for (int t = 0; t < NTILE; ++t) m |= ((valid[t] >> tid) & 1u) << t;
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
            if (p < npairs) pair_off[p] = (uint16_t)(carry + x - c);
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
    const int ablate = g_ablate;
    const bool timing = (ablate & 8) && tid == 0 && blockIdx.x < 1024 && blockIdx.y == 0;
    if (timing) { g_times[blockIdx.x][0] = t_start; g_times[blockIdx.x][1] = clock64(); g_times[blockIdx.x][5] = nitems; }

    if (tid < NPT) {
        // =========================================================== producers: async gather -> smem, then x_lo
        // The gathered fp32 rows are copied global -> shared by cp.async straight into the SWIZZLE_128B K-major UMMA
        // layout and serve AS THEY ARE as the tf32 "hi" operand: the tensor core reads the upper 19 bits of each
        // element, i.e. x_hi = trunc_tf32(x) (checked on B200: results with raw and with pre-truncated operands are
        // identical).  Each thread then reads back the four 16-byte chunks it copied, forms x_lo = x - trunc_tf32(x)
        // (exact; rounded to tf32 by adding half an ulp) and stores it to the "lo" half of the stage.  No gathered
        // value ever waits in a register, so the copies of the next two items fly while this one is converted.
        // Every item (256 rows x 32 channels) is shared by all 16 producer warps: warp w owns tile rows [16w, 16w+16);
        // 8 lanes cover one 128-byte row chunk.
        const int c = lane & 7, rsub = lane >> 3;
        const int rbase = warp * 16;
        // row r = rbase + q*4 + rsub, 16-byte chunk c of its 128-byte line lands at r*128 + ((c ^ (r & 7)) << 4)
        const uint32_t st0 = sX + (uint32_t)(rbase + rsub) * 128u + (uint32_t)((c ^ rsub) << 4);        // q even
        const uint32_t st1 = sX + (uint32_t)(rbase + rsub) * 128u + (uint32_t)((c ^ rsub ^ 4) << 4);    // q odd
        // lanes 0..15 fetch the neighbour indices of the warp's 16 rows (coalesced); rows are handed out by shuffle.
        // Per tile: the table column / input row this lane is responsible for (-1 beyond the last row).
        int colv[NTILE];
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
            const int rr = (tile0 + t) * TR + rbase + (lane & 15);
            colv[t] = -1;
            if (rr < a.n_out) colv[t] = (a.row_perm && !(a.nbr && a.nbr_tiled)) ? __ldg(a.row_perm + rr) : rr;
        }
        const bool identity = a.nbr == nullptr;
        const uint32_t cs0 = (uint32_t)a.c0 * 4u, cs1 = (uint32_t)a.c1 * 4u;
        auto load_idx = [&](uint32_t it) -> int {
            const int col = item_t(it) ? colv[1] : colv[0];
            if (identity || col < 0) return col;
            return __ldg(a.nbr + (size_t)item_k(it) * a.n_out + col);
        };
        // rows without a neighbour are zero-filled by the copy itself (source size 0)
        auto copy_rows = [&](uint32_t it, int idx, uint32_t stage_off) {
            const int cc = item_c(it) * KC;
            const bool first = cc < a.c0;
            const char* src = reinterpret_cast<const char*>(first ? a.in0 : a.in1) + ((first ? cc : cc - a.c0) + c * 4) * 4;
            const uint32_t cs = first ? cs0 : cs1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int v = __shfl_sync(0xffffffffu, idx, q * 4 + rsub);
                const char* ptr = src + (uint64_t)(uint32_t)max(v, 0) * cs;
                const uint32_t dst = ((q & 1) ? st1 : st0) + stage_off + (uint32_t)q * 512u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ptr), "r"(v >= 0 ? 16u : 0u) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto make_lo = [&](uint32_t stage_off) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t addr = ((q & 1) ? st1 : st0) + stage_off + (uint32_t)q * 512u;
                float x0, x1, x2, x3;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(addr) : "memory");
                sts128(addr + XH_BYTES, lo_tf32(x0), lo_tf32(x1), lo_tf32(x2), lo_tf32(x3));
            }
        };
        // Pipeline: copies run NXS-1 items ahead of the x_lo pass (a stage is reusable once the MMAs that read it
        // have completed).  Per item i: wait for its copies, x_lo pass, hand it to the MMA warp; THEN prefetch the
        // neighbour indices of item j+1 (right behind the proxy fence: fence.proxy.async drains this thread's
        // outstanding loads, so a load issued just before it would stall the hand-over for a full L2 round trip)
        // and wait for the stage of item j = i+NXS-1 / issue its copies.
        int idx_j = -1;                                           // indices of the next item to copy
        {
#pragma unroll
            for (int j = 0; j < NXS - 1; ++j) {
                idx_j = j < nitems ? load_idx(items[j]) : -1;
                if (j < nitems && !(ablate & 2)) copy_rows(items[j], idx_j, (uint32_t)j * XS_BYTES);
                else asm volatile("cp.async.commit_group;" ::: "memory");
            }
            idx_j = NXS - 1 < nitems ? load_idx(items[NXS - 1]) : -1;
        }
        for (int i = 0; i < nitems; ++i) {
            const uint32_t si = (uint32_t)(i % NXS);
            asm volatile("cp.async.wait_group %0;" ::"n"(NXS - 2) : "memory");
            if (!(ablate & 2)) make_lo(si * (uint32_t)XS_BYTES);
            fence_proxy_async();
            mbar_arrive(a_full + 8 * si);
            const int j = i + NXS - 1;
            const int idx_j1 = j + 1 < nitems ? load_idx(items[j + 1]) : -1;
            if (j < nitems) {
                const uint32_t sj = (uint32_t)(j % NXS);
                mbar_wait(a_empty + 8 * sj, ((uint32_t)(j / NXS) & 1u) ^ 1u);
                if (!(ablate & 2)) copy_rows(items[j], idx_j, sj * (uint32_t)XS_BYTES);
                else asm volatile("cp.async.commit_group;" ::: "memory");
            } else {
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            idx_j = idx_j1;
        }
        // =========================================================== epilogue: TMEM -> smem transpose -> global
        if (timing) g_times[blockIdx.x][2] = clock64();
        mbar_wait(done_bar, 0);
        if (timing) g_times[blockIdx.x][3] = clock64();
        tc_fence_after();
        const int q4 = warp & 3;                 // TMEM lane quadrant this warp may read
        const int cw = warp >> 2;                // which 64 of the tile's 256 columns (rows) this warp moves
        const int c4n = cpart >> 2;              // float4 per output row
        // channel held by this lane; "lo" lanes (W_lo X partial sums, non-WIDE only) go to a second buffer
        int ch; bool is_lo;
        if (WIDE) { ch = q4 * 32 + lane; is_lo = false; }
        else { ch = (q4 & 1) * 32 + lane; is_lo = q4 >= 2; }
        const bool active = ch < cpart;
        float* const sOutLo = sOut + TR * cpart;                  // non-WIDE: 2 x (256 rows x cout) fp32 <= 128 KB
        for (int t = 0; t < NTILE; ++t) {
            if ((tile0 + t) * TR >= a.n_out) break;
            const bool started = valid[t] != 0;
            if (active) {
                float* const dstbuf = is_lo ? sOutLo : sOut;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t v[16];
                    if (started) tmem_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(t * TR + cw * 64 + c0), v);
                    else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) dstbuf[(size_t)(cw * 64 + c0 + j) * cpart + ch] = __uint_as_float(v[j]);
                }
            }
            bar_sync_producers();
            // row-major pass: fused epilogue + coalesced stores; 4 float4 per thread per batch, loads first
            const int total = TR * c4n;
            for (int e0 = 0; e0 < total; e0 += 4 * NPT) {
                int rowv[4]; float4 resv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * NPT + tid;
                    const int rr = (tile0 + t) * TR + e / c4n;
                    rowv[u] = (e < total && rr < a.n_out) ? (a.row_perm ? __ldg(a.row_perm + rr) : rr) : -1;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * NPT + tid;
                    const int col = part * 128 + (e % c4n) * 4;
                    resv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (a.residual && rowv[u] >= 0)
                        resv[u] = __ldg(reinterpret_cast<const float4*>(a.residual + (size_t)rowv[u] * a.cout + col));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * NPT + tid;
                    if (e >= total) continue;                    // uniform: total is a multiple of NPT
                    const int r = e / c4n, cq = e - r * c4n;
                    const int col = part * 128 + cq * 4;
                    float4 y = *reinterpret_cast<const float4*>(sOut + (size_t)r * cpart + cq * 4);
                    if (!WIDE) {
                        const float4 z = *reinterpret_cast<const float4*>(sOutLo + (size_t)r * cpart + cq * 4);
                        y.x = __fadd_rn(y.x, z.x); y.y = __fadd_rn(y.y, z.y); y.z = __fadd_rn(y.z, z.z); y.w = __fadd_rn(y.w, z.w);
                    }
                    if (a.scale) {
                        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.scale + col));
                        float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (a.shift) sh = __ldg(reinterpret_cast<const float4*>(a.shift + col));
                        y.x = __fmaf_rn(y.x, sc.x, sh.x); y.y = __fmaf_rn(y.y, sc.y, sh.y);
                        y.z = __fmaf_rn(y.z, sc.z, sh.z); y.w = __fmaf_rn(y.w, sc.w, sh.w);
                    } else if (a.shift) {
                        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.shift + col));
                        y.x += sh.x; y.y += sh.y; y.z += sh.z; y.w += sh.w;
                    }
                    if (a.residual) { y.x += resv[u].x; y.y += resv[u].y; y.z += resv[u].z; y.w += resv[u].w; }
                    if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                    if (a.l2norm) {              // host guarantees the whole row sits in this CTA (cout <= 128): c4n lanes per row
                        float ss = __fmaf_rn(y.w, y.w, __fmaf_rn(y.z, y.z, __fmaf_rn(y.y, y.y, __fmul_rn(y.x, y.x))));
                        for (int o = 1; o < c4n; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                        const float nrm = sqrtf(ss);
                        y.x = __fdiv_rn(y.x, nrm); y.y = __fdiv_rn(y.y, nrm); y.z = __fdiv_rn(y.z, nrm); y.w = __fdiv_rn(y.w, nrm);
                    }
                    if (rowv[u] >= 0) *reinterpret_cast<float4*>(a.out + (size_t)rowv[u] * a.cout + col) = y;
                }
            }
            bar_sync_producers();                // the buffers are reused by the next tile
        }
        tc_fence_before();
        if (timing) g_times[blockIdx.x][4] = clock64();
    } else if (warp == wload_warp) {
        // =========================================================== weight slabs: one TMA bulk copy each
        if (lane == 0) {
            uint32_t w_it = 0;
            for (int i = 0; i < nitems; ++i) {
                const uint32_t it = items[i];
                if (!item_first(it)) continue;
                const uint32_t ws = w_it % NSW;
                mbar_wait(w_empty + 8 * ws, ((w_it / NSW) & 1u) ^ 1u);
                if (ablate & 4) { mbar_arrive(w_full + 8 * ws); ++w_it; continue; }
                mbar_expect_tx(w_full + 8 * ws, W_BYTES);
                bulk_g2s(sW + ws * W_BYTES,
                         a.wt_img + (((size_t)item_k(it) * nch + item_c(it)) * nparts + part) * (W_BYTES / 4), W_BYTES,
                         w_full + 8 * ws);
                ++w_it;
            }
        }
        __syncwarp();
    } else {
        // =========================================================== MMA issuer (one thread)
        uint32_t w_it = 0, started = 0;
        uint32_t wh = 0;
        for (int i = 0; i < nitems; ++i) {
            const uint32_t it = items[i];
            if (item_first(it)) {
                if (w_it > 0 && lane == 0) umma_commit(w_empty + 8 * ((w_it - 1) % NSW));   // previous slab fully consumed
                const uint32_t ws = w_it % NSW;
                mbar_wait(w_full + 8 * ws, (w_it / NSW) & 1u);
                wh = sW + ws * W_BYTES;
                ++w_it;
            }
            const int s = i % NXS, t = item_t(it);
            mbar_wait(a_full + 8 * s, (uint32_t)(i / NXS) & 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t xh = sX + s * XS_BYTES, xl = xh + XH_BYTES;
                const uint32_t d = tmem_base + (uint32_t)(t * TR);
                uint32_t acc = (started >> t) & 1u;
#pragma unroll
                for (int j = 0; j < KC / 8; ++j) {
                    if (ablate & 1) break;
                    const uint64_t dxh = make_desc(xh + j * 32), dxl = make_desc(xl + j * 32);
                    const uint64_t dwh = make_desc(wh + j * 32);
                    umma_tf32(d, dwh, dxh, IDESC, acc);                  // (W_hi [; W_lo]) X_hi
                    if (WIDE) {
                        const uint64_t dwl = make_desc(wh + 128 * KC * 4 + j * 32);
                        umma_tf32(d, dwl, dxh, IDESC, 1u);               // W_lo X_hi
                    }
                    umma_tf32(d, dwh, dxl, IDESC, 1u);                   // (W_hi [; W_lo]) X_lo
                    acc = 1u;
                }
                umma_commit(a_empty + 8 * s);
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
        tmem_dealloc(tmem_base, 512);
    }
}

// weight [K, cin, cout] -> shared-memory images.  Per (k, 32-channel chunk ci, 128-channel part p) one slab of
// 2 * RP rows x 32 channels (RP = 64 when cout <= 64, else 128): rows [0, RP) = W_hi (tf32), rows [RP, 2 RP) = W_lo,
// row n = output channel p*128 + n (zero rows when n >= cout); each row's eight 16-byte chunks XOR-swizzled with
// (row & 7) exactly as the SWIZZLE_128B K-major UMMA descriptor expects them.
__global__ void split_weights_kernel(const float* __restrict__ w, int K, int cin, int cout, float* __restrict__ img) {
    const int RP = cout <= 64 ? 64 : 128;
    const int nparts = cout <= 64 ? 1 : cout / 128;
    const size_t total = (size_t)K * (cin / KC) * nparts * RP * KC;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cl = (int)(i % KC);
    const int n = (int)((i / KC) % RP);
    const int p = (int)((i / ((size_t)KC * RP)) % nparts);
    const int ci = (int)((i / ((size_t)KC * RP * nparts)) % (cin / KC));
    const int k = (int)(i / ((size_t)KC * RP * nparts * (cin / KC)));
    const int ch = p * 128 + n;
    float h = 0.f, l = 0.f;
    if (ch < cout) {
        const float x = w[((size_t)k * cin + ci * KC + cl) * cout + ch];
        h = __uint_as_float(to_tf32(x));
        l = __uint_as_float(to_tf32(x - h));
    }
    const size_t slab = (((size_t)k * (cin / KC) + ci) * nparts + p) * (2 * (size_t)RP * KC);
    const int off = n * KC + ((((cl >> 2) ^ (n & 7)) << 2) | (cl & 3));
    img[slab + off] = h;
    img[slab + (size_t)RP * KC + off] = l;
}

template <bool WIDE, int NSW, int NXS>
int launch_tc(const TcArgs& a, cudaStream_t stream) {
    const size_t smem = (size_t)NXS * XS_BYTES + (size_t)NSW * (WIDE ? 256 : 128) * KC * 4;
    EYOC_CUDA(cudaFuncSetAttribute(sparse_conv_tc_kernel<WIDE, NSW, NXS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (a.n_out + TR - 1) / TR;
    dim3 grid((tiles + NTILE - 1) / NTILE, WIDE ? a.cout / 128 : 1);
    sparse_conv_tc_kernel<WIDE, NSW, NXS><<<grid, NPT + 64, smem, stream>>>(a);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

}  // namespace

extern "C" int eyoc_debug_conv_ablate(int flags) {
    EYOC_CUDA(cudaMemcpyToSymbol(g_ablate, &flags, sizeof(int)));
    return EYOC_OK;
}

extern "C" int eyoc_debug_conv_times(long long* host_out_1024x6) {
    EYOC_CUDA(cudaMemcpyFromSymbol(host_out_1024x6, g_times, sizeof(long long) * 1024 * 6));
    return EYOC_OK;
}

extern "C" size_t eyoc_conv_weight_image_floats(int K, int cin, int cout) {
    const size_t RP = cout <= 64 ? 64 : 128, nparts = cout <= 64 ? 1 : cout / 128;
    return (size_t)K * (cin / KC) * nparts * 2 * RP * KC;
}

extern "C" int eyoc_conv_split_weights(const float* weight, int K, int cin, int cout, float* wt_img, cudaStream_t stream) {
    EYOC_CHECK_ARG(weight && wt_img && K >= 1 && cin >= 1 && cout >= 1, "eyoc_conv_split_weights: bad argument");
    EYOC_CHECK_ARG(cin % KC == 0 && (cout == 32 || cout == 64 || cout == 128 || cout == 256),
                   "eyoc_conv_split_weights: cin must be a multiple of 32 and cout one of 32, 64, 128, 256");
    const size_t total = eyoc_conv_weight_image_floats(K, cin, cout) / 2;
    split_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(weight, K, cin, cout, wt_img);
    EYOC_LAUNCH_CHECK();
    return EYOC_OK;
}

extern "C" int eyoc_sparse_conv_tc_supported(int c0, int c1, int cout, int K, int l2norm) {
    const int cin = c0 + c1;
    if (cin % KC || c0 % KC || K < 1 || K > 27 || cin / KC > 12) return 0;
    if (!(cout == 32 || cout == 64 || cout == 128 || cout == 256)) return 0;
    if (l2norm && cout > 128) return 0;
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
    TcArgs a{in0, c0, in1, c1, nbr, row_perm, wt_img, scale, shift, residual, out, K, (int)n_out, cout, relu, l2norm, nbr_tiled};
    if (cout <= 64) return launch_tc<false, 2, 3>(a, stream);      // 3 x 64 KB X stages + 2 x 16 KB weight slabs
    return launch_tc<true, 2, 2>(a, stream);                         // 2 x 64 KB X stages + 2 x 32 KB weight slabs
}
