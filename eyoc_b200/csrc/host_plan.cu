// Host-side index planning of a block of pairs: the six numpy draws the reference makes per pair on the GLOBAL legacy
// RandomState (scripts/test_kitti.py:33-34 find_corr, :69-71 random_sample x2, scripts/SC2_PCR/SC2_PCR.py:288-289
// match_pair), restated on the raw MT19937 state so that the stream is consumed bit for bit as numpy consumes it:
//   choice(n, k, replace=False)  = permutation(n)[:k]           (legacy RandomState.choice)
//   permutation(n)               = Fisher-Yates from the back, j = random_interval(i)
//   random_interval(max)         = masked rejection on 32-bit draws (max <= 2^32-1)
//   choice(n, k) / randint(0, n) = masked rejection on 32-bit draws per element (legacy _rand_int64, use_masked)
// numpy spends ~3.4 ms per pair in these calls (218 ms for a block of 64: more than the whole GPU step).  Here the
// stream-dependent part (the accepted swap partners) is produced sequentially and the memory-bound part (replaying the
// swaps) is spread over worker threads.  tests/test_host_logic.py pins it against numpy itself.  Pure host code, no device work.
#include "common.cuh"
#include "../../include/eyoc_b200.h"
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <thread>
#include <vector>

namespace {

struct MT {
    uint32_t* key;   // [624]
    int pos;
    uint32_t out[624];      // the tempered outputs of the current state block (filled by gen(): vectorisable, unlike next())
    bool have_out;
    MT(uint32_t* k, int p) : key(k), pos(p), have_out(false) {}
    static inline uint32_t temper(uint32_t y) {
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    void temper_all() {
#pragma GCC ivdep
        for (int i = 0; i < 624; ++i) out[i] = temper(key[i]);
        have_out = true;
    }
    // state recurrence in three dependency-free stretches (each reads only words the stretch does not write)
    void gen() {
        const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
        uint32_t* k = key;
#pragma GCC ivdep
        for (int i = 0; i < 227; ++i) {
            const uint32_t y = (k[i] & UPPER) | (k[i + 1] & LOWER);
            k[i] = k[i + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
        }
        // k[i] <- k[i - 227], k[i + 1]: the sources at i - 227 were written >= 227 iterations earlier: blocks of 227 are free
        for (int b0 = 227; b0 < 623; b0 += 227) {
            const int b1 = b0 + 227 < 623 ? b0 + 227 : 623;
#pragma GCC ivdep
            for (int i = b0; i < b1; ++i) {
                const uint32_t y = (k[i] & UPPER) | (k[i + 1] & LOWER);
                k[i] = k[i - 227] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
            }
        }
        const uint32_t y = (k[623] & UPPER) | (k[0] & LOWER);
        k[623] = k[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
        pos = 0;
        temper_all();
    }
    inline uint32_t next() {
        if (pos == 624) gen();
        else if (!have_out) temper_all();
        return out[pos++];
    }
    inline uint32_t bounded(uint32_t max) {       // uniform on [0, max], numpy's masked rejection
        if (max == 0) return 0;
        uint32_t mask = max;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = (next() & mask)) > max) {}
        return v;
    }
};

// The stream-dependent half of permutation(n): the accepted swap partners j_i = random_interval(i), i = n-1 .. 1
// (js[i] = j_i).  Sequential in the MT19937 stream, but touches no large array.
// Written without a data-dependent branch: every raw draw is masked and stored, the index only advances when the draw
// is accepted (numpy's  while ((v = next & mask) > max);  consumes exactly the same raw draws).
void perm_draws(MT& mt, int64_t n, uint32_t* __restrict__ js) {
    int64_t i = n - 1;
    while (i >= 1) {
        // the rejection mask only changes when i crosses a power of two: run each bit range as a tight loop over the block of
        // tempered outputs at hand (no per-draw state checks; the only loop-carried value is i)
        const uint32_t mask = 0xffffffffu >> __builtin_clz((uint32_t)i);
        const int64_t lo = (int64_t)(mask >> 1) + 1;                 // smallest i that still uses this mask
        while (i >= lo) {
            if (mt.pos == 624) mt.gen();
            else if (!mt.have_out) mt.temper_all();
            const uint32_t* __restrict__ o = mt.out + mt.pos;
            const int avail = 624 - mt.pos;
            int c = 0;
            while (c < avail && i >= lo) {
                const uint32_t v = o[c++] & mask;
                js[i] = v;
                i -= (v <= (uint32_t)i);
            }
            mt.pos += c;
        }
    }
}
// The memory-bound half: replay the swaps, keep permutation(n)[:k].  Independent of the stream -> runs on worker threads.
void perm_apply(int64_t n, int64_t k, const uint32_t* js, int64_t* scratch, int64_t* out) {
    for (int64_t i = 0; i < n; ++i) scratch[i] = i;
    for (int64_t i = n - 1; i >= 1; --i) {
        const int64_t j = js[i];
        const int64_t t = scratch[i]; scratch[i] = scratch[j]; scratch[j] = t;
    }
    for (int64_t i = 0; i < k; ++i) out[i] = scratch[i];
}
void with_replacement(MT& mt, int64_t n, int64_t k, int64_t* out) {       // out has room for k + 1
    const uint32_t max = (uint32_t)(n - 1);
    if (max == 0) { for (int64_t i = 0; i < k; ++i) out[i] = 0; return; }     // numpy draws nothing for a range of one
    const uint32_t mask = 0xffffffffu >> __builtin_clz(max);
    int64_t c = 0;
    while (c < k) {
        if (mt.pos == 624) mt.gen();
        else if (!mt.have_out) mt.temper_all();
        const uint32_t* o = mt.out + mt.pos;
        const int avail = 624 - mt.pos;
        int u = 0;
        while (u < avail && c < k) {
            const uint32_t v = o[u++] & mask;
            out[c] = (int64_t)v;
            c += (v <= max);
        }
        mt.pos += u;
    }
}

// everything pair p needs after the sequential pass; the buffers live in an arena that is kept between calls (fresh vectors
// cost 30 MB of page faults and zero fill per block)
struct PairDraws {
    struct U32 { uint32_t* p = nullptr; uint32_t* data() const { return p; } bool empty() const { return p == nullptr; } };
    struct I64 { int64_t* p = nullptr; int64_t* data() const { return p; } int64_t operator[](size_t i) const { return p[i]; } };
    U32 js[4];     // find_corr 0 / 1, random_sample 0 / 1 (null = not a permutation)
    I64 wr[2];     // random_sample with replacement (n < num_sample)
    I64 mp[2];     // match_pair draws
};
struct Arena {
    std::vector<char> buf;
    size_t off = 0;
    void reset(size_t bytes) { if (buf.size() < bytes) buf.resize(bytes); off = 0; }
    template <typename T> T* take(size_t n) { T* r = (T*)(buf.data() + off); off += (n * sizeof(T) + 63) / 64 * 64; return r; }
};

}  // namespace

extern "C" int eyoc_plan_draws(uint32_t* mt_key624, int32_t* mt_pos, int num_pairs, const int64_t* n0, const int64_t* n1,
                               const int64_t* row_offsets, int subsample_size, int num_sample, int num_node, int64_t* fc0,
                               int64_t* fc1, int64_t* src, int64_t* tgt) {
    EYOC_CHECK_ARG(mt_key624 && mt_pos && n0 && n1 && row_offsets && src && tgt, "eyoc_plan_draws: null argument");
    EYOC_CHECK_ARG(*mt_pos >= 0 && *mt_pos <= 624, "eyoc_plan_draws: bad MT19937 position %d", *mt_pos);
    EYOC_CHECK_ARG(num_pairs >= 0 && num_sample >= 1 && num_node >= 1 && subsample_size >= 1, "eyoc_plan_draws: bad sizes");
    EYOC_CHECK_ARG((fc0 != nullptr) == (fc1 != nullptr), "eyoc_plan_draws: fc0 and fc1 go together");
    MT mt(mt_key624, *mt_pos);
    for (int p = 0; p < num_pairs; ++p) {
        EYOC_CHECK_ARG(n0[p] >= 1 && n1[p] >= 1 && n0[p] < (1ll << 32) && n1[p] < (1ll << 32), "eyoc_plan_draws: bad cloud size");
        // the fast path covers the fixed-shape case (every cloud larger than the subsample, as on KITTI)
        EYOC_CHECK_ARG(!fc0 || (n0[p] > subsample_size && n1[p] >= subsample_size), "eyoc_plan_draws: cloud smaller than the find_corr subsample");
    }
    const auto t_start = std::chrono::steady_clock::now();
    // ---- pass 1 (sequential in the RNG stream, reference order): swap partners and with-replacement draws
    std::vector<PairDraws> D((size_t)num_pairs);
    static thread_local Arena arena;
    {
        size_t need = 0;
        for (int p = 0; p < num_pairs; ++p)
            need += (size_t)(n0[p] + n1[p] + 64) * 4 * 2 + (size_t)(num_sample + 9) * 8 * 2 + (size_t)(num_node + 9) * 8 * 2;
        arena.reset(need);
    }
    for (int p = 0; p < num_pairs; ++p) {
        const int64_t nn[2] = {n0[p], n1[p]};
        if (fc0)
            for (int side = 0; side < 2; ++side) { D[p].js[side].p = arena.take<uint32_t>((size_t)nn[side]); perm_draws(mt, nn[side], D[p].js[side].p); }
        for (int side = 0; side < 2; ++side) {
            const int64_t n = nn[side];
            if (n > num_sample) { D[p].js[2 + side].p = arena.take<uint32_t>((size_t)n); perm_draws(mt, n, D[p].js[2 + side].p); }
            else if (n < num_sample) { D[p].wr[side].p = arena.take<int64_t>((size_t)num_sample + 1); with_replacement(mt, n, num_sample, D[p].wr[side].p); }
        }
        for (int side = 0; side < 2; ++side) { D[p].mp[side].p = arena.take<int64_t>((size_t)num_node + 1); with_replacement(mt, num_sample, num_node, D[p].mp[side].p); }
    }
    *mt_pos = mt.pos;
    const auto t_mid = std::chrono::steady_clock::now();
    // ---- pass 2 (independent per pair): replay the swaps, compose the indices
    auto work = [&](int p) {
        const int64_t nn[2] = {n0[p], n1[p]};
        const int64_t off[2] = {row_offsets[2 * p], row_offsets[2 * p + 1]};
        std::vector<int64_t> scratch((size_t)(nn[0] > nn[1] ? nn[0] : nn[1])), rs((size_t)num_sample);
        for (int side = 0; side < 2; ++side) {
            if (fc0) {
                int64_t* out = (side ? fc1 : fc0) + (size_t)p * subsample_size;
                perm_apply(nn[side], subsample_size, D[p].js[side].data(), scratch.data(), out);
                for (int i = 0; i < subsample_size; ++i) out[i] += off[side];
            }
            const int64_t n = nn[side];
            if (n == num_sample) for (int i = 0; i < num_sample; ++i) rs[i] = i;
            else if (n > num_sample) perm_apply(n, num_sample, D[p].js[2 + side].data(), scratch.data(), rs.data());
            else for (int i = 0; i < num_sample; ++i) rs[i] = D[p].wr[side][(size_t)i];
            int64_t* out = (side ? tgt : src) + (size_t)p * num_node;
            for (int i = 0; i < num_node; ++i) out[i] = rs[(size_t)D[p].mp[side][i]] + off[side];
        }
    };
    unsigned nthreads = std::thread::hardware_concurrency();
    nthreads = nthreads < 1 ? 1 : (nthreads > 8 ? 8 : nthreads);
    if (const char* e = getenv("EYOC_PLAN_THREADS")) {          // tuning: worker threads of the replay pass (1..64)
        const int v = atoi(e);
        if (v >= 1 && v <= 64) nthreads = (unsigned)v;
    }
    if ((int)nthreads > num_pairs) nthreads = num_pairs > 0 ? (unsigned)num_pairs : 1u;
    std::atomic<int> next(0);
    auto loop = [&]() { for (int p = next.fetch_add(1); p < num_pairs; p = next.fetch_add(1)) work(p); };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(loop);
    loop();
    for (auto& th : pool) th.join();
    if (getenv("EYOC_PLAN_TIMING")) {
        const auto t_end = std::chrono::steady_clock::now();
        fprintf(stderr, "eyoc_plan_draws: %d pairs, stream pass %.2f ms, replay pass %.2f ms (%u threads)\n", num_pairs,
                std::chrono::duration<double, std::milli>(t_mid - t_start).count(),
                std::chrono::duration<double, std::milli>(t_end - t_mid).count(), nthreads);
    }
    return EYOC_OK;
}
