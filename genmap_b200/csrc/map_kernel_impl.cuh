// map_kernel_impl.cuh — the search kernel template shared by map_kernel.cu (frequencies) and
// locate_kernel.cu (csv lists).  See map_kernel.cu for the mapping to the machine.
#pragma once
#include "map_kernel.cuh"

namespace gmb {

namespace {

constexpr int kThreads = 256;
constexpr int kCounterWords = 14; // instrumented instantiation: words of MapLaunch::fetch_counter
#ifndef GMB_MIN_BLOCKS
#define GMB_MIN_BLOCKS 4 // resident CTAs per SM the register allocation must allow
#endif
#ifndef GMB_MIN_BLOCKS5
#define GMB_MIN_BLOCKS5 3 // Dna5 indices, blocked instantiation (five-way children, 12-word frames): 3 CTAs / 80 registers
#endif                    // measured best (profiles/r01/s15_sweep_dna5_*.txt); the one-k-mer instantiation fits 64 like Dna4

template <int FW> // FW: words per mismatch frame (10 for Dna4, 12 for Dna5)
struct SmemFrames {
    uint32_t* base;  // + threadIdx.x; word i of this chain at base[i * kThreads] (conflict-free)
    uint32_t xoff;   // first word after the E mismatch frames
    __device__ __forceinline__ void set(uint32_t lv, uint32_t i, uint32_t v) { base[(lv * FW + i) * kThreads] = v; }
    __device__ __forceinline__ uint32_t get(uint32_t lv, uint32_t i) const { return base[(lv * FW + i) * kThreads]; }
    __device__ __forceinline__ void xset(uint32_t i, uint32_t v) { base[(xoff + i) * kThreads] = v; }
    __device__ __forceinline__ uint32_t xget(uint32_t i) const { return base[(xoff + i) * kThreads]; }
    // the per-window counters (kept apart from xset / xget: a kernel that lets one lane count for another chain would
    // redirect and atomise exactly these — tried in block_kernel.cu, see there)
    __device__ __forceinline__ void cset(uint32_t i, uint32_t v) { base[(xoff + i) * kThreads] = v; }
    __device__ __forceinline__ uint32_t cget(uint32_t i) const { return base[(xoff + i) * kThreads]; }
    __device__ __forceinline__ void cadd(uint32_t i, uint32_t n, uint32_t maxv)
    {
        uint32_t* p = base + (xoff + i) * kThreads;
        const uint64_t sum = (uint64_t)*p + n;
        *p = sum < maxv ? (uint32_t)sum : maxv; // saturating (src/algo.hpp:48,191)
    }
    __device__ __forceinline__ void cor(uint32_t i, uint32_t bits) { base[(xoff + i) * kThreads] |= bits; }
};

__host__ __device__ inline uint32_t align32(uint32_t x) { return (x + 31u) & ~31u; }
constexpr uint32_t kStartWords = sizeof(SearchStart) / 4;

// MINB: resident CTAs per SM the register allocation must allow (Dna4; Dna5 blocked: GMB_MIN_BLOCKS5)
template <int KW, bool COUNT, typename OutT, bool EP, bool BLK, int SIGMA, bool LOC = false, int MINB = GMB_MIN_BLOCKS>
__global__ void __launch_bounds__(kThreads, (SIGMA == 5 && BLK) ? GMB_MIN_BLOCKS5 : MINB) map_kernel(const MapLaunch L)
{
    // shared memory: step tables | jump-table starts | offsets | per-chain frame store
    extern __shared__ uint32_t smem[];
    const uint32_t n_start_words = (L.cx.B + 1) * kMaxSearches * kStartWords;
    uint32_t* steps_s = smem;
    uint32_t* starts_s = steps_s + align32(L.n_step_words);
    uint32_t* offs_s = starts_s + align32(n_start_words);
    uint32_t* frames_s = offs_s + align32(2 * (kMaxBlockKmers + 1));
    for (uint32_t i = threadIdx.x; i < L.n_step_words; i += kThreads) steps_s[i] = L.cx.steps[i];
    for (uint32_t i = threadIdx.x; i < n_start_words; i += kThreads) starts_s[i] = reinterpret_cast<const uint32_t*>(L.cx.starts)[i];
    if (threadIdx.x == 0) {
#pragma unroll
        for (uint32_t i = 0; i <= kMaxBlockKmers; ++i) { // static indices: the parameter struct stays in constant memory
            offs_s[i] = L.p1_off[i];
            offs_s[kMaxBlockKmers + 1 + i] = L.fl_off[i];
        }
    }
    __syncthreads();

    MapCtx cx = L.cx;
    cx.steps = steps_s;
    cx.starts = reinterpret_cast<const SearchStart*>(starts_s);
    cx.p1_off = offs_s;
    cx.fl_off = offs_s + kMaxBlockKmers + 1;
    SmemFrames<(int)frame_words(SIGMA)> fr{frames_s + threadIdx.x, L.E * frame_words(SIGMA)};

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    OutT* __restrict__ out = static_cast<OutT*>(L.out);
    const unsigned long long B = BLK ? cx.B : 1ull;

    Chain<KW, SIGMA> st;
    uint64_t j = 0; // first position of the chain's block
    bool active = false, exhausted = false;
    // warp-uniform pool of consecutive positions [pool_next, pool_end) and "no more chunks" flag;
    // every refilling lane takes the next B positions (fewer at the end of a chunk)
    unsigned long long pool_next = 0, pool_end = 0;
    bool pool_done = false;
    FetchStats fetches{};
    unsigned long long lut_reads = 0;

    for (;;) {
        // ---- refill: lanes without a block take the next positions of the warp's pool ---------------
        const bool need = !active && !exhausted;
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m) {
            const unsigned cnt = __popc(m), rank = __popc(m & lt_mask);
            const unsigned long long avail = (pool_end - pool_next + B - 1) / B; // blocks left in the pool
            unsigned long long nb = 0, ne = 0;
            if (avail < cnt && !pool_done) {
                unsigned long long cid = 0;
                if (lane == 0) cid = atomicAdd(L.work_counter, 1ull);
                cid = __shfl_sync(0xffffffffu, cid, 0);
                if (cid >= L.n_chunks) {
                    pool_done = true;
                } else {
                    uint32_t lo = 0, hi = L.n_ranges; // largest r with chunk_prefix[r] <= cid (uniform loads)
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (__ldg(L.chunk_prefix + mid) <= cid) lo = mid; else hi = mid;
                    }
                    nb = __ldg(L.range_begin + lo) + (cid - __ldg(L.chunk_prefix + lo)) * L.chunk;
                    ne = nb + L.chunk;
                    const unsigned long long re = __ldg(L.range_end + lo);
                    if (ne > re) ne = re;
                }
            }
            const unsigned long long fresh = (ne - nb + B - 1) / B; // blocks in the new chunk
            unsigned long long jend = 0;
            bool got = false;
            if (need) {
                if (rank < avail) { j = pool_next + rank * B; jend = pool_end; got = true; }
                else if (rank - avail < fresh) { j = nb + (rank - avail) * B; jend = ne; got = true; }
            }
            if (cnt <= avail) {
                pool_next += cnt * B;
                if (pool_next > pool_end) pool_next = pool_end;
            } else {
                const unsigned long long want = cnt - avail;
                pool_next = nb + (want < fresh ? want : fresh) * B;
                if (pool_next > ne) pool_next = ne;
                pool_end = ne;
            }
            if (need) {
                if (got) {
                    st.cnt = (uint32_t)(jend - j < B ? jend - j : B);
                    uint64_t at = L.text_begin + j;
                    if constexpr (LOC) { if (L.loc_list) at = __ldg(L.loc_list + j); }
                    load_pattern(st.pat, L.text, L.nmask, at, cx.K + st.cnt - 1);
                    chain_begin_block<KW, EP, BLK, SIGMA>(st, fr, cx, COUNT ? &lut_reads : nullptr);
                    if (LOC && cx.loc_rows) { // second pass: where this k-mer's two lists start
                        st.loc_at_fwd = L.loc_off[2 * (j - L.loc_pos0)];
                        st.loc_at_rev = L.loc_off[2 * (j - L.loc_pos0) + 1];
                    }
                    active = true;
                } else if (pool_done) {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, active || !exhausted)) break;

        // ---- one node expansion per chain -------------------------------------------------------------
        if (active) {
            if (!chain_step<KW, EP, BLK, SIGMA, decltype(fr), LOC>(st, fr, cx, COUNT ? &fetches : nullptr, COUNT ? &lut_reads : nullptr)) {
                if constexpr (LOC) { // first pass: list lengths of this k-mer, + strand then - strand
                    if (!cx.loc_rows) { out[2 * (j - L.loc_pos0)] = (OutT)st.occ_fwd; out[2 * (j - L.loc_pos0) + 1] = (OutT)st.occ_rev; }
                } else {
                    for (uint32_t w = 0; w < st.cnt; ++w) out[j + w] = (OutT)chain_result<KW, EP, BLK, SIGMA>(st, fr, cx, w);
                }
                active = false;
            }
        }
    }
    if (COUNT) {
        // fetch_counter: [0] rank-block fetches, [1] jump-table reads, [2..9] fetches by interval size, [10] thin paths,
        // [11] state-machine iterations, [12] located entries verified, [13] text reads of those
        unsigned long long v[kCounterWords] = {fetches.total, lut_reads, fetches.by_size[0], fetches.by_size[1], fetches.by_size[2],
                                               fetches.by_size[3], fetches.by_size[4], fetches.by_size[5], fetches.by_size[6],
                                               fetches.by_size[7], fetches.thin_paths, fetches.iterations, fetches.located,
                                               fetches.text_reads};
#pragma unroll
        for (int k = 0; k < kCounterWords; ++k) {
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if (lane == 0 && v[k]) atomicAdd(L.fetch_counter + k, v[k]);
        }
    }
}

template <int KW, bool COUNT, typename OutT, bool EP, bool BLK, int SIGMA, bool LOC = false, int MINB = GMB_MIN_BLOCKS>
cudaError_t launch_b(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    auto kern = map_kernel<KW, COUNT, OutT, EP, BLK, SIGMA, LOC, MINB>;
    const size_t smem = map_kernel_smem_bytes(L.n_step_words, L.E, L.cx.B, EP, SIGMA, BLK);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    // persistent grid: every resident CTA slot of every SM, but never more threads than blocks of work
    unsigned long long want = (L.n_work / L.cx.B + kThreads) / kThreads;
    unsigned long long grid = (unsigned long long)sm_count * per_sm;
    if (want < grid) grid = want ? want : 1;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(L);
    return cudaGetLastError();
}

} // namespace

} // namespace gmb
