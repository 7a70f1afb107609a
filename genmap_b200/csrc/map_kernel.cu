// map_kernel.cu — the (K,E)-frequency kernel for sm_100a.
//
// Replaces the reference's per-position loop computeMappability -> computeMappabilitySingleBlock
// (src/algo.hpp:221-483) and its search-scheme matcher (src/find2_index_approx.hpp:223-457).
//
// Mapping to the machine:
//   * one CUDA thread = one "chain" = one k-mer start at a time; a chain is a strictly dependent series
//     of random 64-byte rank-block reads, so throughput comes from the number of chains in flight
//     (148 SMs x 1024 resident threads ~ 150 k independent 64-B requests), not from intra-chain width.
//   * each loop iteration is one node expansion of gmb_core.h::chain_step — the same code for every
//     chain whatever its depth, error level or search, so warps stay converged although every lane
//     walks a different subtree.
//   * persistent grid (SM count x resident CTAs); lanes that finish their k-mer refill immediately
//     from a warp-local pool fed by one global atomic per 128 positions, so repeats (whose searches are
//     100x longer) never idle a warp.
//   * the <= E backtracking frames live in shared memory, strided by thread (conflict-free); the
//     step table of the search scheme is staged in shared memory once per CTA.
#include "map_kernel.cuh"

namespace gmb {

namespace {

constexpr int kThreads = 256;
#ifndef GMB_MIN_BLOCKS
#define GMB_MIN_BLOCKS 4 // resident CTAs per SM the register allocation must allow
#endif

template <int FW> // words per mismatch frame (10 for Dna4, 12 for Dna5)
struct SmemFrames {
    uint32_t* base;  // + threadIdx.x; word i of this chain at base[i * kThreads] (conflict-free)
    uint32_t xoff;   // first word after the E mismatch frames
    __device__ __forceinline__ void set(uint32_t lv, uint32_t i, uint32_t v) { base[(lv * FW + i) * kThreads] = v; }
    __device__ __forceinline__ uint32_t get(uint32_t lv, uint32_t i) const { return base[(lv * FW + i) * kThreads]; }
    __device__ __forceinline__ void xset(uint32_t i, uint32_t v) { base[(xoff + i) * kThreads] = v; }
    __device__ __forceinline__ uint32_t xget(uint32_t i) const { return base[(xoff + i) * kThreads]; }
};

__host__ __device__ inline uint32_t align32(uint32_t x) { return (x + 31u) & ~31u; }
constexpr uint32_t kStartWords = sizeof(SearchStart) / 4;

template <int KW, bool COUNT, typename OutT, bool EP, bool BLK, int SIGMA>
__global__ void __launch_bounds__(kThreads, SIGMA == 5 ? 2 : GMB_MIN_BLOCKS) map_kernel(const MapLaunch L)
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
    unsigned long long fetches = 0, lut_reads = 0;

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
                    load_pattern(st.pat, L.text, L.nmask, L.text_begin + j, cx.K + st.cnt - 1);
                    chain_begin_block<KW, EP, BLK, SIGMA>(st, fr, cx, COUNT ? &lut_reads : nullptr);
                    active = true;
                } else if (pool_done) {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, active || !exhausted)) break;

        // ---- one node expansion per chain -------------------------------------------------------------
        if (active) {
            if (!chain_step<KW, EP, BLK, SIGMA>(st, fr, cx, COUNT ? &fetches : nullptr, COUNT ? &lut_reads : nullptr)) {
                for (uint32_t w = 0; w < st.cnt; ++w) out[j + w] = (OutT)chain_result<KW, EP, BLK, SIGMA>(st, fr, cx, w);
                active = false;
            }
        }
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            fetches += __shfl_xor_sync(0xffffffffu, fetches, o);
            lut_reads += __shfl_xor_sync(0xffffffffu, lut_reads, o);
        }
        if (lane == 0 && fetches) atomicAdd(L.fetch_counter, fetches);
        if (lane == 0 && lut_reads) atomicAdd(L.fetch_counter + 1, lut_reads);
    }
}

template <int KW, bool COUNT, typename OutT, bool EP, bool BLK, int SIGMA>
cudaError_t launch_b(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    auto kern = map_kernel<KW, COUNT, OutT, EP, BLK, SIGMA>;
    const size_t smem = map_kernel_smem_bytes(L.n_step_words, L.E, L.cx.B, EP, SIGMA);
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

template <int KW, bool COUNT, typename OutT, bool EP>
cudaError_t launch_t(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    // Dna5 indices always run the blocked instantiation (it covers B == 1) to keep the number of kernels down
    if (L.sigma == 5) return launch_b<KW, COUNT, OutT, EP, true, 5>(L, sm_count, stream);
    return L.cx.B > 1 ? launch_b<KW, COUNT, OutT, EP, true, 4>(L, sm_count, stream) : launch_b<KW, COUNT, OutT, EP, false, 4>(L, sm_count, stream);
}

template <int KW>
cudaError_t launch_kw(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.exclude_pseudo) // (the fetch-counting instantiation is not built for --exclude-pseudo)
        return L.value_bits == 16 ? launch_t<KW, false, uint16_t, true>(L, sm_count, stream)
                                  : launch_t<KW, false, uint8_t, true>(L, sm_count, stream);
    if (L.value_bits == 16)
        return L.count_fetches ? launch_t<KW, true, uint16_t, false>(L, sm_count, stream)
                               : launch_t<KW, false, uint16_t, false>(L, sm_count, stream);
    return L.count_fetches ? launch_t<KW, true, uint8_t, false>(L, sm_count, stream)
                           : launch_t<KW, false, uint8_t, false>(L, sm_count, stream);
}

} // namespace

size_t map_kernel_smem_bytes(uint32_t n_step_words, uint32_t E, uint32_t B, bool ep, uint32_t sigma)
{
    const size_t tables = align32(n_step_words) + align32((B + 1) * kMaxSearches * kStartWords) + align32(2 * (kMaxBlockKmers + 1));
    return (tables + (size_t)frame_store_words(E, B, ep, (int)sigma, B > 1 || sigma == 5) * kThreads) * sizeof(uint32_t);
}

cudaError_t launch_map_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.n_work == 0) return cudaSuccess;
    const uint32_t needle = L.cx.K + L.cx.B - 1; // characters a chain keeps in registers
    if (needle <= 32) return launch_kw<1>(L, sm_count, stream);
    if (needle <= 64) return launch_kw<2>(L, sm_count, stream);
    if (needle <= 128) return launch_kw<4>(L, sm_count, stream);
    return launch_kw<9>(L, sm_count, stream);
}

} // namespace gmb
