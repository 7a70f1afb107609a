// block_kernel.cu — the (K,E)-frequency kernel for E >= 1: blocks of adjacent k-mers, every search entered through the
// 16-byte jump-table entries of all the strings it admits.  Dna4 indices, and Dna5 indices in calls whose searches skip
// the text's N (MapCtx::skip_n: the alignments to text windows with an N are counted by the N pass of capi.cu).
//
// Replaces the same reference path as map_kernel.cu (computeMappabilitySingleBlock over the optimum search scheme,
// src/algo.hpp:221-403, src/find2_index_approx.hpp:223-457); the counts are the same by construction: it runs the
// same searches over the same tables, in a different order.
//
// Why a second kernel.  With the substituted keys and the LOCATED entries (gmb_core.h: JtFull) most
// of a block's work is no longer walking the index but reading table entries: at 3 Gbp, K = 30, E = 2 a block of
// three k-mers reads 234 entries per strand, half of them empty, a third located (finished by comparing needle and
// text: verify_located_key), one in seven the root of a real walk.  The general kernel sends every chain through one
// loop body that can do all of it, one entry per pass (~260 thread-instructions), so lanes reading entries wait for
// lanes expanding nodes and the reverse: 10-15 of 32 lanes active, the issue slots the bottleneck
// (profiles/r02/s1_ncu_counters_e{1,2}.csv).  Here a warp moves through two phases TOGETHER:
//   key phase   all lanes enumerate the next 32 keys of their block's current strand from a flat list the host
//               prepared (search, substituted characters as an XOR mask, errors): four entries requested back to back,
//               then looked at — empty: nothing; located: compared on the spot; else: the entry is put aside
//               (8 per chain in shared memory; beyond that a bit remembers the key and the entry is read again);
//   walk phase  every lane walks the subtrees below the entries it put aside with the state machine of gmb_core.h
//               in subtree mode (chain_step<..., SUB>), until no lane has one left.
// (Tried and rejected: dealing the warp's put-aside entries out round robin, walking for the owning chain through a
// shared-memory copy of its needle and atomic window counters — balanced task counts, but E = 1 17.7 -> 28.5 ms and
// E = 2 22.2 -> 32.8 ms: profiles/r02/s8_sweep_shared_walk_rejected.txt; the frame store keeps the counter accessors
// that experiment introduced.)
// One chain = one block of up to B adjacent k-mer starts, as in the general kernel (same step tables, same frame
// store); a warp takes 32 blocks per global atomic.
#include "map_kernel_impl.cuh"

namespace gmb {

namespace {

#ifndef GMB_PEND_SLOTS
#define GMB_PEND_SLOTS 8
#endif
constexpr uint32_t kPendSlots = GMB_PEND_SLOTS; // entries a chain can put aside per round
constexpr uint32_t kPendWords = 4;      // lo_r, size, lo_f, key index
constexpr uint32_t kRound = 32;         // keys per round (one overflow bit each; 64 with a 64-bit mask and 16 slots: slower,
                                        // 22.6 vs 17.8 ms at E = 1 — profiles/r02/s11_sweep_round64.txt)
#ifndef GMB_KEY_BATCH
#define GMB_KEY_BATCH 4
#endif
#ifndef GMB_BLOCK_MINB
#define GMB_BLOCK_MINB 3                // resident CTAs per SM the register allocation must allow (3: 80 registers)
#endif
constexpr uint32_t kKeyBatch = GMB_KEY_BATCH; // table entries requested back to back

template <int KW, bool COUNT, typename OutT, bool EP, int MINB, int SIGMA>
__global__ void __launch_bounds__(kThreads, MINB) block_kernel(const MapLaunch L)
{
    // shared memory: step tables | jump-table starts | offsets | per-chain frame store | per-chain entries put aside
    extern __shared__ uint32_t smem[];
    const uint32_t n_start_words = (L.cx.B + 1) * kMaxSearches * kStartWords;
    uint32_t* steps_s = smem;
    uint32_t* starts_s = steps_s + align32(L.n_step_words);
    uint32_t* offs_s = starts_s + align32(n_start_words);
    uint32_t* frames_s = offs_s + align32(2 * (kMaxBlockKmers + 1));
    uint32_t* pend_s = frames_s + frame_store_words(L.E, L.cx.B, EP, SIGMA, true) * kThreads + threadIdx.x;
    for (uint32_t i = threadIdx.x; i < L.n_step_words; i += kThreads) steps_s[i] = L.cx.steps[i];
    for (uint32_t i = threadIdx.x; i < n_start_words; i += kThreads) starts_s[i] = reinterpret_cast<const uint32_t*>(L.cx.starts)[i];
    if (threadIdx.x == 0) {
#pragma unroll
        for (uint32_t i = 0; i <= kMaxBlockKmers; ++i) {
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
    using Frames = decltype(fr);

    const unsigned lane = threadIdx.x & 31u;
    OutT* __restrict__ out = static_cast<OutT*>(L.out);
    const uint32_t B = cx.B, K = cx.K;
    const uint2* __restrict__ keys = L.keylist;
    FetchStats fetches{};
    unsigned long long lut_reads = 0;
    Chain<KW, SIGMA> st;
    st.has_n = false; st.acc = 0; st.files = 0; st.var = 0; st.sub = 0; st.nsub = 1;

    for (;;) {
        // ---- the warp's next 32 blocks -------------------------------------------------------------------------
        unsigned long long cid = 0;
        if (lane == 0) cid = atomicAdd(L.work_counter, 1ull);
        cid = __shfl_sync(0xffffffffu, cid, 0);
        if (cid >= L.n_chunks) break;
        uint32_t rl = 0, rh = L.n_ranges; // largest r with chunk_prefix[r] <= cid (uniform loads)
        while (rh - rl > 1) {
            const uint32_t mid = (rl + rh) >> 1;
            if (__ldg(L.chunk_prefix + mid) <= cid) rl = mid; else rh = mid;
        }
        const unsigned long long nb = __ldg(L.range_begin + rl) + (cid - __ldg(L.chunk_prefix + rl)) * L.chunk;
        unsigned long long ne = nb + L.chunk;
        const unsigned long long re = __ldg(L.range_end + rl);
        if (ne > re) ne = re;
        const unsigned long long j = nb + (unsigned long long)lane * B; // first position of this lane's block
        const uint32_t cnt = j < ne ? (uint32_t)(ne - j < B ? ne - j : B) : 0u;
        const uint32_t NL = K + cnt - 1;
        uint32_t nk = cnt ? L.key_n[cnt] : 0u;
        const uint32_t koff = cnt ? L.key_off[cnt] : 0u;
        if (cnt) {
            st.cnt = cnt;
            load_pattern(st.pat, L.text, L.nmask, L.text_begin + j, NL);
            const uint32_t per = EP ? 3u : 1u;
            for (uint32_t w = 0; w < cnt * per; ++w) fr.cset(kLeafWords + w, 0u);
            if constexpr (SIGMA == 5) {
                // an N in the common infix is an N in every window of the block: nothing to search (windows with 1..E N
                // get their counts from the N pass, the others have none)
                st.has_n = st.pat.has_n();
                if (st.has_n && st.pat.has_n_in(cnt - 1u, K - cnt + 1u)) nk = 0u;
            }
        }

        for (uint32_t strand = 0; strand < cx.n_strands; ++strand) {
            if (cnt) {
                st.strand = strand;
                if (strand == 1) st.pat.reverse_complement(NL);
            }
            const uint32_t nk_max = __reduce_max_sync(0xffffffffu, nk);
            for (uint32_t g0 = 0; g0 < nk_max; g0 += kRound) {
                // ---- key phase: the next kRound keys of every lane's block ----------------------------------------
                uint32_t n_pend = 0, over = 0; // entries put aside; keys whose entry did not fit (read again below)
                const uint32_t g1 = g0 + kRound < nk ? g0 + kRound : nk;
                for (uint32_t g = g0; g < g1; g += kKeyBatch) {
                    uint32_t e0[kKeyBatch], e1[kKeyBatch], e2[kKeyBatch], e3[kKeyBatch], key[kKeyBatch], meta[kKeyBatch];
#pragma unroll
                    for (uint32_t u = 0; u < kKeyBatch; ++u) {
                        e1[u] = 0u;
                        if (g + u < g1) {
                            const uint2 ke = __ldg(keys + koff + g + u);
                            meta[u] = ke.y;
                            const SearchStart& S = cx.starts[cnt * kMaxSearches + (ke.y & 7u)];
                            key[u] = st.pat.bits(S.a, S.d) ^ ke.x;
                            asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(e0[u]), "=r"(e1[u]), "=r"(e2[u]), "=r"(e3[u]) : "l"(S.full + key[u]));
                            if (COUNT) ++lut_reads;
                        }
                    }
#pragma unroll
                    for (uint32_t u = 0; u < kKeyBatch; ++u) {
                        if (e1[u] == 0u) continue; // no such string in the text (or past the end of the key list)
                        if (e1[u] & kLocated) {
                            st.s = meta[u] & 7u;
                            const SearchStart& S = cx.starts[cnt * kMaxSearches + st.s];
                            verify_located_key<KW, EP, true, SIGMA, Frames>(st, fr, cx, COUNT ? &fetches : nullptr, S, key[u],
                                                                        (meta[u] >> 8) & 1u, e0[u], e2[u], e3[u]);
                        } else if (n_pend < kPendSlots) {
                            uint32_t* p = pend_s + n_pend * kPendWords * kThreads;
                            p[0] = e0[u]; p[kThreads] = e1[u]; p[2 * kThreads] = e2[u]; p[3 * kThreads] = g + u;
                            ++n_pend;
                        } else {
                            over |= 1u << (g + u - g0);
                        }
                    }
                }
                // ---- walk phase: the subtrees below the entries put aside ------------------------------------------
                bool walking = false;
                for (;;) {
                    if (!walking && (n_pend | over)) {
                        uint32_t gi;
                        if (n_pend) {
                            --n_pend;
                            const uint32_t* p = pend_s + n_pend * kPendWords * kThreads;
                            st.lo_r = p[0]; st.size = p[kThreads]; st.lo_f = p[2 * kThreads]; gi = p[3 * kThreads];
                            const uint2 ke = __ldg(keys + koff + gi);
                            st.s = ke.y & 7u; st.e = (ke.y >> 4) & 7u;
                            st.t = cx.starts[cnt * kMaxSearches + st.s].d;
                        } else {
                            gi = g0 + lowest_bit_index(over);
                            over &= over - 1u;
                            const uint2 ke = __ldg(keys + koff + gi);
                            st.s = ke.y & 7u; st.e = (ke.y >> 4) & 7u;
                            const SearchStart& S = cx.starts[cnt * kMaxSearches + st.s];
                            uint32_t pad;
                            jump_lookup(S, st.pat.bits(S.a, S.d) ^ ke.x, st.lo_f, st.lo_r, st.size, pad);
                            st.t = S.d;
                            if (COUNT) ++lut_reads;
                        }
                        st.lvmask = 0; st.win = kNoWin; st.leaf_e = 0; st.thin = false;
                        walking = true;
                    }
                    if (!__any_sync(0xffffffffu, walking)) break;
                    if (walking)
                        walking = chain_step<KW, EP, true, SIGMA, Frames, false, true>(st, fr, cx, COUNT ? &fetches : nullptr, nullptr);
                }
            }
        }
        if (cnt)
            for (uint32_t w = 0; w < cnt; ++w) out[j + w] = (OutT)chain_result<KW, EP, true, SIGMA>(st, fr, cx, w);
    }
    if (COUNT) {
        unsigned long long v[kCounterWords] = {fetches.total, lut_reads, fetches.by_size[0], fetches.by_size[1], fetches.by_size[2],
                                               fetches.by_size[3], fetches.by_size[4], fetches.by_size[5], fetches.by_size[6],
                                               fetches.by_size[7], fetches.thin_paths, fetches.iterations + lut_reads, fetches.located,
                                               fetches.text_reads};
#pragma unroll
        for (int k = 0; k < kCounterWords; ++k) {
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if (lane == 0 && v[k]) atomicAdd(L.fetch_counter + k, v[k]);
        }
    }
}

template <int KW, bool COUNT, typename OutT, bool EP, int MINB, int SIGMA = 4>
cudaError_t launch_blk(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    auto kern = block_kernel<KW, COUNT, OutT, EP, MINB, SIGMA>;
    const size_t smem = block_kernel_smem_bytes(L.n_step_words, L.E, L.cx.B, EP, SIGMA);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned long long grid = (unsigned long long)sm_count * per_sm; // persistent: every resident CTA slot of every SM
    const unsigned long long want = (L.n_chunks * 32 + kThreads - 1) / kThreads; // one warp per chunk at most
    if (want < grid) grid = want ? want : 1;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(L);
    return cudaGetLastError();
}

template <int KW, int MINB>
cudaError_t launch_blk_kw(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.sigma == 5) { // (never with --exclude-pseudo: block_kernel_applies)
        if (L.value_bits == 16)
            return L.count_fetches ? launch_blk<KW, true, uint16_t, false, MINB, 5>(L, sm_count, stream)
                                   : launch_blk<KW, false, uint16_t, false, MINB, 5>(L, sm_count, stream);
        return L.count_fetches ? launch_blk<KW, true, uint8_t, false, MINB, 5>(L, sm_count, stream)
                               : launch_blk<KW, false, uint8_t, false, MINB, 5>(L, sm_count, stream);
    }
    if (L.exclude_pseudo) // (the counting instantiation is not built for --exclude-pseudo, as in map_kernel.cu)
        return L.value_bits == 16 ? launch_blk<KW, false, uint16_t, true, MINB>(L, sm_count, stream)
                                  : launch_blk<KW, false, uint8_t, true, MINB>(L, sm_count, stream);
    if (L.value_bits == 16)
        return L.count_fetches ? launch_blk<KW, true, uint16_t, false, MINB>(L, sm_count, stream)
                               : launch_blk<KW, false, uint16_t, false, MINB>(L, sm_count, stream);
    return L.count_fetches ? launch_blk<KW, true, uint8_t, false, MINB>(L, sm_count, stream)
                           : launch_blk<KW, false, uint8_t, false, MINB>(L, sm_count, stream);
}

} // namespace

size_t block_kernel_smem_bytes(uint32_t n_step_words, uint32_t E, uint32_t B, bool ep, uint32_t sigma)
{
    const size_t tables = align32(n_step_words) + align32((B + 1) * kMaxSearches * kStartWords) + align32(2 * (kMaxBlockKmers + 1));
    return (tables + ((size_t)frame_store_words(E, B, ep, (int)sigma, true) + kPendSlots * kPendWords) * kThreads) * sizeof(uint32_t);
}

bool block_kernel_applies(const MapLaunch& L)
{
    // (E = 3: walks dominate and the general kernel, whose lanes refill one by one, is still the faster of the two, 16.5 vs
    // 18.7 ms; E = 4, with seven searches and 4000 keys per position, is 2.7x faster here: profiles/r02/s9_sweep_*.txt)
    return L.keylist != nullptr && (L.sigma == 4 || (L.sigma == 5 && L.cx.skip_n && !L.exclude_pseudo)) && L.E >= 1 && (L.E != 3 || L.force_block_kernel) && L.cx.K + L.cx.B - 1 <= 64 && L.cx.loc_rows == nullptr && L.loc_off == nullptr &&
           !(L.exclude_pseudo && L.count_fetches);
}

cudaError_t launch_block_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.n_work == 0) return cudaSuccess;
    if (L.chunk != 32u * L.cx.B) return cudaErrorInvalidValue; // one block per lane
    return L.cx.K + L.cx.B - 1 <= 32 ? launch_blk_kw<1, GMB_BLOCK_MINB>(L, sm_count, stream) : launch_blk_kw<2, GMB_BLOCK_MINB>(L, sm_count, stream);
}

} // namespace gmb
