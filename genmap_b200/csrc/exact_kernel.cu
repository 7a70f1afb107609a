// exact_kernel.cu — the (K, 0)-frequency kernel: E = 0, entered through 16-byte jump-table entries (Dna4 and Dna5 indices;
// in a Dna5 index a k-mer with an N has no occurrence at E = 0, on either strand: src/algo.hpp:111-112).
//
// Replaces, for E = 0, the same reference path as map_kernel.cu (computeMappability<0> -> ... -> the exact search of
// src/find2_index_approx.hpp:303-369 for a single block, both strands).  Why a kernel of its own: without errors a
// search is two table reads (the k-mer, its reverse complement) and, for the keys that occur more than once, a short
// run of single-symbol rank steps — no scheme, no backtracking, no frames.  The general state machine of gmb_core.h
// spends ~260 thread-instructions per pass on bookkeeping that is dead here and was bound by instruction issue
// (profiles/r02/s1_ncu_counters_e0.csv: 62 % of the issue slots, 15.7 active lanes); this one is a straight line:
//   * one thread per k-mer start, a warp takes 128 consecutive positions per global atomic: pattern loads and result
//     stores of a warp are coalesced;
//   * both table entries are requested before either is used (two requests in flight per thread);
//   * LOCATED entries (gmb_core.h: JtFull — the key occurs once in the text) end the search at the table read: the
//     forward key's only occurrence is the query itself, the reverse key's is compared with the entry's context
//     characters (K <= d + 16: no memory access) or with the packed text;
//   * other entries are walked with single-symbol ranks on the 32-byte rank blocks (one request per boundary, two
//     only when the interval straddles blocks), the forward strand stopping as soon as the interval is down to the
//     query's own row.
// Counts, fetches and table reads are identical to the general kernel's (tests/test_gpu_parity.py compares them with
// the host-compiled state machine).
#include "map_kernel.cuh"

namespace gmb {

namespace {

constexpr int kThreadsE0 = 256;
#ifndef GMB_EXACT_MINB
#define GMB_EXACT_MINB 4 // resident CTAs per SM the register allocation must allow
#endif

struct ExactCounters { unsigned long long fetches, lut, located, text_reads, steps; };

// rows [lo, lo + size) of SA(T') after matching P[t], t = from .. K-1, rightwards; stop_at_one: the forward strand may
// stop as soon as one row is left (it is the query's own)
template <int KW, bool COUNT, int SIGMA>
__device__ __forceinline__ uint32_t walk_exact(const Pattern<KW, SIGMA>& P, uint32_t from, uint32_t K, uint32_t lo, uint32_t size,
                                               bool stop_at_one, const void* __restrict__ blocks, const uint32_t* __restrict__ SP,
                                               const uint32_t (&C)[5], ExactCounters& ctr)
{
    for (uint32_t t = from; t < K && size != 0u; ++t) {
        if (stop_at_one && size == 1u) break;
        const uint32_t c = P.at(t); // (no N here: k-mers with an N never get this far)
        const uint32_t x = lo, y = lo + size;
        uint32_t r0, r1;
        if constexpr (SIGMA == 4) {
            const RankBlock* B = static_cast<const RankBlock*>(blocks);
            const uint32_t bx = x / kBlockBases, by = y / kBlockBases;
            if (COUNT) { ctr.fetches += 1u + (by != bx); ++ctr.steps; }
            const BlockRegs rbx = load_block(B + bx);
            BlockRegs rby = rbx;
            load_block_if(rby, B + by, by != bx);
            r0 = block_rank_one(rbx, x - bx * kBlockBases, x, c, SP);
            r1 = block_rank_one(rby, y - by * kBlockBases, y, c, SP);
        } else {
            const RankBlock5* B = static_cast<const RankBlock5*>(blocks);
            const uint32_t bx = x / kBlockBases5, by = y / kBlockBases5;
            if (COUNT) { ctr.fetches += 1u + (by != bx); ++ctr.steps; }
            const BlockRegs5 rbx = load_block5(B + bx);
            BlockRegs5 rby = rbx;
            load_block5_if(rby, B + by, by != bx);
            r0 = block_rank5_one(rbx, x - bx * kBlockBases5, x, c, SP);
            r1 = block_rank5_one(rby, y - by * kBlockBases5, y, c, SP);
        }
        size = r1 - r0;
        lo = (c == 0 ? C[0] : (c == 1 ? C[1] : (c == 2 ? C[2] : C[3]))) + r0;
    }
    return size;
}

template <int KW, bool COUNT, typename OutT, int SIGMA>
__global__ void __launch_bounds__(kThreadsE0, GMB_EXACT_MINB) exact_kernel(const MapLaunch L)
{
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t K = L.cx.K, d = L.e0_depth, maxv = L.cx.maxv;
    const JtFull* __restrict__ table = L.e0_table;
    const void* __restrict__ Brev = L.cx.blk[1];
    const uint32_t* __restrict__ SPrev = L.cx.sent[1];
    const uint32_t C[5] = {L.cx.C[0], L.cx.C[1], L.cx.C[2], L.cx.C[3], L.cx.C[4]};
    OutT* __restrict__ out = static_cast<OutT*>(L.out);
    const bool both = L.cx.n_strands > 1;
    ExactCounters ctr{};

    for (;;) {
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

        for (unsigned long long j = nb + lane; j < ne; j += 32) {
            Pattern<KW, SIGMA> pat, rc;
            load_pattern(pat, L.text, L.nmask, L.text_begin + j, K);
            if (SIGMA == 5 && pat.has_n()) { // an N never matches: no occurrence on either strand
                if (COUNT) ctr.lut += both ? 2u : 1u; // (counted like the general kernel counts its dead entries)
                out[j] = (OutT)0;
                continue;
            }
            rc = pat;
            rc.reverse_complement(K);
            // both table entries requested before either is used
            uint32_t f0, f1, f2, f3, r0 = 0, r1 = 0, r2 = 0, r3 = 0;
            {
                const JtFull* pf = table + pat.bits(0, d);
                asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(f0), "=r"(f1), "=r"(f2), "=r"(f3) : "l"(pf));
                if (both) {
                    const JtFull* pr = table + rc.bits(0, d);
                    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(pr));
                }
            }
            if (COUNT) ctr.lut += both ? 2u : 1u;
            uint32_t count = 0;
            // ---- forward strand: the query occurs in the text, so its key does ---------------------------------
            if (f1 & kLocated) {
                count = 1; // the key's only occurrence is the query itself
                if (COUNT) ++ctr.located;
            } else {
                const uint32_t n = walk_exact<KW, COUNT, SIGMA>(pat, d, K, f0, f1, true, Brev, SPrev, C, ctr);
                count = n; // 1 when the walk stopped at the query's own row
            }
            // ---- reverse strand ------------------------------------------------------------------------------------
            if (both && r1 != 0u) {
                if (r1 & kLocated) {
                    if (COUNT) ++ctr.located;
                    const uint32_t q = r0; // text position of the key's only occurrence
                    bool same;
                    if (K - d <= kCtx) { // the entry's right context holds the rest of the k-mer
                        const uint32_t rest = K - d;
                        const uint32_t want = rest ? rc.bits(d, rest) : 0u;
                        const uint32_t have = rest == 16u ? r2 : (r2 & ((1u << (2u * rest)) - 1u));
                        same = want == have;
                    } else {
                        if (COUNT) ++ctr.text_reads;
                        Pattern<KW, SIGMA> tp;
                        load_pattern(tp, L.cx.text, L.cx.nmask, (uint64_t)q, K);
                        same = !tp.has_n(); // (Dna5: a text N matches nothing)
#pragma unroll
                        for (int k = 0; k < KW; ++k) same = same && tp.w[k] == rc.w[k];
                    }
                    if (same) { // inside one sequence?  (the index never matches across a sentinel)
                        uint32_t a = 0, b = L.cx.n_seq;
                        while (b - a > 1) {
                            const uint32_t mid = (a + b) >> 1;
                            if ((uint64_t)__ldg(L.cx.seq_start + mid) - mid <= (uint64_t)q) a = mid; else b = mid;
                        }
                        if ((uint64_t)q + K <= (uint64_t)__ldg(L.cx.seq_start + a + 1) - (a + 1)) count += 1;
                    }
                } else {
                    count += walk_exact<KW, COUNT, SIGMA>(rc, d, K, r0, r1, false, Brev, SPrev, C, ctr);
                }
            }
            out[j] = (OutT)(count < maxv ? count : maxv);
        }
    }
    if (COUNT) {
        // the layout of the general kernel's counters (map_kernel_impl.cuh): [0] fetches, [1] table reads,
        // [11] passes (here: rank steps + table reads), [12] located entries, [13] text reads
        unsigned long long v[5] = {ctr.fetches, ctr.lut, ctr.steps + ctr.lut, ctr.located, ctr.text_reads};
        const int at[5] = {0, 1, 11, 12, 13};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if (lane == 0 && v[k]) atomicAdd(L.fetch_counter + at[k], v[k]);
        }
    }
}

template <int KW, bool COUNT, typename OutT, int SIGMA>
cudaError_t launch_e0(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    auto kern = exact_kernel<KW, COUNT, OutT, SIGMA>;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreadsE0, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned long long grid = (unsigned long long)sm_count * per_sm; // persistent: every resident CTA slot of every SM
    const unsigned long long want = (L.n_chunks * 32 + kThreadsE0 - 1) / kThreadsE0; // one warp per chunk at most
    if (want < grid) grid = want ? want : 1;
    kern<<<(unsigned)grid, kThreadsE0, 0, stream>>>(L);
    return cudaGetLastError();
}

template <int KW, int SIGMA>
cudaError_t launch_e0_kw(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.value_bits == 16)
        return L.count_fetches ? launch_e0<KW, true, uint16_t, SIGMA>(L, sm_count, stream) : launch_e0<KW, false, uint16_t, SIGMA>(L, sm_count, stream);
    return L.count_fetches ? launch_e0<KW, true, uint8_t, SIGMA>(L, sm_count, stream) : launch_e0<KW, false, uint8_t, SIGMA>(L, sm_count, stream);
}

} // namespace

bool exact_kernel_applies(const MapLaunch& L)
{
    return L.E == 0 && (L.sigma == 4 || L.sigma == 5) && !L.exclude_pseudo && L.cx.B == 1 && L.e0_table != nullptr && L.e0_depth >= 1 &&
           L.e0_depth < L.cx.K && L.cx.K <= 64 && L.cx.loc_rows == nullptr && L.loc_off == nullptr;
}

cudaError_t launch_exact_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.n_work == 0) return cudaSuccess;
    if (L.sigma == 5) return L.cx.K <= 32 ? launch_e0_kw<1, 5>(L, sm_count, stream) : launch_e0_kw<2, 5>(L, sm_count, stream);
    return L.cx.K <= 32 ? launch_e0_kw<1, 4>(L, sm_count, stream) : launch_e0_kw<2, 4>(L, sm_count, stream);
}

} // namespace gmb
