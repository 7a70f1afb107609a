// jump_table.cu — builds the jump tables that replace the first d error-free steps of every search.
//
// No counterpart in the reference (which walks every pattern from the root); the idea is the classic
// k-mer lookup table of FM-index mappers, sized for HBM: at depth 15 (3 Gbp default) the table has
// 4^15 entries x 12 bytes = 12.9 GB, about one expected occurrence per entry.
#include "jump_table.cuh"

namespace gmb {

namespace {

template <int SIGMA>
__global__ void k_jump_level(const MapCtx cx, uint32_t d, const JtEntry* __restrict__ prev_uni,
                             const uint32_t* __restrict__ prev_lof, JtEntry* __restrict__ out_uni,
                             uint32_t* __restrict__ out_lof, JtFull* __restrict__ out_full)
{
    const uint64_t n = 1ull << (2 * d);
    const uint64_t key = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= n) return;
    Node par;
    if (d == 1) {
        par.lo_f = 0; par.lo_r = 0; par.size = cx.n_bwt;
    } else {
        const uint64_t pk = key & ((1ull << (2 * (d - 1))) - 1ull);
        const JtEntry e = prev_uni[pk];
        par.lo_f = prev_lof ? prev_lof[pk] : 0u; par.lo_r = e.lo_r; par.size = e.size;
    }
    const Node m = extend_right<SIGMA>(par, (uint32_t)(key >> (2 * (d - 1))), cx); // keys are A,C,G,T only
    if (out_full) {
        JtFull f;
        f.lo_r = m.lo_r; f.size = m.size; f.lo_f = m.lo_f; f.pad = 0;
        out_full[key] = f;
        return;
    }
    JtEntry o;
    o.lo_r = m.lo_r; o.size = m.size;
    out_uni[key] = o;
    if (out_lof) out_lof[key] = m.lo_f;
}

// Text pass over a finished level of 16-byte entries: the entry of every key that occurs exactly once is rewritten as
// a LOCATED entry (gmb_core.h: JtFull) — text position and 2 x kCtx characters of context instead of two one-row
// intervals.  One thread per text position; a key with one occurrence has one writer.  Positions whose key window
// leaves its sequence are not occurrences (the index never matches across a sentinel); keys within kLocateMargin of
// either end of the text keep their intervals, so verify_located may read around the occurrence without range checks.
__global__ void k_locate_singletons(const uint64_t* __restrict__ text, const uint64_t* __restrict__ nmask, uint64_t n_text,
                                    const uint32_t* __restrict__ seq_start, uint32_t n_seq, uint32_t d, JtFull* __restrict__ full)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + kLocateMargin;
    if (q + d + kLocateMargin > n_text) return;
    if (nmask) { // Dna5: a window with an N is no key, and a key is only located if its context holds no N either
        const uint64_t b = q - kCtx, len = d + 2 * kCtx; // <= 48 positions: at most two mask words
        const uint64_t w0 = nmask[b >> 6], w1 = nmask[(b >> 6) + 1];
        const uint32_t sh = (uint32_t)(b & 63u);
        const uint64_t bits = sh ? (w0 >> sh) | (w1 << (64u - sh)) : w0;
        if (bits & ((1ull << len) - 1ull)) return;
    }
    auto chars = [&](uint64_t p, uint32_t len) { // len <= 16 characters starting at text position p
        const uint64_t w = text[p >> 5], w2 = text[(p >> 5) + 1];
        const uint32_t sh = 2u * (uint32_t)(p & 31u);
        const uint64_t v = sh ? (w >> sh) | (w2 << (64u - sh)) : w;
        return (uint32_t)(v & ((1ull << (2u * len)) - 1ull));
    };
    const uint32_t key = chars(q, d);
    JtFull* e = full + key;
    if (e->size != 1u) return;
    uint32_t a = 0, b = n_seq; // largest s with limits[s] <= q (limits[s] = seq_start[s] - s)
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if ((uint64_t)seq_start[mid] - mid <= q) a = mid; else b = mid;
    }
    if (q + d > (uint64_t)seq_start[a + 1] - (a + 1)) return; // the window crosses into the next sequence
    JtFull o;
    o.lo_r = (uint32_t)q;
    o.size = kLocated | 1u;
    o.lo_f = chars(q + d, kCtx);
    o.pad = chars(q - kCtx, kCtx);
    *e = o;
}

} // namespace

cudaError_t locate_jump_singletons(const uint64_t* text, const uint64_t* nmask, uint64_t n_text, const uint32_t* seq_start, uint32_t n_seq,
                                   uint32_t d, JtFull* full, cudaStream_t stream)
{
    if (n_text < 2ull * kLocateMargin + d + 1) return cudaSuccess; // too small a text: nothing gets located
    const uint64_t n = n_text - 2ull * kLocateMargin - d + 1;
    const unsigned threads = 256;
    k_locate_singletons<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(text, nmask, n_text, seq_start, n_seq, d, full);
    return cudaGetLastError();
}

cudaError_t build_jump_level(const MapCtx& cx, uint32_t sigma, uint32_t d, const JtEntry* prev_uni, const uint32_t* prev_lof,
                             JtEntry* out_uni, uint32_t* out_lof, JtFull* out_full, cudaStream_t stream)
{
    const uint64_t n = 1ull << (2 * d);
    const unsigned threads = 256;
    const unsigned long long blocks = (n + threads - 1) / threads;
    if (sigma == 5) k_jump_level<5><<<(unsigned)blocks, threads, 0, stream>>>(cx, d, prev_uni, prev_lof, out_uni, out_lof, out_full);
    else k_jump_level<4><<<(unsigned)blocks, threads, 0, stream>>>(cx, d, prev_uni, prev_lof, out_uni, out_lof, out_full);
    return cudaGetLastError();
}

} // namespace gmb
