// gmb_core.h — rank arithmetic on the 32-byte rank blocks and the per-k-mer search state machine.
//
// One "chain" (a CUDA thread in map_kernel.cu) owns one k-mer start at a time and walks the
// bidirectional FM-index with the optimum-search-scheme bounds.  The recursion of the reference
// (_optimalSearchSchemeGM/-ChildrenGM/-ExactGM, src/find2_index_approx.hpp:223-428, plus the by-value
// iterator copies that act as its stack) is restated as an explicit state machine:
//   * every loop iteration expands exactly ONE node: two rank-block reads (one if both interval
//     ends fall into the same block) give the ranks of all four symbols at both ends, hence all
//     children and every `smaller` of index_fm_stree.h:256-278 / index_bifm_stree.h:46-58 at once;
//   * mismatching children are visited before the matching one, so the walk at error level e needs
//     exactly one saved frame: at most E frames per chain, kept in shared memory;
//   * a full-length node adds its interval size to the k-mer's count (countOccurrences,
//     src/algo.hpp:48,191), saturating at the value type's maximum.
// The same header is compiled for the host by the CPU tests (tests/hostsim) to debug the logic and to
// count rank-block fetches without a GPU; the product library only instantiates it in device code.
#pragma once
#include "gmb_layout.h"

namespace gmb {

struct Ranks { uint32_t a, c, g, t, s; };

// one rank block in registers: header + the two bit planes as six 32-symbol pieces each
constexpr int kPieces = 2 * kBlockWords; // 32-symbol pieces per plane
struct BlockRegs {
    uint32_t h[4];
    uint32_t p0[kPieces], p1[kPieces];
};

GMB_HD uint32_t popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

GMB_HD BlockRegs load_block(const RankBlock* p)
{
    BlockRegs b;
#if defined(__CUDA_ARCH__)
    // 32 bytes per 256-bit read-only load (LDG.E.256 on sm_100a): one load for a 32-byte block, two for 64.
    // L2::64B: without the hint a 256-bit load makes the memory system fetch the whole 128-byte line
    // (measured: 2.9 DRAM sectors per requested sector; 1.45 with the hint — profiles/r01/s5_ldflavor_ncu.csv)
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "l"(p));
    b.h[0] = r0; b.h[1] = r1; b.h[2] = r2; b.h[3] = r3;
    b.p0[0] = r4; b.p0[1] = r5; b.p1[0] = r6; b.p1[1] = r7;
    if constexpr (kBlockWords == 3) {
        uint32_t s0, s1, s2, s3, s4, s5, s6, s7;
        asm volatile("ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4), "=r"(s5), "=r"(s6), "=r"(s7)
                     : "l"(reinterpret_cast<const char*>(p) + 32));
        b.p0[kPieces - 4] = s0; b.p0[kPieces - 3] = s1; b.p1[kPieces - 4] = s2; b.p1[kPieces - 3] = s3;
        b.p0[kPieces - 2] = s4; b.p0[kPieces - 1] = s5; b.p1[kPieces - 2] = s6; b.p1[kPieces - 1] = s7;
    }
#else
    b.h[0] = p->cnt[0]; b.h[1] = p->cnt[1]; b.h[2] = p->cnt[2]; b.h[3] = p->sent;
    for (uint32_t k = 0; k < kBlockWords; ++k) {
        b.p0[2 * k] = (uint32_t)p->w[k][0]; b.p0[2 * k + 1] = (uint32_t)(p->w[k][0] >> 32);
        b.p1[2 * k] = (uint32_t)p->w[k][1]; b.p1[2 * k + 1] = (uint32_t)(p->w[k][1] >> 32);
    }
#endif
    return b;
}

// b = (pred ? *p : b): the load is predicated, not branched around, so that both blocks of a node
// expansion are requested back to back by every lane of the warp
GMB_HD void load_block_if(BlockRegs& b, const RankBlock* p, bool pred)
{
#if defined(__CUDA_ARCH__)
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t"
        "@q ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t}"
        : "+r"(b.h[0]), "+r"(b.h[1]), "+r"(b.h[2]), "+r"(b.h[3]), "+r"(b.p0[0]), "+r"(b.p0[1]), "+r"(b.p1[0]), "+r"(b.p1[1])
        : "l"(p), "r"((uint32_t)pred));
    if constexpr (kBlockWords == 3) {
        asm volatile(
            "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t"
            "@q ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];\n\t}"
            : "+r"(b.p0[kPieces - 4]), "+r"(b.p0[kPieces - 3]), "+r"(b.p1[kPieces - 4]), "+r"(b.p1[kPieces - 3]),
              "+r"(b.p0[kPieces - 2]), "+r"(b.p0[kPieces - 1]), "+r"(b.p1[kPieces - 2]), "+r"(b.p1[kPieces - 1])
            : "l"(p), "r"((uint32_t)pred));
    }
#else
    if (pred) b = load_block(p);
#endif
}

// true iff `x` holds for every lane of the warp that is executing this call (host: this one chain)
GMB_HD bool warp_all(bool x)
{
#if defined(__CUDA_ARCH__)
    return __all_sync(__activemask(), x) != 0;
#else
    return x;
#endif
}

// mask of the symbols of piece q (32 symbols) that lie before in-block offset r
GMB_HD uint32_t piece_mask(uint32_t r, int q)
{
    const int w = (int)r - 32 * q;
#if defined(__CUDA_ARCH__)
    uint32_t m;
    const int wc = w > 0 ? w : 0;
    asm("shl.b32 %0, 1, %1;" : "=r"(m) : "r"(wc)); // PTX shl clamps shift amounts > 31: 1 << 32 == 0
    return m - 1u;
#else
    return w <= 0 ? 0u : (w >= 32 ? ~0u : ((1u << w) - 1u));
#endif
}

// sentinel rows of this block before BWT position i (rare path: the block holds a sentinel)
GMB_HD uint32_t sentinels_in_block_before(const BlockRegs& b, uint32_t i, const uint32_t* sent_pos)
{
    const uint32_t s_before = b.h[3] >> 8, s_in = b.h[3] & 0xffu;
    uint32_t s = 0;
    for (uint32_t k = 0; k < s_in; ++k) s += sent_pos[s_before + k] < i;
    return s;
}

// ranks of all symbols at BWT position i = blk*kBlockBases + r  (rank_c(i) = #c in bwt[0,i))
GMB_HD Ranks block_rank(const BlockRegs& b, uint32_t r, uint32_t i, const uint32_t* sent_pos)
{
    uint32_t a = 0, c = 0, g = 0;
#pragma unroll
    for (int q = 0; q < kPieces; ++q) {
        const uint32_t m = piece_mask(r, q);
        const uint32_t x0 = b.p0[q], x1 = b.p1[q];
        a += popc32(~x0 & ~x1 & m);
        c += popc32(x0 & ~x1 & m);
        g += popc32(~x0 & x1 & m);
    }
    uint32_t s = 0;
    if (b.h[3] & 0xffu) s = sentinels_in_block_before(b, i, sent_pos);
    Ranks R;
    R.a = b.h[0] + a - s;
    R.c = b.h[1] + c;
    R.g = b.h[2] + g;
    R.s = (b.h[3] >> 8) + s;
    R.t = i - R.a - R.c - R.g - R.s;
    return R;
}

// rank of ONE symbol (exact-match steps): a third of the popcounts of block_rank
GMB_HD uint32_t block_rank_one(const BlockRegs& b, uint32_t r, uint32_t i, uint32_t sym, const uint32_t* sent_pos)
{
    const uint32_t k0 = (sym & 1u) ? 0u : ~0u, k1 = (sym & 2u) ? 0u : ~0u;
    uint32_t n = 0;
#pragma unroll
    for (int q = 0; q < kPieces; ++q) n += popc32((b.p0[q] ^ k0) & (b.p1[q] ^ k1) & piece_mask(r, q));
    const uint32_t s_before = b.h[3] >> 8;
    uint32_t base;
    if (sym == 3u) base = (i - r) - b.h[0] - b.h[1] - b.h[2] - s_before; // T is the derived counter
    else base = sym == 0u ? b.h[0] : (sym == 1u ? b.h[1] : b.h[2]);
    if (sym == 0u && (b.h[3] & 0xffu)) n -= sentinels_in_block_before(b, i, sent_pos);
    return base + n;
}

// ---- Dna5 (sigma = 5) rank block: three planes, 32 symbols, one 256-bit load -----------------------------
struct BlockRegs5 {
    uint32_t h[5];            // A, C, G, T before the block; sent
    uint32_t p0, p1, p2;
};

GMB_HD BlockRegs5 load_block5(const RankBlock5* p)
{
    BlockRegs5 b;
#if defined(__CUDA_ARCH__)
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(p));
    b.h[0] = r0; b.h[1] = r1; b.h[2] = r2; b.h[3] = r3; b.h[4] = r4;
    b.p0 = r5; b.p1 = r6; b.p2 = r7;
#else
    for (int c = 0; c < 4; ++c) b.h[c] = p->cnt[c];
    b.h[4] = p->sent;
    b.p0 = p->plane[0]; b.p1 = p->plane[1]; b.p2 = p->plane[2];
#endif
    return b;
}

// b = (pred ? *p : b), predicated like load_block_if
GMB_HD void load_block5_if(BlockRegs5& b, const RankBlock5* p, bool pred)
{
#if defined(__CUDA_ARCH__)
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t"
        "@q ld.global.nc.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t}"
        : "+r"(b.h[0]), "+r"(b.h[1]), "+r"(b.h[2]), "+r"(b.h[3]), "+r"(b.h[4]), "+r"(b.p0), "+r"(b.p1), "+r"(b.p2)
        : "l"(p), "r"((uint32_t)pred));
#else
    if (pred) b = load_block5(p);
#endif
}

// ranks of all symbols (c[4] = N) at BWT position i = blk*32 + r
struct Ranks5 { uint32_t c[5]; uint32_t s; };
GMB_HD Ranks5 block_rank5(const BlockRegs5& b, uint32_t r, uint32_t i, const uint32_t* sent_pos)
{
    const uint32_t m = piece_mask(r, 0);
    const uint32_t x0 = b.p0, x1 = b.p1, x2 = b.p2;
    const uint32_t a = popc32(~x0 & ~x1 & ~x2 & m);
    const uint32_t c = popc32(x0 & ~x1 & m);
    const uint32_t g = popc32(~x0 & x1 & m);
    const uint32_t t = popc32(x0 & x1 & m);
    const uint32_t s_before = b.h[4] >> 8, s_in = b.h[4] & 0xffu;
    uint32_t s = 0;
    for (uint32_t k = 0; k < s_in; ++k) s += sent_pos[s_before + k] < i;
    Ranks5 R;
    R.c[0] = b.h[0] + a - s;
    R.c[1] = b.h[1] + c;
    R.c[2] = b.h[2] + g;
    R.c[3] = b.h[3] + t;
    R.s = s_before + s;
    R.c[4] = i - R.c[0] - R.c[1] - R.c[2] - R.c[3] - R.s;
    return R;
}

// rank of ONE of A, C, G, T at BWT position i = blk*32 + r of a Dna5 block (exact-match steps)
GMB_HD uint32_t block_rank5_one(const BlockRegs5& b, uint32_t r, uint32_t i, uint32_t sym, const uint32_t* sent_pos)
{
    const uint32_t m = piece_mask(r, 0);
    const uint32_t k0 = (sym & 1u) ? b.p0 : ~b.p0, k1 = (sym & 2u) ? b.p1 : ~b.p1;
    uint32_t n = popc32(k0 & k1 & ~b.p2 & m);
    if (sym == 0u) { // sentinel rows are stored as code 0
        const uint32_t s_before = b.h[4] >> 8, s_in = b.h[4] & 0xffu;
        for (uint32_t k = 0; k < s_in; ++k) n -= sent_pos[s_before + k] < i;
    }
    return (sym == 0u ? b.h[0] : (sym == 1u ? b.h[1] : (sym == 2u ? b.h[2] : b.h[3]))) + n;
}

// ---- pattern -----------------------------------------------------------------------------------------
GMB_HD uint64_t reverse_groups64(uint64_t x) // reverse the order of the 32 two-bit groups
{
#if defined(__CUDA_ARCH__)
    x = __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    x = __builtin_bswap64(x);
#endif
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

GMB_HD uint32_t reverse_bits32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(x);
#endif
}

// KW words of 32 two-bit characters; SIGMA == 5 adds one N-mask bit per character (nm[k] covers w[k])
template <int KW, int SIGMA = 4>
struct Pattern {
    uint64_t w[KW];
    uint32_t nm[SIGMA == 5 ? KW : 1];
    GMB_HD uint32_t at(uint32_t i) const
    {
        if (SIGMA == 5) {
            const uint32_t nbit = KW == 1 ? (nm[0] >> i) & 1u : (nm[i >> 5] >> (i & 31)) & 1u;
            if (nbit) return 4u; // N (src/algo.hpp:111-112: a pattern N never matches)
        }
        if (KW == 1) return (uint32_t)(w[0] >> (2 * i)) & 3u;
        if (KW == 2) return (uint32_t)((i < 32 ? w[0] : w[KW - 1]) >> (2 * (i & 31))) & 3u; // select: stays in registers
        return (uint32_t)(w[i >> 5] >> (2 * (i & 31))) & 3u;
    }
    // does any of the d <= 16 characters starting at offset a equal N?
    GMB_HD bool has_n(uint32_t a, uint32_t d) const
    {
        if constexpr (SIGMA == 5) {
            const uint32_t m = (1u << d) - 1u; // d <= 16
            if (KW == 1) return ((nm[0] >> a) & m) != 0;
            const uint32_t wi = a >> 5, sh = a & 31u;
            uint32_t v = nm[wi] >> sh;
            if (sh && wi + 1 < (uint32_t)KW) v |= nm[wi + 1] << (32u - sh);
            return (v & m) != 0;
        }
        return false;
    }
    // does any of the characters [a, a + d) equal N?  (any d)
    GMB_HD bool has_n_in(uint32_t a, uint32_t d) const
    {
        if constexpr (SIGMA == 5) {
#pragma unroll
            for (int k = 0; k < KW; ++k) {
                const uint32_t lo = 32u * (uint32_t)k, hi = lo + 32u; // the characters of word k
                const uint32_t b = a > lo ? a : lo, e = a + d < hi ? a + d : hi;
                if (b < e) {
                    const uint32_t len = e - b, m = (len == 32u ? ~0u : ((1u << len) - 1u)) << (b - lo);
                    if (nm[k] & m) return true;
                }
            }
        }
        return false;
    }
    GMB_HD bool has_n() const
    {
        uint32_t any = 0;
        if constexpr (SIGMA == 5)
            for (int k = 0; k < KW; ++k) any |= nm[k];
        return any != 0;
    }
    // the d <= 16 characters starting at offset a as an integer, character a in the low bits
    GMB_HD uint32_t bits(uint32_t a, uint32_t d) const
    {
        uint64_t v;
        if (KW == 1) {
            v = w[0] >> (2 * a);
        } else if (KW == 2) {
            const uint32_t sh = 2 * (a & 31);
            v = a < 32 ? ((w[0] >> sh) | (sh ? w[KW - 1] << (64 - sh) : 0ull)) : (w[KW - 1] >> sh);
        } else {
            const uint32_t wi = a >> 5, sh = 2 * (a & 31);
            v = w[wi] >> sh;
            if (sh && wi + 1 < (uint32_t)KW) v |= w[wi + 1] << (64 - sh);
        }
        return (uint32_t)(v & ((1ull << (2 * d)) - 1ull));
    }
    // in place: the reverse complement of the K-character pattern (src/algo.hpp:284-305)
    GMB_HD void reverse_complement(uint32_t K)
    {
        if constexpr (SIGMA == 5) { // N stays N: reverse the mask over the K characters (bit i -> bit K-1-i)
            if constexpr (KW == 1) {
                nm[0] = reverse_bits32(nm[0]) >> (32u - K);
            } else {
                uint32_t y[KW + 1];
#pragma unroll
                for (int k = 0; k < KW; ++k) y[k] = reverse_bits32(nm[KW - 1 - k]);
                y[KW] = 0;
                const uint32_t sft = 32u * KW - K, ws = sft >> 5, bs = sft & 31u; // drop the 32*KW - K unused low bits
#pragma unroll
                for (int k = 0; k < KW; ++k) {
                    const uint32_t a = k + ws;
                    uint32_t v = a <= (uint32_t)KW ? y[a < (uint32_t)KW ? a : KW] >> bs : 0u;
                    if (bs && a + 1 <= (uint32_t)KW) v |= y[a + 1] << (32u - bs);
                    nm[k] = v;
                }
            }
        }
        if constexpr (KW == 1) { // registers only
            w[0] = reverse_groups64(~w[0]) >> (64u - 2u * K);
        } else {
        uint64_t y[KW + 1];
#pragma unroll
        for (int k = 0; k < KW; ++k) y[k] = reverse_groups64(~w[KW - 1 - k]);
        y[KW] = 0;
        const uint32_t s = 2u * (32u * KW - K), ws = s >> 6, bs = s & 63u;
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            const uint32_t a = k + ws;
            uint64_t v = a <= (uint32_t)KW ? y[a] >> bs : 0ull;
            if (bs && a + 1 <= (uint32_t)KW) v |= y[a + 1] << (64 - bs);
            w[k] = v;
        }
        // clear the bits above 2K
        const uint32_t top = K >> 5, rem = K & 31u;
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            if ((uint32_t)k > top || ((uint32_t)k == top && rem == 0)) w[k] = 0;
            else if ((uint32_t)k == top) w[k] &= (1ull << (2 * rem)) - 1ull;
        }
        }
    }
};

// 2-bit packed text (32 bases per uint64, base i in bits 2*(i&31)) -> the K bases starting at `pos`;
// nmask (Dna5 only): bit i of the N mask, 64 positions per uint64
template <int KW, int SIGMA>
GMB_HD void load_pattern(Pattern<KW, SIGMA>& p, const uint64_t* text, const uint64_t* nmask, uint64_t pos, uint32_t K)
{
    if constexpr (SIGMA == 5) {
        for (int k = 0; k < KW; ++k) {
            uint32_t v = 0;
            if (32u * k < K) {
                const uint64_t q = pos + 32u * k;
                const uint64_t lo = nmask[q >> 6], hi = nmask[(q >> 6) + 1];
                const uint32_t sh = (uint32_t)(q & 63);
                v = (uint32_t)(sh ? (lo >> sh) | (hi << (64 - sh)) : lo);
                const uint32_t left = K - 32u * k;
                if (left < 32) v &= (1u << left) - 1u;
            }
            p.nm[k] = v;
        }
    } else {
        p.nm[0] = 0;
    }
    uint64_t wi = pos >> 5;
    uint32_t sh = 2u * (uint32_t)(pos & 31);
#pragma unroll
    for (int k = 0; k < KW; ++k) {
        if (32u * k >= K) { p.w[k] = 0; continue; }
        uint64_t lo = text[wi + k], hi = text[wi + k + 1];
        uint64_t v = sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
        uint32_t left = K - 32u * k;
        if (left < 32) v &= (1ull << (2 * left)) - 1ull;
        p.w[k] = v;
    }
}

// ---- search context ------------------------------------------------------------------------------------
// Jump table of one search: every search starts with an error-free, rightwards run (U[0] = 0 in every
// scheme, first direction Rev: src/find2_index_approx.hpp:441); the node reached after its first `d`
// characters is looked up instead of walked.  key = those characters, the first one in the low bits.
// Beyond the error-free prefix a search may be entered once per ADMISSIBLE STRING of length d: the mismatches the
// scheme allows inside the key window are substituted into the key and every resulting key is read (`var`: the
// admissible sets of substituted key offsets, one byte each, 0xff = unused; a set of m offsets stands for 3^m
// keys).  One table read replaces the walk through the dense top of the trie (JumpPlan, gmb_host.h).
struct JtEntry { uint32_t lo_r, size; };
struct JtFull { uint32_t lo_r, size, lo_f, pad; }; // both intervals in one 16-byte entry: one memory request
// LOCATED entries (set by the text pass of jump_table.cu; in a Dna5 index only keys without an N in window or context):
// a key that occurs exactly ONCE in the text carries where — and what stands around it — instead of its two one-row
// intervals:
//     lo_r = q, the text position of the occurrence (concatenated text, no sentinels)
//     size = kLocated | 1
//     lo_f = the kCtx characters right of the key window, text[q+d+i] in bits 2i
//     pad  = the kCtx characters left of it,            text[q-kCtx+i] in bits 2i
// A search entered through such an entry has ONE candidate alignment: it is finished by comparing the needle with
// the text (verify_located) — from the 2*kCtx context characters when they cover the needle (no memory access at
// all), else from one read of the packed text — instead of walking the index through one-row intervals, one
// dependent rank-block fetch per character.  At 3 Gbp 69 % of the existing depth-16 entries are of this kind.
// (kLocated, kCtx, kLocateMargin: gmb_layout.h)
struct SearchStart {
    const JtEntry* uni;   // [4^d] interval in SA(T') + size            (nullptr: no table, start at the root)
    const uint32_t* lof;  // [4^d] start of the interval in SA(T)       (nullptr when never needed again or `full` is set)
    const JtFull* full;   // [4^d] both intervals                       (set instead of uni/lof when SA(T) is needed)
    const uint32_t* var;  // n_var sets of substituted key offsets
    uint32_t a;           // pattern offset of the first character of the key window
    uint32_t d;           // depth of the table
    uint32_t n_var;
    uint32_t set0;        // var[0], kept inline: the first entry needs no extra read
};

struct MapCtx {
    const void* blk[2];       // [0]: BWT of T (extend left), [1]: BWT of T' (extend right); RankBlock or RankBlock5
    const uint32_t* sent[2];
    uint32_t C[5];            // C[c] = #symbols smaller than base c (sentinels included); C[4] only for Dna5
    uint32_t n_bwt;
    // search tables (gmb_host.h: BlockTables), indexed by cnt = k-mers in the block (1..B):
    //   infix search : steps[p1_off[cnt] + search * (K - cnt + 1) + t]
    //   flank walks  : steps[fl_off[cnt] + window * (cnt - 1) + t]
    const uint32_t* steps;
    const uint32_t* p1_off;
    const uint32_t* fl_off;
    const SearchStart* starts; // [cnt * kMaxSearches + search]
    uint32_t K, B, E, n_search, n_strands, maxv;
    // --exclude-pseudo only (src/algo.hpp:351-361): locate every hit and count distinct FASTA files
    const uint32_t* sa;          // full suffix array of T
    const uint32_t* seq_start;   // n_seq + 1 sequence starts inside T
    const uint32_t* seq_to_file; // n_seq file ids (mappingSeqIdFile, src/mappability.hpp:230-250)
    uint32_t n_seq, own_file;
    uint64_t all_files;          // mask with one bit per FASTA file
    // locate instantiation only (csv output, src/algo.hpp:311-343): second pass writes the SA value of every
    // occurrence; nullptr in the first (counting) pass
    uint32_t* loc_rows;
    // located table entries (verify_located): the packed text, its N mask (Dna5 indices) and its length
    const uint64_t* text;
    const uint64_t* nmask;
    uint64_t n_text;
    // Dna5 indices: the searches never match a text N — the call counts the alignments to text windows WITHOUT an N,
    // which lets every search enter through substituted keys as on a Dna4 index; the alignments to the (few) text
    // windows with 1..E N are added afterwards by a pass of their own (capi.cu: NFix)
    uint32_t skip_n;
};

struct Node { uint32_t lo_f, lo_r, size; };

// Counters of the instrumented (count_fetches) instantiation: rank-block fetches in total and by the size of
// the interval being expanded (1, 2, 3-4, 5-8, 9-16, 17-32, 33-64, 65+), and the number of "thin paths"
// (maximal runs of expansions of one-row intervals: what a locate + text comparison could replace).
struct FetchStats {
    unsigned long long total;
    unsigned long long by_size[8];
    unsigned long long thin_paths;
    unsigned long long iterations; // passes through the state machine (expansions, table reads, window switches)
    unsigned long long located;    // searches finished by verify_located (table entry of a key that occurs once)
    unsigned long long text_reads; // ... of which needed a read of the packed text (the entry's context did not cover the needle)
};
constexpr int kFetchStatWords = 13;
GMB_HD uint32_t size_bucket(uint32_t n)
{
    if (n <= 1u) return 0u;
#if defined(__CUDA_ARCH__)
    const uint32_t b = 32u - (uint32_t)__clz((int)(n - 1u));
#else
    const uint32_t b = 32u - (uint32_t)__builtin_clz(n - 1u);
#endif
    return b < 7u ? b : 7u;
}
GMB_HD void count_fetch(FetchStats* f, uint32_t size, uint32_t n)
{
    if (f) { f->total += n; f->by_size[size_bucket(size)] += n; }
}

// Children of a node in one direction: for every symbol c of the alphabet the size n[c] of the child
// interval and its start l[c] in the ACTIVE index; oth0 = start of child 0 in the OTHER index (child c
// starts at oth0 + n[0] + ... + n[c-1]: `smaller` of index_fm_stree.h:256-278 with the sentinels first).
template <int SIGMA>
struct Children { uint32_t n[SIGMA], l[SIGMA], oth0; };

// Both rank blocks of the node are requested before either is used (one memory latency per expansion).
// `dir` selects the index: 1 = BWT of T' (extend right), 0 = BWT of T (extend left).
template <int SIGMA>
GMB_HD Children<SIGMA> expand_node(const MapCtx& cx, uint32_t dir, uint32_t x, uint32_t size, uint32_t z, FetchStats* fetches)
{
    Children<SIGMA> ch;
    const uint32_t y = x + size;
    const uint32_t* SP = dir ? cx.sent[1] : cx.sent[0];
    if constexpr (SIGMA == 4) {
        const RankBlock* B = static_cast<const RankBlock*>(dir ? cx.blk[1] : cx.blk[0]);
        const uint32_t bx = x / kBlockBases, by = y / kBlockBases;
        count_fetch(fetches, size, 1u + (by != bx));
        const BlockRegs rbx = load_block(B + bx);
        BlockRegs rby = rbx;
        load_block_if(rby, B + by, by != bx);
        const Ranks R0 = block_rank(rbx, x - bx * kBlockBases, x, SP);
        const Ranks R1 = block_rank(rby, y - by * kBlockBases, y, SP);
        ch.n[0] = R1.a - R0.a; ch.n[1] = R1.c - R0.c; ch.n[2] = R1.g - R0.g; ch.n[3] = R1.t - R0.t;
        ch.l[0] = cx.C[0] + R0.a; ch.l[1] = cx.C[1] + R0.c; ch.l[2] = cx.C[2] + R0.g; ch.l[3] = cx.C[3] + R0.t;
        ch.oth0 = z + (R1.s - R0.s);
    } else {
        const RankBlock5* B = static_cast<const RankBlock5*>(dir ? cx.blk[1] : cx.blk[0]);
        const uint32_t bx = x / kBlockBases5, by = y / kBlockBases5;
        count_fetch(fetches, size, 1u + (by != bx));
        const BlockRegs5 rbx = load_block5(B + bx);
        BlockRegs5 rby = rbx;
        load_block5_if(rby, B + by, by != bx);
        const Ranks5 R0 = block_rank5(rbx, x - bx * kBlockBases5, x, SP);
        const Ranks5 R1 = block_rank5(rby, y - by * kBlockBases5, y, SP);
#pragma unroll
        for (int c = 0; c < 5; ++c) { ch.n[c] = R1.c[c] - R0.c[c]; ch.l[c] = cx.C[c] + R0.c[c]; }
        ch.oth0 = z + (R1.s - R0.s);
    }
    return ch;
}

// select element c of a small array with compile-time indices only (keeps the array in registers)
template <int SIGMA>
GMB_HD uint32_t pick(const uint32_t (&v)[SIGMA], uint32_t c)
{
    uint32_t r = v[SIGMA - 1];
#pragma unroll
    for (int k = SIGMA - 2; k >= 0; --k) r = c == (uint32_t)k ? v[k] : r;
    return r;
}
// n[0] + ... + n[c-1]
template <int SIGMA>
GMB_HD uint32_t sum_below(const uint32_t (&n)[SIGMA], uint32_t c)
{
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < SIGMA - 1; ++k) r += c > (uint32_t)k ? n[k] : 0u;
    return r;
}

// P -> Pc on the bidirectional index (goDown(it, c, Rev()): index_bidirectional_stree.h:250-265)
template <int SIGMA>
GMB_HD Node extend_right(const Node& n, uint32_t c, const MapCtx& cx)
{
    Node m;
    m.lo_f = 0; m.lo_r = 0; m.size = 0;
    if (n.size == 0) return m;
    const Children<SIGMA> ch = expand_node<SIGMA>(cx, 1, n.lo_r, n.size, n.lo_f, nullptr);
    m.size = pick<SIGMA>(ch.n, c);
    m.lo_r = pick<SIGMA>(ch.l, c);
    m.lo_f = ch.oth0 + sum_below<SIGMA>(ch.n, c);
    return m;
}

// ---- one chain = one block of cnt <= B adjacent k-mer starts -------------------------------------------
// The cnt k-mers share the infix needle[cnt-1 .. K-1] (needle = the K+cnt-1 text characters they cover).
// As in the reference (computeMappabilitySingleBlock, src/algo.hpp:221-308) the infix is searched once
// with the search scheme; every infix hit with e errors is then completed, window by window, through the
// window's flank characters with the remaining E - e errors (the job of extend/approxSearch/extendExact,
// src/algo.hpp:26-218, here a plain bounded walk per window).  cnt = 1 is the one-k-mer-per-chain case.
constexpr uint32_t kNoWin = 0xffu;

template <int KW, int SIGMA = 4>
struct Chain {
    Pattern<KW, SIGMA> pat;    // the needle; on the reverse strand its reverse complement
    uint32_t lo_f, lo_r, size; // current node: [lo_f, lo_f+size) in SA(T), [lo_r, lo_r+size) in SA(T')
    uint32_t acc;              // B == 1: occurrences so far (saturating); B > 1: counts live in the frame store
    uint32_t t, e, s, strand;  // step, errors, search, strand of the current walk
    uint32_t lvmask;           // error levels holding a frame with pending children
    uint32_t cnt, win, leaf_e; // k-mers in this block; window being completed (kNoWin: infix search); errors of the infix hit
    uint64_t files;            // B == 1, --exclude-pseudo: FASTA files seen so far (one bit each)
    uint32_t pre_lo_f, pre_lo_r, pre_size, pre_ctx; // jump-table entry of (reverse strand, search 0), fetched early
    uint32_t ctx_l;            // located entry (size & kLocated): its left context; lo_r = text position, lo_f = right context
    bool has_n;                // Dna5: the needle contains N (then no occurrence is error-free)
    // locate instantiation: occurrences found so far per strand (not saturated) and where this k-mer's two
    // lists start in cx.loc_rows
    uint32_t occ_fwd, occ_rev;
    uint64_t loc_at_fwd, loc_at_rev;
    bool thin;                 // instrumented instantiation: the previous expansion was of a one-row interval
    uint32_t var, sub, nsub;   // entry into the current search: set of substituted key offsets, which of its 3^m keys, 3^m
};

// --exclude-pseudo: mark the FASTA file of every occurrence in SA rows [lo, lo+n)
// (getOccurrences -> CompressedSA::value, index_fm_compressed_sa.h:478-513, on the full SA kept in HBM;
// file of a sequence: src/algo.hpp:354-358).  Stops as soon as every file has been seen.
GMB_HD void ep_mark_rows(uint64_t& mask, uint32_t lo, uint32_t n, const MapCtx& cx)
{
    for (uint32_t r = 0; r < n && mask != cx.all_files; ++r) {
        const uint32_t pos = cx.sa[lo + r];
        uint32_t a = 0, b = cx.n_seq; // largest s with seq_start[s] <= pos
        while (b - a > 1) {
            const uint32_t mid = (a + b) >> 1;
            if (cx.seq_start[mid] <= pos) a = mid; else b = mid;
        }
        mask |= 1ull << cx.seq_to_file[a];
    }
}

// Frame store of a chain (shared memory on the device):
//   E mismatch frames x (2*SIGMA + 2) words: child lo in the active index [SIGMA], child sizes [SIGMA],
//                                            lo of child 0 in the other index, t | pending << 8
//   then kLeafWords for the infix hit being completed (lo_f, lo_r, size),
//   then, in the blocked instantiation, one counter per window (+ two words of file mask per window under
//   --exclude-pseudo).
// Accessors: set/get(level, word) for the mismatch frames, xset/xget(word) for the leaf frame; the per-window
// counters go through cget / cset / cadd (saturating add) / cor, which a kernel may point at ANOTHER chain's store
// and make atomic (block_kernel.cu: a lane walks a subtree for the chain that owns the block).
constexpr int kFrameWords = 10;  // SIGMA == 4
constexpr int kFrameWords5 = 12; // SIGMA == 5
constexpr int kLeafWords = 3;
GMB_HD constexpr uint32_t frame_words(int sigma) { return sigma == 5 ? kFrameWords5 : kFrameWords; }
GMB_HD uint32_t frame_store_words(uint32_t E, uint32_t B, bool ep, int sigma, bool blocked)
{
    return E * frame_words(sigma) + kLeafWords + (blocked ? B * (ep ? 3u : 1u) : 0u);
}

GMB_HD void jump_lookup(const SearchStart& S, uint32_t key, uint32_t& lo_f, uint32_t& lo_r, uint32_t& size, uint32_t& pad)
{
    pad = 0u;
#if defined(__CUDA_ARCH__)
    if (S.full) {
        asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo_r), "=r"(size), "=r"(lo_f), "=r"(pad) : "l"(S.full + key));
        return;
    }
    asm volatile("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(lo_r), "=r"(size) : "l"(S.uni + key));
    lo_f = 0u;
    if (S.lof) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(lo_f) : "l"(S.lof + key));
#else
    if (S.full) { lo_r = S.full[key].lo_r; size = S.full[key].size; lo_f = S.full[key].lo_f; pad = S.full[key].pad; return; }
    lo_r = S.uni[key].lo_r; size = S.uni[key].size;
    lo_f = S.lof ? S.lof[key] : 0u;
#endif
}

// the one-k-mer instantiation reads 8-byte entries (+ the separate SA(T) array when needed): its tables never hold
// 16-byte entries
GMB_HD void jump_lookup_lean(const SearchStart& S, uint32_t key, uint32_t& lo_f, uint32_t& lo_r, uint32_t& size)
{
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.nc.L2::64B.v2.u32 {%0,%1}, [%2];" : "=r"(lo_r), "=r"(size) : "l"(S.uni + key));
    lo_f = 0u;
    if (S.lof) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(lo_f) : "l"(S.lof + key));
#else
    lo_r = S.uni[key].lo_r; size = S.uni[key].size;
    lo_f = S.lof ? S.lof[key] : 0u;
#endif
}

// start the infix search number st.s on the current strand
// (BLK = false is the one-k-mer-per-chain instantiation: cnt == 1 is a compile-time fact there, so all the
// window bookkeeping disappears and the count stays in a register)
// (the one-k-mer instantiation, BLK = false, enters every search through its error-free prefix only: the host
// plans no substituted keys for it, which keeps the E = 0 path lean)
template <int KW, bool BLK, int SIGMA>
GMB_HD void chain_start(Chain<KW, SIGMA>& st, const MapCtx& cx, unsigned long long* lut_reads)
{
    if constexpr (!BLK) {
        const SearchStart& S = cx.starts[kMaxSearches + st.s];
        st.e = 0; st.lvmask = 0; st.win = kNoWin; st.leaf_e = 0; st.thin = false;
        if (S.uni == nullptr && S.full == nullptr) {
            st.lo_f = 0; st.lo_r = 0; st.size = cx.n_bwt; st.t = 0;
        } else {
            if (st.strand == 1 && st.s == 0) { st.lo_f = st.pre_lo_f; st.lo_r = st.pre_lo_r; st.size = st.pre_size; st.ctx_l = st.pre_ctx; }
            else if (S.set0 == kDeadVariant || (SIGMA == 5 && (st.pat.has_n(S.a, S.d) || (cx.E == 0 && st.has_n)))) { st.lo_f = 0; st.lo_r = 0; st.size = 0; } // N never matches (E = 0: anywhere in the k-mer)
            else if (S.full) jump_lookup(S, st.pat.bits(S.a, S.d), st.lo_f, st.lo_r, st.size, st.ctx_l);
            else jump_lookup_lean(S, st.pat.bits(S.a, S.d), st.lo_f, st.lo_r, st.size);
            st.t = S.d;
            if (lut_reads) *lut_reads += 1;
        }
        return;
    } else {
    const SearchStart S = cx.starts[st.cnt * kMaxSearches + st.s];
    st.e = 0; st.lvmask = 0; st.win = kNoWin; st.leaf_e = 0; st.thin = false;
    if (S.uni == nullptr && S.full == nullptr) {
        st.lo_f = 0; st.lo_r = 0; st.size = cx.n_bwt; st.t = 0; st.nsub = 1;
    } else {
#if defined(__CUDA_ARCH__)
        const uint32_t set = st.var == 0 ? S.set0 : __ldg(S.var + st.var);
#else
        const uint32_t set = st.var == 0 ? S.set0 : S.var[st.var];
#endif
        if (st.strand == 1 && st.s == 0 && set == 0xffffffffu) { st.lo_f = st.pre_lo_f; st.lo_r = st.pre_lo_r; st.size = st.pre_size; st.ctx_l = st.pre_ctx; st.nsub = 1; }
        else if (set == kDeadVariant || (SIGMA == 5 && st.pat.has_n(S.a, S.d))) { st.lo_f = 0; st.lo_r = 0; st.size = 0; st.nsub = 1; } // N never matches
        else {
            // substitute: offset p of the set gets one of the three other characters (XOR with 1..3), chosen by the
            // base-3 digits of st.sub
            uint32_t key = st.pat.bits(S.a, S.d), e = 0, n3 = 1;
            if (set != 0xffffffffu) {
                uint32_t q = st.sub;
#pragma unroll
                for (uint32_t k = 0; k < kMaxE; ++k) {
                    const uint32_t p = (set >> (8 * k)) & 0xffu;
                    if (p != 0xffu) { key ^= (1u + q % 3u) << (2u * p); q /= 3u; n3 *= 3u; ++e; }
                }
            }
            jump_lookup(S, key, st.lo_f, st.lo_r, st.size, st.ctx_l);
            st.e = e; st.nsub = n3;
        }
        st.t = S.d;
        if (lut_reads) *lut_reads += 1;
    }
    }
}

// st.pat (needle of K + cnt - 1 characters) and st.cnt are set by the caller
template <int KW, bool EP, bool BLK, int SIGMA, class Frames>
GMB_HD void chain_begin_block(Chain<KW, SIGMA>& st, Frames& fr, const MapCtx& cx, unsigned long long* lut_reads)
{
    st.acc = 0; st.s = 0; st.strand = 0; st.files = 0; st.occ_fwd = 0; st.occ_rev = 0; st.var = 0; st.sub = 0; st.nsub = 1;
    if (!BLK) st.cnt = 1;
    st.has_n = st.pat.has_n();
    if (BLK) {
        const uint32_t per = EP ? 3u : 1u;
        for (uint32_t w = 0; w < st.cnt * per; ++w) fr.cset(kLeafWords + w, 0u);
    }
    // the reverse strand's first jump-table entry does not depend on the forward search: request it now so
    // that its latency overlaps the forward strand instead of starting the reverse strand with a stall
    const SearchStart S0 = cx.starts[(BLK ? st.cnt : 1u) * kMaxSearches];
    if (cx.n_strands > 1 && (S0.uni != nullptr || S0.full != nullptr)) {
        Pattern<KW, SIGMA> rc = st.pat;
        rc.reverse_complement(cx.K + (BLK ? st.cnt : 1u) - 1);
        st.pre_ctx = 0;
        if (SIGMA == 5 && (rc.has_n(S0.a, S0.d) || (!BLK && cx.E == 0 && st.has_n))) { st.pre_lo_f = 0; st.pre_lo_r = 0; st.pre_size = 0; }
        else if (BLK || S0.full) jump_lookup(S0, rc.bits(S0.a, S0.d), st.pre_lo_f, st.pre_lo_r, st.pre_size, st.pre_ctx);
        else jump_lookup_lean(S0, rc.bits(S0.a, S0.d), st.pre_lo_f, st.pre_lo_r, st.pre_size);
    }
    chain_start<KW, BLK, SIGMA>(st, cx, lut_reads);
}

// result of window w (position j0 + w) once chain_step has returned false
template <int KW, bool EP, bool BLK, int SIGMA, class Frames>
GMB_HD uint32_t chain_result(const Chain<KW, SIGMA>& st, const Frames& fr, const MapCtx& cx, uint32_t w)
{
    if (!BLK) return st.acc;
    if (!EP) { const uint32_t v = fr.cget(kLeafWords + w); return v < cx.maxv ? v : cx.maxv; }
    const uint64_t m = (uint64_t)fr.cget(kLeafWords + st.cnt + 2 * w) | ((uint64_t)fr.cget(kLeafWords + st.cnt + 2 * w + 1) << 32);
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popcll(m);
#else
    return (uint32_t)__builtin_popcountll(m);
#endif
}

GMB_HD uint32_t lowest_bit_index(uint32_t m)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)m) - 1u;
#else
    return (uint32_t)__builtin_ctz(m);
#endif
}

GMB_HD uint32_t highest_bit_index(uint32_t m)
{
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)m);
#else
    return 31u - (uint32_t)__builtin_clz(m);
#endif
}

// add `n` occurrences (or, under --exclude-pseudo, the files of SA rows [lo, lo+n)) to window `w` of the strand
template <int KW, bool EP, bool BLK, int SIGMA, class Frames, bool LOC = false>
GMB_HD void chain_count(Chain<KW, SIGMA>& st, Frames& fr, const MapCtx& cx, uint32_t w, uint32_t lo, uint32_t n, bool own_only)
{
    if constexpr (LOC) {
        // csv lists (src/algo.hpp:327-343): the occurrences are SA rows [lo, lo+n) of T; counted per strand in
        // the first pass, written behind the ones already found in the second
        const uint32_t have = st.strand ? st.occ_rev : st.occ_fwd;
        if (cx.loc_rows) {
            const uint64_t at = (st.strand ? st.loc_at_rev : st.loc_at_fwd) + have;
            for (uint32_t r = 0; r < n; ++r) cx.loc_rows[at + r] = cx.sa[lo + r];
        }
        if (st.strand) st.occ_rev = have + n; else st.occ_fwd = have + n;
        return;
    }
    const uint32_t widx = !BLK ? 0u : (st.strand ? st.cnt - 1u - w : w); // reverse-strand windows run backwards (src/algo.hpp:304)
    if (EP) {
        const uint32_t at = kLeafWords + st.cnt + 2 * widx;
        const uint64_t before = !BLK ? st.files : (uint64_t)fr.cget(at) | ((uint64_t)fr.cget(at + 1) << 32);
        uint64_t m = before;
        if (own_only) m |= 1ull << cx.own_file;
        else ep_mark_rows(m, lo, n, cx);
        if (!BLK) st.files = m;
        else if (m != before) { fr.cor(at, (uint32_t)(m & ~before)); fr.cor(at + 1, (uint32_t)((m & ~before) >> 32)); }
    } else {
        if (!BLK) {
            const uint64_t sum = (uint64_t)st.acc + n;
            st.acc = sum < cx.maxv ? (uint32_t)sum : cx.maxv; // saturating (src/algo.hpp:48,191)
        } else {
            fr.cadd(kLeafWords + widx, n, cx.maxv);
        }
    }
}

// ---- located table entries: one candidate alignment, finished by comparison with the text ------------------------
// mismatch bits of the needle against the text in 2-bit spacing (bit 2i of word k: character 32k + i differs)
template <int KW>
GMB_HD uint32_t count_mismatches(const uint64_t (&mm)[KW], uint32_t b, uint32_t e) // characters [b, e) of the needle
{
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < KW; ++k) {
        const uint32_t w0 = 32u * (uint32_t)k;
        const uint32_t lo = b > w0 ? b - w0 : 0u, hi = e < w0 + 32u ? (e > w0 ? e - w0 : 0u) : 32u;
        if (lo < hi) {
            const uint64_t upto = hi == 32u ? ~0ull : ((1ull << (2u * hi)) - 1ull);
            const uint64_t m = mm[k] & upto & ~((1ull << (2u * lo)) - 1ull);
#if defined(__CUDA_ARCH__)
            c += (uint32_t)__popcll(m);
#else
            c += (uint32_t)__builtin_popcountll(m);
#endif
        }
    }
    return c;
}

// mark FASTA file `file` for window w of the strand (--exclude-pseudo)
template <int KW, bool BLK, int SIGMA, class Frames>
GMB_HD void chain_mark_file(Chain<KW, SIGMA>& st, Frames& fr, uint32_t w, uint32_t file)
{
    const uint32_t widx = !BLK ? 0u : (st.strand ? st.cnt - 1u - w : w);
    if (!BLK) { st.files |= 1ull << file; return; }
    const uint32_t at = kLeafWords + st.cnt + 2 * widx;
    if (file < 32u) fr.cor(at, 1u << file);
    else fr.cor(at + 1, 1u << (file - 32u));
}

// bit i of a 32-bit mask -> bit 2i of a 64-bit word (the spacing of the mismatch bits)
GMB_HD uint64_t spread_bits(uint32_t m)
{
    uint64_t x = m;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

// The current search of the chain was entered through a LOCATED entry (see JtFull): the key of the entry occurs once
// in the text, at position st.lo_r, so the search has one candidate alignment: needle[x] <-> text[q - a + x].  What the
// index walk would find below this node is decided here by direct comparison:
//   1. mismatch bits of the whole needle against the text — from the entry's 2 x kCtx context characters when they
//      cover the needle (no memory access), else from one read of the packed text;
//   2. more than E mismatches inside the infix: nothing (almost every chance hit ends here);
//   3. the steps of the search over the infix with these mismatches, under the scheme's bounds (the same test per step
//      as chain_step: an alignment is found by exactly one search of the scheme);
//   4. every window whose K characters hold at most E mismatches and lie inside one sequence counts one occurrence
//      (the index walk cannot leave a sequence: sentinels; here the sequence limits are checked).
// `key`: the table key the entry was read with (the needle's window with this entry's substitutions); own_key: it is
// the needle's own window, nothing substituted; q / ctx_r / ctx_l: the entry.  st.s, st.strand, st.cnt, st.pat are read.
// Dna5 indices: a pattern N never matches and a text N matches nothing (src/algo.hpp:111-112) — both are mismatch bits;
// the table builder locates a key only if neither its window nor its context holds an N, so the context path needs no
// text mask.
template <int KW, bool EP, bool BLK, int SIGMA, class Frames>
GMB_HD void verify_located_key(Chain<KW, SIGMA>& st, Frames& fr, const MapCtx& cx, FetchStats* fetches, const SearchStart& S,
                               uint32_t key, bool own_key, uint32_t q, uint32_t ctx_r, uint32_t ctx_l)
{
    const uint32_t K = cx.K, cnt = BLK ? st.cnt : 1u;
    const uint32_t Li = K - cnt + 1, NL = K + cnt - 1, E = cx.E;
    const uint32_t tab = (BLK ? cx.p1_off[cnt] : 0u) + st.s * Li;
    if (fetches) ++fetches->located;
    [[maybe_unused]] uint64_t tn[KW] = {}; // Dna5, skip_n: the text's N (a window holding one is not counted)
    if (st.strand == 0 && own_key && !(SIGMA == 5 && st.has_n)) {
        // the query's own window is an occurrence of its own key, and the key occurs once: this is the query itself
        // (a needle with an N outside the key window is compared like any other candidate: its N never matches)
        if (step_exact_ok(cx.steps[tab]))
            for (uint32_t w = 0; w < cnt; ++w) chain_count<KW, EP, BLK, SIGMA>(st, fr, cx, w, 0, 1, true);
        return;
    }
    const uint64_t t0 = (uint64_t)q - S.a; // the table builder leaves keys near the ends of the text unlocated: no underflow
    uint64_t mm[KW];
    bool covered = false;
    if constexpr (KW <= 2) covered = S.a <= kCtx && NL - S.a - S.d <= kCtx;
    if (covered) {
        // text[q - kCtx, q + d + kCtx) as one bit string T (character c in bits 2c), shifted so that character 0 is the needle's
        const uint32_t sh = 32u + 2u * S.d; // 34..64
        const uint64_t t_lo = (uint64_t)ctx_l | ((uint64_t)key << 32) | (sh < 64u ? (uint64_t)ctx_r << sh : 0ull);
        const uint64_t t_hi = (uint64_t)ctx_r >> (64u - sh);
        const uint32_t s2 = 2u * (kCtx - S.a); // 0..32
        uint64_t tw[2];
        tw[0] = s2 ? (t_lo >> s2) | (t_hi << (64u - s2)) : t_lo;
        tw[1] = t_hi >> s2;
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            const uint64_t x = st.pat.w[k] ^ tw[k < 2 ? k : 1];
            mm[k] = (x | (x >> 1)) & 0x5555555555555555ull;
            if constexpr (SIGMA == 5) mm[k] |= spread_bits(st.pat.nm[k]);
        }
    } else {
        if (fetches) ++fetches->text_reads;
        Pattern<KW, SIGMA> tp;
        load_pattern(tp, cx.text, cx.nmask, t0, NL);
#pragma unroll
        for (int k = 0; k < KW; ++k) {
            const uint64_t x = st.pat.w[k] ^ tp.w[k];
            mm[k] = (x | (x >> 1)) & 0x5555555555555555ull;
            if constexpr (SIGMA == 5) { mm[k] |= spread_bits(st.pat.nm[k] | tp.nm[k]); tn[k] = spread_bits(tp.nm[k]); }
        }
    }
    if (count_mismatches<KW>(mm, cnt - 1u, K) > E) return;
    // the search's steps over the infix, with the text's mismatches (admissible child: e' <= ub and e' + rem >= lb)
    uint32_t e = 0;
    for (uint32_t t = 0; t < Li; ++t) {
        const uint32_t ent = cx.steps[tab + t], x = step_pos(ent);
        const uint64_t word = KW == 1 ? mm[0] : (KW == 2 ? (x < 32u ? mm[0] : mm[KW - 1]) : mm[x >> 5]); // selects keep mm in registers
        e += (uint32_t)(word >> (2u * (x & 31u))) & 1u;
        if (e > step_ub(ent) || e + step_rem(ent) < step_lb(ent)) return;
    }
    // the sequence holding the occurrence (largest s with limits[s] <= q, limits[s] = seq_start[s] - s), looked up
    // only when a window qualifies
    uint32_t sq = 0xffffffffu;
    uint64_t sb = 0, se = 0;
    for (uint32_t w = 0; w < cnt; ++w) {
        if (count_mismatches<KW>(mm, w, w + K) > E) continue;
        if constexpr (SIGMA == 5) { if (cx.skip_n && count_mismatches<KW>(tn, w, w + K) != 0u) continue; }
        if (sq == 0xffffffffu) {
            uint32_t a = 0, b = cx.n_seq;
            while (b - a > 1) {
                const uint32_t mid = (a + b) >> 1;
                if ((uint64_t)cx.seq_start[mid] - mid <= (uint64_t)q) a = mid; else b = mid;
            }
            sq = a;
            sb = (uint64_t)cx.seq_start[a] - a;
            se = (uint64_t)cx.seq_start[a + 1] - (a + 1);
        }
        if (t0 + w < sb || t0 + w + K > se) continue;
        if constexpr (EP) chain_mark_file<KW, BLK, SIGMA>(st, fr, w, cx.seq_to_file[sq]);
        else chain_count<KW, EP, BLK, SIGMA>(st, fr, cx, w, 0, 1, false);
    }
}

// the same for the search the chain is in (general kernel: the entry sits in st.lo_r / lo_f / ctx_l)
template <int KW, bool EP, bool BLK, int SIGMA, class Frames>
GMB_HD void verify_located(Chain<KW, SIGMA>& st, Frames& fr, const MapCtx& cx, FetchStats* fetches)
{
    const uint32_t cnt = BLK ? st.cnt : 1u;
    const SearchStart& S = cx.starts[cnt * kMaxSearches + st.s];
    uint32_t set = 0xffffffffu, key = st.pat.bits(S.a, S.d);
    if constexpr (BLK) {
#if defined(__CUDA_ARCH__)
        set = st.var == 0 ? S.set0 : __ldg(S.var + st.var);
#else
        set = st.var == 0 ? S.set0 : S.var[st.var];
#endif
        if (set != 0xffffffffu) {
            uint32_t sub = st.sub;
#pragma unroll
            for (uint32_t k = 0; k < kMaxE; ++k) {
                const uint32_t p = (set >> (8 * k)) & 0xffu;
                if (p != 0xffu) { key ^= (1u + sub % 3u) << (2u * p); sub /= 3u; }
            }
        }
    }
    verify_located_key<KW, EP, BLK, SIGMA>(st, fr, cx, fetches, S, key, set == 0xffffffffu, st.lo_r, st.lo_f, st.ctx_l);
}

// One state-machine iteration.  Returns false when the block is finished (results via chain_result).
// `fetches` counts rank-block reads (the roofline's algorithmic unit), when non-null.
// LOC (locate instantiation, one k-mer per chain, EP tables): every occurrence is reported, not counted.
// SUB (block_kernel.cu): the chain walks the subtree below ONE table entry; when that is exhausted the call returns
// false instead of moving on to the search's next key (the caller enumerates the keys itself).
template <int KW, bool EP, bool BLK, int SIGMA, class Frames, bool LOC = false, bool SUB = false>
GMB_HD bool chain_step(Chain<KW, SIGMA>& st, Frames& fr, const MapCtx& cx, FetchStats* fetches,
                       unsigned long long* lut_reads)
{
    static_assert(!LOC || (EP && !BLK), "the locate instantiation keeps both intervals in step and owns one k-mer");
    constexpr uint32_t kNone = 0xffu; // "no symbol": the pattern character is N
    const uint32_t K = cx.K, cnt = BLK ? st.cnt : 1u;
    const uint32_t Li = K - cnt + 1; // infix length
    if (fetches) ++fetches->iterations;

    if constexpr (!LOC) {
        // the search was entered through the table entry of a key that occurs once: one candidate alignment, finished
        // by comparing needle and text; the node is done after that
        if (st.size & kLocated) { verify_located<KW, EP, BLK, SIGMA>(st, fr, cx, fetches); st.size = 0; }
    }

    if (BLK && st.win == kNoWin && st.t == Li) {
        // the whole infix is matched (only reached when cnt > 1): this node is an infix hit; complete it for
        // every window, starting with window 0.  Frames of levels >= e are free here.
        fr.xset(0, st.lo_f); fr.xset(1, st.lo_r); fr.xset(2, st.size);
        st.leaf_e = st.e; st.win = 0; st.t = 0;
    }
    const bool in_flank = BLK && st.win != kNoWin;
    const uint32_t T = in_flank ? cnt - 1u : Li; // steps of the current walk
    const uint32_t tab = in_flank ? cx.fl_off[cnt] + st.win * (cnt - 1u) : (BLK ? cx.p1_off[cnt] : 0u) + st.s * Li;
    const uint32_t ent = cx.steps[tab + st.t];
    const uint32_t dir = step_dir(ent);
    // on the reverse strand st.pat already holds the reverse complement; a pattern N matches nothing
    const uint32_t pc = st.pat.at(step_pos(ent));
    const uint32_t p = (SIGMA == 5 && pc == 4u) ? kNone : pc;

    bool descend = false;
    uint32_t c = 0, csize = 0, cact = 0, coth = 0, ce = 0, ct = 0, cdir = dir;

    if (st.size == 0) {
        // an empty node can only come out of a jump table: nothing to search here
    } else if (!LOC && st.strand == 0 && st.e == 0 && st.size == 1 && !(SIGMA == 5 && st.has_n)) {
        // Forward strand, no error so far, one occurrence left: it is the query's own position in the
        // indexed text, so the rest of the pattern matches it exactly and no mismatching extension
        // exists.  The subtree holds exactly one full-length node, the query itself without any error: one
        // occurrence per window if the search admits an error-free completion, nothing if a later part of the
        // scheme demands an error — no need to walk it either way.
        if (step_exact_ok(ent)) {
            if (in_flank) chain_count<KW, EP, BLK, SIGMA>(st, fr, cx, st.win, 0, 1, true);
            else for (uint32_t w = 0; w < cnt; ++w) chain_count<KW, EP, BLK, SIGMA>(st, fr, cx, w, 0, 1, true);
        }
    } else {
        // ---- expand the node: ranks at both interval ends of the active index -----------------------
        if (fetches) { // a run of one-row expansions starts here unless the parent was one already
            if (st.size == 1 && !st.thin) ++fetches->thin_paths;
            st.thin = st.size == 1;
        }
        const uint32_t x = dir ? st.lo_r : st.lo_f;
        const uint32_t z = dir ? st.lo_f : st.lo_r;

        // admissible children (search-scheme bounds, find2_index_approx.hpp:388-389,254-258)
        const uint32_t ub = step_ub(ent), lb = step_lb(ent), rem = step_rem(ent);
        const uint32_t em = st.e + 1; // errors after a mismatching child
        const bool mis_ok = em <= ub && em + rem >= lb;
        const bool hit_ok = st.e + rem >= lb; // st.e <= ub is an invariant of the walk

        Children<SIGMA> ch;
        bool fast = false;
        if constexpr (SIGMA == 4) {
            // exact step whose other-index interval is never needed again: one symbol's rank suffices.  Taken
            // only when the whole warp agrees, so that lanes never serialise two differently shaped paths.
            fast = warp_all(!mis_ok && !step_sync(ent));
            if (fast) {
                const uint32_t y = x + st.size;
                const RankBlock* B = static_cast<const RankBlock*>(dir ? cx.blk[1] : cx.blk[0]);
                const uint32_t* SP = dir ? cx.sent[1] : cx.sent[0];
                const uint32_t bx = x / kBlockBases, by = y / kBlockBases;
                count_fetch(fetches, st.size, 1u + (by != bx));
                const BlockRegs rbx = load_block(B + bx);
                BlockRegs rby = rbx;
                load_block_if(rby, B + by, by != bx);
                const uint32_t r0 = block_rank_one(rbx, x - bx * kBlockBases, x, p, SP);
                const uint32_t r1 = block_rank_one(rby, y - by * kBlockBases, y, p, SP);
                const uint32_t np = r1 - r0, lp = (p == 0 ? cx.C[0] : (p == 1 ? cx.C[1] : (p == 2 ? cx.C[2] : cx.C[3]))) + r0;
#pragma unroll
                for (int k = 0; k < 4; ++k) { ch.n[k] = p == (uint32_t)k ? np : 0u; ch.l[k] = lp; }
                ch.oth0 = z;
            }
        }
        if (!fast) ch = expand_node<SIGMA>(cx, dir, x, st.size, z, fetches);

        uint32_t ok = 0;
#pragma unroll
        for (int k = 0; k < SIGMA; ++k)
            if (ch.n[k] && (p == (uint32_t)k ? hit_ok : mis_ok)) ok |= 1u << k;
        if (SIGMA == 5 && cx.skip_n) ok &= 0xfu; // the N child: text windows with an N are not this pass's
        const uint32_t hit_bit = p == kNone ? 0u : (1u << p);

        if (st.t + 1 == T && (in_flank || cnt == 1)) {
            // children are full-length matches of one window: count them (src/algo.hpp:48,191)
            const uint32_t w = in_flank ? st.win : 0u;
            if (EP) {
                // rows in SA(T): the active index's children when extending left, else the synchronised side
#pragma unroll
                for (int k = 0; k < SIGMA; ++k)
                    if (ok & (1u << k))
                        chain_count<KW, EP, BLK, SIGMA, Frames, LOC>(st, fr, cx, w, dir ? ch.oth0 + sum_below<SIGMA>(ch.n, (uint32_t)k) : ch.l[k], ch.n[k], false);
            } else {
                uint64_t sum = 0;
#pragma unroll
                for (int k = 0; k < SIGMA; ++k) sum += (ok & (1u << k)) ? ch.n[k] : 0u;
                chain_count<KW, EP, BLK, SIGMA>(st, fr, cx, w, 0, sum < cx.maxv ? (uint32_t)sum : cx.maxv, false);
            }
        } else if (ok) {
            const uint32_t mm = ok & ~hit_bit;
            c = mm ? lowest_bit_index(mm) : p; // mismatching children first, the matching child last
            const uint32_t pending = ok & ~(1u << c);
            if (pending) {
                const uint32_t lv = st.e;
#pragma unroll
                for (int k = 0; k < SIGMA; ++k) { fr.set(lv, k, ch.l[k]); fr.set(lv, SIGMA + k, ch.n[k]); }
                fr.set(lv, 2 * SIGMA, ch.oth0);
                fr.set(lv, 2 * SIGMA + 1, st.t | (pending << 8));
                st.lvmask |= 1u << lv;
            }
            csize = pick<SIGMA>(ch.n, c);
            cact = pick<SIGMA>(ch.l, c);
            coth = ch.oth0 + sum_below<SIGMA>(ch.n, c);
            ce = st.e + (c != p);
            ct = st.t + 1;
            descend = true;
        }
    }

    if (!descend) {
        // ---- backtrack ---------------------------------------------------------------------------------
        st.thin = false;
        // While an infix hit is being completed, the frames of its window walks sit at levels >= leaf_e and
        // the infix search's own pending frames below that.
        uint32_t cand = st.lvmask;
        if (BLK && st.win != kNoWin) {
            cand = (st.lvmask >> st.leaf_e) << st.leaf_e;
            if (cand == 0) {
                // this window is done: next window from the saved infix hit, or back to the infix search
                if (++st.win < cnt) {
                    st.lo_f = fr.xget(0); st.lo_r = fr.xget(1); st.size = fr.xget(2);
                    st.e = st.leaf_e; st.t = 0;
                    return true;
                }
                st.win = kNoWin;
                cand = st.lvmask;
            }
        }
        if (cand == 0) {
            if constexpr (SUB) return false; // the subtree below this table entry is done
            // this entry into the search is exhausted: its next key, the next set of substituted offsets, ...
            if constexpr (BLK) {
                if (++st.sub < st.nsub) { chain_start<KW, BLK, SIGMA>(st, cx, lut_reads); return true; }
                st.sub = 0;
                if (++st.var < cx.starts[st.cnt * kMaxSearches + st.s].n_var) { chain_start<KW, BLK, SIGMA>(st, cx, lut_reads); return true; }
                st.var = 0;
            }
            // ... then the next search, the next strand, or done
            if (++st.s == cx.n_search) {
                st.s = 0;
                if (++st.strand == cx.n_strands) {
                    if (EP && !BLK) { // distinct FASTA files with at least one occurrence on either strand (:360)
#if defined(__CUDA_ARCH__)
                        st.acc = (uint32_t)__popcll(st.files);
#else
                        st.acc = (uint32_t)__builtin_popcountll(st.files);
#endif
                    }
                    return false;
                }
                st.pat.reverse_complement(K + cnt - 1);
            }
            chain_start<KW, BLK, SIGMA>(st, cx, lut_reads);
            return true;
        }
        const uint32_t lv = highest_bit_index(cand);
        const uint32_t meta = fr.get(lv, 2 * SIGMA + 1);
        const uint32_t tf = meta & 0xffu;
        uint32_t pending = meta >> 8;
        const uint32_t tabf = (BLK && st.win != kNoWin) ? cx.fl_off[cnt] + st.win * (cnt - 1u) : (BLK ? cx.p1_off[cnt] : 0u) + st.s * Li;
        const uint32_t entf = cx.steps[tabf + tf];
        const uint32_t pcf = st.pat.at(step_pos(entf));
        const uint32_t pf = (SIGMA == 5 && pcf == 4u) ? kNone : pcf;
        const uint32_t mm = pending & ~(pf == kNone ? 0u : (1u << pf));
        c = mm ? lowest_bit_index(mm) : pf;
        pending &= ~(1u << c);
        if (pending) fr.set(lv, 2 * SIGMA + 1, tf | (pending << 8));
        else st.lvmask &= ~(1u << lv);
        uint32_t below = 0;
#pragma unroll
        for (int k = 0; k < SIGMA - 1; ++k) below += c > (uint32_t)k ? fr.get(lv, SIGMA + k) : 0u;
        csize = fr.get(lv, SIGMA + c);
        cact = fr.get(lv, c);
        coth = fr.get(lv, 2 * SIGMA) + below;
        ce = lv + (c != pf);
        ct = tf + 1;
        cdir = step_dir(entf);
    }

    st.size = csize;
    st.e = ce;
    st.t = ct;
    if (cdir) { st.lo_r = cact; st.lo_f = coth; }
    else      { st.lo_f = cact; st.lo_r = coth; }
    return true;
}

} // namespace gmb
