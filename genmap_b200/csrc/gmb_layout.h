// gmb_layout.h — the HBM-resident index layout and the per-(K,E) search step tables.
//
// Everything here is plain data shared by the host builder, the GPU builder and the kernels.
// What it replaces in the reference: SeqAn's EPR rank dictionary (14-byte unaligned entries of 32
// symbols + superblocks, SEQAN/index/index_fm_rank_dictionary_levels.h:197-209,484-491), the separate
// sentinel bit-vector dictionary consulted whenever c == 'A' (index_fm_lf_table.h:468-491) and the
// C array (`lf.sums`, src/seqan_libdivsufsort.h:231-233).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GMB_HD __host__ __device__ __forceinline__
#else
#define GMB_HD inline
#endif

namespace gmb {

// ---- rank block: 16-byte header + W x 16 bytes of bit planes (64 BWT symbols per plane pair) --------
//   bytes  0..15 : cntA (sentinels NOT counted), cntC, cntG  = occurrences before this block
//                  sent  = (#sentinels before this block) << 8 | (#sentinels inside this block)
//   then W x { plane0 (low code bit) u64, plane1 (high code bit) u64 }, 64 symbols each
// codes: A=0 C=1 G=2 T=3; a sentinel row is stored as code 0 and listed in `sent_pos`.
// cntT is derived: T(i) = i - A(i) - C(i) - G(i) - $(i).
// W = 1 (default): a block is ONE 32-byte sector = one 256-bit load = one memory request per rank boundary
//                  (measured: the memory system sustains ~2x more random 32-byte requests than 64-byte blocks
//                  fetched as two 32-byte requests, profiles/r01/s1_randread.txt), 1.5 GB per direction at 3 Gbp.
// W = 3          : 64-byte block, 192 symbols, 1.0 GB per direction (the first layout measured; build with
//                  -DGMB_BLOCK_WORDS=3 to compare).
#ifndef GMB_BLOCK_WORDS
#define GMB_BLOCK_WORDS 1
#endif
constexpr uint32_t kBlockWords = GMB_BLOCK_WORDS;
constexpr uint32_t kBlockBases = 64 * kBlockWords;
constexpr uint32_t kBlockBytes = 16 + 16 * kBlockWords;
constexpr uint32_t kMaxSeq = (1u << 24) - 1; // sentinel counter has 24 bits
constexpr uint32_t kMaxK = 255;              // step tables keep pattern offsets in 8 bits
constexpr uint32_t kMaxE = 4;                // src/mappability.hpp:187
constexpr uint32_t kMaxSearches = 7;         // src/find2_index_approx.hpp:121-131
constexpr uint32_t kMaxBlockKmers = 16;      // adjacent k-mers searched together through their common infix
constexpr uint32_t kLocated = 0x80000000u;   // size flag of a LOCATED jump-table entry (gmb_core.h: JtFull)
constexpr uint32_t kCtx = 16;                // context characters a located entry carries on either side of its key
constexpr uint32_t kLocateMargin = 512;      // keys this close to either end of the text are never located (> kMaxK + 16 + 32 * 9)
constexpr uint32_t kDeadVariant = 0xfffffffeu; // jump-table entry set of a search that admits no string of the key's length

struct alignas(kBlockBytes) RankBlock {
    uint32_t cnt[3];
    uint32_t sent;
    uint64_t w[kBlockWords][2];
};
static_assert(sizeof(RankBlock) == kBlockBytes, "rank block size");
static_assert(kBlockWords == 1 || kBlockWords == 3, "supported layouts: 32-byte and 64-byte blocks");

// ---- Dna5 rank block (genomes containing N): 32 bytes, 32 BWT symbols, three bit planes ----------------
// codes A=0 C=1 G=2 T=3 N=4 (plane 2 set only for N); counters for A,C,G,T, N is derived:
// N(i) = i - A - C - G - T - $.  Sentinel rows as in the Dna4 block (code 0 + side list).
// Like the Dna4 block it is ONE 32-byte sector = one memory request per rank boundary: the path is bound by the
// request rate, not by bytes (the first Dna5 layout, 64 bytes / 96 symbols = two requests per boundary, ran
// 3x slower than Dna4 at E = 0: profiles/r01/s10_sweep_dna5.txt).  1 byte per symbol and direction.
constexpr uint32_t kBlockBases5 = 32;
struct alignas(32) RankBlock5 {
    uint32_t cnt[4];
    uint32_t sent;
    uint32_t plane[3]; // 32 symbols per plane
};
static_assert(sizeof(RankBlock5) == 32, "Dna5 rank block size");

// ---- on-disk / in-HBM blob --------------------------------------------------------------------------
// One contiguous, 256-byte aligned blob; offsets are relative to its start so the same bytes serve as
// file, pinned host copy and device copy (copied verbatim, broadcast verbatim).
constexpr uint64_t kMagic = 0x3130584449424d47ULL; // "GMBIDX01"
constexpr uint32_t kVersion = 4 + 16 * kBlockWords; // the block layouts are part of the format

struct IndexHeader {
    uint64_t magic;
    uint32_t version;
    uint32_t sigma;          // 4 (Dna4: RankBlock) or 5 (Dna5, the text contains N: RankBlock5 + N mask)
    uint64_t n_bwt;          // N = text length + one sentinel per sequence
    uint64_t n_text;         // concatenated text length (no sentinels)
    uint32_t n_seq;
    uint32_t n_blocks;       // N / kBlockBases + 1 per direction
    uint64_t C[6];           // C[c] = #symbols smaller than base c in T (sentinels included); C[4] = N
    uint64_t off_fwd;        // RankBlock[n_blocks]   BWT of T      (extend left,  reference: Fwd)
    uint64_t off_rev;        // RankBlock[n_blocks]   BWT of T'     (extend right, reference: Rev)
    uint64_t off_sent_fwd;   // uint32[n_seq]  sorted BWT rows holding a sentinel
    uint64_t off_sent_rev;
    uint64_t off_text;       // uint64[n_text/32 + 2]  2-bit packed concatenated text
    uint64_t off_limits;     // uint64[n_seq + 1]      sequence limits in the concatenated text
    uint64_t off_sa;         // uint32[n_bwt] FULL suffix array of T (0 = absent).  The reference samples
                             // it every 10th text position (src/seqan_libdivsufsort.h:135) to save host
                             // RAM; 4 bytes/row (12 GB at 3 Gbp) is affordable in 180 GB of HBM and
                             // makes locate one read instead of an LF walk.  Only -ep / csv need it.
    uint64_t off_seq_start;  // uint32[n_seq + 1] start of every sequence inside T (limits[i] + i)
    uint64_t off_nmask;      // sigma == 5: uint64[n_text/64 + 2], bit i set <=> text position i is N (0 = absent)
    uint64_t total_bytes;
    uint64_t reserved[7];
};
static_assert(sizeof(IndexHeader) % 8 == 0, "header alignment");

// ---- search step tables -----------------------------------------------------------------------------
// A search of an optimum search scheme (pi, L, U over nb blocks; src/find2_index_approx.hpp:41-62) is
// flattened into K steps.  Step t consumes pattern offset `pos` extending right (dir=1, BWT of T') or
// left (dir=0, BWT of T); a child with e' errors is admissible iff e' <= ub and e' + rem >= lb, where
// rem = characters of the current block still unread after this one.
//   bits  0..7  pos      bits  8..15 rem      bits 16..19 ub     bits 20..23 lb    bit 24 dir
//   bit 25 = the other index's interval is still needed after this step (a direction switch follows)
//   bit 26 = every block from this step on has lb == 0: an error-free completion is admissible
GMB_HD uint32_t step_pos(uint32_t s) { return s & 0xffu; }
GMB_HD uint32_t step_rem(uint32_t s) { return (s >> 8) & 0xffu; }
GMB_HD uint32_t step_ub(uint32_t s) { return (s >> 16) & 0xfu; }
GMB_HD uint32_t step_lb(uint32_t s) { return (s >> 20) & 0xfu; }
GMB_HD uint32_t step_dir(uint32_t s) { return (s >> 24) & 1u; }
GMB_HD uint32_t step_sync(uint32_t s) { return (s >> 25) & 1u; }
GMB_HD uint32_t step_exact_ok(uint32_t s) { return (s >> 26) & 1u; }

struct StepTables {
    uint32_t n_search;
    uint32_t K;
    uint32_t step[kMaxSearches * (kMaxK + 1)]; // [search * K + t]
};

} // namespace gmb
