// gmb_host.h — host-side pieces of the library that need no CUDA: search-scheme step tables, the host
// index builder (SA-IS) and the blob packer shared with the GPU builder.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "gmb_layout.h"

namespace gmb {

// Flatten the optimum search scheme for E errors over a pattern of K characters into step tables
// (scheme tables: src/find2_index_approx.hpp:67-134; block lengths :164-176; start/direction :149-162).
// Returns false with `err` set if (K,E) is unsupported.
// force_sync: keep both intervals of the bidirectional index in step at every step (--exclude-pseudo
// needs the interval in SA(T) at every full-length match).
// n_bwt != 0: the part lengths are chosen for a text of n_bwt symbols (see choose_part_lengths in gmb_host.cpp);
// n_bwt == 0: the reference's equal split.  Results never depend on the split.
bool build_step_tables(uint32_t K, uint32_t E, StepTables& out, std::string& err, bool force_sync = false, uint64_t n_bwt = 0,
                       uint32_t block_kmers = 1, uint32_t block_bases = 64, bool nfree = false);

// Search tables of a (K,E) configuration for blocks of up to B adjacent k-mers (see Chain in gmb_core.h):
// for every block size cnt = 1..B the scheme's step table over the common infix (K - cnt + 1 characters,
// offsets in needle coordinates) and, per window, the steps through its flank characters (left flank from
// the infix outwards, then right flank) with the error budget E and no lower bound.
struct BlockTables {
    uint32_t K = 0, E = 0, B = 0, n_search = 0;
    uint32_t p1_off[kMaxBlockKmers + 1] = {};
    uint32_t fl_off[kMaxBlockKmers + 1] = {};
    std::vector<uint32_t> steps;
    std::vector<StepTables> infix; // [cnt] the per-cnt infix tables (index 0 unused), kept for jump-table planning
};
// B == 0 picks the default for (K,E).  force_sync: see build_step_tables.  nfree (Dna5 indices): plan for searches that
// never match a text N and are therefore entered through substituted keys like on a Dna4 index (MapCtx::skip_n).
bool build_block_tables(uint32_t K, uint32_t E, uint32_t B, bool force_sync, BlockTables& out, std::string& err, uint64_t n_bwt = 0,
                        uint32_t block_bases = 64, bool nfree = false);
uint32_t default_block_kmers(uint32_t K, uint32_t E);
// B for a text of n_bwt symbols by the expected-fetch model (gmb_host.cpp); falls back to default_block_kmers
uint32_t model_block_kmers(uint32_t K, uint32_t E, uint64_t n_bwt, uint32_t block_bases, bool nfree = false);

// How every search of a (K,E) configuration is entered through the jump tables.  depth[s]: length of the key
// (0 = no table: start at the root); a[s]: pattern offset of the key window [a, a + depth) (the region the first
// depth steps consume, contiguous whatever their directions).  Up to the search's error-free prefix one key is read;
// deeper (chosen by the expected-fetch model when the text size is known) every string of that length the scheme
// admits is read instead of walking to it: variants[var_off[s] .. + n_var[s]) lists the admissible sets of
// substituted key offsets (one byte each, 0xff = unused; the error-free set first), each standing for 3^|set| keys.
// need_lof[s]: the interval in SA(T) is needed after the jump (a later step extends to the left / both are kept in step).
struct JumpPlan {
    uint32_t depth[kMaxSearches];
    uint32_t a[kMaxSearches];
    bool need_lof[kMaxSearches];
    uint32_t var_off[kMaxSearches], n_var[kMaxSearches];
    std::vector<uint32_t> variants;
    uint32_t max_depth;
};
// allow_variants: the launch uses the blocked instantiation of the kernel (the one-k-mer instantiation enters every
// search through its error-free prefix only)
void plan_jump_tables(const StepTables& tabs, uint32_t max_depth, JumpPlan& plan, uint32_t E = 0, uint64_t n_bwt = 0, uint32_t sigma = 4,
                      uint32_t block_kmers = 1, bool allow_variants = false, bool nfree = false);
// ceil(log4(n_bwt)) clamped to [1,16]: less than one expected occurrence per table entry
uint32_t default_jump_depth(uint64_t n_bwt);

// The table keys of one strand of a block as a flat list, per block size cnt = 1..B (block_kernel.cu): every search,
// every admissible set of substituted offsets, the 3^m substitutions of a set of m offsets spelled out as an XOR mask
// on the key window.  x = XOR mask, y = search | errors << 4 | (nothing substituted) << 8.  Returns false (and an empty
// list) when some search starts at the root instead of a table.
struct KeyLists {
    std::vector<uint32_t> xy; // pairs (x, y)
    uint32_t off[kMaxBlockKmers + 1] = {}, n[kMaxBlockKmers + 1] = {};
};
bool build_key_lists(const BlockTables& tabs, const std::vector<JumpPlan>& plans, KeyLists& out);

// 256-byte aligned growable byte buffer for the index blob
struct Blob {
    std::vector<uint64_t> storage;
    uint64_t bytes = 0;
    uint8_t* data() { return reinterpret_cast<uint8_t*>(storage.data()); }
    const uint8_t* data() const { return reinterpret_cast<const uint8_t*>(storage.data()); }
    void resize(uint64_t n) { storage.assign((n + 7) / 8, 0); bytes = n; }
};

// section sizes/offsets for a text of n_text bases in n_seq sequences
struct BlobPlan {
    IndexHeader h;
};
BlobPlan plan_blob(uint64_t n_text, uint32_t n_seq, bool with_sa, uint32_t sigma = 4);
inline uint32_t block_bases(uint32_t sigma) { return sigma == 5 ? kBlockBases5 : kBlockBases; }
inline uint32_t block_bytes(uint32_t sigma) { return sigma == 5 ? (uint32_t)sizeof(RankBlock5) : kBlockBytes; }

// Pack one direction's BWT (symbols: 0/1 = sentinel, 2..5 = A,C,G,T, 6 = N) into rank blocks + sentinel list.
// Returns per-base totals in tot[].  The Dna5 variant fills RankBlock5.
void pack_bwt_blocks(const uint8_t* bwt, uint64_t n, RankBlock* blocks, uint32_t n_blocks, uint32_t* sent_pos,
                     uint32_t n_seq, uint64_t tot[4]);
void pack_bwt_blocks5(const uint8_t* bwt, uint64_t n, RankBlock5* blocks, uint32_t n_blocks, uint32_t* sent_pos,
                      uint32_t n_seq, uint64_t tot[5]);

// Build the whole index on the host.  codes: 0..3 = ACGT, 4 = N (any N makes it a Dna5 index, as in
// src/indexing.hpp:459-473); limits: n_seq+1 cumulative offsets.
bool build_index_host(const uint8_t* codes, const uint64_t* limits, uint32_t n_seq, bool with_sa, Blob& blob,
                      std::string& err);

// Import an index written by the reference's own `genmap index` (SeqAn fibres index.lf.drv/.drp,
// index.rev.lf.*, index.txt.*, index.lf.pst; src/genmap_helper.hpp:71-98 lists what `map` opens) into the
// HBM blob layout, so that pre-built GenMap indices can be used as they are.  Dna4 and Dna5, every width class of
// the reference as long as the text has fewer than 2^32 - 1 rows; the sampled suffix array is not imported
// (no --exclude-pseudo / csv on imported indices).
bool import_reference_index(const std::string& dir, Blob& blob, std::string& err);

// The reverse (seqan_export.cpp): the reference's own index directory from a blob that holds the suffix array.
// ids: one "file;length;name" line per sequence (index.ids, src/indexing.hpp:399-401).
bool export_reference_index(const uint8_t* blob, uint64_t bytes, const std::string& dir, const std::vector<std::string>& ids,
                            bool fasta_directory, uint32_t sampling, std::string& err);

// Positions whose k-mer is actually searched: inside a sequence with at least K bases left
// (everything else stays 0: resetLimits, src/algo.hpp:10-22), inside a selection interval if any
// (src/algo.hpp:441-476), inside [pos_begin, pos_end) (multi-GPU sharding).  Sorted, disjoint.
struct WorkRange { uint64_t begin, end; };
void build_work_ranges(uint64_t text_len, uint32_t K, const uint64_t* chrom_cum, uint32_t n_chrom,
                       const uint64_t* intervals, uint64_t n_intervals, uint64_t pos_begin, uint64_t pos_end,
                       std::vector<WorkRange>& out);

// sanity-check a blob (magic, version, offsets inside total_bytes)
bool validate_blob(const uint8_t* blob, uint64_t bytes, std::string& err);
bool validate_header(const IndexHeader& h, uint64_t bytes, std::string& err);

} // namespace gmb
