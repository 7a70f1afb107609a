/*
 * genmap_b200.h — C ABI of libgenmap_b200.so: the B200-native drop-in for the `genmap map` hot path.
 *
 * What each entry point replaces in the reference (paths relative to cpockrandt/genmap):
 *
 *   gmb_map_frequencies      <- the call site  run(...) { std::vector<value_type> c(length(text), 0);
 *                               switch (opt.errors) computeMappability<E>(index, text, c, searchParams, ...) }
 *                               src/mappability.hpp:157-189, i.e. computeMappability (src/algo.hpp:405-483)
 *                               and everything below it (src/find2_index_approx.hpp, SeqAn FM index).
 *   gmb_index_open/_from_blob<- open(index, path, OPEN_RDONLY)            src/mappability.hpp:221-223,
 *                               src/genmap_helper.hpp:71-98
 *   gmb_index_build          <- buildIndex / indexCreate(fwd+rev)         src/indexing.hpp:72-149,
 *                               src/seqan_libdivsufsort.h:35-240
 *   gmb_last_error           <- the reference prints to stderr and exit(1)s (src/mappability.hpp:187-188)
 *
 * Conventions (same as the reference call site):
 *   - `text` of a FASTA file is the infix [text_begin, text_begin + text_len) of the index's
 *     concatenated text; positions in `out`, `chrom_cum_lengths` and `intervals` are file-local.
 *   - occurrences are counted over the WHOLE index (all files), both strands unless revcompl == 0.
 *   - out[j] = min(MAXV, #occurrences with <= E mismatches), 0 for the last K-1 positions of every
 *     sequence and outside the selection intervals; MAXV = 255 (value_bits 8) or 65535 (16).
 *   - every function returns 0 on success or a negative gmb_status; nothing throws across the ABI;
 *     the message is available from gmb_last_error() (thread-local).
 *   - there is NO CPU fallback: without a CUDA device (or if the kernels fail to load) the compute
 *     entry points return GMB_ERR_CUDA.
 *   - ONE call in flight per handle: a handle keeps per-call device scratch (work ranges, work counter) and
 *     may rebuild or drop its cached jump tables when the configuration changes.  Calls on one handle must come
 *     from one host thread at a time and, for the asynchronous gmb_map_frequencies_device, on one CUDA stream
 *     (stream order then protects the scratch).  Use one handle per stream / host thread (gmb_index_replicate
 *     onto the same device gives an independent handle).
 * Plain pointers and sizes only; no torch / STL types in any signature.
 */
#ifndef GENMAP_B200_H
#define GENMAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmb_index gmb_index;

typedef enum gmb_status {
    GMB_OK = 0,
    GMB_ERR_ARG = -1,         /* bad argument (K, E, sizes, NULL) */
    GMB_ERR_UNSUPPORTED = -2, /* E > 4, K > 255, index >= 2^32-1 symbols */
    GMB_ERR_CUDA = -3,        /* no device / CUDA runtime error */
    GMB_ERR_IO = -4,          /* index file missing / malformed */
    GMB_ERR_NOMEM = -5
} gmb_status;

typedef struct gmb_params {
    uint32_t K;              /* -K  k-mer length, E+2 <= K <= 255 */
    uint32_t E;              /* -E  mismatches, 0..4 (src/mappability.hpp:175-188) */
    uint32_t revcompl;       /* 0 = -nc (forward strand only), else both strands */
    uint32_t exclude_pseudo; /* -ep: count distinct FASTA files instead (needs the SA section) */
    uint32_t value_bits;     /* 8 (-fs) or 16 (-fl and the default float output) */
    uint32_t count_fetches;  /* 1: also count rank-block fetches (instrumented kernel, slower) */
    uint32_t block_kmers;    /* adjacent k-mers searched together through their common infix (the reference's
                                K - overlap + 1, src/algo.hpp:416); 0 = default for (K,E), 1 = one k-mer at a
                                time; results do not depend on it (tests/tests.sh:47-60 checks the same for -xo) */
    uint32_t reserved;
} gmb_params;

typedef struct gmb_index_info {
    uint64_t n_text;      /* concatenated text length */
    uint64_t n_bwt;       /* text + one sentinel per sequence */
    uint32_t n_seq;
    uint32_t has_sa;
    uint64_t blob_bytes;  /* size of the index blob in HBM */
    uint64_t rank_block_bytes; /* one 32-byte sector: 64 symbols (Dna4) or 32 symbols (Dna5) per block */
    void *device_blob;    /* device address of the blob (for broadcast / diagnostics) */
    int32_t device;
    int32_t alphabet_size; /* 4 = Dna4; 5 = Dna5, chosen when the text contains N (src/indexing.hpp:459-473).  A Dna5 index
                            * that holds the suffix array is searched faster at E >= 1: same results */
    uint64_t jump_table_bytes; /* HBM currently held by the handle's jump tables (built lazily by the map calls) */
} gmb_index_info;

typedef struct gmb_map_stats {
    double kernel_ms;            /* CUDA-event time of the search kernel */
    uint64_t positions;          /* k-mer starts actually searched */
    uint64_t rank_block_fetches; /* only with count_fetches: 32-byte rank blocks read */
    uint64_t jump_table_reads;   /* only with count_fetches: jump-table entries read (8, 12 or 16 bytes each) */
    uint32_t kernel_launches;
    uint32_t jump_depth;         /* deepest jump table used by this call (0 = none) */
    /* only with count_fetches: the fetches split by the size of the suffix-array interval being expanded
     * (1, 2, 3-4, 5-8, 9-16, 17-32, 33-64, 65+) and the number of maximal runs of one-row expansions */
    uint64_t fetches_by_size[8];
    uint64_t thin_paths;
    uint64_t iterations;         /* only with count_fetches: passes of all chains through the search state machine */
    uint64_t located_entries;    /* only with count_fetches: searches finished at a table entry whose key occurs once */
    uint64_t text_reads;         /* only with count_fetches: ... of which had to read the packed text (<= 32 bytes each) */
    uint32_t block_kmers;        /* adjacent k-mers searched together by this call (the planner's choice when params.block_kmers == 0) */
    uint32_t reserved;
} gmb_map_stats;

/* flags for gmb_index_build */
#define GMB_BUILD_WITH_SA 1u  /* keep the full suffix array (needed by -ep) */
#define GMB_BUILD_ON_GPU  2u  /* suffix-sort on the device (prefix doubling) instead of host SA-IS */

const char *gmb_last_error(void);
const char *gmb_version(void);
int gmb_device_count(void);

/* Build an index blob in host memory from code text (0..3 = ACGT, 4 = N; any N makes it a Dna5 index).
 * limits: n_seq+1 cumulative offsets.  The blob is released with gmb_blob_free. */
int gmb_index_build(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq, uint32_t flags,
                    int device, void **blob_out, uint64_t *bytes_out);
void gmb_blob_free(void *blob);

/* Build on the GPU and keep the blob in HBM: returns an opened index that owns its device blob.
 * timings_ms (optional, 4 doubles): H2D, suffix sorting, BWT + block packing, total. */
int gmb_index_build_device(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq, uint32_t flags,
                           int device, gmb_index **out, double *timings_ms);
int gmb_blob_save(const void *blob, uint64_t bytes, const char *path);

/* Open an index for searching on `device`: from a directory written by `genmap index`
 * (<dir>/index.gmb), from a host blob (copied to HBM) or by adopting a blob that already sits in
 * device memory (e.g. after an NCCL broadcast; not freed by gmb_index_close). */
int gmb_index_open(const char *dir, int device, gmb_index **out);
/* Convert an index directory written by the reference's own `genmap index` (SeqAn fibres) into a host
 * blob (release with gmb_blob_free).  gmb_index_open does this automatically when <dir>/index.gmb is
 * absent but <dir>/index.lf.drv exists.  Dna4 and Dna5 indices of every width class of the reference
 * (src/indexing.hpp:151-170) with fewer than 2^32 - 1 rows; the sampled suffix array is not imported. */
int gmb_index_import_reference(const char *dir, void **blob_out, uint64_t *bytes_out);
/* The reverse: write a host blob that holds the suffix array as an index directory in the reference's own format
 * (<dir>/index.lf.drv, index.sa.val, ...: everything `genmap map` of the reference opens, src/genmap_helper.hpp:71-127),
 * byte-identical to what the reference's `genmap index` writes for the same FASTA input.  ids: one
 * "file;length;name" string per indexed sequence (index.ids, src/indexing.hpp:399-401); sampling: suffix-array
 * sampling rate (the reference's -S, default 10).  Dna4 indices with at most 65535 sequences. */
int gmb_blob_export_reference(const void *blob, uint64_t bytes, const char *dir, const char *const *ids, uint32_t n_ids,
                              int fasta_directory, uint32_t sampling);
int gmb_index_from_blob(const void *host_blob, uint64_t bytes, int device, gmb_index **out);
int gmb_index_adopt_device(void *device_blob, uint64_t bytes, int device, gmb_index **out);
/* Replicate an opened index into the HBM of another GPU of the same node with a peer-to-peer copy (NVLink /
 * NVSwitch when the devices can address each other): how a single-process host (the `genmap` CLI with --gpus N)
 * puts the index on every GPU after reading it once.  Multi-process hosts broadcast the blob themselves (NCCL)
 * and call gmb_index_adopt_device. */
int gmb_index_replicate(const gmb_index *src, int device, gmb_index **out);
int gmb_index_close(gmb_index *idx);
int gmb_index_get_info(const gmb_index *idx, gmb_index_info *info);
/* Maximum depth of the jump tables that replace the first steps of every search:
 * -1 = automatic (ceil(log4 N), at most 16, less when HBM is short), 0 = off, 1..16 = fixed.  Tables are built
 * lazily on the device by the first map call that needs them and cached in the handle. */
int gmb_index_set_jump_depth(gmb_index *idx, int depth);
/* Plan the searches (part lengths of the search scheme, k-mers per block, jump-table entry depths) as if the
 * text had n_symbols symbols; 0 = the index's own size (default).  Results never depend on the plan, only the
 * cost does: this is how the plan of a 3 Gbp genome is exercised on a small one (tests), or a plan pinned. */
int gmb_index_set_plan_text_size(gmb_index *idx, uint64_t n_symbols);
/* Progress of the map call in flight on this handle, readable from another host thread (what the reference prints
 * as "Progress: x%", src/common.hpp:94-131, src/algo.hpp:478-481): k-mer start positions handed out so far and
 * the total of the call.  Both are 0 before the first call. */
int gmb_progress(gmb_index *idx, uint64_t *done, uint64_t *total);
/* Diagnostics / test support: decode one direction's BWT (rev = 0: of T, 1: of T') to one byte per
 * row (0 = sentinel, 1..5 = A,C,G,T,N) into host memory (n_bwt bytes). */
int gmb_index_export_bwt(gmb_index *idx, int rev, uint8_t *out_host);
/* Copy the full suffix array of T (n_bwt uint32 positions inside the sentinel-separated text) to host
 * memory; fails with GMB_ERR_UNSUPPORTED when the index was built without GMB_BUILD_WITH_SA. */
int gmb_index_export_sa(gmb_index *idx, uint32_t *out_host);

/* The hot path with HOST output (what the reference's run() does for one FASTA file).
 * out: text_len elements of value_bits/8 bytes, overwritten.  seq_to_file / n_seq only under -ep. */
int gmb_map_frequencies(gmb_index *idx, const gmb_params *params, uint64_t text_begin, uint64_t text_len,
                        const uint64_t *chrom_cum_lengths, uint32_t n_chrom,
                        const uint64_t (*intervals)[2], uint64_t n_intervals,
                        const uint32_t *seq_to_file, uint32_t n_seq, void *out, gmb_map_stats *stats);

/* Same, but only for file-local positions [pos_begin, pos_end): `out` receives pos_end - pos_begin
 * elements (the slice).  This is what a multi-GPU host driver calls per GPU with the slices of one
 * pinned host vector c. */
int gmb_map_frequencies_range(gmb_index *idx, const gmb_params *params, uint64_t text_begin, uint64_t text_len,
                              const uint64_t *chrom_cum_lengths, uint32_t n_chrom,
                              const uint64_t (*intervals)[2], uint64_t n_intervals,
                              const uint32_t *seq_to_file, uint32_t n_seq, uint64_t pos_begin,
                              uint64_t pos_end, void *out, gmb_map_stats *stats);

/* Same, restricted to file-local positions [pos_begin, pos_end) (range-sharding across GPUs) and
 * writing into DEVICE memory: out_device has text_len elements; only positions inside the range are
 * written (the caller zero-fills).  Launches on `cuda_stream` (a cudaStream_t; NULL = default
 * stream) and does not synchronise unless stats != NULL. */
int gmb_map_frequencies_device(gmb_index *idx, const gmb_params *params, uint64_t text_begin,
                               uint64_t text_len, const uint64_t *chrom_cum_lengths, uint32_t n_chrom,
                               const uint64_t (*intervals)[2], uint64_t n_intervals,
                               const uint32_t *seq_to_file, uint32_t n_seq, uint64_t pos_begin,
                               uint64_t pos_end, void *out_device, void *cuda_stream, gmb_map_stats *stats);

/* ---- locations: what the csv output needs (the csvComputation branch of computeMappabilitySingleBlock,
 * src/algo.hpp:311-343, and the `locations` map it fills; consumed by saveCsv, src/output.hpp:189-288) ----
 * For every file-local position j of [pos_begin, pos_end') the sorted list of all occurrences (<= E mismatches)
 * of the k-mer at j on the + strand and, unless revcompl == 0, of its reverse complement (the "- strand" column):
 *     list (j, s) = loc[offsets[2*(j-pos_begin)+s] .. offsets[2*(j-pos_begin)+s+1])      s = 0 (+), 1 (-)
 * Each occurrence is the reference's Pair<TSeqNo,TSeqPos> (src/common.hpp:59-63): sequence number over the whole
 * index and offset inside that sequence; lists are sorted by (seq, pos) like src/algo.hpp:335,346.  Lists are NOT
 * saturated.  Positions that are not searched (window leaves its sequence, outside the selection) have empty
 * lists — exactly the k-mers the reference does not emit (src/algo.hpp:377-385).
 * One call covers positions [pos_begin, out->pos_end) with out->pos_end <= pos_end chosen so that at most
 * `max_locations` occurrences (0 = default, 64 Mi) are returned, but always at least one position; the caller
 * continues from out->pos_end.  Needs an index with the suffix-array section.  The arrays are owned by the
 * library until gmb_locations_free(). */
typedef struct gmb_location {
    uint32_t seq;
    uint32_t pos;
} gmb_location;

typedef struct gmb_locations {
    uint64_t pos_begin, pos_end;
    uint64_t n_locations;
    uint64_t *offsets;     /* 2 * (pos_end - pos_begin) + 1 */
    gmb_location *loc;     /* n_locations */
    double kernel_ms;      /* both search passes */
} gmb_locations;

int gmb_map_locations(gmb_index *idx, const gmb_params *params, uint64_t text_begin, uint64_t text_len,
                      const uint64_t *chrom_cum_lengths, uint32_t n_chrom,
                      const uint64_t (*intervals)[2], uint64_t n_intervals, uint64_t pos_begin,
                      uint64_t pos_end, uint64_t max_locations, gmb_locations *out);
void gmb_locations_free(gmb_locations *locations);

/* ---- runs: the frequency vector of one FASTA file, run-length encoded on the device ----------------------
 * What the track writers consume (saveWig / saveBedGraph scan c for maximal runs of equal values inside every
 * sequence, src/output.hpp:73-187).  Same arguments and semantics as gmb_map_frequencies_range, but instead of
 * the vector the call returns its runs inside [pos_begin, pos_end): run r covers file-local positions
 * [start[r], start[r+1]) (the last one ends at pos_end) and has value[r] (0 = not computed / tail, like c).
 * A new run starts at every sequence start, so no run spans two sequences.  Only the runs leave the GPU
 * (10 bytes per run instead of value_bits/8 bytes per position).  Released with gmb_runs_free(). */
typedef struct gmb_runs {
    uint64_t pos_begin, pos_end;
    uint64_t n_runs;
    uint64_t *start;   /* n_runs file-local start positions, ascending */
    uint16_t *value;   /* n_runs values */
    double kernel_ms;  /* the search kernel */
    double rle_ms;     /* run detection + compaction on the device */
} gmb_runs;

int gmb_map_runs(gmb_index *idx, const gmb_params *params, uint64_t text_begin, uint64_t text_len,
                 const uint64_t *chrom_cum_lengths, uint32_t n_chrom,
                 const uint64_t (*intervals)[2], uint64_t n_intervals,
                 const uint32_t *seq_to_file, uint32_t n_seq, uint64_t pos_begin, uint64_t pos_end,
                 gmb_runs *out, gmb_map_stats *stats);
void gmb_runs_free(gmb_runs *runs);

#ifdef __cplusplus
}
#endif
#endif
