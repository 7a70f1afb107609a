/*
 * gm_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the `genmap map` hot path of cpockrandt/genmap, used only as the
 * parity checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 * Nothing under genmap_b200/ may include, link or call this.
 *
 * Parity status: PINNED — checked against the reference's golden vectors
 * (tests/test_cases/case_*: raw_freq16 / raw_freq8, copied as fixtures to tests/golden/) and
 * against outputs of the unmodified reference binary (oracle/_ref/genmap_ref) on seeded synthetic
 * genomes; see tests/test_oracle_golden.py and tests/golden/make_fixtures.py.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef GM_ORACLE_H
#define GM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmo_index gmo_index;

typedef struct gmo_params {
    uint32_t K;               /* k-mer length                       (-K, src/mappability.hpp:514) */
    uint32_t E;               /* Hamming errors 0..4                (-E, src/mappability.hpp:175-188) */
    uint32_t revcompl;        /* 1 unless -nc                       (src/mappability.hpp:516) */
    uint32_t exclude_pseudo;  /* -ep                                (src/mappability.hpp:517) */
    uint32_t value_bits;      /* 8 (-fs) or 16 (-fl and default)    (src/mappability.hpp:387-394) */
    uint32_t infix_len;       /* length of the common infix = K - overlap; 0 -> reference default
                                 (src/mappability.hpp:519-543) */
    uint32_t threads;         /* OpenMP threads, 0 -> all           (-T) */
    uint32_t copy_shortcut;   /* 1: copy value to all exact occurrences like the reference does for
                                 single-FASTA indices (src/algo.hpp:389-396); needs the full SA */
} gmo_params;

/* Build an index from code text: codes 0..3 = ACGT, 4 = N; `limits` has n_seq+1 cumulative offsets
 * into `codes` (no sentinels).  Naive suffix sorting — meant for inputs up to a few Mbp.
 * Restates src/indexing.hpp:72-149 + src/seqan_libdivsufsort.h:35-240 (BWT of text and of the
 * per-sequence reversed text, C array incl. sentinels, full SA kept for locate). */
gmo_index *gmo_index_build(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq);

/* Adopt BWTs computed elsewhere (symbols 0 = sentinel, 1..sigma = bases), e.g. exported from the
 * product's index so that the CPU baseline can run on a genome too large to suffix-sort here.
 * `sa` may be NULL (then exclude_pseudo / copy_shortcut are unavailable). */
gmo_index *gmo_index_from_bwt(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq,
                              const uint8_t *bwt_fwd, const uint8_t *bwt_rev, uint32_t sigma,
                              const uint64_t *sa);

void gmo_index_free(gmo_index *ix);
uint64_t gmo_index_bwt_len(const gmo_index *ix);
uint32_t gmo_index_sigma(const gmo_index *ix);
/* copy out the oracle's own BWTs / SA so tests can compare the product's index builder against them */
void gmo_index_get_bwt(const gmo_index *ix, int rev, uint8_t *out);
int gmo_index_get_sa(const gmo_index *ix, uint64_t *out);

/* Restates computeMappability<E> (src/algo.hpp:405-483) for one FASTA file whose text is the
 * infix [text_begin, text_begin+text_len) of the index's concatenated text.
 * chrom_cum: n_chrom+1 file-local cumulative lengths (first = 0).
 * intervals: n_intervals pairs [begin,end) file-local (may be NULL/0).
 * seq_to_file: global sequence number -> file id (only read under exclude_pseudo).
 * out: text_len values of value_bits/8 bytes, overwritten.  Returns 0 or a negative error. */
int gmo_map(const gmo_index *ix, const gmo_params *p, uint64_t text_begin, uint64_t text_len,
            const uint64_t *chrom_cum, uint32_t n_chrom, const uint64_t *intervals,
            uint64_t n_intervals, const uint32_t *seq_to_file, void *out);

/* Definition-level counter: sliding windows over every indexed sequence,
 * Hamming <= E on both strands, tails zeroed, saturation, -ep.  O(text_len * N * K): tiny inputs. */
int gmo_brute(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq, const gmo_params *p,
              uint64_t text_begin, uint64_t text_len, const uint64_t *chrom_cum, uint32_t n_chrom,
              const uint64_t *intervals, uint64_t n_intervals, const uint32_t *seq_to_file,
              void *out);

/* Definition-level csv lists of ONE k-mer (src/algo.hpp:311-346): all occurrences (<= E mismatches) of the
 * k-mer at concatenated-text position `pos` (strand 0) or of its reverse complement (strand 1) as sorted
 * (sequence, offset) pairs.  Returns the number of occurrences; at most `cap` are written. */
uint64_t gmo_brute_locations(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq, uint32_t K, uint32_t E,
                             uint64_t pos, int strand, uint32_t *seq_out, uint32_t *pos_out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
