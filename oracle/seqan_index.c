/*
 * seqan_index.c — TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Writes the big fibres of the reference's on-disk index (SeqAn 2.4 FM index as configured by GenMap,
 * src/common.hpp:38-52) from a BWT and a suffix array computed elsewhere, so that the UNMODIFIED reference
 * binary (oracle/_ref/genmap_ref) can run `map` on a genome whose divsufsort-based `genmap index` would
 * take ~45 minutes (3 Gbp).  Layouts (verified byte for byte against indices written by the reference,
 * tests/test_seqan_index_writer.py):
 *   index[.rev].lf.drv      ceil(N/32) x { u64 word: 32 Dna values, value k in bits 62-2k ; u16 prefix[3] }
 *                           prefix[c] = #values <= c before this block inside its superblock
 *                           (LevelsPrefixRDConfig, SEQAN/index/index_fm_rank_dictionary_levels.h:197-209,1663-1680)
 *   index[.rev].lf.drv.sbl  one u32 prefix[3] per 65504 values
 *   index[.rev].lf.drp      ceil(N/64) x { u64 bits, bit k at 63-k ; u16 ones before the block in its superblock }
 *   index[.rev].lf.drp.sbl  one u32 per 65472 values
 *   index.sa.ind            ceil(N/64) x { u64 bits ; u64 ones before the block }       (1-level dictionary)
 *   index.sa.val            one { u16 seqNo ; u32 seqPos } (6 bytes, packed) per sampled row, in row order;
 *                           a row is sampled iff seqPos % sampling == 0 (src/seqan_libdivsufsort.h:135)
 *   index.txt.concat        u64 length, then ceil(n/32) u64 words, value k in bits 62-2k
 * Small text fibres (.info .ids .limits .pst .drs .len) are written by the Python driver.
 * Limits of this writer: Dna4, (uint16 seqNo, uint32 seqPos, uint32 BWT) index class.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#pragma pack(push, 1)
typedef struct { uint64_t word; uint16_t prefix[3]; } lf_entry;      /* 14 bytes */
typedef struct { uint64_t bits; uint16_t ones; } bool_entry16;       /* 10 bytes */
typedef struct { uint64_t bits; uint64_t ones; } bool_entry64;       /* 16 bytes */
typedef struct { uint16_t i1; uint32_t i2; } sa_pair;                /* 6 bytes */
#pragma pack(pop)

static int dump(const char *path, const void *p, size_t bytes)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    size_t w = bytes ? fwrite(p, 1, bytes, f) : 0;
    return (fclose(f) != 0 || w != bytes) ? -1 : 0;
}

/* bwt: N symbols, 0 = sentinel row, 1..4 = A,C,G,T.  counts_out[4] receives #A,#C,#G,#T (without sentinels). */
int gmo_seqan_write_lf(const char *prefix, const uint8_t *bwt, uint64_t n, uint64_t *counts_out)
{
    const uint64_t nb = (n + 31) / 32, nsb = (n + 65503) / 65504;
    const uint64_t nb2 = (n + 63) / 64, nsb2 = (n + 65471) / 65472;
    lf_entry *drv = (lf_entry *)calloc(nb ? nb : 1, sizeof(lf_entry));
    uint32_t *sbl = (uint32_t *)calloc((nsb ? nsb : 1) * 3, sizeof(uint32_t));
    bool_entry16 *drp = (bool_entry16 *)calloc(nb2 ? nb2 : 1, sizeof(bool_entry16));
    uint32_t *sbl2 = (uint32_t *)calloc(nsb2 ? nsb2 : 1, sizeof(uint32_t));
    if (!drv || !sbl || !drp || !sbl2) return -2;
    /* two passes over chunks of 65504*64 values (a multiple of both superblock sizes' block counts is not
     * needed: chunk starts are multiples of 32 and 64 values): per-chunk counts, prefix sums, then fill */
    const uint64_t CH = 65504ull * 64ull;
    const uint64_t nch = (n + CH - 1) / CH;
    uint64_t *ch_tot = (uint64_t *)calloc((nch + 1) * 4, sizeof(uint64_t));   /* stored values (sentinel rows as A) */
    uint64_t *ch_real = (uint64_t *)calloc((nch + 1) * 4, sizeof(uint64_t));
    uint64_t *ch_ones = (uint64_t *)calloc(nch + 1, sizeof(uint64_t));
    if (!ch_tot || !ch_real || !ch_ones) return -2;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < (int64_t)nch; ++c) {
        uint64_t t[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0}, o = 0;
        const uint64_t e = ((uint64_t)c + 1) * CH < n ? ((uint64_t)c + 1) * CH : n;
        for (uint64_t i = (uint64_t)c * CH; i < e; ++i) {
            const uint32_t v = bwt[i] ? (uint32_t)bwt[i] - 1u : 0u;
            t[v]++;
            if (bwt[i]) r[v]++; else o++;
        }
        for (int k = 0; k < 4; ++k) { ch_tot[4 * (c + 1) + k] = t[k]; ch_real[4 * (c + 1) + k] = r[k]; }
        ch_ones[c + 1] = o;
    }
    for (uint64_t c = 1; c <= nch; ++c) {
        for (int k = 0; k < 4; ++k) { ch_tot[4 * c + k] += ch_tot[4 * (c - 1) + k]; ch_real[4 * c + k] += ch_real[4 * (c - 1) + k]; }
        ch_ones[c] += ch_ones[c - 1];
    }
    uint64_t real[4];
    for (int k = 0; k < 4; ++k) real[k] = ch_real[4 * nch + k];
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t c = 0; c < (int64_t)nch; ++c) {
        uint64_t tot[4] = {ch_tot[4 * c], ch_tot[4 * c + 1], ch_tot[4 * c + 2], ch_tot[4 * c + 3]};
        uint64_t sb_base[3] = {0, 0, 0};
        const uint64_t b0 = (uint64_t)c * CH / 32, b1 = ((uint64_t)c + 1) * CH / 32 < nb ? ((uint64_t)c + 1) * CH / 32 : nb;
        /* CH is a multiple of 65504, so every chunk starts on a superblock boundary */
        for (uint64_t b = b0; b < b1; ++b) {
            if (b % (65504 / 32) == 0) {
                const uint64_t s = b / (65504 / 32);
                sb_base[0] = tot[0]; sb_base[1] = tot[0] + tot[1]; sb_base[2] = tot[0] + tot[1] + tot[2];
                for (int k = 0; k < 3; ++k) sbl[3 * s + k] = (uint32_t)sb_base[k];
            }
            drv[b].prefix[0] = (uint16_t)(tot[0] - sb_base[0]);
            drv[b].prefix[1] = (uint16_t)(tot[0] + tot[1] - sb_base[1]);
            drv[b].prefix[2] = (uint16_t)(tot[0] + tot[1] + tot[2] - sb_base[2]);
            uint64_t w = 0;
            for (uint32_t k = 0; k < 32; ++k) {
                const uint64_t i = b * 32 + k;
                if (i >= n) break;
                const uint32_t v = bwt[i] ? (uint32_t)bwt[i] - 1u : 0u; /* sentinel substitute = A (lf.drs = 0) */
                w |= (uint64_t)v << (62 - 2 * k);
                tot[v]++;
            }
            drv[b].word = w;
        }
    }
    /* sentinel bit vector: bits in parallel, then one cheap sequential sweep over the blocks for the counts
     * (its superblocks of 65472 values do not align with the chunks) */
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)nb2; ++b) {
        uint64_t w = 0;
        for (uint32_t k = 0; k < 64; ++k) {
            const uint64_t i = (uint64_t)b * 64 + k;
            if (i >= n) break;
            if (bwt[i] == 0) w |= 1ull << (63 - k);
        }
        drp[b].bits = w;
    }
    {
        uint64_t ones = 0, sb_ones = 0;
        for (uint64_t b = 0; b < nb2; ++b) {
            if (b % (65472 / 64) == 0) { sb_ones = ones; sbl2[b / (65472 / 64)] = (uint32_t)sb_ones; }
            drp[b].ones = (uint16_t)(ones - sb_ones);
            ones += (uint64_t)__builtin_popcountll(drp[b].bits);
        }
    }
    free(ch_tot); free(ch_real); free(ch_ones);
    char path[4096];
    int rc = 0;
    snprintf(path, sizeof path, "%s.drv", prefix);     rc |= dump(path, drv, nb * sizeof(lf_entry));
    snprintf(path, sizeof path, "%s.drv.sbl", prefix); rc |= dump(path, sbl, nsb * 3 * sizeof(uint32_t));
    snprintf(path, sizeof path, "%s.drp", prefix);     rc |= dump(path, drp, nb2 * sizeof(bool_entry16));
    snprintf(path, sizeof path, "%s.drp.sbl", prefix); rc |= dump(path, sbl2, nsb2 * sizeof(uint32_t));
    if (counts_out) for (int c = 0; c < 4; ++c) counts_out[c] = real[c];
    free(drv); free(sbl); free(drp); free(sbl2);
    return rc;
}

/* sa: N rows, positions inside the sentinel-separated text; seq_start: n_seq+1 starts (limits[i] + i). */
int gmo_seqan_write_sa(const char *prefix, const uint32_t *sa, uint64_t n, const uint64_t *seq_start,
                       uint32_t n_seq, uint32_t sampling)
{
    const uint64_t nb = (n + 63) / 64;
    bool_entry64 *ind = (bool_entry64 *)calloc(nb ? nb : 1, sizeof(bool_entry64));
    if (!ind) return -2;
    uint64_t n_val = 0;
    for (uint32_t s = 0; s < n_seq; ++s) n_val += (seq_start[s + 1] - seq_start[s] - 1 + sampling - 1) / sampling;
    sa_pair *val = (sa_pair *)malloc((n_val ? n_val : 1) * sizeof(sa_pair));
    if (!val) { free(ind); return -2; }
    /* pass 1: indicator bits per block (parallel); pass 2: prefix sums; pass 3: values (parallel) */
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)nb; ++b) {
        uint64_t w = 0;
        for (uint32_t k = 0; k < 64; ++k) {
            const uint64_t i = (uint64_t)b * 64 + k;
            if (i >= n) break;
            const uint64_t p = sa[i];
            uint32_t lo = 0, hi = n_seq; /* largest s with seq_start[s] <= p */
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seq_start[mid] <= p) lo = mid; else hi = mid; }
            if (p + 1 == seq_start[lo + 1]) continue; /* a sentinel position is never sampled */
            if ((p - seq_start[lo]) % sampling == 0) w |= 1ull << (63 - k);
        }
        ind[b].bits = w;
    }
    uint64_t ones = 0;
    for (uint64_t b = 0; b < nb; ++b) { ind[b].ones = ones; ones += (uint64_t)__builtin_popcountll(ind[b].bits); }
    if (ones != n_val) { free(ind); free(val); return -4; }
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)nb; ++b) {
        uint64_t o = ind[b].ones;
        for (uint32_t k = 0; k < 64; ++k) {
            if (!((ind[b].bits >> (63 - k)) & 1ull)) continue;
            const uint64_t p = sa[(uint64_t)b * 64 + k];
            uint32_t lo = 0, hi = n_seq;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seq_start[mid] <= p) lo = mid; else hi = mid; }
            val[o].i1 = (uint16_t)lo;
            val[o].i2 = (uint32_t)(p - seq_start[lo]);
            ++o;
        }
    }
    if (ones != n_val) { free(ind); free(val); return -4; }
    char path[4096];
    int rc = 0;
    snprintf(path, sizeof path, "%s.ind", prefix); rc |= dump(path, ind, nb * sizeof(bool_entry64));
    snprintf(path, sizeof path, "%s.val", prefix); rc |= dump(path, val, n_val * sizeof(sa_pair));
    free(ind); free(val);
    return rc;
}

/* codes: n values 0..3 */
int gmo_seqan_write_packed_text(const char *path, const uint8_t *codes, uint64_t n)
{
    const uint64_t nw = (n + 31) / 32;
    uint64_t *buf = (uint64_t *)calloc(nw + 1, sizeof(uint64_t));
    if (!buf) return -2;
    buf[0] = n;
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)nw; ++b) {
        uint64_t w = 0;
        for (uint32_t k = 0; k < 32; ++k) {
            const uint64_t i = (uint64_t)b * 32 + k;
            if (i >= n) break;
            w |= (uint64_t)(codes[i] & 3u) << (62 - 2 * k);
        }
        buf[b + 1] = w;
    }
    int rc = dump(path, buf, (nw + 1) * sizeof(uint64_t));
    free(buf);
    return rc;
}
