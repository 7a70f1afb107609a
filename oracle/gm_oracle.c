/*
 * gm_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see gm_oracle.h).
 *
 * Plain-C restatement of the reference's `genmap map` hot path:
 *   computeMappability / computeMappabilitySingleBlock / extend / approxSearch / extendExact
 *       (src/algo.hpp:10-483)
 *   optimum search schemes, GenMap variant (src/find2_index_approx.hpp:41-457)
 *   bidirectional FM-index iterator (SeqAn index_fm_stree.h:256-341, index_bifm_stree.h:46-70,
 *       index_fm_lf_table.h:468-491) and rank dictionary (index_fm_rank_dictionary_levels.h:1489-1523)
 *   index construction semantics (src/seqan_libdivsufsort.h:35-240, src/indexing.hpp:72-149)
 * plus a definition-level brute-force counter.
 *
 * Parity: PINNED against tests/golden (reference golden vectors + outputs of oracle/_ref/genmap_ref).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
 */
#include "gm_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GMO_N 4u      /* code of 'N' in the text */
#define GMO_MAXSYM 6u /* sentinel + up to 5 bases */

/* ------------------------------------------------------------------------------------------------
 * Rank dictionary.  The reference uses SeqAn's 2-level EPR dictionary (packed 64-bit words + prefix
 * counters, index_fm_rank_dictionary_levels.h:1489-1523); only its VALUE — rank_c(i) = #c in
 * bwt[0,i) — is part of the contract, so the oracle keeps a simpler layout of its own: symbols as
 * nibbles, one counter row per 64 symbols.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t n;
    uint32_t sigma;    /* number of base symbols (4 or 5); BWT symbols 0 = sentinel, 1..sigma */
    uint64_t *nib;     /* 4 words per block of 64 symbols */
    uint32_t *cnt;     /* GMO_MAXSYM counters per block: occurrences before the block */
    uint64_t C[GMO_MAXSYM + 1]; /* C[s] = #symbols < s in the text (sentinels included,
                                   src/seqan_libdivsufsort.h:231-233) */
} gmo_fm;

struct gmo_index {
    gmo_fm fwd, rev;
    uint8_t *codes;    /* concatenated text, no sentinels */
    uint64_t *limits;  /* n_seq+1 */
    uint32_t n_seq;
    uint64_t n_text;
    uint64_t *sa;      /* full suffix array over the sentinel-separated text, or NULL */
    uint64_t *cum_sent; /* limits[i] + i: start of sequence i in the sentinel-separated text */
};

static inline uint32_t nib_eq(uint64_t w, uint32_t c, uint32_t nn)
{
    uint64_t x = w ^ (0x1111111111111111ULL * c);
    x |= x >> 1;
    x |= x >> 2;
    x = ~x & 0x1111111111111111ULL;
    if (nn < 16) x &= (nn == 0) ? 0 : ((1ULL << (4 * nn)) - 1);
    return (uint32_t)__builtin_popcountll(x);
}

static inline uint64_t fm_rank(const gmo_fm *fm, uint32_t c, uint64_t i)
{
    uint64_t blk = i >> 6;
    uint32_t r = (uint32_t)(i & 63);
    uint64_t res = fm->cnt[blk * GMO_MAXSYM + c];
    const uint64_t *w = fm->nib + blk * 4;
    for (uint32_t k = 0; k < 4 && r > 0; ++k) {
        uint32_t nn = r >= 16 ? 16 : r;
        res += nib_eq(w[k], c, nn);
        r -= nn;
    }
    return res;
}

static int fm_init(gmo_fm *fm, const uint8_t *bwt, uint64_t n, uint32_t sigma)
{
    uint64_t nblk = n / 64 + 1;
    fm->n = n;
    fm->sigma = sigma;
    fm->nib = (uint64_t *)calloc(nblk * 4, sizeof(uint64_t));
    fm->cnt = (uint32_t *)calloc(nblk * GMO_MAXSYM, sizeof(uint32_t));
    if (!fm->nib || !fm->cnt) return -1;
    uint64_t run[GMO_MAXSYM] = {0};
    for (uint64_t b = 0; b < nblk; ++b) {
        for (uint32_t s = 0; s < GMO_MAXSYM; ++s) fm->cnt[b * GMO_MAXSYM + s] = (uint32_t)run[s];
        for (uint32_t k = 0; k < 64; ++k) {
            uint64_t i = b * 64 + k;
            if (i >= n) break;
            uint32_t s = bwt[i];
            fm->nib[b * 4 + (k >> 4)] |= (uint64_t)s << (4 * (k & 15));
            run[s]++;
        }
    }
    fm->C[0] = 0;
    for (uint32_t s = 0; s < GMO_MAXSYM; ++s) fm->C[s + 1] = fm->C[s] + run[s];
    return 0;
}

static void fm_free(gmo_fm *fm)
{
    free(fm->nib);
    free(fm->cnt);
    fm->nib = NULL;
    fm->cnt = NULL;
}

/* ------------------------------------------------------------------------------------------------
 * Suffix sorting (naive prefix doubling).  Order = the one divsufsort gives the reference on
 * ctext = ord+1 with all sentinels 0 (src/seqan_libdivsufsort.h:80-96): plain lexicographic order
 * of the suffixes of s1$s2$...sm$ with '$' < A and a shorter suffix smaller than its extensions.
 * ---------------------------------------------------------------------------------------------- */
static const uint32_t *g_rank;
static uint64_t g_h, g_n;

static int cmp_suffix(const void *pa, const void *pb)
{
    uint32_t a = *(const uint32_t *)pa, b = *(const uint32_t *)pb;
    if (g_rank[a] != g_rank[b]) return g_rank[a] < g_rank[b] ? -1 : 1;
    uint32_t ra = (a + g_h < g_n) ? g_rank[a + g_h] : 0;
    uint32_t rb = (b + g_h < g_n) ? g_rank[b + g_h] : 0;
    if (ra != rb) return ra < rb ? -1 : 1;
    return 0;
}

static uint64_t *suffix_sort(const uint8_t *t, uint64_t n)
{
    if (n >= 0xFFFFFFF0ULL) return NULL;
    uint32_t *sa = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *rank = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *tmp = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint64_t *out = (uint64_t *)malloc(n * sizeof(uint64_t));
    if (!sa || !rank || !tmp || !out) return NULL;
    /* initial rank: first 8 symbols, base 7 (0 = beyond the end, symbols shifted by one) */
    const uint32_t H0 = 8;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t v = 0;
        for (uint32_t k = 0; k < H0; ++k) v = v * 7u + ((i + k < n) ? (uint32_t)t[i + k] + 1u : 0u);
        rank[i] = v + 1;
        sa[i] = (uint32_t)i;
    }
    g_n = n;
    for (uint64_t h = H0;; h *= 2) {
        g_rank = rank;
        g_h = h;
        qsort(sa, n, sizeof(uint32_t), cmp_suffix);
        tmp[sa[0]] = 1;
        uint32_t r = 1;
        for (uint64_t i = 1; i < n; ++i) {
            if (cmp_suffix(&sa[i - 1], &sa[i]) != 0) ++r;
            tmp[sa[i]] = r;
        }
        memcpy(rank, tmp, n * sizeof(uint32_t));
        if (r == n || h >= n) break;
    }
    for (uint64_t i = 0; i < n; ++i) out[i] = sa[i];
    free(sa);
    free(rank);
    free(tmp);
    return out;
}

/* text with sentinels: symbols 0 = '$', 1..5 = A,C,G,T,N */
static uint8_t *with_sentinels(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq, int reversed,
                               uint64_t *n_out)
{
    uint64_t n = limits[n_seq] + n_seq;
    uint8_t *t = (uint8_t *)malloc(n ? n : 1);
    if (!t) return NULL;
    uint64_t o = 0;
    for (uint32_t s = 0; s < n_seq; ++s) {
        uint64_t b = limits[s], e = limits[s + 1];
        if (!reversed)
            for (uint64_t k = b; k < e; ++k) t[o++] = (uint8_t)(codes[k] + 1);
        else /* src/indexing.hpp:130 — every sequence reversed in place */
            for (uint64_t k = e; k > b; --k) t[o++] = (uint8_t)(codes[k - 1] + 1);
        t[o++] = 0;
    }
    *n_out = n;
    return t;
}

/* BWT from SA: src/seqan_libdivsufsort.h:165-229 (bwt[i] = text[sa[i]-1]; the row of a sequence
 * start holds a sentinel — the reference stores a substitute + marker, the value is what counts). */
static uint8_t *bwt_from_sa(const uint8_t *t, const uint64_t *sa, uint64_t n)
{
    uint8_t *bwt = (uint8_t *)malloc(n ? n : 1);
    if (!bwt) return NULL;
    for (uint64_t i = 0; i < n; ++i) bwt[i] = sa[i] ? t[sa[i] - 1] : t[n - 1];
    return bwt;
}

static gmo_index *index_alloc(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq)
{
    gmo_index *ix = (gmo_index *)calloc(1, sizeof(gmo_index));
    if (!ix) return NULL;
    ix->n_seq = n_seq;
    ix->n_text = limits[n_seq];
    ix->codes = (uint8_t *)malloc(ix->n_text ? ix->n_text : 1);
    ix->limits = (uint64_t *)malloc((n_seq + 1) * sizeof(uint64_t));
    ix->cum_sent = (uint64_t *)malloc((n_seq + 1) * sizeof(uint64_t));
    memcpy(ix->codes, codes, ix->n_text);
    memcpy(ix->limits, limits, (n_seq + 1) * sizeof(uint64_t));
    for (uint32_t i = 0; i <= n_seq; ++i) ix->cum_sent[i] = limits[i] + i;
    return ix;
}

gmo_index *gmo_index_build(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq)
{
    gmo_index *ix = index_alloc(codes, limits, n_seq);
    if (!ix) return NULL;
    uint32_t sigma = 4;
    for (uint64_t i = 0; i < ix->n_text; ++i)
        if (codes[i] >= GMO_N) { sigma = 5; break; } /* src/indexing.hpp:459-473 */
    for (int rev = 0; rev < 2; ++rev) {
        uint64_t n = 0;
        uint8_t *t = with_sentinels(codes, limits, n_seq, rev, &n);
        uint64_t *sa = suffix_sort(t, n);
        if (!sa) { free(t); gmo_index_free(ix); return NULL; }
        uint8_t *bwt = bwt_from_sa(t, sa, n);
        fm_init(rev ? &ix->rev : &ix->fwd, bwt, n, sigma);
        free(bwt);
        free(t);
        if (!rev) ix->sa = sa; else free(sa);
    }
    return ix;
}

gmo_index *gmo_index_from_bwt(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq,
                              const uint8_t *bwt_fwd, const uint8_t *bwt_rev, uint32_t sigma,
                              const uint64_t *sa)
{
    gmo_index *ix = index_alloc(codes, limits, n_seq);
    if (!ix) return NULL;
    uint64_t n = limits[n_seq] + n_seq;
    if (fm_init(&ix->fwd, bwt_fwd, n, sigma) || fm_init(&ix->rev, bwt_rev, n, sigma)) {
        gmo_index_free(ix);
        return NULL;
    }
    if (sa) {
        ix->sa = (uint64_t *)malloc(n * sizeof(uint64_t));
        memcpy(ix->sa, sa, n * sizeof(uint64_t));
    }
    return ix;
}

void gmo_index_free(gmo_index *ix)
{
    if (!ix) return;
    fm_free(&ix->fwd);
    fm_free(&ix->rev);
    free(ix->codes);
    free(ix->limits);
    free(ix->cum_sent);
    free(ix->sa);
    free(ix);
}

uint64_t gmo_index_bwt_len(const gmo_index *ix) { return ix->fwd.n; }
uint32_t gmo_index_sigma(const gmo_index *ix) { return ix->fwd.sigma; }

void gmo_index_get_bwt(const gmo_index *ix, int rev, uint8_t *out)
{
    const gmo_fm *fm = rev ? &ix->rev : &ix->fwd;
    for (uint64_t i = 0; i < fm->n; ++i)
        out[i] = (uint8_t)((fm->nib[(i >> 6) * 4 + ((i & 63) >> 4)] >> (4 * (i & 15))) & 15);
}

int gmo_index_get_sa(const gmo_index *ix, uint64_t *out)
{
    if (!ix->sa) return -1;
    memcpy(out, ix->sa, ix->fwd.n * sizeof(uint64_t));
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Bidirectional iterator.  fwd range = suffixes of T starting with P, rev range = suffixes of T'
 * starting with reverse(P); equal sizes (index_bidirectional_stree.h:217-265).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t flo, fhi, rlo, rhi; } gmo_it;
enum { DIR_FWD = 0 /* extend to the left, uses the forward BWT */,
       DIR_REV = 1 /* extend to the right, uses the BWT of the reversed text */ };

static inline void it_root(const gmo_index *ix, gmo_it *it)
{
    it->flo = it->rlo = 0;
    it->fhi = it->rhi = ix->fwd.n;
}

/* goDown(it, c, Dir): index_fm_stree.h:256-308 (_getNodeByChar: two LF queries + "smaller"),
 * index_fm_lf_table.h:468-491 (C[c] + rank, sentinels never counted as a base),
 * index_bifm_stree.h:46-58 (_update: shift the opposite range by `smaller`). */
static inline int it_down(const gmo_index *ix, gmo_it *it, uint32_t c, int dir)
{
    const gmo_fm *fm = dir == DIR_FWD ? &ix->fwd : &ix->rev;
    uint64_t lo = dir == DIR_FWD ? it->flo : it->rlo;
    uint64_t hi = dir == DIR_FWD ? it->fhi : it->rhi;
    uint32_t sym = c + 1;
    if (sym > fm->sigma) return 0;
    uint64_t rl = fm_rank(fm, sym, lo), rh = fm_rank(fm, sym, hi);
    if (rl >= rh) return 0;
    uint64_t smaller = 0;
    for (uint32_t x = 0; x < sym; ++x) smaller += fm_rank(fm, x, hi) - fm_rank(fm, x, lo);
    uint64_t nlo = fm->C[sym] + rl, nhi = fm->C[sym] + rh;
    if (dir == DIR_FWD) {
        it->flo = nlo; it->fhi = nhi;
        it->rlo += smaller; it->rhi = it->rlo + (nhi - nlo);
    } else {
        it->rlo = nlo; it->rhi = nhi;
        it->flo += smaller; it->fhi = it->flo + (nhi - nlo);
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Optimum search schemes, GenMap variant: tables src/find2_index_approx.hpp:67-134.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t nb;
    uint8_t pi[6], l[6], u[6];
    uint32_t bl[6]; /* cumulative block lengths in search order */
    uint32_t start;
} gmo_search;

static const struct { uint8_t ns, nb; uint8_t pi[7][6], l[7][6], u[7][6]; } SCHEMES[5] = {
    {1, 1, {{1}}, {{0}}, {{0}}},
    {2, 2, {{1, 2}, {2, 1}}, {{0, 0}, {0, 1}}, {{0, 1}, {0, 1}}},
    {3, 4, {{1, 2, 3, 4}, {3, 2, 1, 4}, {4, 3, 2, 1}},
           {{0, 0, 1, 1}, {0, 0, 0, 0}, {0, 0, 0, 2}},
           {{0, 0, 2, 2}, {0, 1, 1, 2}, {0, 1, 2, 2}}},
    {4, 5, {{1, 2, 3, 4, 5}, {2, 3, 4, 5, 1}, {3, 4, 5, 2, 1}, {5, 4, 3, 2, 1}},
           {{0, 0, 0, 0, 3}, {0, 0, 0, 2, 2}, {0, 0, 1, 1, 1}, {0, 0, 0, 0, 0}},
           {{0, 1, 2, 3, 3}, {0, 1, 2, 2, 3}, {0, 1, 1, 3, 3}, {0, 0, 3, 3, 3}}},
    {7, 6, {{1, 2, 3, 4, 5, 6}, {3, 4, 5, 6, 2, 1}, {2, 3, 4, 5, 6, 1}, {3, 2, 4, 5, 6, 1},
            {4, 3, 2, 5, 6, 1}, {4, 3, 2, 5, 6, 1}, {6, 5, 4, 3, 2, 1}},
           {{0, 0, 0, 0, 0, 4}, {0, 0, 0, 1, 4, 4}, {0, 0, 0, 0, 0, 0}, {0, 1, 1, 1, 1, 1},
            {0, 0, 2, 2, 2, 2}, {0, 1, 2, 2, 2, 2}, {0, 0, 0, 0, 3, 3}},
           {{0, 2, 3, 3, 4, 4}, {0, 0, 1, 1, 4, 4}, {0, 2, 2, 3, 3, 4}, {0, 1, 2, 3, 3, 4},
            {0, 0, 2, 3, 3, 4}, {0, 1, 2, 3, 3, 4}, {0, 0, 4, 4, 4, 4}}},
};

/* _optimalSearchSchemeComputeFixedBlocklengthGM + SetBlockLengthGM + InitGM
 * (src/find2_index_approx.hpp:139-176) */
static uint32_t scheme_make(uint32_t E, uint32_t needle_len, gmo_search *out)
{
    uint32_t ns = SCHEMES[E].ns, nb = SCHEMES[E].nb;
    uint32_t base = needle_len / nb, rest = needle_len - nb * base, len[6];
    for (uint32_t i = 0; i < nb; ++i) len[i] = base + (i < rest);
    for (uint32_t s = 0; s < ns; ++s) {
        gmo_search *S = &out[s];
        S->nb = (uint8_t)nb;
        for (uint32_t i = 0; i < nb; ++i) {
            S->pi[i] = SCHEMES[E].pi[s][i];
            S->l[i] = SCHEMES[E].l[s][i];
            S->u[i] = SCHEMES[E].u[s][i];
            S->bl[i] = len[S->pi[i] - 1] + (i ? S->bl[i - 1] : 0);
        }
        S->start = 0;
        for (uint32_t i = 0; i < nb; ++i)
            if (S->pi[i] < S->pi[0]) S->start += S->bl[i] - S->bl[i - 1];
    }
    return ns;
}

/* ------------------------------------------------------------------------------------------------
 * Per-block state (the std::vectors of src/algo.hpp:251-254)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t lo, hi; } gmo_range;
typedef struct { gmo_range *r; uint32_t n, cap; } gmo_rlist;

typedef struct {
    const gmo_index *ix;
    uint32_t K, E;
    const uint8_t *needles;  /* current strand, length nl = K + cnt - 1 */
    const uint8_t *infix;    /* needles + K - infix_len */
    uint32_t infix_len;
    uint64_t bb;             /* last needle index */
    uint64_t *hits;          /* per window of the current strand orientation */
    uint64_t maxv;
    int report_exact;        /* reportExactMatch template flag */
    gmo_range *it_exact;     /* per window */
    uint8_t *has_exact;
    gmo_rlist *it_all;       /* per window, current strand; NULL unless csvComputation */
} blk_ctx;

static void rlist_push(gmo_rlist *l, uint64_t lo, uint64_t hi)
{
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 4;
        l->r = (gmo_range *)realloc(l->r, l->cap * sizeof(gmo_range));
    }
    l->r[l->n].lo = lo;
    l->r[l->n].hi = hi;
    l->n++;
}

/* the full-length branch shared by extendExact (src/algo.hpp:38-50) and extend (:180-193) */
static inline void report_hit(blk_ctx *cx, const gmo_it *it, uint64_t a, uint32_t errors_left)
{
    if (cx->report_exact && errors_left == cx->E) {
        cx->it_exact[a].lo = it->flo;
        cx->it_exact[a].hi = it->fhi;
        cx->has_exact[a] = 1;
    }
    if (cx->it_all) rlist_push(&cx->it_all[a], it->flo, it->fhi);
    uint64_t v = (it->fhi - it->flo) + cx->hits[a];
    cx->hits[a] = v < cx->maxv ? v : cx->maxv;
}

/* extendExact: src/algo.hpp:26-79 */
static void extend_exact(blk_ctx *cx, gmo_it it, uint64_t a, uint64_t b)
{
    const uint32_t K = cx->K;
    if (b - a + 1 == K) {
        report_hit(cx, &it, a, 0);
        return;
    }
    gmo_it it2 = it;
    uint64_t brm = a + K - 1;
    uint64_t b_new = b + (((brm - b) + 2 - 1) >> 1);
    if (b_new <= cx->bb) {
        int ok = 1;
        for (uint64_t i = b + 1; i <= b_new && ok; ++i)
            ok = cx->needles[i] != GMO_N && it_down(cx->ix, &it2, cx->needles[i], DIR_REV);
        if (ok) extend_exact(cx, it2, a, b_new);
    }
    if (a >= 1) {
        int64_t alm = (int64_t)b + 1 - (int64_t)K;
        int64_t half = (((int64_t)a - alm) - 1) >> 1;
        uint64_t a_new = (uint64_t)(alm + (half > 0 ? half : 0));
        for (int64_t i = (int64_t)a - 1; i >= (int64_t)a_new; --i)
            if (cx->needles[i] == GMO_N || !it_down(cx->ix, &it, cx->needles[i], DIR_FWD)) return;
        extend_exact(cx, it, a_new, b);
    }
}

static void extend(blk_ctx *cx, gmo_it it, uint32_t errors_left, uint64_t a, uint64_t b);

/* approxSearch, both directions: src/algo.hpp:90-163.  The children loop restates
 * goDown(it,Dir) / goRight(it,Dir) (index_fm_stree.h:323-341,398-430): all symbols in alphabet
 * order, empty ranges skipped. */
static void approx_search(blk_ctx *cx, gmo_it it, uint32_t errors_left, uint64_t a, uint64_t b,
                          uint64_t target, int dir)
{
    if ((dir == DIR_REV && b == target) || (dir == DIR_FWD && a == target)) {
        extend(cx, it, errors_left, a, b);
        return;
    }
    if (errors_left > 0) {
        uint8_t nc = dir == DIR_REV ? cx->needles[b + 1] : cx->needles[a - 1];
        for (uint32_t c = 0; c < cx->ix->fwd.sigma; ++c) {
            gmo_it ch = it;
            if (!it_down(cx->ix, &ch, c, dir)) continue;
            uint32_t delta = (c != nc) || (nc == GMO_N);
            if (dir == DIR_REV) approx_search(cx, ch, errors_left - delta, a, b + 1, target, dir);
            else                approx_search(cx, ch, errors_left - delta, a - 1, b, target, dir);
        }
    } else if (dir == DIR_REV) {
        for (uint64_t i = b + 1; i <= target; ++i)
            if (cx->needles[i] == GMO_N || !it_down(cx->ix, &it, cx->needles[i], DIR_REV)) return;
        extend_exact(cx, it, a, target);
    } else {
        for (int64_t i = (int64_t)a - 1; i >= (int64_t)target; --i)
            if (cx->needles[i] == GMO_N || !it_down(cx->ix, &it, cx->needles[i], DIR_FWD)) return;
        extend_exact(cx, it, target, b);
    }
}

/* extend: src/algo.hpp:165-218 */
static void extend(blk_ctx *cx, gmo_it it, uint32_t errors_left, uint64_t a, uint64_t b)
{
    const uint32_t K = cx->K;
    if (errors_left == 0) {
        extend_exact(cx, it, a, b);
        return;
    }
    if (b - a + 1 == K) {
        report_hit(cx, &it, a, errors_left);
        return;
    }
    uint64_t brm = a + K - 1;
    uint64_t b_new = b + (((brm - b) + 2 - 1) >> 1);
    if (b_new <= cx->bb) approx_search(cx, it, errors_left, a, b, b_new, DIR_REV);
    if (a >= 1) {
        int64_t alm = (int64_t)b + 1 - (int64_t)K;
        int64_t half = (((int64_t)a - alm) - 1) >> 1;
        uint64_t a_new = (uint64_t)(alm + (half > 0 ? half : 0));
        approx_search(cx, it, errors_left, a, b, a_new, DIR_FWD);
    }
}

/* the delegate lambdas of src/algo.hpp:262-298 */
static inline void delegate(blk_ctx *cx, const gmo_it *it, uint32_t errors_spent, int is_fwd_strand)
{
    cx->report_exact = is_fwd_strand && errors_spent == 0;
    extend(cx, *it, cx->E - errors_spent, cx->K - cx->infix_len, cx->K - 1);
}

static void scheme_rec(blk_ctx *cx, gmo_it it, int32_t left, uint32_t right, uint32_t errors,
                       const gmo_search *s, uint32_t bi, int dir, int is_fwd_strand);

/* _optimalSearchSchemeChildrenGM (Hamming): src/find2_index_approx.hpp:223-301 */
static void scheme_children(blk_ctx *cx, gmo_it it, int32_t left, uint32_t right, uint32_t errors,
                            const gmo_search *s, uint32_t bi, uint32_t min_err, int dir, int is_fwd_strand)
{
    const int go_right = dir == DIR_REV;
    uint8_t nc = cx->infix[go_right ? right - 1 : (uint32_t)left - 1];
    uint32_t chars_left = s->bl[bi] - (right - (uint32_t)left - 1);
    for (uint32_t c = 0; c < cx->ix->fwd.sigma; ++c) {
        gmo_it ch = it;
        if (!it_down(cx->ix, &ch, c, dir)) continue;
        uint32_t delta = (c != nc) || (nc == GMO_N);
        if (min_err > 0 && chars_left + delta < min_err + 1u) continue; /* :254-258 */
        int32_t left2 = left - !go_right;
        uint32_t right2 = right + go_right;
        if (right - (uint32_t)left == s->bl[bi]) {
            uint32_t bi2 = bi + 1 < s->nb ? bi + 1 : s->nb - 1u;
            int go_right2 = s->pi[bi2] > s->pi[bi2 - 1];
            scheme_rec(cx, ch, left2, right2, errors + delta, s, bi2, go_right2 ? DIR_REV : DIR_FWD,
                       is_fwd_strand);
        } else {
            scheme_rec(cx, ch, left2, right2, errors + delta, s, bi, dir, is_fwd_strand);
        }
    }
}

/* _optimalSearchSchemeExactGM: src/find2_index_approx.hpp:303-369 */
static void scheme_exact(blk_ctx *cx, gmo_it it, int32_t left, uint32_t right, uint32_t errors,
                         const gmo_search *s, uint32_t bi, int dir, int is_fwd_strand)
{
    int go_right2 = (bi + 1 < s->nb) && s->pi[bi + 1] > s->pi[bi];
    uint32_t bi2 = bi + 1 < s->nb ? bi + 1 : s->nb - 1u;
    if (dir == DIR_REV) {
        uint32_t pl = right - 1, pr = (uint32_t)left + s->bl[bi] - 1;
        while (pl <= pr) {
            if (cx->infix[pl] == GMO_N || !it_down(cx->ix, &it, cx->infix[pl], DIR_REV)) return;
            ++pl;
        }
        scheme_rec(cx, it, left, pr + 2, errors, s, bi2, go_right2 ? DIR_REV : DIR_FWD, is_fwd_strand);
    } else {
        int32_t pl = (int32_t)right - (int32_t)s->bl[bi] - 1, pr = left - 1;
        while (pl <= pr) {
            if (cx->infix[pr] == GMO_N || !it_down(cx->ix, &it, cx->infix[pr], DIR_FWD)) return;
            --pr;
        }
        scheme_rec(cx, it, pl, right, errors, s, bi2, go_right2 ? DIR_REV : DIR_FWD, is_fwd_strand);
    }
}

/* _optimalSearchSchemeGM: src/find2_index_approx.hpp:371-428 (HammingDistance only — the
 * EditDistance branches are never instantiated by GenMap, src/algo.hpp:301,308) */
static void scheme_rec(blk_ctx *cx, gmo_it it, int32_t left, uint32_t right, uint32_t errors,
                       const gmo_search *s, uint32_t bi, int dir, int is_fwd_strand)
{
    uint32_t max_err = (uint32_t)s->u[bi] - errors;
    uint32_t min_err = s->l[bi] > errors ? s->l[bi] - errors : 0;
    if (min_err == 0 && left == 0 && right == cx->infix_len + 1)
        delegate(cx, &it, errors, is_fwd_strand);
    else if (max_err == 0 && right - (uint32_t)left - 1 != s->bl[bi])
        scheme_exact(cx, it, left, right, errors, s, bi, dir, is_fwd_strand);
    else
        scheme_children(cx, it, left, right, errors, s, bi, min_err, dir, is_fwd_strand);
}

/* ------------------------------------------------------------------------------------------------
 * computeMappabilitySingleBlock: src/algo.hpp:221-403
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const gmo_index *ix;
    const gmo_params *p;
    const uint8_t *text;     /* file text = codes + text_begin */
    uint64_t text_len;
    uint32_t infix_len;      /* params.overlap in the reference (length of the common infix) */
    int value16;
    void *c;
    const uint32_t *seq_to_file;
    int csv_computation;     /* opt.csvFile || excludePseudo (src/mappability.hpp:172) */
    int copy_ok;             /* !directory && SA present && requested */
} map_ctx;

static inline uint64_t c_get(const map_ctx *m, uint64_t i)
{
    return m->value16 ? ((const uint16_t *)m->c)[i] : ((const uint8_t *)m->c)[i];
}
static inline void c_set(const map_ctx *m, uint64_t i, uint64_t v)
{
    if (m->value16) ((uint16_t *)m->c)[i] = (uint16_t)v; else ((uint8_t *)m->c)[i] = (uint8_t)v;
}

/* locate: CompressedSA::value (index_fm_compressed_sa.h:478-513) restated on the full SA;
 * returns the sequence number (Pair.i1) of a row and its in-sequence offset (Pair.i2). */
static inline uint32_t locate_seq(const gmo_index *ix, uint64_t row, uint64_t *pos_in_seq)
{
    uint64_t p = ix->sa[row];
    uint32_t lo = 0, hi = ix->n_seq; /* largest s with cum_sent[s] <= p */
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) / 2;
        if (ix->cum_sent[mid] <= p) lo = mid; else hi = mid;
    }
    *pos_in_seq = p - ix->cum_sent[lo];
    return lo;
}

static void single_block(const map_ctx *m, uint64_t i, uint64_t j, int complete_same_kmers, int have_intervals)
{
    const uint32_t K = m->p->K, E = m->p->E;
    uint64_t max_pos = i + K - m->infix_len;
    if (max_pos > m->text_len - K) max_pos = m->text_len - K;
    max_pos += 1;
    if (max_pos > j) max_pos = j;

    uint64_t begin = i;
    while (begin < max_pos && c_get(m, begin) != 0) ++begin; /* :236-238 */
    uint64_t end = max_pos;
    while (i > 0 && end >= 1 && end - 1 >= i && c_get(m, end - 1) != 0) --end; /* :240-242 */
    if (begin >= end) return;

    uint32_t cnt = (uint32_t)(end - begin);
    uint32_t infix_len = K - cnt + 1; /* :246 */
    gmo_search scheme[7];
    uint32_t ns = scheme_make(E, infix_len, scheme);

    uint32_t nl = K + cnt - 1;
    uint8_t *needles_rc = (uint8_t *)malloc(nl);
    uint64_t *hits = (uint64_t *)calloc(cnt, sizeof(uint64_t));
    gmo_range *it_exact = (gmo_range *)calloc(cnt, sizeof(gmo_range));
    uint8_t *has_exact = (uint8_t *)calloc(cnt, 1);
    gmo_rlist *all_f = NULL, *all_r = NULL;
    if (m->csv_computation) {
        all_f = (gmo_rlist *)calloc(cnt, sizeof(gmo_rlist));
        all_r = (gmo_rlist *)calloc(cnt, sizeof(gmo_rlist));
    }

    blk_ctx cx;
    cx.ix = m->ix; cx.K = K; cx.E = E;
    cx.infix_len = infix_len;
    cx.bb = nl - 1; /* :260 */
    cx.hits = hits;
    cx.maxv = m->value16 ? 65535u : 255u;
    cx.it_exact = it_exact; cx.has_exact = has_exact;
    cx.report_exact = 0;

    gmo_it root;
    if (m->p->revcompl) { /* :284-305 */
        const uint8_t *nd = m->text + begin;
        for (uint32_t k = 0; k < nl; ++k) {
            uint8_t ch = nd[nl - 1 - k];
            needles_rc[k] = ch < GMO_N ? (uint8_t)(3 - ch) : ch;
        }
        cx.needles = needles_rc;
        cx.infix = needles_rc + (K - infix_len);
        cx.it_all = all_r;
        for (uint32_t s = 0; s < ns; ++s) {
            it_root(m->ix, &root);
            scheme_rec(&cx, root, (int32_t)scheme[s].start, scheme[s].start + 1, 0, &scheme[s], 0, DIR_REV, 0);
        }
        for (uint32_t k = 0; k < cnt / 2; ++k) { /* std::reverse(hits) :304 */
            uint64_t t = hits[k]; hits[k] = hits[cnt - 1 - k]; hits[cnt - 1 - k] = t;
        }
    }
    cx.needles = m->text + begin;
    cx.infix = cx.needles + (K - infix_len);
    cx.it_all = all_f;
    for (uint32_t s = 0; s < ns; ++s) { /* :307-308 */
        it_root(m->ix, &root);
        scheme_rec(&cx, root, (int32_t)scheme[s].start, scheme[s].start + 1, 0, &scheme[s], 0, DIR_REV, 1);
    }

    for (uint64_t q = begin; q < end; ++q) { /* :309-401 */
        uint32_t w = (uint32_t)(q - begin);
        if (m->csv_computation && m->p->exclude_pseudo) { /* :351-361 */
            uint8_t seen[8192] = {0}; /* file ids < 65536 */
            uint32_t distinct = 0;
            const gmo_rlist *ls[2] = { &all_f[w], &all_r[cnt - 1 - w] }; /* :340 reversed order */
            for (int st = 0; st < 2; ++st)
                for (uint32_t r = 0; r < ls[st]->n; ++r)
                    for (uint64_t row = ls[st]->r[r].lo; row < ls[st]->r[r].hi; ++row) {
                        uint64_t pos;
                        uint32_t f = m->seq_to_file[locate_seq(m->ix, row, &pos)];
                        if (!(seen[f >> 3] & (1u << (f & 7)))) { seen[f >> 3] |= (uint8_t)(1u << (f & 7)); ++distinct; }
                    }
            hits[w] = distinct; /* :360 */
        }
        if (m->copy_ok && (!have_intervals || complete_same_kmers) && has_exact[w] &&
            it_exact[w].hi - it_exact[w].lo > 1) { /* :389-396 */
            for (uint64_t row = it_exact[w].lo; row < it_exact[w].hi; ++row) {
                uint64_t pos;
                uint32_t s = locate_seq(m->ix, row, &pos);
                c_set(m, m->ix->limits[s] + pos, hits[w]); /* posGlobalize; single file => local == global */
            }
        } else {
            c_set(m, q, hits[w]);
        }
    }
    if (m->csv_computation) {
        for (uint32_t k = 0; k < cnt; ++k) { free(all_f[k].r); free(all_r[k].r); }
        free(all_f); free(all_r);
    }
    free(needles_rc); free(hits); free(it_exact); free(has_exact);
}

static uint32_t default_infix_len(uint32_t K, uint32_t E)
{
    /* src/mappability.hpp:519-543: user-facing overlap, clamped, then converted to infix length */
    uint64_t overlap;
    if (E == 0) overlap = (uint64_t)(K * 0.7);
    else {
        uint32_t kk = K < 30u ? 30u : (K > 100u ? 100u : K);
        float pw = 1.0f;
        for (uint32_t i = 0; i < E; ++i) pw *= 0.7f;
        overlap = (uint64_t)(K * kk * pw / 100.0);
    }
    uint64_t a = K - 1, b = (uint64_t)K - E - 2; /* unsigned arithmetic as in the reference */
    uint64_t max_overlap = a < b ? a : b;
    if (overlap > max_overlap) overlap = max_overlap;
    return (uint32_t)(K - overlap);
}

/* resetLimits: src/algo.hpp:10-22 */
static void reset_limits(const map_ctx *m, const uint64_t *cum, uint32_t n_chrom)
{
    for (uint32_t i = 1; i <= n_chrom; ++i) {
        uint64_t len1 = cum[i] - cum[i - 1] + 1;
        uint64_t lim = m->p->K < len1 ? m->p->K : len1;
        for (uint64_t j = 1; j < lim; ++j) c_set(m, cum[i] - j, 0);
    }
}

int gmo_map(const gmo_index *ix, const gmo_params *p, uint64_t text_begin, uint64_t text_len,
            const uint64_t *chrom_cum, uint32_t n_chrom, const uint64_t *intervals,
            uint64_t n_intervals, const uint32_t *seq_to_file, void *out)
{
    if (p->E > 4) return -2;                        /* src/mappability.hpp:187 */
    if (p->K < p->E + 2) return -3;                 /* undefined in the reference */
    if (p->value_bits != 8 && p->value_bits != 16) return -4;
    if (text_begin + text_len > ix->n_text) return -5;
    if (p->exclude_pseudo && (!ix->sa || !seq_to_file)) return -6;
    map_ctx m;
    m.ix = ix; m.p = p;
    m.text = ix->codes + text_begin;
    m.text_len = text_len;
    m.value16 = p->value_bits == 16;
    m.c = out;
    m.seq_to_file = seq_to_file;
    m.csv_computation = p->exclude_pseudo != 0;
    m.copy_ok = p->copy_shortcut && ix->sa != NULL && text_begin == 0 && text_len == ix->n_text;
    m.infix_len = p->infix_len ? p->infix_len : default_infix_len(p->K, p->E);
    if (m.infix_len > p->K || m.infix_len < SCHEMES[p->E].nb) return -7;
    memset(out, 0, text_len * (p->value_bits / 8));
    if (text_len < p->K) return 0; /* the reference's unsigned numberOfKmers would wrap; nothing to search */

    uint64_t n_kmers = text_len - p->K + 1;
    uint64_t step = p->K - m.infix_len + 1; /* src/algo.hpp:416 */
    int threads = (int)p->threads;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    if (n_intervals == 0) { /* :420-440 */
        int64_t n_blocks = (int64_t)((n_kmers + step - 1) / step);
        int64_t chunk = (int64_t)(n_kmers / (step * (uint64_t)threads * 50));
        if (chunk < 1) chunk = 1;
        #pragma omp parallel for schedule(dynamic, chunk) num_threads(threads)
        for (int64_t b = 0; b < n_blocks; ++b) {
            uint64_t i = (uint64_t)b * step;
            single_block(&m, i, i + step, 1, 0);
        }
    } else { /* :441-476 */
        uint64_t sum = 0, n_det = 0, cap = 0;
        uint64_t *det = NULL;
        for (uint64_t k = 0; k < n_intervals; ++k) {
            uint64_t b = intervals[2 * k], e = intervals[2 * k + 1];
            sum += e - b;
            for (uint64_t i = b; i < e; i += step) {
                if (n_det == cap) { cap = cap ? cap * 2 : 64; det = (uint64_t *)realloc(det, cap * 2 * sizeof(uint64_t)); }
                det[2 * n_det] = i;
                det[2 * n_det + 1] = i + step < e ? i + step : e;
                ++n_det;
            }
        }
        float fraction = (float)sum / (float)text_len;
        int complete = fraction > 0.5f;
        int64_t chunk = (int64_t)(n_det / ((uint64_t)threads * 50));
        if (chunk < 1) chunk = 1;
        #pragma omp parallel for schedule(dynamic, chunk) num_threads(threads)
        for (int64_t d = 0; d < (int64_t)n_det; ++d) {
            if (det[2 * d] + p->K > text_len) continue; /* maxPos would precede i: nothing to compute */
            single_block(&m, det[2 * d], det[2 * d + 1], complete, 1);
        }
        free(det);
        if (complete) { /* outputMappability's re-zeroing, src/mappability.hpp:83-99 */
            uint64_t last_end = 0;
            for (uint64_t k = 0; k < n_intervals; ++k) {
                for (uint64_t i = last_end; i < intervals[2 * k] && i < text_len; ++i) c_set(&m, i, 0);
                last_end = intervals[2 * k + 1];
            }
            for (uint64_t i = last_end; i < text_len; ++i) c_set(&m, i, 0);
        }
    }
    reset_limits(&m, chrom_cum, n_chrom);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Brute force (the reference's own unit test does the same with a trivial
 * backtracker, tests/tests.cpp:104-131,195).
 * ---------------------------------------------------------------------------------------------- */
int gmo_brute(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq, const gmo_params *p,
              uint64_t text_begin, uint64_t text_len, const uint64_t *chrom_cum, uint32_t n_chrom,
              const uint64_t *intervals, uint64_t n_intervals, const uint32_t *seq_to_file, void *out)
{
    const uint32_t K = p->K, E = p->E;
    if (p->value_bits != 8 && p->value_bits != 16) return -4;
    const int v16 = p->value_bits == 16;
    const uint64_t maxv = v16 ? 65535u : 255u;
    memset(out, 0, text_len * (p->value_bits / 8));
    if (text_len < K) return 0;
    uint8_t *sel = NULL;
    if (n_intervals) {
        sel = (uint8_t *)calloc(text_len, 1);
        for (uint64_t k = 0; k < n_intervals; ++k)
            for (uint64_t i = intervals[2 * k]; i < intervals[2 * k + 1] && i < text_len; ++i) sel[i] = 1;
    }
    uint8_t *pat = (uint8_t *)malloc(K);
    for (uint64_t j = 0; j + K <= text_len; ++j) {
        if (sel && !sel[j]) continue;
        uint64_t total = 0;
        uint8_t seen[8192];
        if (p->exclude_pseudo) memset(seen, 0, sizeof seen);
        uint32_t distinct = 0;
        for (int strand = 0; strand < (p->revcompl ? 2 : 1); ++strand) {
            for (uint32_t k = 0; k < K; ++k) {
                uint8_t ch = strand ? codes[text_begin + j + K - 1 - k] : codes[text_begin + j + k];
                pat[k] = (strand && ch < GMO_N) ? (uint8_t)(3 - ch) : ch;
            }
            for (uint32_t s = 0; s < n_seq; ++s) {
                uint64_t b = limits[s], e = limits[s + 1];
                for (uint64_t q = b; q + K <= e; ++q) {
                    uint32_t mm = 0;
                    for (uint32_t k = 0; k < K && mm <= E; ++k)
                        mm += (pat[k] == GMO_N) || (codes[q + k] != pat[k]); /* pattern N never matches */
                    if (mm <= E) {
                        ++total;
                        if (p->exclude_pseudo) {
                            uint32_t f = seq_to_file[s];
                            if (!(seen[f >> 3] & (1u << (f & 7)))) { seen[f >> 3] |= (uint8_t)(1u << (f & 7)); ++distinct; }
                        }
                    }
                }
            }
        }
        uint64_t v = p->exclude_pseudo ? distinct : (total < maxv ? total : maxv);
        if (v16) ((uint16_t *)out)[j] = (uint16_t)v; else ((uint8_t *)out)[j] = (uint8_t)v;
    }
    for (uint32_t i = 1; i <= n_chrom; ++i) { /* tails: src/algo.hpp:10-22 */
        uint64_t len1 = chrom_cum[i] - chrom_cum[i - 1] + 1;
        uint64_t lim = K < len1 ? K : len1;
        for (uint64_t j = 1; j < lim; ++j) {
            if (v16) ((uint16_t *)out)[chrom_cum[i] - j] = 0; else ((uint8_t *)out)[chrom_cum[i] - j] = 0;
        }
    }
    free(pat);
    free(sel);
    return 0;
}

/* Definition-level csv lists (the csvComputation branch, src/algo.hpp:311-346): every occurrence with <= E
 * mismatches of the k-mer at concatenated-text position `pos` (strand 0) or of its reverse complement
 * (strand 1: the "- strand" column), as (sequence, offset) pairs in the order std::sort gives them (:335,:346).
 * The k-mer is read from the concatenated text like the reference does (it may span two sequences; the caller
 * decides which positions are emitted, :377-385).  Returns the number of occurrences; at most `cap` are written. */
uint64_t gmo_brute_locations(const uint8_t *codes, const uint64_t *limits, uint32_t n_seq, uint32_t K, uint32_t E,
                             uint64_t pos, int strand, uint32_t *seq_out, uint32_t *pos_out, uint64_t cap)
{
    uint8_t pat[256];
    uint64_t n = 0;
    if (K > 255 || pos + K > limits[n_seq]) return 0;
    for (uint32_t k = 0; k < K; ++k) {
        uint8_t ch = strand ? codes[pos + K - 1 - k] : codes[pos + k];
        pat[k] = (strand && ch < GMO_N) ? (uint8_t)(3 - ch) : ch;
    }
    for (uint32_t s = 0; s < n_seq; ++s)
        for (uint64_t q = limits[s]; q + K <= limits[s + 1]; ++q) {
            uint32_t mm = 0;
            for (uint32_t k = 0; k < K && mm <= E; ++k)
                mm += (pat[k] == GMO_N) || (codes[q + k] != pat[k]); /* a pattern N never matches (src/algo.hpp:111-112) */
            if (mm <= E) {
                if (n < cap) { seq_out[n] = s; pos_out[n] = (uint32_t)(q - limits[s]); }
                ++n;
            }
        }
    return n;
}
