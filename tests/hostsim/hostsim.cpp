// hostsim.cpp — TEST INFRASTRUCTURE.  Compiles the kernel's search state machine (gmb_core.h) and the
// host index builder for the CPU, so that `pytest -m "not gpu"` can check the logic against the oracle
// and count rank-block fetches without a GPU.  It is never linked into libgenmap_b200.so and is not
// reachable from the product API: the product path has no CPU fallback.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "../../genmap_b200/csrc/gmb_core.h"
#include "../../genmap_b200/csrc/gmb_host.h"

using namespace gmb;

namespace {
struct HostFrames {
    uint32_t w[kMaxE][kFrameWords5];
    uint32_t x[kLeafWords + 3 * kMaxBlockKmers];
    inline void set(uint32_t lv, uint32_t i, uint32_t v) { w[lv][i] = v; }
    inline uint32_t get(uint32_t lv, uint32_t i) const { return w[lv][i]; }
    inline void xset(uint32_t i, uint32_t v) { x[i] = v; }
    inline uint32_t xget(uint32_t i) const { return x[i]; }
    inline void cset(uint32_t i, uint32_t v) { x[i] = v; }
    inline uint32_t cget(uint32_t i) const { return x[i]; }
    inline void cadd(uint32_t i, uint32_t n, uint32_t maxv) { const uint64_t sum = (uint64_t)x[i] + n; x[i] = sum < maxv ? (uint32_t)sum : maxv; }
    inline void cor(uint32_t i, uint32_t bits) { x[i] |= bits; }
};

template <int KW, bool EP, bool BLK, int SIGMA>
void run_ranges(const MapCtx& cx, const uint64_t* text, const uint64_t* nmask, uint64_t text_begin,
                const std::vector<WorkRange>& ranges, int value_bits, void* out, FetchStats* fetches,
                unsigned long long* lut_reads)
{
    for (const WorkRange& r : ranges)
        for (uint64_t j0 = r.begin; j0 < r.end; j0 += cx.B) { // blocks of up to B adjacent k-mers, never across ranges
            Chain<KW, SIGMA> st;
            HostFrames fr;
            st.cnt = (uint32_t)std::min<uint64_t>(cx.B, r.end - j0);
            load_pattern(st.pat, text, nmask, text_begin + j0, cx.K + st.cnt - 1);
            chain_begin_block<KW, EP, BLK, SIGMA>(st, fr, cx, lut_reads);
            while (chain_step<KW, EP, BLK, SIGMA>(st, fr, cx, fetches, lut_reads)) {}
            for (uint32_t w = 0; w < st.cnt; ++w) {
                const uint32_t v = chain_result<KW, EP, BLK, SIGMA>(st, fr, cx, w);
                if (value_bits == 16) static_cast<uint16_t*>(out)[j0 + w] = (uint16_t)v;
                else static_cast<uint8_t*>(out)[j0 + w] = (uint8_t)v;
            }
        }
}
// The driver of block_kernel.cu on the host, one block at a time: per strand the keys of the flat list, empty
// entries skipped, located ones verified, the others walked in subtree mode (chain_step<..., SUB>).
template <int KW, bool EP, int SIGMA>
void run_ranges_blockdriver(const MapCtx& cx, const KeyLists& kl, const uint64_t* text, const uint64_t* nmask, uint64_t text_begin,
                            const std::vector<WorkRange>& ranges, int value_bits, void* out, FetchStats* fetches,
                            unsigned long long* lut_reads)
{
    for (const WorkRange& r : ranges)
        for (uint64_t j0 = r.begin; j0 < r.end; j0 += cx.B) {
            Chain<KW, SIGMA> st;
            HostFrames fr;
            st.has_n = false; st.acc = 0; st.files = 0; st.var = 0; st.sub = 0; st.nsub = 1;
            const uint32_t cnt = (uint32_t)std::min<uint64_t>(cx.B, r.end - j0), NL = cx.K + cnt - 1;
            st.cnt = cnt;
            load_pattern(st.pat, text, nmask, text_begin + j0, NL);
            if (SIGMA == 5) st.has_n = st.pat.has_n();
            for (uint32_t w = 0; w < cnt * (EP ? 3u : 1u); ++w) fr.cset(kLeafWords + w, 0u);
            // Dna5: an N in the common infix is an N in every window of the block — nothing to search (the windows with
            // 1..E N get their counts from the N pass, the others have none)
            const bool dead = SIGMA == 5 && st.pat.has_n_in(cnt - 1, cx.K - cnt + 1);
            for (uint32_t strand = 0; strand < (dead ? 0u : cx.n_strands); ++strand) {
                st.strand = strand;
                if (strand == 1) st.pat.reverse_complement(NL);
                for (uint32_t g = 0; g < kl.n[cnt]; ++g) {
                    const uint32_t x = kl.xy[2 * (kl.off[cnt] + g)], y = kl.xy[2 * (kl.off[cnt] + g) + 1];
                    st.s = y & 7u;
                    const SearchStart& S = cx.starts[cnt * kMaxSearches + st.s];
                    const uint32_t key = st.pat.bits(S.a, S.d) ^ x;
                    uint32_t pad;
                    jump_lookup(S, key, st.lo_f, st.lo_r, st.size, pad);
                    if (lut_reads) ++*lut_reads;
                    if (st.size == 0) continue;
                    if (st.size & kLocated) {
                        verify_located_key<KW, EP, true, SIGMA>(st, fr, cx, fetches, S, key, (y >> 8) & 1u, st.lo_r, st.lo_f, pad);
                        continue;
                    }
                    st.e = (y >> 4) & 7u; st.t = S.d; st.lvmask = 0; st.win = kNoWin; st.leaf_e = 0; st.thin = false;
                    while (chain_step<KW, EP, true, SIGMA, HostFrames, false, true>(st, fr, cx, fetches, nullptr)) {}
                }
            }
            for (uint32_t w = 0; w < cnt; ++w) {
                const uint32_t v = chain_result<KW, EP, true, SIGMA>(st, fr, cx, w);
                if (value_bits == 16) static_cast<uint16_t*>(out)[j0 + w] = (uint16_t)v;
                else static_cast<uint8_t*>(out)[j0 + w] = (uint8_t)v;
            }
        }
}

// The N pass of a Dna5 call whose searches skip the text's N (MapCtx::skip_n), by brute force: the product locates the
// text windows with 1..E N through the index (capi.cu: NFix) — here every such window is compared with every query.
//   * a query window with 1..E N: its whole count (every text window, N mismatching everything), overwriting;
//   * a query window without N: + 1 per text window with 1..E N it matches with <= E mismatches, per strand.
void nfix_bruteforce(const IndexHeader& h, const uint8_t* base, uint32_t K, uint32_t E, bool revcompl, uint64_t text_begin,
                     const std::vector<WorkRange>& ranges, int value_bits, void* out)
{
    const uint64_t* text = reinterpret_cast<const uint64_t*>(base + h.off_text);
    const uint64_t* nmask = reinterpret_cast<const uint64_t*>(base + h.off_nmask);
    const uint32_t* seq_start = reinterpret_cast<const uint32_t*>(base + h.off_seq_start);
    const uint64_t n = h.n_text;
    std::vector<uint8_t> c(n);
    for (uint64_t i = 0; i < n; ++i)
        c[i] = ((nmask[i >> 6] >> (i & 63)) & 1ull) ? 4 : (uint8_t)((text[i >> 5] >> (2 * (i & 31))) & 3ull);
    std::vector<uint32_t> nn(n + 1, 0); // prefix counts of N
    for (uint64_t i = 0; i < n; ++i) nn[i + 1] = nn[i] + (c[i] == 4);
    std::vector<uint64_t> starts; // every window start inside one sequence
    std::vector<uint64_t> nwin;   // ... whose window holds 1..E N
    for (uint32_t q = 0; q < h.n_seq; ++q) {
        const uint64_t b = (uint64_t)seq_start[q] - q, e = (uint64_t)seq_start[q + 1] - (q + 1);
        for (uint64_t t = b; t + K <= e; ++t) {
            starts.push_back(t);
            const uint32_t k = nn[t + K] - nn[t];
            if (k >= 1 && k <= E) nwin.push_back(t);
        }
    }
    const uint32_t maxv = value_bits == 16 ? 65535u : 255u;
    auto matches = [&](uint64_t j, uint64_t t, bool rc) { // query window j (its reverse complement) against text window t
        uint32_t e = 0;
        for (uint32_t i = 0; i < K && e <= E; ++i) {
            const uint8_t a = rc ? c[j + K - 1 - i] : c[j + i], b = c[t + i];
            e += a == 4 || b == 4 || (rc ? 3 - a : a) != b;
        }
        return e <= E;
    };
    for (const WorkRange& r : ranges)
        for (uint64_t j0 = r.begin; j0 < r.end; ++j0) {
            const uint64_t j = text_begin + j0;
            const uint32_t k = nn[j + K] - nn[j];
            if (k > E) continue;
            uint64_t v = 0;
            if (k >= 1) {
                for (uint64_t t : starts) v += matches(j, t, false) + (revcompl && matches(j, t, true));
            } else {
                v = value_bits == 16 ? static_cast<uint16_t*>(out)[j0] : static_cast<uint8_t*>(out)[j0];
                for (uint64_t t : nwin) v += matches(j, t, false) + (revcompl && matches(j, t, true));
            }
            if (v > maxv) v = maxv;
            if (value_bits == 16) static_cast<uint16_t*>(out)[j0] = (uint16_t)v;
            else static_cast<uint8_t*>(out)[j0] = (uint8_t)v;
        }
}
// the locate instantiation (csv lists): counting pass, prefix sums, filling pass, per-list sort — the
// same sequence locate_kernel.cu / gmb_map_locations run on the device
template <int KW, int SIGMA>
void run_locate(MapCtx cx, const uint64_t* text, const uint64_t* nmask, uint64_t text_begin, const std::vector<WorkRange>& ranges,
                uint64_t pos0, uint64_t npos, std::vector<uint64_t>& off, std::vector<uint32_t>& rows)
{
    std::vector<uint32_t> counts(2 * npos + 1, 0);
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            off.assign(2 * npos + 1, 0);
            for (uint64_t i = 0; i < 2 * npos; ++i) off[i + 1] = off[i] + counts[i];
            rows.assign(off[2 * npos], 0);
            cx.loc_rows = rows.data();
        }
        for (const WorkRange& r : ranges)
            for (uint64_t j = r.begin; j < r.end; ++j) {
                Chain<KW, SIGMA> st;
                HostFrames fr;
                st.cnt = 1;
                load_pattern(st.pat, text, nmask, text_begin + j, cx.K);
                chain_begin_block<KW, true, false, SIGMA>(st, fr, cx, nullptr);
                if (pass == 1) { st.loc_at_fwd = off[2 * (j - pos0)]; st.loc_at_rev = off[2 * (j - pos0) + 1]; }
                while (chain_step<KW, true, false, SIGMA, HostFrames, true>(st, fr, cx, nullptr, nullptr)) {}
                if (pass == 0) { counts[2 * (j - pos0)] = st.occ_fwd; counts[2 * (j - pos0) + 1] = st.occ_rev; }
            }
    }
    for (uint64_t i = 0; i < 2 * npos; ++i) std::sort(rows.begin() + off[i], rows.begin() + off[i + 1]);
}
void build_host_jump_levels(const MapCtx& cx, uint32_t sigma, uint32_t max_depth, std::vector<std::vector<JtFull>>& full)
{
    for (uint32_t d = 1; d <= max_depth; ++d) {
        const uint64_t n = 1ull << (2 * d), pmask = (1ull << (2 * (d - 1))) - 1;
        full[d].resize(n);
        for (uint64_t key = 0; key < n; ++key) {
            Node par;
            if (d == 1) { par.lo_f = 0; par.lo_r = 0; par.size = cx.n_bwt; }
            else { const JtFull& q = full[d - 1][key & pmask]; par.lo_f = q.lo_f; par.lo_r = q.lo_r; par.size = q.size; }
            const Node m = sigma == 5 ? extend_right<5>(par, (uint32_t)(key >> (2 * (d - 1))), cx)
                                      : extend_right<4>(par, (uint32_t)(key >> (2 * (d - 1))), cx);
            full[d][key] = JtFull{m.lo_r, m.size, m.lo_f, 0u};
        }
    }
}

// the text pass of jump_table.cu (k_locate_singletons) on the host: keys that occur once become LOCATED entries
void locate_host_singletons(const MapCtx& cx, uint32_t d, std::vector<JtFull>& full)
{
    const uint64_t n_text = cx.n_text;
    if (n_text < 2ull * kLocateMargin + d + 1) return;
    auto chars = [&](uint64_t p, uint32_t len) {
        const uint64_t w = cx.text[p >> 5], w2 = cx.text[(p >> 5) + 1];
        const uint32_t sh = 2u * (uint32_t)(p & 31u);
        const uint64_t v = sh ? (w >> sh) | (w2 << (64u - sh)) : w;
        return (uint32_t)(v & ((1ull << (2u * len)) - 1ull));
    };
    for (uint64_t q = kLocateMargin; q + d + kLocateMargin <= n_text; ++q) {
        if (cx.nmask) { // Dna5: no N in the key window or its context
            bool any = false;
            for (uint64_t i = q - kCtx; i < q + d + kCtx; ++i) any = any || ((cx.nmask[i >> 6] >> (i & 63)) & 1ull);
            if (any) continue;
        }
        JtFull& e = full[chars(q, d)];
        if (e.size != 1u) continue;
        uint32_t a = 0, b = cx.n_seq;
        while (b - a > 1) {
            const uint32_t mid = (a + b) >> 1;
            if ((uint64_t)cx.seq_start[mid] - mid <= q) a = mid; else b = mid;
        }
        if (q + d > (uint64_t)cx.seq_start[a + 1] - (a + 1)) continue;
        e = JtFull{(uint32_t)q, kLocated | 1u, chars(q + d, kCtx), chars(q - kCtx, kCtx)};
    }
}
} // namespace

// Pattern<KW, 5>::has_n_in / has_n against a per-character loop, for every (offset, length) of random N masks
// (block_kernel.cu asks it about common infixes of 32 and more characters); returns the number of disagreements
template <int KW>
static int has_n_selftest(uint64_t seed)
{
    int bad = 0;
    for (int round = 0; round < 64; ++round) {
        Pattern<KW, 5> p;
        for (int k = 0; k < KW; ++k) {
            seed = seed * 6364136223846793005ull + 1442695040888963407ull;
            p.w[k] = seed;
            // sparse masks (and a few empty / full ones): single bits are what a wrong shift loses
            p.nm[k] = round % 8 == 0 ? 0u : (round % 8 == 1 ? ~0u : (1u << ((seed >> 40) & 31u)) | ((seed >> 13) & 1u ? 1u << ((seed >> 20) & 31u) : 0u));
        }
        const uint32_t n = 32u * KW;
        for (uint32_t a = 0; a < n; ++a)
            for (uint32_t d = 0; a + d <= n; ++d) {
                bool want = false;
                for (uint32_t i = a; i < a + d; ++i) want = want || ((p.nm[i >> 5] >> (i & 31u)) & 1u);
                bad += p.has_n_in(a, d) != want;
                if (d >= 1 && d <= 16) bad += p.has_n(a, d) != want; // the table-key form (d <= 16)
            }
        bool any = false;
        for (uint32_t i = 0; i < n; ++i) any = any || ((p.nm[i >> 5] >> (i & 31u)) & 1u);
        bad += p.has_n() != any;
    }
    return bad;
}

extern "C" {

int hs_build(const uint8_t* codes, const uint64_t* limits, uint32_t n_seq, int with_sa, void** blob_out, uint64_t* bytes)
{
    Blob b;
    std::string err;
    if (!build_index_host(codes, limits, n_seq, with_sa != 0, b, err)) return -1;
    void* p = std::malloc(b.bytes);
    std::memcpy(p, b.data(), b.bytes);
    *blob_out = p;
    *bytes = b.bytes;
    return 0;
}

void hs_free(void* p) { std::free(p); }

// decode one direction's rank blocks back to symbols (0 = sentinel, 1..4 = A,C,G,T)
void hs_export_bwt(const void* blob, int rev, uint8_t* out)
{
    const uint8_t* base = static_cast<const uint8_t*>(blob);
    const IndexHeader& h = *reinterpret_cast<const IndexHeader*>(base);
    const uint32_t* sent = reinterpret_cast<const uint32_t*>(base + (rev ? h.off_sent_rev : h.off_sent_fwd));
    if (h.sigma == 5) {
        const RankBlock5* B = reinterpret_cast<const RankBlock5*>(base + (rev ? h.off_rev : h.off_fwd));
        for (uint64_t i = 0; i < h.n_bwt; ++i) {
            const RankBlock5& b = B[i / kBlockBases5];
            const uint32_t k = (uint32_t)(i % kBlockBases5);
            uint32_t c = 0;
            for (int pl = 0; pl < 3; ++pl) c |= ((b.plane[pl] >> k) & 1u) << pl;
            out[i] = (uint8_t)(1 + c);
        }
    } else {
        const RankBlock* B = reinterpret_cast<const RankBlock*>(base + (rev ? h.off_rev : h.off_fwd));
        for (uint64_t i = 0; i < h.n_bwt; ++i) {
            const RankBlock& b = B[i / kBlockBases];
            const uint32_t k = (uint32_t)(i % kBlockBases);
            out[i] = (uint8_t)(1 + (((b.w[k >> 6][0] >> (k & 63)) & 1) | (((b.w[k >> 6][1] >> (k & 63)) & 1) << 1)));
        }
    }
    for (uint32_t s = 0; s < h.n_seq; ++s) out[sent[s]] = 0;
}

int hs_export_sa(const void* blob, uint32_t* out)
{
    const uint8_t* base = static_cast<const uint8_t*>(blob);
    const IndexHeader& h = *reinterpret_cast<const IndexHeader*>(base);
    if (!h.off_sa) return -1;
    std::memcpy(out, base + h.off_sa, h.n_bwt * 4);
    return 0;
}

int hs_step_tables(uint32_t K, uint32_t E, uint32_t* n_search, uint32_t* steps)
{
    StepTables* t = new StepTables;
    std::string err;
    bool ok = build_step_tables(K, E, *t, err);
    if (ok) { *n_search = t->n_search; std::memcpy(steps, t->step, sizeof(uint32_t) * t->n_search * K); }
    delete t;
    return ok ? 0 : -1;
}

// jump_depth: -1 = default for the index size, 0 = no jump tables, else the maximum depth
// fetches (optional): 13 words — total, by interval size [8], thin paths, iterations, located entries, text reads (gmb_core.h: FetchStats)
int hs_map(const void* blob, uint32_t K, uint32_t E, int revcompl, int value_bits, uint64_t text_begin,
           uint64_t text_len, const uint64_t* chrom_cum, uint32_t n_chrom, const uint64_t* intervals,
           uint64_t n_intervals, uint64_t pos_begin, uint64_t pos_end, void* out, unsigned long long* fetches,
           int jump_depth, unsigned long long* lut_reads_out, const uint32_t* seq_to_file, uint32_t own_file,
           uint32_t block_kmers)
{
    const bool ep = seq_to_file != nullptr;
    const uint8_t* base = static_cast<const uint8_t*>(blob);
    std::string err;
    const IndexHeader& h = *reinterpret_cast<const IndexHeader*>(base);
    BlockTables tabs;
    // as capi.cu (dna5_nfree): on a Dna5 index with the suffix array the searches of an E >= 1 call skip the text's N
    const char* nf_env = std::getenv("GMB_DNA5_NFREE");
    const bool nfree = h.sigma == 5 && E >= 1 && !ep && h.off_sa != 0 && !(nf_env && nf_env[0] == '0');
    if (!build_block_tables(K, E, block_kmers, ep, tabs, err, h.n_bwt, block_bases(h.sigma), nfree)) return -2;
    const uint32_t B = tabs.B;
    MapCtx cx;
    cx.skip_n = nfree ? 1u : 0u;
    cx.blk[0] = base + h.off_fwd;
    cx.blk[1] = base + h.off_rev;
    cx.sent[0] = reinterpret_cast<const uint32_t*>(base + h.off_sent_fwd);
    cx.sent[1] = reinterpret_cast<const uint32_t*>(base + h.off_sent_rev);
    for (int c = 0; c < 5; ++c) cx.C[c] = (uint32_t)h.C[c];
    const uint32_t sigma = h.sigma;
    cx.n_bwt = (uint32_t)h.n_bwt;
    cx.steps = tabs.steps.data();
    cx.p1_off = tabs.p1_off;
    cx.fl_off = tabs.fl_off;
    cx.K = K; cx.B = B; cx.n_search = tabs.n_search; cx.n_strands = revcompl ? 2 : 1;
    cx.maxv = value_bits == 16 ? 65535u : 255u;
    cx.sa = nullptr; cx.seq_start = nullptr; cx.seq_to_file = nullptr; cx.n_seq = h.n_seq; cx.own_file = own_file; cx.all_files = 0;
    cx.loc_rows = nullptr; cx.text = nullptr; cx.nmask = nullptr; cx.n_text = 0; cx.E = E;
    if (ep) {
        if (!h.off_sa) return -3;
        cx.sa = reinterpret_cast<const uint32_t*>(base + h.off_sa);
        cx.seq_start = reinterpret_cast<const uint32_t*>(base + h.off_seq_start);
        cx.seq_to_file = seq_to_file;
        uint32_t nf = 0;
        for (uint32_t q = 0; q < h.n_seq; ++q) nf = std::max(nf, seq_to_file[q] + 1);
        if (nf > 64) return -4;
        cx.all_files = nf == 64 ? ~0ull : ((1ull << nf) - 1ull);
    }
    // jump tables, level by level (the device builder does the same with one thread per entry)
    const uint32_t want_depth = jump_depth < 0 ? default_jump_depth(h.n_bwt) : (uint32_t)jump_depth;
    std::vector<JumpPlan> plans(B + 1);
    uint32_t max_depth = 0;
    for (uint32_t cnt = 1; cnt <= B; ++cnt) {
        plan_jump_tables(tabs.infix[cnt], want_depth, plans[cnt], E, h.n_bwt, sigma, cnt, B > 1, nfree);
        max_depth = std::max(max_depth, plans[cnt].max_depth);
    }
    std::vector<std::vector<JtFull>> full(max_depth + 1);
    build_host_jump_levels(cx, sigma, max_depth, full);
    std::vector<std::vector<JtEntry>> uni(max_depth + 1);   // what the one-k-mer instantiation reads on Dna5 indices
    std::vector<std::vector<uint32_t>> lof(max_depth + 1);
    const char* loc_env = std::getenv("GMB_LOCATE");
    const bool all_full = !(loc_env && loc_env[0] == '0'); // as capi.cu: every search reads 16-byte entries
    if (B == 1 && !all_full)
        for (uint32_t d = 1; d <= max_depth; ++d)
            for (const JtFull& q : full[d]) { uni[d].push_back(JtEntry{q.lo_r, q.size}); lof[d].push_back(q.lo_f); }
    cx.seq_start = reinterpret_cast<const uint32_t*>(base + h.off_seq_start);
    cx.text = reinterpret_cast<const uint64_t*>(base + h.off_text);
    cx.nmask = sigma == 5 ? reinterpret_cast<const uint64_t*>(base + h.off_nmask) : nullptr;
    cx.n_text = h.n_text;
    cx.E = E;
    if (all_full)
        for (uint32_t d = 1; d <= max_depth; ++d) locate_host_singletons(cx, d, full[d]);
    std::vector<SearchStart> starts((B + 1) * kMaxSearches);
    for (uint32_t cnt = 1; cnt <= B; ++cnt)
        for (uint32_t s = 0; s < kMaxSearches; ++s) {
            const uint32_t d = plans[cnt].depth[s];
            SearchStart& S = starts[cnt * kMaxSearches + s];
            std::memset(&S, 0, sizeof(S));
            if (B == 1 && !all_full) { S.uni = d ? uni[d].data() : nullptr; S.lof = (d && plans[cnt].need_lof[s]) ? lof[d].data() : nullptr; }
            else S.full = d ? full[d].data() : nullptr; // (the device uses 8-byte entries where SA(T) is not needed: same values)
            S.a = plans[cnt].a[s]; S.d = d;
            S.n_var = std::max(1u, plans[cnt].n_var[s]);
            S.var = plans[cnt].variants.data() + plans[cnt].var_off[s];
            S.set0 = S.var[0];
        }
    cx.starts = starts.data();
    std::memset(out, 0, text_len * (value_bits / 8));
    std::vector<WorkRange> ranges;
    build_work_ranges(text_len, K, chrom_cum, n_chrom, intervals, n_intervals, pos_begin, pos_end, ranges);
    const uint64_t* text = reinterpret_cast<const uint64_t*>(base + h.off_text);
    const uint64_t* nmask = sigma == 5 ? reinterpret_cast<const uint64_t*>(base + h.off_nmask) : nullptr;
    FetchStats f{};
    unsigned long long lr = 0;
#define RUN_KS(KW, BLK, SG) (ep ? run_ranges<KW, true, BLK, SG>(cx, text, nmask, text_begin, ranges, value_bits, out, &f, &lr) \
                                : run_ranges<KW, false, BLK, SG>(cx, text, nmask, text_begin, ranges, value_bits, out, &f, &lr))
#define RUN_KB(KW, BLK) (sigma == 5 ? RUN_KS(KW, BLK, 5) : RUN_KS(KW, BLK, 4))
#define RUN_KW(KW) (B > 1 ? RUN_KB(KW, true) : RUN_KB(KW, false))
    const uint32_t needle = K + B - 1; // characters a chain keeps in registers
    // the two-phase driver of block_kernel.cu, where the library would launch it (capi.cu: get_plan)
    KeyLists kl;
    const char* bk_env = std::getenv("GMB_BLOCK_KERNEL");
    const bool block_driver = E >= 1 && all_full && (sigma == 4 || nfree) && needle <= 64 && !(bk_env && bk_env[0] == '0') && build_key_lists(tabs, plans, kl);
    if (block_driver && sigma == 5) {
        if (needle <= 32) run_ranges_blockdriver<1, false, 5>(cx, kl, text, nmask, text_begin, ranges, value_bits, out, &f, &lr);
        else run_ranges_blockdriver<2, false, 5>(cx, kl, text, nmask, text_begin, ranges, value_bits, out, &f, &lr);
    } else if (block_driver) {
        if (needle <= 32) { if (ep) run_ranges_blockdriver<1, true, 4>(cx, kl, text, nullptr, text_begin, ranges, value_bits, out, &f, &lr);
                            else run_ranges_blockdriver<1, false, 4>(cx, kl, text, nullptr, text_begin, ranges, value_bits, out, &f, &lr); }
        else { if (ep) run_ranges_blockdriver<2, true, 4>(cx, kl, text, nullptr, text_begin, ranges, value_bits, out, &f, &lr);
               else run_ranges_blockdriver<2, false, 4>(cx, kl, text, nullptr, text_begin, ranges, value_bits, out, &f, &lr); }
    }
    else if (needle <= 32) RUN_KW(1);
    else if (needle <= 64) RUN_KW(2);
    else if (needle <= 128) RUN_KW(4);
    else RUN_KW(9);
    if (nfree) nfix_bruteforce(h, base, K, E, revcompl != 0, text_begin, ranges, value_bits, out);
    if (fetches) { fetches[0] = f.total; for (int k = 0; k < 8; ++k) fetches[1 + k] = f.by_size[k]; fetches[9] = f.thin_paths; fetches[10] = f.iterations; fetches[11] = f.located; fetches[12] = f.text_reads; }
    if (lut_reads_out) *lut_reads_out = lr;
    return 0;
}

// csv lists of the file-local positions [pos_begin, pos_end): offsets_out has 2*(pos_end-pos_begin)+1 entries;
// rows (sorted positions inside the sentinel-separated text T) are malloc'ed into *rows_out (hs_free).
int hs_locate(const void* blob, uint32_t K, uint32_t E, int revcompl, uint64_t text_begin, uint64_t text_len,
              const uint64_t* chrom_cum, uint32_t n_chrom, const uint64_t* intervals, uint64_t n_intervals,
              uint64_t pos_begin, uint64_t pos_end, int jump_depth, uint64_t* offsets_out, uint32_t** rows_out)
{
    const uint8_t* base = static_cast<const uint8_t*>(blob);
    std::string err;
    const IndexHeader& h = *reinterpret_cast<const IndexHeader*>(base);
    if (!h.off_sa) return -3;
    BlockTables tabs;
    if (!build_block_tables(K, E, 1, true, tabs, err, h.n_bwt, block_bases(h.sigma))) return -2;
    MapCtx cx;
    cx.blk[0] = base + h.off_fwd;
    cx.blk[1] = base + h.off_rev;
    cx.sent[0] = reinterpret_cast<const uint32_t*>(base + h.off_sent_fwd);
    cx.sent[1] = reinterpret_cast<const uint32_t*>(base + h.off_sent_rev);
    for (int c = 0; c < 5; ++c) cx.C[c] = (uint32_t)h.C[c];
    const uint32_t sigma = h.sigma;
    cx.n_bwt = (uint32_t)h.n_bwt;
    cx.steps = tabs.steps.data();
    cx.p1_off = tabs.p1_off;
    cx.fl_off = tabs.fl_off;
    cx.K = K; cx.B = 1; cx.n_search = tabs.n_search; cx.n_strands = revcompl ? 2 : 1;
    cx.maxv = 65535u;
    cx.sa = reinterpret_cast<const uint32_t*>(base + h.off_sa);
    cx.seq_start = reinterpret_cast<const uint32_t*>(base + h.off_seq_start);
    cx.seq_to_file = nullptr; cx.n_seq = h.n_seq; cx.own_file = 0; cx.all_files = 0;
    cx.loc_rows = nullptr; cx.text = reinterpret_cast<const uint64_t*>(base + h.off_text); cx.nmask = nullptr; cx.n_text = h.n_text; cx.E = E;
    cx.skip_n = 0;
    const uint32_t want_depth = jump_depth < 0 ? default_jump_depth(h.n_bwt) : (uint32_t)jump_depth;
    JumpPlan plan;
    plan_jump_tables(tabs.infix[1], want_depth, plan, E, h.n_bwt, sigma, 1, false);
    std::vector<std::vector<JtFull>> full(plan.max_depth + 1);
    build_host_jump_levels(cx, sigma, plan.max_depth, full);
    std::vector<std::vector<JtEntry>> uni(plan.max_depth + 1);
    std::vector<std::vector<uint32_t>> lof(plan.max_depth + 1);
    for (uint32_t d = 1; d <= plan.max_depth; ++d)
        for (const JtFull& q : full[d]) { uni[d].push_back(JtEntry{q.lo_r, q.size}); lof[d].push_back(q.lo_f); }
    std::vector<SearchStart> starts(2 * kMaxSearches);
    for (uint32_t s = 0; s < kMaxSearches; ++s) {
        const uint32_t d = plan.depth[s];
        SearchStart& S = starts[kMaxSearches + s];
        std::memset(&S, 0, sizeof(S));
        S.uni = d ? uni[d].data() : nullptr;
        S.lof = (d && plan.need_lof[s]) ? lof[d].data() : nullptr;
        S.a = plan.a[s]; S.d = d;
        S.n_var = std::max(1u, plan.n_var[s]);
        S.var = plan.variants.data() + plan.var_off[s];
        S.set0 = S.var[0];
    }
    cx.starts = starts.data();
    std::vector<WorkRange> ranges;
    build_work_ranges(text_len, K, chrom_cum, n_chrom, intervals, n_intervals, pos_begin, pos_end, ranges);
    const uint64_t* text = reinterpret_cast<const uint64_t*>(base + h.off_text);
    const uint64_t* nmask = sigma == 5 ? reinterpret_cast<const uint64_t*>(base + h.off_nmask) : nullptr;
    std::vector<uint64_t> off;
    std::vector<uint32_t> rows;
    const uint64_t npos = pos_end - pos_begin;
#define LOC_KW(KW) (sigma == 5 ? run_locate<KW, 5>(cx, text, nmask, text_begin, ranges, pos_begin, npos, off, rows) \
                               : run_locate<KW, 4>(cx, text, nmask, text_begin, ranges, pos_begin, npos, off, rows))
    if (K <= 32) LOC_KW(1);
    else if (K <= 64) LOC_KW(2);
    else if (K <= 128) LOC_KW(4);
    else LOC_KW(9);
    std::memcpy(offsets_out, off.data(), off.size() * 8);
    uint32_t* r = static_cast<uint32_t*>(std::malloc(rows.size() * 4 + 4));
    std::memcpy(r, rows.data(), rows.size() * 4);
    *rows_out = r;
    return 0;
}

int hs_has_n_selftest(uint64_t seed) { return has_n_selftest<1>(seed) + has_n_selftest<2>(seed + 1) + has_n_selftest<4>(seed + 2) + has_n_selftest<9>(seed + 3); }

} // extern "C"
