// gmb_host.cpp — step tables, host index builder, blob packing (no CUDA).
#include "gmb_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "sais.hpp"

namespace gmb {

// ---------------------------------------------------------------------------------------------------
// Optimum search schemes (Kianfar et al.) in the variant GenMap ships: pi / L / U per search.
// Values: src/find2_index_approx.hpp:67-134 (they are the algorithm's specification).
// ---------------------------------------------------------------------------------------------------
namespace {
struct SchemeDef {
    uint8_t n_search, n_blocks;
    uint8_t pi[7][6], lo[7][6], up[7][6];
};
const SchemeDef kSchemes[5] = {
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
} // namespace

// ---------------------------------------------------------------------------------------------------
// Part lengths.  The reference splits the pattern into nb parts of equal length (floor(K/nb), the first K mod nb
// one longer: src/find2_index_approx.hpp:164-176).  Any split into nb non-empty parts keeps a scheme exhaustive
// and non-redundant (its validity depends on the errors per part, not on the part lengths), so the counts do
// not depend on it — but the size of the search tree does, strongly: on a text of N symbols every node of depth
// < log4 N exists, so errors allowed close to the root are expensive and the best split depends on N.  Measured
// at 3 Gbp, K = 30, E = 2: 406 rank-block fetches per position with the equal split, 281 with (5,6,7,7)
// (profiles/r01/s18_parts_e2.txt).  expected_fetches() is the model used to pick the split: expected number of
// existing trie nodes (x 1..2 blocks each) on an iid text, the query's own occurrence counted as certain; it
// tracks the measured fetch counts within 5 %.
// ---------------------------------------------------------------------------------------------------
namespace {
// Expected cost of one search on an iid text of N symbols, depth by depth (see the comment above).
struct SearchProfile {
    uint32_t Li = 0, run = 0;          // steps; leading steps that allow no error
    double walk[2][kMaxK + 1];         // expected rank-block fetches of the nodes of depth t, strand 0 / 1
    double nstr[kMaxK + 2];            // strings of length t the scheme admits (error placements x 3^errors)
    double leaf[2];                    // window completions per infix hit
    double N = 0;                      // text size the profile was made for
    uint32_t needle = 0;               // characters of the block's needle (K + B - 1)
};

void profile_search(const uint32_t* ub, const uint32_t* lb, const uint32_t* rem, uint32_t Li, uint32_t E, double N, uint32_t B,
                    uint32_t block_bases, SearchProfile& P)
{
    P.Li = Li;
    P.N = N;
    P.needle = Li + 2 * (B - 1);
    P.run = 0;
    while (P.run < Li && ub[P.run] == 0) ++P.run;
    std::vector<char> exact_ok(Li + 1);
    exact_ok[Li] = 1;
    for (uint32_t q = Li; q-- > 0;) exact_ok[q] = exact_ok[q + 1] && lb[q] == 0;
    double cnt[kMaxE + 2] = {1.0, 0, 0, 0, 0, 0};
    double lam = N;
    for (uint32_t t = 0; t < Li; ++t, lam *= 0.25) {
        const double pex = lam > 30 ? 1.0 : 1.0 - std::exp(-lam);   // a given string of length t occurs
        const double size = pex > 0 ? std::max(1.0, lam / pex) : 1.0; // rows of an existing node
        const double fc = 1.0 + std::min(1.0, size / block_bases);   // blocks per expansion
        P.walk[0][t] = P.walk[1][t] = 0.0;
        P.nstr[t] = 0.0;
        for (uint32_t e = 0; e <= E; ++e) {
            if (cnt[e] == 0) continue;
            P.nstr[t] += cnt[e];
            P.walk[1][t] += cnt[e] * pex * fc;
            if (e == 0) P.walk[0][t] += (size <= 1.0 && exact_ok[t]) ? 0.0 : fc; // the query's own path
            else P.walk[0][t] += cnt[e] * pex * fc;
        }
        double nxt[kMaxE + 2] = {0, 0, 0, 0, 0, 0};
        for (uint32_t e = 0; e <= E; ++e) {
            if (cnt[e] == 0) continue;
            if (e + rem[t] >= lb[t]) nxt[e] += cnt[e];
            if (e + 1 <= ub[t] && e + 1 + rem[t] >= lb[t]) nxt[e + 1] += 3.0 * cnt[e];
        }
        for (uint32_t e = 0; e <= E + 1; ++e) cnt[e] = nxt[e];
    }
    const double pleaf = lam > 30 ? 1.0 : 1.0 - std::exp(-lam); // infix hits are completed window by window
    P.nstr[Li] = 0;
    P.leaf[0] = P.leaf[1] = 0.0;
    for (uint32_t e = 0; e <= E; ++e) {
        P.nstr[Li] += cnt[e];
        const double w = B * (B - 1) * 0.5 * (e == E ? 1.0 : 1.5);
        P.leaf[0] += (e == 0 ? 1.0 : cnt[e] * pleaf) * w;
        P.leaf[1] += cnt[e] * pleaf * w;
    }
}

// Depth at which the search is entered through the jump table.  Up to `run` the table replaces an error-free
// walk (one entry); deeper, every admissible string of that length is looked up (its mismatches substituted into
// the key: one table read per string, both strands) instead of walking the dense top of the trie.  Returns the
// depth that minimises expected table reads + expected fetches below it; *cost = that minimum (both strands).
constexpr double kMaxVariantStrings = 3000.0;
// located: the tables hold LOCATED entries (gmb_core.h: JtFull) — a key that occurs once ends its search at the table
// read (verify_located).  Of the occurrences of a random d-mer a fraction e^(-N/4^d) are alone, so what is walked
// below an entry depth d shrinks by that fraction; a located entry whose context does not cover the needle costs one
// read of the text instead.
uint32_t best_jump_depth(const SearchProfile& P, uint32_t dmax, bool allow_variants, double* cost, bool located = false)
{
    const uint32_t top = std::min(dmax, P.Li > 0 ? P.Li - 1 : 0u);
    const uint32_t first = std::min(P.run, top);
    std::vector<double> below(P.Li + 1, 0.0); // expected fetches of all depths >= t, both strands
    for (uint32_t t = P.Li; t-- > 0;) below[t] = below[t + 1] + P.walk[0][t] + P.walk[1][t];
    auto entered_at = [&](uint32_t d, double strings) { // expected accesses of entering at depth d through `strings` keys per strand
        if (d == 0) return below[0] + P.leaf[0] + P.leaf[1];
        double walked = 1.0, text = 0.0;
        if (located) {
            const double lam = P.N / std::pow(4.0, (double)d);
            const double alone = lam > 30 ? 0.0 : std::exp(-lam);
            walked = 1.0 - alone;
            if (P.needle > d + 2 * kCtx) text = (2.0 * strings - 1.0) * lam * alone; // (the query's own key needs no comparison)
        }
        return 2.0 * strings + text + walked * (below[d] + P.leaf[0] + P.leaf[1]);
    };
    uint32_t best_d = first;
    double best = entered_at(first, 1.0);
    if (allow_variants)
        for (uint32_t d = first + 1; d <= top; ++d) {
            if (P.nstr[d] > kMaxVariantStrings) break;
            const double c = entered_at(d, P.nstr[d]);
            if (c < best) { best = c; best_d = d; }
        }
    if (cost) *cost = best;
    return best_d;
}

bool located_enabled(uint64_t n_bwt, uint32_t sigma)
{
    const char* env = std::getenv("GMB_LOCATE"); // "0": tables without located entries (every search walks the index)
    return n_bwt != 0 && (sigma == 4 || sigma == 5) && !(env && env[0] == '0');
}

// nfree: the searches of this call never match a text N (gmb_core.h: MapCtx::skip_n; the alignments to text windows with
// an N are counted by a pass of their own, capi.cu), so a Dna5 index is entered through substituted keys like a Dna4 one
bool variants_enabled(uint64_t n_bwt, uint32_t sigma, bool nfree)
{
    const char* env = std::getenv("GMB_JUMP_VARIANTS"); // "0": enter every search through its error-free prefix only
    return n_bwt != 0 && (sigma == 4 || (sigma == 5 && nfree)) && !(env && env[0] == '0');
}

double expected_fetches(const SchemeDef& sd, const uint32_t* len, uint32_t E, double N, uint32_t B, uint32_t jump_max, uint32_t block_bases,
                        bool allow_variants, bool located)
{
    const uint32_t nb = sd.n_blocks;
    uint32_t Li = 0;
    for (uint32_t b = 0; b < nb; ++b) Li += len[b];
    double total = 0.0;
    std::vector<uint32_t> ub(Li), lb(Li), rem(Li);
    SearchProfile P;
    for (uint32_t s = 0; s < sd.n_search; ++s) {
        uint32_t t = 0;
        for (uint32_t i = 0; i < nb; ++i) {
            const uint32_t n = len[sd.pi[s][i] - 1u];
            for (uint32_t k = 0; k < n; ++k, ++t) { ub[t] = sd.up[s][i]; lb[t] = sd.lo[s][i]; rem[t] = n - 1 - k; }
        }
        profile_search(ub.data(), lb.data(), rem.data(), Li, E, N, B, block_bases, P);
        double c = 0;
        best_jump_depth(P, jump_max, allow_variants, &c, located);
        total += c;
    }
    return total / B;
}

// steepest descent over single-character moves between parts, from the reference's equal split
void choose_part_lengths(const SchemeDef& sd, uint32_t K, uint32_t E, uint64_t n_bwt, uint32_t B, uint32_t block_bases, uint32_t* len,
                         bool nfree)
{
    const bool av = B > 1 && variants_enabled(n_bwt, block_bases == kBlockBases5 ? 5u : 4u, nfree); // (the one-k-mer kernel has none)
    const bool loc = located_enabled(n_bwt, block_bases == kBlockBases5 ? 5u : 4u);
    const uint32_t nb = sd.n_blocks;
    if (nb < 2 || n_bwt == 0 || K < 2 * nb) return;
    const uint32_t jump_max = default_jump_depth(n_bwt);
    double best = expected_fetches(sd, len, E, (double)n_bwt, B, jump_max, block_bases, av, loc);
    for (int round = 0; round < 256; ++round) {
        int bi = -1, bj = -1;
        double bc = best * (1.0 - 1e-6);
        for (uint32_t i = 0; i < nb; ++i)
            for (uint32_t j = 0; j < nb; ++j) {
                if (i == j || len[i] <= 1) continue;
                --len[i]; ++len[j];
                const double c = expected_fetches(sd, len, E, (double)n_bwt, B, jump_max, block_bases, av, loc);
                ++len[i]; --len[j];
                if (c < bc) { bc = c; bi = (int)i; bj = (int)j; }
            }
        if (bi < 0) break;
        --len[bi]; ++len[bj];
        best = bc;
    }
}
} // namespace

bool build_step_tables(uint32_t K, uint32_t E, StepTables& out, std::string& err, bool force_sync, uint64_t n_bwt, uint32_t block_kmers,
                       uint32_t block_bases, bool nfree)
{
    if (E > kMaxE) { err = "E > 4 not yet supported."; return false; } // src/mappability.hpp:187
    if (K < E + 2) { err = "K must be at least E + 2."; return false; } // undefined in the reference (rc 139)
    if (K > kMaxK) { err = "K > 255 is not supported."; return false; }
    const SchemeDef& sd = kSchemes[E];
    const uint32_t nb = sd.n_blocks;
    // block lengths over the whole k-mer: floor(K/nb), the first K mod nb blocks one longer (:164-176)
    uint32_t len[6], begin[6];
    for (uint32_t b = 0; b < nb; ++b) len[b] = K / nb + (b < K % nb);
    const char* model_env = std::getenv("GMB_PART_MODEL"); // "0": keep the reference's equal split
    if (n_bwt != 0 && !(model_env && model_env[0] == '0')) choose_part_lengths(sd, K, E, n_bwt, block_kmers ? block_kmers : 1, block_bases, len, nfree);
    // Tuning knob (results never depend on it: any split into nb non-empty parts keeps the scheme exhaustive and
    // non-redundant, only the size of the search tree changes): relative part lengths, e.g. GMB_PART_WEIGHTS=5,5,8,8
    if (const char* env = std::getenv("GMB_PART_WEIGHTS")) {
        double w[6], sum = 0;
        uint32_t got = 0;
        for (const char* q = env; *q && got < 6;) {
            char* end = nullptr;
            const double v = std::strtod(q, &end);
            if (end == q) break;
            w[got++] = v > 0 ? v : 0;
            q = *end == ',' ? end + 1 : end;
        }
        if (got == nb && K >= 2 * nb) {
            for (uint32_t b = 0; b < nb; ++b) sum += w[b];
            uint32_t used = 0;
            for (uint32_t b = 0; b < nb && sum > 0; ++b) {
                len[b] = std::max<uint32_t>(1, (uint32_t)(K * w[b] / sum + 0.5));
                used += len[b];
            }
            // fix rounding on the longest part so that the parts add up to K
            while (sum > 0 && used != K) {
                uint32_t big = 0;
                for (uint32_t b = 1; b < nb; ++b) if (len[b] > len[big]) big = b;
                if (used > K) { --len[big]; --used; } else { ++len[big]; ++used; }
            }
        }
    }
    for (uint32_t b = 0, o = 0; b < nb; ++b) {
        begin[b] = o;
        o += len[b];
    }
    out.n_search = sd.n_search;
    out.K = K;
    for (uint32_t s = 0; s < sd.n_search; ++s) {
        uint32_t* st = out.step + s * K;
        uint32_t t = 0;
        uint32_t l = begin[sd.pi[s][0] - 1], r = l; // consumed window [l, r)
        for (uint32_t i = 0; i < nb; ++i) {
            const uint32_t b = sd.pi[s][i] - 1u;
            const bool right = (i == 0) || sd.pi[s][i] > sd.pi[s][i - 1]; // :273-285,321
            for (uint32_t k = 0; k < len[b]; ++k, ++t) {
                const uint32_t pos = right ? r++ : --l;
                st[t] = pos | ((len[b] - 1 - k) << 8) | ((uint32_t)sd.up[s][i] << 16) |
                        ((uint32_t)sd.lo[s][i] << 20) | ((right ? 1u : 0u) << 24);
            }
            if (right ? (r != begin[b] + len[b]) : (l != begin[b])) { err = "scheme is not contiguous"; return false; }
        }
        if (t != K || l != 0 || r != K) { err = "scheme does not cover the pattern"; return false; }
        for (uint32_t a = 0; a < K; ++a) { // does a later step go the other way?
            bool sw = false;
            for (uint32_t b2 = a + 1; b2 < K; ++b2) sw |= step_dir(st[b2]) != step_dir(st[a]);
            if (sw || force_sync) st[a] |= 1u << 25;
            bool zero_lb = true; // may the rest of the pattern be matched without any further error?
            for (uint32_t b2 = a; b2 < K; ++b2) zero_lb &= step_lb(st[b2]) == 0;
            if (zero_lb) st[a] |= 1u << 26;
        }
    }
    return true;
}

uint32_t default_block_kmers(uint32_t K, uint32_t E)
{
    // E = 0: with the jump tables a k-mer costs 2-3 rank-block reads, sharing an infix cannot beat that.
    // E >= 1: the infix search dominates and is shared by the block; more k-mers per block shorten the infix,
    // which makes its search bushier (same trade-off as the reference's overlap, src/mappability.hpp:519-543).
    if (E == 0) return 1;
    uint32_t b = E == 1 ? 4u : (E == 2 ? 6u : 4u);    // measured at 3 Gbp, K = 30: profiles/r01/s7_sweep_block_sizes.txt
    const uint32_t cap = K > E + 1 ? K - E - 1 : 1; // the infix must keep one character per scheme block
    while (b > 1 && (b > cap || b * 4 > K)) --b;    // and stay most of the k-mer
    return b < 1 ? 1 : b;
}

// Block size by the same model: expected fetches per position of the best split for every B; the largest B within
// 3 % of the minimum (equal fetch counts favour the larger block: fewer chain start-ups).  Measured at 3 Gbp, K = 30
// (profiles/r01/s19_sweep_model1.txt): the model is 4 % above the counted fetches for B = 3..8 and ranks them alike.
uint32_t model_block_kmers(uint32_t K, uint32_t E, uint64_t n_bwt, uint32_t block_bases, bool nfree)
{
    if (E == 0 || n_bwt == 0) return default_block_kmers(K, E);
    const SchemeDef& sd = kSchemes[E];
    double cost[13];
    uint32_t top = 0;
    double best = 0;
    for (uint32_t B = 1; B <= 12; ++B) {
        if (B + E + 1 > K || K + B - 2 > 255 || K - B + 1 < 2 * sd.n_blocks) break;
        const uint32_t Li = K - B + 1;
        uint32_t len[6];
        for (uint32_t b = 0; b < sd.n_blocks; ++b) len[b] = Li / sd.n_blocks + (b < Li % sd.n_blocks);
        choose_part_lengths(sd, Li, E, n_bwt, B, block_bases, len, nfree);
        cost[B] = expected_fetches(sd, len, E, (double)n_bwt, B, default_jump_depth(n_bwt), block_bases,
                                   B > 1 && variants_enabled(n_bwt, block_bases == kBlockBases5 ? 5u : 4u, nfree),
                                   located_enabled(n_bwt, block_bases == kBlockBases5 ? 5u : 4u));
        // a Dna5 needle longer than 32 characters no longer fits one register word per plane: measured 1.55x slower per
        // fetch (profiles/r01/s25_sweep_dna5.txt vs s20_sweep_dna5.txt)
        if (block_bases == kBlockBases5 && K + B - 1 > 32 && K <= 32) cost[B] *= 1.5;
        if (top == 0 || cost[B] < best) best = cost[B];
        top = B;
    }
    if (top == 0) return 1;
    uint32_t pick = 1;
    for (uint32_t B = 1; B <= top; ++B)
        if (cost[B] <= 1.03 * best) pick = B;
    return pick;
}

static bool build_block_tables_uncached(uint32_t K, uint32_t E, uint32_t B, bool force_sync, BlockTables& out, std::string& err,
                                        uint64_t n_bwt, uint32_t block_bases, bool nfree);

// The tables of one (K, E, B, text size) are asked for by every map call (every pipeline piece): keep the last few.
bool build_block_tables(uint32_t K, uint32_t E, uint32_t B, bool force_sync, BlockTables& out, std::string& err, uint64_t n_bwt,
                        uint32_t block_bases, bool nfree)
{
    struct Entry { std::string key; BlockTables tabs; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    const char* e1 = std::getenv("GMB_PART_MODEL");
    const char* e2 = std::getenv("GMB_PART_WEIGHTS");
    char buf[160];
    std::snprintf(buf, sizeof buf, "%u/%u/%u/%d/%llu/%u/%d/", K, E, B, force_sync ? 1 : 0, (unsigned long long)n_bwt, block_bases, nfree ? 1 : 0);
    const char* e3 = std::getenv("GMB_JUMP_VARIANTS");
    const char* e4 = std::getenv("GMB_LOCATE");
    const std::string key = std::string(buf) + (e1 ? e1 : "") + "/" + (e2 ? e2 : "") + "/" + (e3 ? e3 : "") + "/" + (e4 ? e4 : "");
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const Entry& en : cache)
            if (en.key == key) { out = en.tabs; return true; }
    }
    if (!build_block_tables_uncached(K, E, B, force_sync, out, err, n_bwt, block_bases, nfree)) return false;
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() >= 8) cache.erase(cache.begin());
    cache.push_back(Entry{key, out});
    return true;
}

static bool build_block_tables_uncached(uint32_t K, uint32_t E, uint32_t B, bool force_sync, BlockTables& out, std::string& err,
                                        uint64_t n_bwt, uint32_t block_bases, bool nfree)
{
    if (E > kMaxE) { err = "E > 4 not yet supported."; return false; }
    if (K < E + 2) { err = "K must be at least E + 2."; return false; }
    if (B == 0) {
        const char* model_env = std::getenv("GMB_PART_MODEL");
        B = (n_bwt != 0 && !(model_env && model_env[0] == '0')) ? model_block_kmers(K, E, n_bwt, block_bases, nfree) : default_block_kmers(K, E);
    }
    if (B > kMaxBlockKmers) B = kMaxBlockKmers;
    while (B > 1 && (B + E + 1 > K || K + B - 2 > 255)) --B; // the infix keeps >= E + 2 characters; offsets fit 8 bits
    if (K + B - 2 > 255) { err = "K > 255 is not supported."; return false; }
    out.K = K; out.E = E; out.B = B;
    out.steps.clear();
    out.infix.assign(B + 1, StepTables());
    for (uint32_t cnt = 1; cnt <= B; ++cnt) {
        const uint32_t Li = K - cnt + 1;
        StepTables& t = out.infix[cnt];
        if (!build_step_tables(Li, E, t, err, force_sync || cnt > 1, n_bwt, cnt, block_bases, nfree)) return false; // flanks need both intervals
        out.n_search = t.n_search;
        out.p1_off[cnt] = (uint32_t)out.steps.size();
        for (uint32_t i = 0; i < t.n_search * Li; ++i) {
            t.step[i] += cnt - 1; // pattern offsets in needle coordinates: the infix starts at cnt - 1
            out.steps.push_back(t.step[i]);
        }
        out.fl_off[cnt] = (uint32_t)out.steps.size();
        for (uint32_t w = 0; w < cnt; ++w) {
            // window w = needle[w .. w+K): left flank needle[cnt-2 .. w] (right to left), then right flank needle[K .. K+w-1]
            for (uint32_t q = cnt - 1; q > w; --q) {
                const bool sync = force_sync || w > 0; // right-flank steps follow
                out.steps.push_back((q - 1) | (E << 16) | (0u << 24) | ((sync ? 1u : 0u) << 25) | (1u << 26));
            }
            for (uint32_t q = 0; q < w; ++q)
                out.steps.push_back((K + q) | (E << 16) | (1u << 24) | ((force_sync ? 1u : 0u) << 25) | (1u << 26));
        }
    }
    return true;
}

namespace {
// every set of mismatching steps among the first d steps that the scheme admits (the nodes of depth d of the
// search's trie, up to the choice of the substituted characters), as up to kMaxE key offsets per set
void enumerate_variants(const uint32_t* st, uint32_t d, uint32_t E, uint32_t a, uint32_t t, uint32_t e, uint32_t set, std::vector<uint32_t>& out)
{
    if (t == d) { out.push_back(set); return; }
    const uint32_t ub = step_ub(st[t]), lb = step_lb(st[t]), rem = step_rem(st[t]);
    if (e + rem >= lb) enumerate_variants(st, d, E, a, t + 1, e, set, out); // the pattern's own character
    if (e + 1 <= ub && e + 1 + rem >= lb && e < kMaxE) {                    // a substituted character here
        const uint32_t off = step_pos(st[t]) - a;
        const uint32_t with = (set & ~(0xffu << (8 * e))) | (off << (8 * e));
        enumerate_variants(st, d, E, a, t + 1, e + 1, with, out);
    }
}
} // namespace

void plan_jump_tables(const StepTables& tabs, uint32_t max_depth, JumpPlan& plan, uint32_t E, uint64_t n_bwt, uint32_t sigma,
                      uint32_t block_kmers, bool allow_variants, bool nfree)
{
    plan.max_depth = 0;
    plan.variants.clear();
    for (uint32_t s = 0; s < kMaxSearches; ++s) { plan.depth[s] = 0; plan.a[s] = 0; plan.need_lof[s] = false; plan.var_off[s] = 0; plan.n_var[s] = 0; }
    const uint32_t K = tabs.K;
    const bool allow = allow_variants && variants_enabled(n_bwt, sigma, nfree);
    if (max_depth > 16) max_depth = 16;
    std::vector<uint32_t> ub(K), lb(K), rem(K);
    SearchProfile P;
    for (uint32_t s = 0; s < tabs.n_search; ++s) {
        const uint32_t* st = tabs.step + s * K;
        for (uint32_t t = 0; t < K; ++t) { ub[t] = step_ub(st[t]); lb[t] = step_lb(st[t]); rem[t] = step_rem(st[t]); }
        profile_search(ub.data(), lb.data(), rem.data(), K, E, n_bwt ? (double)n_bwt : 1.0, block_kmers ? block_kmers : 1,
                       block_bases(sigma), P);
        const uint32_t d = best_jump_depth(P, max_depth, allow, nullptr, located_enabled(n_bwt, sigma));
        plan.depth[s] = d;
        plan.var_off[s] = (uint32_t)plan.variants.size();
        if (d == 0) { plan.n_var[s] = 1; plan.variants.push_back(0xffffffffu); continue; }
        uint32_t a = step_pos(st[0]);
        for (uint32_t t = 0; t < d; ++t) a = std::min(a, step_pos(st[t])); // the consumed window is contiguous: [a, a + d)
        plan.a[s] = a;
        enumerate_variants(st, d, E, a, 0, 0, 0xffffffffu, plan.variants);
        if (plan.variants.size() == plan.var_off[s]) plan.variants.push_back(kDeadVariant); // no admissible string of this length
        plan.n_var[s] = (uint32_t)plan.variants.size() - plan.var_off[s];
        // the error-free string first (the reverse strand's first read is requested early under that assumption)
        for (uint32_t v = plan.var_off[s]; v < plan.variants.size(); ++v)
            if (plan.variants[v] == 0xffffffffu) { std::swap(plan.variants[v], plan.variants[plan.var_off[s]]); break; }
        bool left_later = false; // the interval in SA(T) is the active one for every later step to the left
        for (uint32_t t = d; t < K; ++t) left_later |= step_dir(st[t]) == 0;
        plan.need_lof[s] = step_sync(st[d - 1]) != 0 || left_later;
        plan.max_depth = std::max(plan.max_depth, d);
    }
}


bool build_key_lists(const BlockTables& tabs, const std::vector<JumpPlan>& plans, KeyLists& out)
{
    out.xy.clear();
    for (uint32_t cnt = 1; cnt <= tabs.B; ++cnt) {
        out.off[cnt] = (uint32_t)(out.xy.size() / 2);
        for (uint32_t s = 0; s < tabs.n_search; ++s) {
            if (plans[cnt].depth[s] == 0) { out.xy.clear(); return false; } // a search that starts at the root
            for (uint32_t v = 0; v < std::max(1u, plans[cnt].n_var[s]); ++v) {
                const uint32_t set = plans[cnt].variants.empty() ? 0xffffffffu : plans[cnt].variants[plans[cnt].var_off[s] + v];
                if (set == kDeadVariant) continue;
                uint32_t m = 0, n3 = 1;
                for (uint32_t k = 0; k < kMaxE; ++k)
                    if (((set >> (8 * k)) & 0xffu) != 0xffu) { ++m; n3 *= 3; }
                for (uint32_t sub = 0; sub < n3; ++sub) {
                    uint32_t xr = 0, q = sub;
                    for (uint32_t k = 0; k < kMaxE; ++k) {
                        const uint32_t off = (set >> (8 * k)) & 0xffu;
                        if (off != 0xffu) { xr ^= (1u + q % 3u) << (2u * off); q /= 3u; }
                    }
                    out.xy.push_back(xr);
                    out.xy.push_back(s | (m << 4) | ((set == 0xffffffffu ? 1u : 0u) << 8));
                }
            }
        }
        out.n[cnt] = (uint32_t)(out.xy.size() / 2) - out.off[cnt];
    }
    return true;
}

uint32_t default_jump_depth(uint64_t n_bwt)
{
    uint32_t d = 1; // ceil(log4(n_bwt)): the first depth at which a random d-mer is expected less than once
    while (d < 16 && (n_bwt >> (2 * d)) != 0) ++d;
    return d;
}

// ---------------------------------------------------------------------------------------------------
// blob layout
// ---------------------------------------------------------------------------------------------------
static uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

BlobPlan plan_blob(uint64_t n_text, uint32_t n_seq, bool with_sa, uint32_t sigma)
{
    BlobPlan p;
    std::memset(&p.h, 0, sizeof(p.h));
    IndexHeader& h = p.h;
    h.magic = kMagic;
    h.version = kVersion;
    h.sigma = sigma;
    h.n_text = n_text;
    h.n_seq = n_seq;
    h.n_bwt = n_text + n_seq;
    h.n_blocks = (uint32_t)(h.n_bwt / block_bases(sigma) + 1);
    uint64_t o = align_up(sizeof(IndexHeader), 256);
    h.off_fwd = o;        o = align_up(o + (uint64_t)h.n_blocks * block_bytes(sigma), 256);
    h.off_rev = o;        o = align_up(o + (uint64_t)h.n_blocks * block_bytes(sigma), 256);
    h.off_sent_fwd = o;   o = align_up(o + (uint64_t)n_seq * 4, 256);
    h.off_sent_rev = o;   o = align_up(o + (uint64_t)n_seq * 4, 256);
    h.off_text = o;       o = align_up(o + (n_text / 32 + 2) * 8, 256);
    h.off_limits = o;     o = align_up(o + ((uint64_t)n_seq + 1) * 8, 256);
    h.off_seq_start = o;  o = align_up(o + ((uint64_t)n_seq + 1) * 4, 256);
    if (sigma == 5) { h.off_nmask = o; o = align_up(o + (n_text / 64 + 2) * 8, 256); }
    if (with_sa) { h.off_sa = o; o = align_up(o + h.n_bwt * 4, 256); }
    h.total_bytes = o;
    return p;
}

void pack_bwt_blocks(const uint8_t* bwt, uint64_t n, RankBlock* blocks, uint32_t n_blocks, uint32_t* sent_pos,
                     uint32_t n_seq, uint64_t tot[4])
{
    uint64_t cnt[4] = {0, 0, 0, 0};
    uint32_t n_sent = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        RankBlock& B = blocks[b];
        std::memset(&B, 0, sizeof(B));
        B.cnt[0] = (uint32_t)cnt[0];
        B.cnt[1] = (uint32_t)cnt[1];
        B.cnt[2] = (uint32_t)cnt[2];
        const uint32_t sent_before = n_sent;
        for (uint32_t k = 0; k < kBlockBases; ++k) {
            const uint64_t i = (uint64_t)b * kBlockBases + k;
            if (i >= n) break;
            const uint8_t s = bwt[i];
            if (s < 2) { // sentinel row: stored as code 0, listed on the side
                if (n_sent < n_seq) sent_pos[n_sent] = (uint32_t)i;
                ++n_sent;
                continue;
            }
            const uint32_t c = s - 2u;
            ++cnt[c];
            B.w[k >> 6][0] |= (uint64_t)(c & 1u) << (k & 63);
            B.w[k >> 6][1] |= (uint64_t)(c >> 1) << (k & 63);
        }
        B.sent = (sent_before << 8) | (n_sent - sent_before);
    }
    for (int c = 0; c < 4; ++c) tot[c] = cnt[c];
}

void pack_bwt_blocks5(const uint8_t* bwt, uint64_t n, RankBlock5* blocks, uint32_t n_blocks, uint32_t* sent_pos,
                      uint32_t n_seq, uint64_t tot[5])
{
    uint64_t cnt[5] = {0, 0, 0, 0, 0};
    uint32_t n_sent = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        RankBlock5& B = blocks[b];
        std::memset(&B, 0, sizeof(B));
        for (int c = 0; c < 4; ++c) B.cnt[c] = (uint32_t)cnt[c];
        const uint32_t sent_before = n_sent;
        for (uint32_t k = 0; k < kBlockBases5; ++k) {
            const uint64_t i = (uint64_t)b * kBlockBases5 + k;
            if (i >= n) break;
            const uint8_t s = bwt[i];
            if (s < 2) { // sentinel row: stored as code 0, listed on the side
                if (n_sent < n_seq) sent_pos[n_sent] = (uint32_t)i;
                ++n_sent;
                continue;
            }
            const uint32_t c = s - 2u;
            ++cnt[c];
            for (int pl = 0; pl < 3; ++pl) B.plane[pl] |= ((c >> pl) & 1u) << k;
        }
        B.sent = (sent_before << 8) | (n_sent - sent_before);
    }
    for (int c = 0; c < 5; ++c) tot[c] = cnt[c];
}

namespace {

// T (or T') as SA-IS input: '#' = 0 (last symbol, unique smallest), '$' = 1, A,C,G,T = 2..5.
// Order = the one the reference gets from divsufsort on ord+1 with 0 sentinels
// (src/seqan_libdivsufsort.h:80-96): '$' < A, comparisons run through sentinels, shorter is smaller.
void make_symbols(const uint8_t* codes, const uint64_t* limits, uint32_t n_seq, bool reversed, std::vector<uint8_t>& t)
{
    const uint64_t n = limits[n_seq] + n_seq;
    t.resize(n);
    uint64_t o = 0;
    for (uint32_t s = 0; s < n_seq; ++s) {
        const uint64_t b = limits[s], e = limits[s + 1];
        if (!reversed) for (uint64_t k = b; k < e; ++k) t[o++] = (uint8_t)(codes[k] + 2);
        else           for (uint64_t k = e; k > b; --k) t[o++] = (uint8_t)(codes[k - 1] + 2); // src/indexing.hpp:130
        t[o++] = 1;
    }
    t[n - 1] = 0;
}

template <class Idx>
void build_direction(const std::vector<uint8_t>& t, void* blocks, uint32_t n_blocks, uint32_t* sent_pos,
                     uint32_t n_seq, uint64_t tot[5], uint32_t* sa_out, uint32_t sigma)
{
    const uint64_t n = t.size();
    std::vector<Idx> sa(n);
    suffix_array<Idx>(t.data(), sa.data(), (Idx)n, (Idx)7);
    std::vector<uint8_t> bwt(n);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t p = (uint64_t)sa[i];
        bwt[i] = p ? t[p - 1] : t[n - 1]; // src/seqan_libdivsufsort.h:165-229
        if (sa_out) sa_out[i] = (uint32_t)p;
    }
    if (sigma == 5) pack_bwt_blocks5(bwt.data(), n, static_cast<RankBlock5*>(blocks), n_blocks, sent_pos, n_seq, tot);
    else { tot[4] = 0; pack_bwt_blocks(bwt.data(), n, static_cast<RankBlock*>(blocks), n_blocks, sent_pos, n_seq, tot); }
}

} // namespace

bool build_index_host(const uint8_t* codes, const uint64_t* limits, uint32_t n_seq, bool with_sa, Blob& blob,
                      std::string& err)
{
    if (n_seq == 0 || limits[n_seq] == 0) { err = "There is no non-empty sequence in the fasta file(s)."; return false; }
    if (n_seq > kMaxSeq) { err = "too many sequences (limit 2^24 - 1)"; return false; }
    const uint64_t n_text = limits[n_seq];
    const uint64_t n = n_text + n_seq;
    if (n >= 0xFFFFFFFFull) { err = "index too large: text + sentinels must stay below 2^32 - 1"; return false; }
    for (uint32_t s = 0; s < n_seq; ++s)
        if (limits[s + 1] <= limits[s]) { err = "empty sequence in input (skip empty records before indexing)"; return false; }
    uint32_t sigma = 4;
    for (uint64_t i = 0; i < n_text; ++i) {
        if (codes[i] > 4) { err = "invalid base code (expected 0..3 = ACGT, 4 = N)"; return false; }
        if (codes[i] == 4) sigma = 5; // src/indexing.hpp:459-473: any N makes it a Dna5 index
    }

    BlobPlan plan = plan_blob(n_text, n_seq, with_sa, sigma);
    blob.resize(plan.h.total_bytes);
    uint8_t* base = blob.data();
    IndexHeader& h = *reinterpret_cast<IndexHeader*>(base);
    h = plan.h;

    std::vector<uint8_t> t;
    uint64_t tot[5] = {0, 0, 0, 0, 0};
    for (int rev = 0; rev < 2; ++rev) {
        make_symbols(codes, limits, n_seq, rev != 0, t);
        void* blocks = base + (rev ? h.off_rev : h.off_fwd);
        uint32_t* sent = reinterpret_cast<uint32_t*>(base + (rev ? h.off_sent_rev : h.off_sent_fwd));
        uint32_t* sa_out = (!rev && with_sa) ? reinterpret_cast<uint32_t*>(base + h.off_sa) : nullptr;
        if (n < 0x7FFFFFF0ull) build_direction<int32_t>(t, blocks, h.n_blocks, sent, n_seq, tot, sa_out, sigma);
        else                   build_direction<int64_t>(t, blocks, h.n_blocks, sent, n_seq, tot, sa_out, sigma);
    }
    // C array with the sentinels counted as smallest symbols (src/seqan_libdivsufsort.h:231-233)
    h.C[0] = n_seq;
    for (int c = 0; c < 5; ++c) h.C[c + 1] = h.C[c] + tot[c];

    uint64_t* text = reinterpret_cast<uint64_t*>(base + h.off_text);
    for (uint64_t i = 0; i < n_text; ++i) text[i >> 5] |= (uint64_t)(codes[i] & 3u) << (2 * (i & 31));
    if (sigma == 5) {
        uint64_t* nm = reinterpret_cast<uint64_t*>(base + h.off_nmask);
        for (uint64_t i = 0; i < n_text; ++i)
            if (codes[i] == 4) nm[i >> 6] |= 1ull << (i & 63);
    }
    uint64_t* lim = reinterpret_cast<uint64_t*>(base + h.off_limits);
    uint32_t* sst = reinterpret_cast<uint32_t*>(base + h.off_seq_start);
    for (uint32_t s = 0; s <= n_seq; ++s) { lim[s] = limits[s]; sst[s] = (uint32_t)(limits[s] + s); }
    return true;
}

namespace {

bool slurp(const std::string& path, std::vector<uint8_t>& out)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? (size_t)sz : 0);
    const size_t r = out.empty() ? 0 : std::fread(out.data(), 1, out.size(), f);
    std::fclose(f);
    return r == out.size();
}

// SeqAn fibres of one direction -> BWT symbols (0/1 = sentinel, 2..6 = A,C,G,T,N as pack_bwt_blocks[5] expect).
// Rank entries (SEQAN/index/index_fm_rank_dictionary_levels.h:197-209): one 64-bit word of packed values, value k in
// the bits starting at (per_word-1-k)*bits, followed by sigma-1 uint16 block counters: 32 two-bit values + 3 counters
// = 14 bytes (Dna4), 21 three-bit values + 4 counters = 16 bytes (Dna5).
bool decode_reference_bwt(const std::string& prefix, uint64_t n, uint32_t sigma, std::vector<uint8_t>& bwt, std::string& err)
{
    std::vector<uint8_t> drv, drp;
    if (!slurp(prefix + ".drv", drv) || !slurp(prefix + ".drp", drp)) { err = "cannot read " + prefix + ".drv/.drp"; return false; }
    // the block counters behind every word are uint16 in the reference's 32-bit BWT classes and uint32 in its 64-bit
    // ones (bwt_dimensions:64 — more than 65535 sequences, or a text beyond 2^32; src/indexing.hpp:151-170): entries of
    // 14 / 20 bytes (Dna4), 16 / 24 bytes (Dna5); the sentinel bit vector's 10 / 12 bytes.  Only the words are read.
    const uint32_t bits = sigma == 5 ? 3 : 2, per_word = 64 / bits;
    const uint64_t n_words = (n + per_word - 1) / per_word, n_words2 = (n + 63) / 64;
    const uint64_t entry = n_words ? drv.size() / n_words : 0, entry2 = n_words2 ? drp.size() / n_words2 : 0;
    if (n_words == 0 || drv.size() != n_words * entry || drp.size() != n_words2 * entry2 ||
        (entry != 8 + 2 * (sigma - 1) && entry != 8 + 4 * (sigma - 1)) || (entry2 != 10 && entry2 != 12)) {
        err = "unsupported reference index: unexpected size of the rank dictionary";
        return false;
    }
    bwt.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t w, s;
        std::memcpy(&w, drv.data() + (i / per_word) * entry, 8);
        std::memcpy(&s, drp.data() + (i / 64) * entry2, 8); // sentinel marker bits, bit k at 63-k
        const uint32_t v = (uint32_t)(w >> ((per_word - 1 - i % per_word) * bits)) & ((1u << bits) - 1u);
        if (v >= sigma) { err = "corrupt rank dictionary in the reference index"; return false; }
        bwt[i] = ((s >> (63 - (i % 64))) & 1u) ? 1 : (uint8_t)(v + 2);
    }
    return true;
}

} // namespace

bool import_reference_index(const std::string& dir, Blob& blob, std::string& err)
{
    std::string base = dir;
    if (!base.empty() && base.back() != '/') base += '/';
    base += "index";
    std::vector<uint8_t> info, limits_raw, text_raw;
    if (!slurp(base + ".info.concat", info)) { err = "cannot read " + base + ".info.concat"; return false; }
    const std::string info_s(info.begin(), info.end());
    const uint32_t sigma = info_s.find("alphabet_size:5") != std::string::npos ? 5u : 4u;
    if (sigma == 4 && info_s.find("alphabet_size:4") == std::string::npos) { err = "only Dna4 / Dna5 reference indices can be imported"; return false; }
    // every width class of the reference (src/indexing.hpp:151-170: (16,32,32), (16,32,64), (32,16,64), (64,64,64)) as long
    // as the text fits this layout's 32-bit rows (checked below); the classes differ in the counter widths of the rank
    // dictionaries (decode_reference_bwt) and in the sampled suffix array, which is not imported
    if (info_s.find("bwt_dimensions:") == std::string::npos || info_s.find("packed_text:true") == std::string::npos) {
        err = "not an index written by genmap >= 1.3 (index.info lacks bwt_dimensions / packed_text:true)";
        return false;
    }
    if (!slurp(base + ".txt.limits", limits_raw) || limits_raw.size() < 16 || limits_raw.size() % 8) { err = "cannot read " + base + ".txt.limits"; return false; }
    const uint32_t n_seq = (uint32_t)(limits_raw.size() / 8 - 1);
    std::vector<uint64_t> limits(n_seq + 1);
    std::memcpy(limits.data(), limits_raw.data(), limits_raw.size());
    const uint64_t n_text = limits[n_seq], n = n_text + n_seq;
    if (n_seq > kMaxSeq || n >= 0xFFFFFFFFull) { err = "reference index too large for the 32-bit HBM layout"; return false; }
    // packed text (src/indexing.hpp: packed_text:true): a length word, then 32 two-bit (Dna4) or 21 three-bit (Dna5) values per word
    const uint32_t tbits = sigma == 5 ? 3 : 2, tper = 64 / tbits;
    if (!slurp(base + ".txt.concat", text_raw) || text_raw.size() != ((n_text + tper - 1) / tper + 1) * 8) { err = "cannot read " + base + ".txt.concat (packed text expected)"; return false; }

    BlobPlan plan = plan_blob(n_text, n_seq, false, sigma);
    blob.resize(plan.h.total_bytes);
    uint8_t* b = blob.data();
    IndexHeader& h = *reinterpret_cast<IndexHeader*>(b);
    h = plan.h;
    uint64_t tot[5] = {0, 0, 0, 0, 0};
    std::vector<uint8_t> bwt;
    for (int rev = 0; rev < 2; ++rev) {
        if (!decode_reference_bwt(base + (rev ? ".rev.lf" : ".lf"), n, sigma, bwt, err)) return false;
        void* blocks = b + (rev ? h.off_rev : h.off_fwd);
        uint32_t* sent = reinterpret_cast<uint32_t*>(b + (rev ? h.off_sent_rev : h.off_sent_fwd));
        if (sigma == 5) pack_bwt_blocks5(bwt.data(), n, static_cast<RankBlock5*>(blocks), h.n_blocks, sent, n_seq, tot);
        else pack_bwt_blocks(bwt.data(), n, static_cast<RankBlock*>(blocks), h.n_blocks, sent, n_seq, tot);
    }
    h.C[0] = n_seq;
    for (int c = 0; c < 5; ++c) h.C[c + 1] = h.C[c] + tot[c];
    uint64_t* text = reinterpret_cast<uint64_t*>(b + h.off_text);
    uint64_t* nm = sigma == 5 ? reinterpret_cast<uint64_t*>(b + h.off_nmask) : nullptr;
    const uint8_t* words = text_raw.data() + 8; // first word = length
    for (uint64_t i = 0; i < n_text; ++i) {
        uint64_t w;
        std::memcpy(&w, words + (i / tper) * 8, 8);
        const uint32_t v = (uint32_t)(w >> ((tper - 1 - i % tper) * tbits)) & ((1u << tbits) - 1u);
        text[i >> 5] |= (uint64_t)(v & 3u) << (2 * (i & 31));
        if (v == 4 && nm) nm[i >> 6] |= 1ull << (i & 63);
    }
    uint64_t* lim = reinterpret_cast<uint64_t*>(b + h.off_limits);
    uint32_t* sst = reinterpret_cast<uint32_t*>(b + h.off_seq_start);
    for (uint32_t s = 0; s <= n_seq; ++s) { lim[s] = limits[s]; sst[s] = (uint32_t)(limits[s] + s); }
    return true;
}

void build_work_ranges(uint64_t text_len, uint32_t K, const uint64_t* chrom_cum, uint32_t n_chrom,
                       const uint64_t* intervals, uint64_t n_intervals, uint64_t pos_begin, uint64_t pos_end,
                       std::vector<WorkRange>& out)
{
    out.clear();
    if (pos_end > text_len) pos_end = text_len;
    std::vector<WorkRange> sel;
    if (n_intervals) {
        for (uint64_t k = 0; k < n_intervals; ++k) {
            WorkRange r{intervals[2 * k], std::min(intervals[2 * k + 1], text_len)};
            if (r.begin < r.end) sel.push_back(r);
        }
        std::sort(sel.begin(), sel.end(), [](const WorkRange& a, const WorkRange& b) { return a.begin < b.begin; });
        std::vector<WorkRange> merged;
        for (const WorkRange& r : sel) {
            if (!merged.empty() && r.begin <= merged.back().end) merged.back().end = std::max(merged.back().end, r.end);
            else merged.push_back(r);
        }
        sel.swap(merged);
    } else {
        sel.push_back(WorkRange{0, text_len});
    }
    size_t si = 0;
    for (uint32_t c = 0; c < n_chrom; ++c) {
        const uint64_t cb = chrom_cum[c], ce = chrom_cum[c + 1];
        if (ce - cb < K) continue;
        const uint64_t vb = std::max(cb, pos_begin), ve = std::min(ce - K + 1, pos_end);
        if (vb >= ve) continue;
        while (si < sel.size() && sel[si].end <= vb) ++si;
        for (size_t k = si; k < sel.size() && sel[k].begin < ve; ++k) {
            const uint64_t b = std::max(vb, sel[k].begin), e = std::min(ve, sel[k].end);
            if (b < e) out.push_back(WorkRange{b, e});
        }
    }
}

bool validate_header(const IndexHeader& h, uint64_t bytes, std::string& err)
{
    if (h.magic != kMagic) { err = "not a genmap-b200 index (bad magic)"; return false; }
    if (h.version != kVersion) { err = "index version mismatch: rebuild the index"; return false; }
    if (h.sigma != 4 && h.sigma != 5) { err = "unsupported alphabet size in the index header"; return false; }
    if (h.total_bytes > bytes) { err = "index blob truncated"; return false; }
    if (h.n_bwt != h.n_text + h.n_seq || h.n_blocks != h.n_bwt / block_bases(h.sigma) + 1) { err = "index header inconsistent"; return false; }
    if (h.sigma == 5 && !h.off_nmask) { err = "Dna5 index without N mask"; return false; }
    if (h.n_seq == 0 || h.n_seq > kMaxSeq || h.n_bwt >= 0xffffffffull) { err = "index header inconsistent (sizes)"; return false; }
    // every section where the layout puts it for these sizes (so every section fits and none overlaps), and the
    // cumulative counts consistent with the text size: a truncated or corrupt blob must not reach the kernels
    const IndexHeader want = plan_blob(h.n_text, h.n_seq, h.off_sa != 0, h.sigma).h;
    if (h.off_fwd != want.off_fwd || h.off_rev != want.off_rev || h.off_sent_fwd != want.off_sent_fwd || h.off_sent_rev != want.off_sent_rev ||
        h.off_text != want.off_text || h.off_limits != want.off_limits || h.off_seq_start != want.off_seq_start || h.off_sa != want.off_sa ||
        h.off_nmask != want.off_nmask || h.total_bytes != want.total_bytes) {
        err = "index header offsets do not match the layout of an index of this size";
        return false;
    }
    bool counts_ok = h.C[0] == h.n_seq && h.C[h.sigma] == h.n_bwt; // C[c] = symbols smaller than base c, sentinels first
    for (uint32_t c = 0; c < h.sigma; ++c) counts_ok = counts_ok && h.C[c] <= h.C[c + 1];
    if (!counts_ok) { err = "index header inconsistent (cumulative counts)"; return false; }
    return true;
}

bool validate_blob(const uint8_t* blob, uint64_t bytes, std::string& err)
{
    if (bytes < sizeof(IndexHeader)) { err = "index blob too small"; return false; }
    IndexHeader h;
    std::memcpy(&h, blob, sizeof(h));
    return validate_header(h, bytes, err);
}

} // namespace gmb
