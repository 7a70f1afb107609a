// seqan_export.cpp — writes an index in the REFERENCE's own on-disk format from an index blob of this library
// (gmb_blob_export_reference): the directory `genmap_ref map` opens (src/genmap_helper.hpp:71-127 lists the fibres,
// src/indexing.hpp:277-510 writes them).  Together with gmb_index_import_reference this makes the two index formats
// interchangeable: an index built on the GPU in seconds can be handed to the reference, or to tools built on it.
//
// Fibres and layouts (SeqAn 2.4 FM index as GenMap configures it, src/common.hpp:38-52; verified byte for byte
// against indices written by the reference itself: tests/test_seqan_index_writer.py):
//   index.info / index.ids       string sets: <name>.concat = the strings back to back, <name>.limits = u64 offsets
//   index.txt.concat / .limits   u64 length + ceil(n/32) u64 words, value k of a word in bits 62-2k ; u64 sequence limits
//   index[.rev].lf.drv           ceil(N/32) x { u64 word of 32 values (sentinel rows stored as A), u16 prefix[3] } = 14 bytes,
//                                prefix[c] = #values <= c before the block inside its superblock of 65504 values
//                                (LevelsPrefixRDConfig, SEQAN/index/index_fm_rank_dictionary_levels.h:197-209,1663-1680)
//   index[.rev].lf.drv.sbl       u32 prefix[3] per superblock
//   index[.rev].lf.drp / .sbl    sentinel bit vector: ceil(N/64) x { u64 bits (bit k at 63-k), u16 ones before the block in its
//                                superblock of 65472 values } ; u32 per superblock
//   index[.rev].lf.pst / .drs    u32 C array {#$, +A, +C, +G, +T} ; the sentinel substitute (one byte, 0 = A)
//   index.sa.ind / .val / .len   sampled suffix array: indicator bits ceil(N/64) x { u64 bits, u64 ones before } ;
//                                { u16 seqNo, u32 seqPos } per sampled row (seqPos % sampling == 0,
//                                src/seqan_libdivsufsort.h:135) ; u64 N
// Limits: Dna4 indices with the suffix-array section, at most 65535 sequences (the reference's (16,32,32) class).
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gmb_host.h"

namespace gmb {

namespace {

#pragma pack(push, 1)
struct LfEntry { uint64_t word; uint16_t prefix[3]; };   // 14 bytes
struct BitEntry16 { uint64_t bits; uint16_t ones; };     // 10 bytes
struct BitEntry64 { uint64_t bits; uint64_t ones; };     // 16 bytes
struct SaPair { uint16_t seq; uint32_t pos; };           // 6 bytes
#pragma pack(pop)

constexpr uint64_t kSuper = 65504, kSuperBits = 65472; // values per superblock: rank dictionary / sentinel bit vector

bool dump(const std::string& path, const void* p, size_t bytes, std::string& err)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    const size_t w = bytes ? std::fwrite(p, 1, bytes, f) : 0;
    if (std::fclose(f) != 0 || w != bytes) { err = "short write to " + path; return false; }
    return true;
}

template <class F>
void parallel_chunks(uint64_t n_chunks, F&& f)
{
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nt = (unsigned)std::min<uint64_t>(hw, n_chunks);
    if (nt <= 1) { for (uint64_t c = 0; c < n_chunks; ++c) f(c); return; }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&, t] { for (uint64_t c = t; c < n_chunks; c += nt) f(c); });
    for (std::thread& w : pool) w.join();
}

bool string_set(const std::string& path, const std::vector<std::string>& strings, std::string& err)
{
    std::string concat;
    std::vector<uint64_t> limits{0};
    for (const std::string& s : strings) { concat += s; limits.push_back(concat.size()); }
    return dump(path + ".concat", concat.data(), concat.size(), err) && dump(path + ".limits", limits.data(), limits.size() * 8, err);
}

// value of BWT row i: 0 = A .. 3 = T (sentinel rows are stored as A in the blob too); is_sentinel from the side list
inline uint32_t bwt_value(const RankBlock* B, uint64_t i)
{
    const RankBlock& b = B[i / kBlockBases];
    const uint32_t k = (uint32_t)(i % kBlockBases);
    return (uint32_t)((b.w[k >> 6][0] >> (k & 63)) & 1u) + 2u * (uint32_t)((b.w[k >> 6][1] >> (k & 63)) & 1u);
}

// one direction's LF fibres; tot[c] = #A,#C,#G,#T without the sentinel rows
bool write_lf(const std::string& prefix, const RankBlock* B, const uint32_t* sent_rows, uint32_t n_seq, uint64_t n, uint64_t tot[4],
              std::string& err)
{
    const uint64_t nb = (n + 31) / 32, nsb = (n + kSuper - 1) / kSuper, nb2 = (n + 63) / 64, nsb2 = (n + kSuperBits - 1) / kSuperBits;
    std::vector<LfEntry> drv(nb);
    std::vector<uint32_t> sbl(nsb * 3);
    std::vector<BitEntry16> drp(nb2);
    std::vector<uint32_t> sbl2(nsb2);
    // chunks of whole superblocks: stored-value counts per chunk, prefix sums, then the entries
    const uint64_t CH = kSuper * 64, nch = (n + CH - 1) / CH;
    std::vector<uint64_t> before((nch + 1) * 4, 0);
    parallel_chunks(nch, [&](uint64_t c) {
        uint64_t t[4] = {0, 0, 0, 0};
        const uint64_t e = std::min(n, (c + 1) * CH);
        for (uint64_t i = c * CH; i < e; ++i) ++t[bwt_value(B, i)];
        for (int k = 0; k < 4; ++k) before[4 * (c + 1) + k] = t[k];
    });
    for (uint64_t c = 1; c <= nch; ++c)
        for (int k = 0; k < 4; ++k) before[4 * c + k] += before[4 * (c - 1) + k];
    for (int k = 0; k < 4; ++k) tot[k] = before[4 * nch + k];
    tot[0] -= n_seq; // sentinel rows were counted as A
    parallel_chunks(nch, [&](uint64_t c) {
        uint64_t cnt[4] = {before[4 * c], before[4 * c + 1], before[4 * c + 2], before[4 * c + 3]};
        uint64_t base[3] = {0, 0, 0};
        const uint64_t b0 = c * CH / 32, b1 = std::min(nb, (c + 1) * CH / 32);
        for (uint64_t b = b0; b < b1; ++b) {
            if (b % (kSuper / 32) == 0) { // every chunk starts on a superblock boundary
                base[0] = cnt[0]; base[1] = cnt[0] + cnt[1]; base[2] = cnt[0] + cnt[1] + cnt[2];
                for (int k = 0; k < 3; ++k) sbl[3 * (b / (kSuper / 32)) + k] = (uint32_t)base[k];
            }
            drv[b].prefix[0] = (uint16_t)(cnt[0] - base[0]);
            drv[b].prefix[1] = (uint16_t)(cnt[0] + cnt[1] - base[1]);
            drv[b].prefix[2] = (uint16_t)(cnt[0] + cnt[1] + cnt[2] - base[2]);
            uint64_t w = 0;
            for (uint32_t k = 0; k < 32 && b * 32 + k < n; ++k) {
                const uint32_t v = bwt_value(B, b * 32 + k);
                w |= (uint64_t)v << (62 - 2 * k);
                ++cnt[v];
            }
            drv[b].word = w;
        }
    });
    for (uint32_t s = 0; s < n_seq; ++s) drp[sent_rows[s] / 64].bits |= 1ull << (63 - sent_rows[s] % 64);
    uint64_t ones = 0, sb_ones = 0;
    for (uint64_t b = 0; b < nb2; ++b) {
        if (b % (kSuperBits / 64) == 0) { sb_ones = ones; sbl2[b / (kSuperBits / 64)] = (uint32_t)sb_ones; }
        drp[b].ones = (uint16_t)(ones - sb_ones);
        ones += (uint64_t)__builtin_popcountll(drp[b].bits);
    }
    uint32_t pst[5];
    pst[0] = n_seq;
    for (int c = 0; c < 4; ++c) pst[c + 1] = pst[c] + (uint32_t)tot[c];
    const char drs = 0;
    return dump(prefix + ".drv", drv.data(), nb * sizeof(LfEntry), err) && dump(prefix + ".drv.sbl", sbl.data(), sbl.size() * 4, err) &&
           dump(prefix + ".drp", drp.data(), nb2 * sizeof(BitEntry16), err) && dump(prefix + ".drp.sbl", sbl2.data(), sbl2.size() * 4, err) &&
           dump(prefix + ".pst", pst, sizeof pst, err) && dump(prefix + ".drs", &drs, 1, err);
}

bool write_sa(const std::string& prefix, const uint32_t* sa, uint64_t n, const uint32_t* seq_start, uint32_t n_seq, uint32_t sampling,
              std::string& err)
{
    const uint64_t nb = (n + 63) / 64;
    std::vector<BitEntry64> ind(nb);
    auto seq_of = [&](uint64_t p) { // largest s with seq_start[s] <= p
        uint32_t lo = 0, hi = n_seq;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (seq_start[mid] <= p) lo = mid; else hi = mid; }
        return lo;
    };
    const uint64_t per = 1u << 14, nch = (nb + per - 1) / per; // blocks per chunk
    parallel_chunks(nch, [&](uint64_t c) {
        for (uint64_t b = c * per; b < std::min(nb, (c + 1) * per); ++b) {
            uint64_t w = 0;
            for (uint32_t k = 0; k < 64 && b * 64 + k < n; ++k) {
                const uint64_t p = sa[b * 64 + k];
                const uint32_t s = seq_of(p);
                if (p + 1 == seq_start[s + 1]) continue; // a sentinel position is never sampled
                if ((p - seq_start[s]) % sampling == 0) w |= 1ull << (63 - k);
            }
            ind[b].bits = w;
        }
    });
    uint64_t ones = 0;
    for (uint64_t b = 0; b < nb; ++b) { ind[b].ones = ones; ones += (uint64_t)__builtin_popcountll(ind[b].bits); }
    std::vector<SaPair> val(ones);
    parallel_chunks(nch, [&](uint64_t c) {
        for (uint64_t b = c * per; b < std::min(nb, (c + 1) * per); ++b) {
            uint64_t o = ind[b].ones;
            for (uint32_t k = 0; k < 64; ++k) {
                if (!((ind[b].bits >> (63 - k)) & 1ull)) continue;
                const uint64_t p = sa[b * 64 + k];
                const uint32_t s = seq_of(p);
                val[o].seq = (uint16_t)s;
                val[o].pos = (uint32_t)(p - seq_start[s]);
                ++o;
            }
        }
    });
    const uint64_t len = n;
    return dump(prefix + ".ind", ind.data(), nb * sizeof(BitEntry64), err) && dump(prefix + ".val", val.data(), val.size() * sizeof(SaPair), err) &&
           dump(prefix + ".len", &len, 8, err);
}

} // namespace

bool export_reference_index(const uint8_t* blob, uint64_t bytes, const std::string& dir, const std::vector<std::string>& ids,
                            bool fasta_directory, uint32_t sampling, std::string& err)
{
    if (!validate_blob(blob, bytes, err)) return false;
    IndexHeader h;
    std::memcpy(&h, blob, sizeof h);
    if (h.sigma != 4) { err = "the reference-format writer handles Dna4 indices only"; return false; }
    if (!h.off_sa) { err = "the reference format stores a sampled suffix array: build the index with the suffix array (without --no-sa)"; return false; }
    if (h.n_seq > 65535) { err = "more than 65535 sequences: not the reference's (16,32,32) index class"; return false; }
    if (ids.size() != h.n_seq) { err = "one id line per indexed sequence is needed"; return false; }
    if (sampling == 0) sampling = 10;
    const std::string base = dir + (dir.empty() || dir.back() == '/' ? "" : "/") + "index";
    char num[32];
    std::snprintf(num, sizeof num, "%u", sampling);
    if (!string_set(base + ".info", {"alphabet_size:4", "sa_dimensions_i1:16", "sa_dimensions_i2:32", "bwt_dimensions:32",
                                     std::string("sampling_rate:") + num, std::string("fasta_directory:") + (fasta_directory ? "true" : "false"),
                                     "packed_text:true"}, err) ||
        !string_set(base + ".ids", ids, err))
        return false;
    // text: our words hold character i in bits 2(i & 31), the reference's value k in bits 62 - 2k: the 2-bit groups reversed
    const uint64_t* text = reinterpret_cast<const uint64_t*>(blob + h.off_text);
    const uint64_t nw = (h.n_text + 31) / 32;
    std::vector<uint64_t> packed(nw + 1);
    packed[0] = h.n_text;
    for (uint64_t w = 0; w < nw; ++w) {
        uint64_t x = text[w];
        if (w == nw - 1 && h.n_text % 32) x &= (1ull << (2 * (h.n_text % 32))) - 1ull;
        x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
        x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
        packed[w + 1] = __builtin_bswap64(x);
    }
    if (!dump(base + ".txt.concat", packed.data(), packed.size() * 8, err) ||
        !dump(base + ".txt.limits", blob + h.off_limits, ((uint64_t)h.n_seq + 1) * 8, err))
        return false;
    uint64_t tot[4];
    if (!write_lf(base + ".lf", reinterpret_cast<const RankBlock*>(blob + h.off_fwd), reinterpret_cast<const uint32_t*>(blob + h.off_sent_fwd),
                  h.n_seq, h.n_bwt, tot, err) ||
        !write_lf(base + ".rev.lf", reinterpret_cast<const RankBlock*>(blob + h.off_rev), reinterpret_cast<const uint32_t*>(blob + h.off_sent_rev),
                  h.n_seq, h.n_bwt, tot, err))
        return false;
    return write_sa(base + ".sa", reinterpret_cast<const uint32_t*>(blob + h.off_sa), h.n_bwt, reinterpret_cast<const uint32_t*>(blob + h.off_seq_start),
                    h.n_seq, sampling, err);
}

} // namespace gmb
