// sais.hpp — host suffix-array construction by induced sorting (SA-IS, Nong/Zhang/Chan 2009),
// written for this repository.  Used by the host index builder (`genmap index` without --gpu and the
// CPU-side tests); the reference uses libdivsufsort / SeqAn Skew7 for the same job
// (src/seqan_libdivsufsort.h:96, SEQAN/index/index_fm.h:446).  Only the resulting order matters.
//
// Contract: s[n-1] must be the unique smallest symbol of s.  Idx is a signed integer type wide
// enough for n (int32_t below 2^31, int64_t otherwise).
#pragma once
#include <cstdint>
#include <vector>

namespace gmb {
namespace sais_detail {

struct BitVec {
    std::vector<uint64_t> w;
    explicit BitVec(uint64_t n) : w((n + 63) / 64, 0) {}
    inline bool get(uint64_t i) const { return (w[i >> 6] >> (i & 63)) & 1u; }
    inline void set(uint64_t i, bool v)
    {
        if (v) w[i >> 6] |= (1ull << (i & 63));
        else w[i >> 6] &= ~(1ull << (i & 63));
    }
};

template <class Idx, class Sym>
void buckets(const Sym* s, Idx n, Idx K, std::vector<Idx>& bkt, bool end)
{
    for (Idx i = 0; i < K; ++i) bkt[i] = 0;
    for (Idx i = 0; i < n; ++i) ++bkt[(Idx)s[i]];
    Idx sum = 0;
    for (Idx i = 0; i < K; ++i) {
        sum += bkt[i];
        bkt[i] = end ? sum : sum - bkt[i];
    }
}

template <class Idx, class Sym>
void induce(const Sym* s, Idx* SA, Idx n, Idx K, std::vector<Idx>& bkt, const BitVec& t)
{
    buckets(s, n, K, bkt, false);
    for (Idx i = 0; i < n; ++i) { // L-type suffixes, left to right
        Idx j = SA[i] - 1;
        if (SA[i] > 0 && !t.get((uint64_t)j)) SA[bkt[(Idx)s[j]]++] = j;
    }
    buckets(s, n, K, bkt, true);
    for (Idx i = n - 1; i >= 0; --i) { // S-type suffixes, right to left
        Idx j = SA[i] - 1;
        if (SA[i] > 0 && t.get((uint64_t)j)) SA[--bkt[(Idx)s[j]]] = j;
    }
}

template <class Idx, class Sym>
void sais(const Sym* s, Idx* SA, Idx n, Idx K)
{
    if (n == 1) { SA[0] = 0; return; }
    BitVec t((uint64_t)n); // 1 = S-type
    t.set((uint64_t)n - 1, true);
    for (Idx i = n - 2; i >= 0; --i)
        t.set((uint64_t)i, s[i] < s[i + 1] || (s[i] == s[i + 1] && t.get((uint64_t)i + 1)));
    auto is_lms = [&](Idx i) { return i > 0 && t.get((uint64_t)i) && !t.get((uint64_t)i - 1); };

    std::vector<Idx> bkt((size_t)K);
    // stage 1: sort the LMS substrings
    buckets(s, n, K, bkt, true);
    for (Idx i = 0; i < n; ++i) SA[i] = -1;
    for (Idx i = 1; i < n; ++i)
        if (is_lms(i)) SA[--bkt[(Idx)s[i]]] = i;
    induce(s, SA, n, K, bkt, t);

    Idx n1 = 0;
    for (Idx i = 0; i < n; ++i)
        if (is_lms(SA[i])) SA[n1++] = SA[i];
    for (Idx i = n1; i < n; ++i) SA[i] = -1;
    Idx name = 0, prev = -1;
    for (Idx i = 0; i < n1; ++i) {
        Idx pos = SA[i];
        bool diff = false;
        for (Idx d = 0; d < n; ++d) {
            if (prev == -1 || s[pos + d] != s[prev + d] || t.get((uint64_t)(pos + d)) != t.get((uint64_t)(prev + d))) {
                diff = true;
                break;
            }
            if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
        }
        if (diff) { ++name; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (Idx i = n - 1, j = n - 1; i >= n1; --i)
        if (SA[i] >= 0) SA[j--] = SA[i];

    // stage 2: order of the LMS suffixes from the reduced string
    Idx* SA1 = SA;
    Idx* s1 = SA + n - n1;
    if (name < n1) sais<Idx, Idx>(s1, SA1, n1, name);
    else for (Idx i = 0; i < n1; ++i) SA1[s1[i]] = i;

    // stage 3: induce the full order
    buckets(s, n, K, bkt, true);
    {
        Idx j = 0;
        for (Idx i = 1; i < n; ++i)
            if (is_lms(i)) s1[j++] = i;
    }
    for (Idx i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
    for (Idx i = n1; i < n; ++i) SA[i] = -1;
    for (Idx i = n1 - 1; i >= 0; --i) {
        Idx j = SA[i];
        SA[i] = -1;
        SA[--bkt[(Idx)s[j]]] = j;
    }
    induce(s, SA, n, K, bkt, t);
}

} // namespace sais_detail

// s: n symbols in [0,K), s[n-1] unique smallest.  SA receives the n suffix start positions in order.
template <class Idx>
inline void suffix_array(const uint8_t* s, Idx* SA, Idx n, Idx K)
{
    sais_detail::sais<Idx, uint8_t>(s, SA, n, K);
}

} // namespace gmb
