// map_kernel.cu — the GENERAL (K,E)-frequency kernel for sm_100a: every configuration the two specialised kernels do
// not take (exact_kernel.cu: E = 0, K <= 64; block_kernel.cu: E = 1, 2, 4 on Dna4, needle <= 64): E = 3, Dna5 indices at E >= 1, K > 64,
// jump tables switched off; its locate instantiation serves the csv lists (locate_kernel.cu).
//
// Replaces the reference's per-position loop computeMappability -> computeMappabilitySingleBlock
// (src/algo.hpp:221-483) and its search-scheme matcher (src/find2_index_approx.hpp:223-457).
//
// Mapping to the machine:
//   * one CUDA thread = one "chain" = one block of k-mer starts at a time; a chain is a strictly dependent series
//     of random 32-byte rank-block reads, so throughput comes from the number of chains in flight
//     (148 SMs x 768-1024 resident threads), not from intra-chain width.
//   * each loop iteration is one node expansion of gmb_core.h::chain_step — the same code for every
//     chain whatever its depth, error level or search, so warps stay converged although every lane
//     walks a different subtree.
//   * persistent grid (SM count x resident CTAs); lanes that finish their k-mer refill immediately
//     from a warp-local pool fed by one global atomic per 128 positions, so repeats (whose searches are
//     100x longer) never idle a warp.
//   * the <= E backtracking frames live in shared memory, strided by thread (conflict-free); the
//     step table of the search scheme is staged in shared memory once per CTA.
#include "map_kernel_impl.cuh"

namespace gmb {

namespace {

template <int KW, bool COUNT, typename OutT, bool EP>
cudaError_t launch_t(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.sigma == 5)
        return L.cx.B > 1 ? launch_b<KW, COUNT, OutT, EP, true, 5>(L, sm_count, stream) : launch_b<KW, COUNT, OutT, EP, false, 5>(L, sm_count, stream);
    if (L.cx.B <= 1) return launch_b<KW, COUNT, OutT, EP, false, 4>(L, sm_count, stream);
    // Blocked Dna4 kernels: up to E = 2 they are bound by instruction issue since the jump tables answer for the dense
    // top of the trie, and 3 CTAs x 80 registers (no spills) beat 4 x 64; E >= 3 still wants the extra chains in flight
    // (profiles/r01/s28_sweep_*.txt: E=1 35.0 -> 31.9 ms, E=2 35.2 -> 33.9 ms, E=3 19.7 -> 22.6 ms with 3 CTAs)
    return L.E <= 2 ? launch_b<KW, COUNT, OutT, EP, true, 4, false, 3>(L, sm_count, stream)
                    : launch_b<KW, COUNT, OutT, EP, true, 4, false, GMB_MIN_BLOCKS>(L, sm_count, stream);
}

template <int KW>
cudaError_t launch_kw(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.exclude_pseudo) // (the fetch-counting instantiation is not built for --exclude-pseudo)
        return L.value_bits == 16 ? launch_t<KW, false, uint16_t, true>(L, sm_count, stream)
                                  : launch_t<KW, false, uint8_t, true>(L, sm_count, stream);
    if (L.value_bits == 16)
        return L.count_fetches ? launch_t<KW, true, uint16_t, false>(L, sm_count, stream)
                               : launch_t<KW, false, uint16_t, false>(L, sm_count, stream);
    return L.count_fetches ? launch_t<KW, true, uint8_t, false>(L, sm_count, stream)
                           : launch_t<KW, false, uint8_t, false>(L, sm_count, stream);
}

} // namespace

size_t map_kernel_smem_bytes(uint32_t n_step_words, uint32_t E, uint32_t B, bool ep, uint32_t sigma, bool blocked)
{
    const size_t tables = align32(n_step_words) + align32((B + 1) * kMaxSearches * kStartWords) + align32(2 * (kMaxBlockKmers + 1));
    return (tables + (size_t)frame_store_words(E, B, ep, (int)sigma, blocked) * kThreads) * sizeof(uint32_t);
}

cudaError_t launch_map_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.n_work == 0) return cudaSuccess;
    if (exact_kernel_applies(L)) return launch_exact_kernel(L, sm_count, stream);
    if (block_kernel_applies(L)) return launch_block_kernel(L, sm_count, stream);
    const uint32_t needle = L.cx.K + L.cx.B - 1; // characters a chain keeps in registers
    if (needle <= 32) return launch_kw<1>(L, sm_count, stream);
    if (needle <= 64) return launch_kw<2>(L, sm_count, stream);
    if (needle <= 128) return launch_kw<4>(L, sm_count, stream);
    return launch_kw<9>(L, sm_count, stream);
}

} // namespace gmb
