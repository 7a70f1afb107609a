// jump_table.cu — builds the jump tables that replace the first d error-free steps of every search.
//
// No counterpart in the reference (which walks every pattern from the root); the idea is the classic
// k-mer lookup table of FM-index mappers, sized for HBM: at depth 15 (3 Gbp default) the table has
// 4^15 entries x 12 bytes = 12.9 GB, about one expected occurrence per entry.
#include "jump_table.cuh"

namespace gmb {

namespace {

template <int SIGMA>
__global__ void k_jump_level(const MapCtx cx, uint32_t d, const JtEntry* __restrict__ prev_uni,
                             const uint32_t* __restrict__ prev_lof, JtEntry* __restrict__ out_uni,
                             uint32_t* __restrict__ out_lof, JtFull* __restrict__ out_full)
{
    const uint64_t n = 1ull << (2 * d);
    const uint64_t key = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= n) return;
    Node par;
    if (d == 1) {
        par.lo_f = 0; par.lo_r = 0; par.size = cx.n_bwt;
    } else {
        const uint64_t pk = key & ((1ull << (2 * (d - 1))) - 1ull);
        const JtEntry e = prev_uni[pk];
        par.lo_f = prev_lof ? prev_lof[pk] : 0u; par.lo_r = e.lo_r; par.size = e.size;
    }
    const Node m = extend_right<SIGMA>(par, (uint32_t)(key >> (2 * (d - 1))), cx); // keys are A,C,G,T only
    if (out_full) {
        JtFull f;
        f.lo_r = m.lo_r; f.size = m.size; f.lo_f = m.lo_f; f.pad = 0;
        out_full[key] = f;
        return;
    }
    JtEntry o;
    o.lo_r = m.lo_r; o.size = m.size;
    out_uni[key] = o;
    if (out_lof) out_lof[key] = m.lo_f;
}

} // namespace

cudaError_t build_jump_level(const MapCtx& cx, uint32_t sigma, uint32_t d, const JtEntry* prev_uni, const uint32_t* prev_lof,
                             JtEntry* out_uni, uint32_t* out_lof, JtFull* out_full, cudaStream_t stream)
{
    const uint64_t n = 1ull << (2 * d);
    const unsigned threads = 256;
    const unsigned long long blocks = (n + threads - 1) / threads;
    if (sigma == 5) k_jump_level<5><<<(unsigned)blocks, threads, 0, stream>>>(cx, d, prev_uni, prev_lof, out_uni, out_lof, out_full);
    else k_jump_level<4><<<(unsigned)blocks, threads, 0, stream>>>(cx, d, prev_uni, prev_lof, out_uni, out_lof, out_full);
    return cudaGetLastError();
}

} // namespace gmb
