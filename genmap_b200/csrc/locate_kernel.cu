// locate_kernel.cu — the locate instantiation of the search kernel and its post-processing: what the csv
// output (`genmap map -d`) needs.
//
// Replaces the csvComputation branch of computeMappabilitySingleBlock (src/algo.hpp:311-343): for every
// k-mer the reference keeps the iterators of all hits (itAll / itAllrevCompl), locates every occurrence
// through the sampled suffix array (getOccurrences -> CompressedSA::value,
// SEQAN/index/index_fm_compressed_sa.h:478-513) into two std::vectors and sorts them.  Here:
//   pass 1  the search kernel (one k-mer per chain, both intervals kept in step) counts the occurrences of
//           every k-mer per strand -> two uint32 per position;
//   scan    exclusive prefix sum of the counts = where every list starts (cub::DeviceScan);
//   pass 2  the same search again; every full-length node copies its SA rows (the full suffix array sits
//           in HBM: one read per occurrence, no LF walk) to its k-mer's list;
//   sort    cub::DeviceSegmentedSort orders every list by text position (= by (sequence, offset), the
//           order std::sort gives the reference's Pair<seqNo, seqPos>, src/algo.hpp:335,346);
//   convert one thread per occurrence turns the position inside T into (sequence number, offset).
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "map_kernel_impl.cuh"
#include "locate.cuh"

namespace gmb {

namespace {

template <int KW>
cudaError_t launch_loc_kw(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    return L.sigma == 5 ? launch_b<KW, false, uint32_t, true, false, 5, true>(L, sm_count, stream)
                        : launch_b<KW, false, uint32_t, true, false, 4, true>(L, sm_count, stream);
}

// position inside the sentinel-separated text T -> (sequence, offset); seq_start[s] = start of sequence s in T
__global__ void k_rows_to_locations(const uint32_t* __restrict__ rows, uint64_t n, const uint32_t* __restrict__ seq_start,
                                    uint32_t n_seq, uint2* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pos = rows[i];
    uint32_t a = 0, b = n_seq; // largest s with seq_start[s] <= pos
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if (__ldg(seq_start + mid) <= pos) a = mid; else b = mid;
    }
    out[i] = make_uint2(a, pos - __ldg(seq_start + a));
}

__device__ __forceinline__ uint32_t file_of_row(uint32_t pos, const uint32_t* __restrict__ seq_start, uint32_t n_seq,
                                                const uint32_t* __restrict__ seq_to_file)
{
    uint32_t a = 0, b = n_seq;
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if (__ldg(seq_start + mid) <= pos) a = mid; else b = mid;
    }
    return __ldg(seq_to_file + a);
}

// distinct files in the union of the + and - lists of one position (src/algo.hpp:351-361); both lists are sorted by
// text position, so their file ids do not decrease: a two-pointer merge counts the changes
template <typename OutT>
__global__ void k_distinct_files(const uint32_t* __restrict__ rows, const uint64_t* __restrict__ off, uint64_t n_pos,
                                 const uint32_t* __restrict__ seq_start, uint32_t n_seq, const uint32_t* __restrict__ seq_to_file,
                                 OutT* __restrict__ out, uint64_t pos0)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pos) return;
    uint64_t a = off[2 * j], b = off[2 * j + 1];
    const uint64_t ea = b, eb = off[2 * j + 2];
    if (a == eb) return; // not searched / no occurrence: the caller's zero stays
    constexpr uint32_t kNone = 0xffffffffu;
    uint32_t fa = a < ea ? file_of_row(rows[a], seq_start, n_seq, seq_to_file) : kNone;
    uint32_t fb = b < eb ? file_of_row(rows[b], seq_start, n_seq, seq_to_file) : kNone;
    uint32_t last = kNone, cnt = 0;
    while (fa != kNone || fb != kNone) {
        const uint32_t f = fa < fb ? fa : fb;
        if (f != last) { ++cnt; last = f; }
        if (fa == f) { ++a; fa = a < ea ? file_of_row(rows[a], seq_start, n_seq, seq_to_file) : kNone; }
        else { ++b; fb = b < eb ? file_of_row(rows[b], seq_start, n_seq, seq_to_file) : kNone; }
    }
    out[pos0 + j] = (OutT)cnt;
}

// ---- the N pass of Dna5 calls whose searches skip the text's N (MapCtx::skip_n; capi.cu: NFix) ---------------------
// number of N among the text positions [t, t + K) (nmask: one bit per position)
__device__ __forceinline__ uint32_t n_in_window(const uint64_t* __restrict__ nmask, uint64_t t, uint32_t K)
{
    uint32_t cnt = 0;
    for (uint64_t pos = t, end = t + K; pos < end;) {
        const uint32_t off = (uint32_t)(pos & 63u);
        const uint64_t left = end - pos;
        const uint32_t take = left < 64u - off ? (uint32_t)left : 64u - off;
        const uint64_t bits = (__ldg(nmask + (pos >> 6)) >> off) & (take == 64u ? ~0ull : ((1ull << take) - 1ull));
        cnt += (uint32_t)__popcll(bits);
        pos += take;
    }
    return cnt;
}

// every window start t whose K characters lie inside one sequence and hold 1..E N -> out_pos (any order)
__global__ void k_nwin_collect(const uint64_t* __restrict__ nmask, uint64_t n_text, const uint32_t* __restrict__ seq_start, uint32_t n_seq,
                               uint32_t K, uint32_t E, uint32_t* __restrict__ out_pos, unsigned long long* __restrict__ counter,
                               unsigned long long cap)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t + K > n_text) return;
    const uint32_t cnt = n_in_window(nmask, t, K);
    if (cnt < 1u || cnt > E) return;
    uint32_t a = 0, b = n_seq; // largest s with limits[s] <= t, limits[s] = seq_start[s] - s
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if ((uint64_t)__ldg(seq_start + mid) - mid <= t) a = mid; else b = mid;
    }
    if (t + K > (uint64_t)__ldg(seq_start + a + 1) - (a + 1)) return; // the window crosses into the next sequence
    const unsigned long long at = atomicAdd(counter, 1ull);
    if (at < cap) out_pos[at] = (uint32_t)t;
}

// the located lists of m N windows: counts[i] = occurrences of window i on both strands; every occurrence whose own
// window holds no N -> hits (position in the concatenated text; any order)
__global__ void k_nfix_collect(const uint32_t* __restrict__ rows, const uint64_t* __restrict__ off, uint64_t m, uint64_t n_rows,
                               const uint32_t* __restrict__ seq_start, uint32_t n_seq, const uint64_t* __restrict__ nmask, uint32_t K,
                               uint32_t* __restrict__ counts, uint32_t* __restrict__ hits, unsigned long long* __restrict__ counter)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const uint64_t c = off[2 * i + 2] - off[2 * i];
        counts[i] = c < 0xffffffffull ? (uint32_t)c : 0xffffffffu;
    }
    if (i >= n_rows) return;
    const uint32_t pos = rows[i];
    uint32_t a = 0, b = n_seq; // largest s with seq_start[s] <= pos
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if (__ldg(seq_start + mid) <= pos) a = mid; else b = mid;
    }
    const uint32_t j = pos - a; // s sentinels precede sequence s in T
    if (n_in_window(nmask, j, K) != 0u) return; // a window with N gets its whole count from its own lists
    hits[atomicAdd(counter, 1ull)] = j;
}

template <typename OutT>
__device__ __forceinline__ void saturating_inc(OutT* p, uint32_t maxv)
{
    constexpr uint32_t bits = sizeof(OutT) * 8u, mask = (1u << bits) - 1u;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    unsigned int* w = reinterpret_cast<unsigned int*>(addr & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(addr & 3u) * 8u;
    unsigned int old = *w, assumed;
    do {
        assumed = old;
        const uint32_t v = (assumed >> sh) & mask;
        if (v >= maxv) return;
        old = atomicCAS(w, assumed, (assumed & ~(mask << sh)) | ((v + 1u) << sh));
    } while (old != assumed);
}

// phase 0: out[t] = count of N window t (overwrites what the search left there); phase 1: out[j] += 1 per hit.
// Only positions of this call: inside [text_begin, ...) and inside one of its work ranges.
template <typename OutT>
__global__ void k_nfix_apply(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ counts, uint64_t n, uint64_t text_begin,
                             const uint64_t* __restrict__ range_begin, const uint64_t* __restrict__ range_end, uint32_t n_ranges,
                             OutT* __restrict__ out, uint32_t maxv)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t g = pos[i];
    if (g < text_begin) return;
    const uint64_t j = g - text_begin;
    uint32_t a = 0, b = n_ranges; // largest r with range_begin[r] <= j
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if (__ldg(range_begin + mid) <= j) a = mid; else b = mid;
    }
    if (j < __ldg(range_begin + a) || j >= __ldg(range_end + a)) return;
    if (counts) { const uint32_t c = counts[i]; out[j] = (OutT)(c < maxv ? c : maxv); }
    else saturating_inc(out + j, maxv);
}

struct CountToU64 {
    __host__ __device__ uint64_t operator()(uint32_t x) const { return x; }
};

} // namespace

cudaError_t launch_locate_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream)
{
    if (L.n_work == 0) return cudaSuccess;
    if (L.cx.K <= 32) return launch_loc_kw<1>(L, sm_count, stream);
    if (L.cx.K <= 64) return launch_loc_kw<2>(L, sm_count, stream);
    if (L.cx.K <= 128) return launch_loc_kw<4>(L, sm_count, stream);
    return launch_loc_kw<9>(L, sm_count, stream);
}

cudaError_t locate_scan_counts(const uint32_t* counts, uint64_t n_lists, uint64_t* offsets, void* temp, size_t& temp_bytes,
                               cudaStream_t stream)
{
    // offsets[0 .. n_lists] = exclusive sums of counts[0 .. n_lists) followed by the total: scan n_lists + 1
    // items (the caller keeps counts[n_lists] == 0)
    cub::TransformInputIterator<uint64_t, CountToU64, const uint32_t*> in(counts, CountToU64());
    return cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, offsets, (int64_t)(n_lists + 1), stream);
}

cudaError_t locate_sort_lists(const uint32_t* rows_in, uint32_t* rows_out, uint64_t n_rows, const uint64_t* offsets,
                              uint64_t n_lists, void* temp, size_t& temp_bytes, cudaStream_t stream)
{
    return cub::DeviceSegmentedSort::SortKeys(temp, temp_bytes, rows_in, rows_out, (int64_t)n_rows, (int64_t)n_lists, offsets,
                                              offsets + 1, stream);
}

cudaError_t locate_convert(const uint32_t* rows, uint64_t n_rows, const uint32_t* seq_start, uint32_t n_seq, void* out,
                           cudaStream_t stream)
{
    if (n_rows == 0) return cudaSuccess;
    k_rows_to_locations<<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>(rows, n_rows, seq_start, n_seq, static_cast<uint2*>(out));
    return cudaGetLastError();
}

cudaError_t locate_distinct_files(const uint32_t* rows, const uint64_t* offsets, uint64_t n_pos, const uint32_t* seq_start,
                                  uint32_t n_seq, const uint32_t* seq_to_file, void* out, uint32_t value_bits, uint64_t pos0,
                                  cudaStream_t stream)
{
    if (n_pos == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_pos + 255) / 256);
    if (value_bits == 16) k_distinct_files<<<grid, 256, 0, stream>>>(rows, offsets, n_pos, seq_start, n_seq, seq_to_file, static_cast<uint16_t*>(out), pos0);
    else k_distinct_files<<<grid, 256, 0, stream>>>(rows, offsets, n_pos, seq_start, n_seq, seq_to_file, static_cast<uint8_t*>(out), pos0);
    return cudaGetLastError();
}

} // namespace gmb

namespace gmb {

cudaError_t nfix_collect_windows(const uint64_t* nmask, uint64_t n_text, const uint32_t* seq_start, uint32_t n_seq, uint32_t K, uint32_t E,
                                 uint32_t* out_pos, unsigned long long* counter, uint64_t cap, cudaStream_t stream)
{
    if (n_text < K) return cudaSuccess;
    const uint64_t n = n_text - K + 1;
    k_nwin_collect<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(nmask, n_text, seq_start, n_seq, K, E, out_pos, counter, cap);
    return cudaGetLastError();
}

cudaError_t nfix_collect_hits(const uint32_t* rows, const uint64_t* offsets, uint64_t m, uint64_t n_rows, const uint32_t* seq_start,
                              uint32_t n_seq, const uint64_t* nmask, uint32_t K, uint32_t* counts, uint32_t* hits,
                              unsigned long long* counter, cudaStream_t stream)
{
    const uint64_t n = m > n_rows ? m : n_rows;
    if (n == 0) return cudaSuccess;
    k_nfix_collect<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rows, offsets, m, n_rows, seq_start, n_seq, nmask, K, counts, hits, counter);
    return cudaGetLastError();
}

cudaError_t nfix_apply(const uint32_t* pos, const uint32_t* counts, uint64_t n, uint64_t text_begin, const uint64_t* range_begin,
                       const uint64_t* range_end, uint32_t n_ranges, void* out, uint32_t value_bits, cudaStream_t stream)
{
    if (n == 0 || n_ranges == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (value_bits == 16)
        k_nfix_apply<<<grid, 256, 0, stream>>>(pos, counts, n, text_begin, range_begin, range_end, n_ranges, static_cast<uint16_t*>(out), 65535u);
    else
        k_nfix_apply<<<grid, 256, 0, stream>>>(pos, counts, n, text_begin, range_begin, range_end, n_ranges, static_cast<uint8_t*>(out), 255u);
    return cudaGetLastError();
}

} // namespace gmb
