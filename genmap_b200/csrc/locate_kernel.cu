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
