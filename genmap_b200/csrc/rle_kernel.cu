// rle_kernel.cu — run-length encoding of the frequency vector on the device.
//
// The track writers of the reference (saveWig / saveBedGraph, src/output.hpp:73-187) scan the whole vector c on
// the host for maximal runs of equal values inside every sequence.  At 3 Gbp that is a 6 GB device-to-host copy
// followed by a serial scan; here the runs are found where c already lives and only (start, value) pairs leave
// the GPU.  A run starts at position i iff i is the first position of the range, the first position of a
// sequence, or c[i] != c[i-1].
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include <cub/device/device_reduce.cuh>

#include "rle.cuh"

namespace gmb {

namespace {

template <typename T>
struct RunHead {
    const T* c;               // biased so that c[i] is file-local position i
    const uint64_t* cum;      // n_chrom + 1 cumulative sequence lengths (device)
    uint32_t n_chrom;
    uint64_t begin;
    __device__ bool operator()(uint64_t i) const
    {
        if (i == begin || c[i] != c[i - 1]) return true;
        uint32_t lo = 0, hi = n_chrom; // is i the start of a sequence?  largest s with cum[s] <= i
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (cum[mid] <= i) lo = mid; else hi = mid;
        }
        return cum[lo] == i;
    }
};

template <typename T>
struct HeadCount {
    RunHead<T> h;
    __device__ unsigned long long operator()(uint64_t i) const { return h(i) ? 1ull : 0ull; }
};

template <typename T>
__global__ void k_gather_values(const T* __restrict__ c, const uint64_t* __restrict__ start, uint64_t n, uint16_t* __restrict__ value)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) value[r] = (uint16_t)c[start[r]];
}

template <typename T>
cudaError_t count_t(const void* c, const uint64_t* cum, uint32_t n_chrom, uint64_t begin, uint64_t end, unsigned long long* d_count,
                    void* temp, size_t& temp_bytes, cudaStream_t stream)
{
    HeadCount<T> f{RunHead<T>{static_cast<const T*>(c), cum, n_chrom, begin}};
    cub::CountingInputIterator<uint64_t> idx(begin);
    cub::TransformInputIterator<unsigned long long, HeadCount<T>, cub::CountingInputIterator<uint64_t>> in(idx, f);
    return cub::DeviceReduce::Sum(temp, temp_bytes, in, d_count, (int64_t)(end - begin), stream);
}

template <typename T>
cudaError_t select_t(const void* c, const uint64_t* cum, uint32_t n_chrom, uint64_t begin, uint64_t end, uint64_t* d_start,
                     unsigned long long* d_count, void* temp, size_t& temp_bytes, cudaStream_t stream)
{
    RunHead<T> f{static_cast<const T*>(c), cum, n_chrom, begin};
    cub::CountingInputIterator<uint64_t> idx(begin);
    return cub::DeviceSelect::If(temp, temp_bytes, idx, d_start, d_count, (int64_t)(end - begin), f, stream);
}

} // namespace

cudaError_t rle_count(const void* c, uint32_t value_bits, const uint64_t* cum, uint32_t n_chrom, uint64_t begin, uint64_t end,
                      unsigned long long* d_count, void* temp, size_t& temp_bytes, cudaStream_t stream)
{
    return value_bits == 16 ? count_t<uint16_t>(c, cum, n_chrom, begin, end, d_count, temp, temp_bytes, stream)
                            : count_t<uint8_t>(c, cum, n_chrom, begin, end, d_count, temp, temp_bytes, stream);
}

cudaError_t rle_select(const void* c, uint32_t value_bits, const uint64_t* cum, uint32_t n_chrom, uint64_t begin, uint64_t end,
                       uint64_t* d_start, unsigned long long* d_count, void* temp, size_t& temp_bytes, cudaStream_t stream)
{
    return value_bits == 16 ? select_t<uint16_t>(c, cum, n_chrom, begin, end, d_start, d_count, temp, temp_bytes, stream)
                            : select_t<uint8_t>(c, cum, n_chrom, begin, end, d_start, d_count, temp, temp_bytes, stream);
}

cudaError_t rle_gather(const void* c, uint32_t value_bits, const uint64_t* d_start, uint64_t n_runs, uint16_t* d_value, cudaStream_t stream)
{
    if (n_runs == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_runs + 255) / 256);
    if (value_bits == 16) k_gather_values<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(c), d_start, n_runs, d_value);
    else k_gather_values<<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(c), d_start, n_runs, d_value);
    return cudaGetLastError();
}

} // namespace gmb
