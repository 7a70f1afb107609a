// rle.cuh — device run-length encoding of the frequency vector (rle_kernel.cu); temp == nullptr queries temp_bytes.
// `c` is biased so that c[i] is file-local position i; runs are found inside [begin, end) and break at every
// sequence start listed in cum (device, n_chrom + 1 entries).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gmb {

cudaError_t rle_count(const void* c, uint32_t value_bits, const uint64_t* cum, uint32_t n_chrom, uint64_t begin, uint64_t end,
                      unsigned long long* d_count, void* temp, size_t& temp_bytes, cudaStream_t stream);
cudaError_t rle_select(const void* c, uint32_t value_bits, const uint64_t* cum, uint32_t n_chrom, uint64_t begin, uint64_t end,
                       uint64_t* d_start, unsigned long long* d_count, void* temp, size_t& temp_bytes, cudaStream_t stream);
cudaError_t rle_gather(const void* c, uint32_t value_bits, const uint64_t* d_start, uint64_t n_runs, uint16_t* d_value, cudaStream_t stream);

} // namespace gmb
