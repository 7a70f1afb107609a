// locate.cuh — post-processing steps of the locate path (locate_kernel.cu); temp == nullptr queries temp_bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gmb {

// offsets[0..n_lists] = exclusive prefix sums of counts[0..n_lists] (counts[n_lists] must be 0)
cudaError_t locate_scan_counts(const uint32_t* counts, uint64_t n_lists, uint64_t* offsets, void* temp, size_t& temp_bytes,
                               cudaStream_t stream);
// sort every list rows[offsets[l] .. offsets[l+1]) ascending
cudaError_t locate_sort_lists(const uint32_t* rows_in, uint32_t* rows_out, uint64_t n_rows, const uint64_t* offsets,
                              uint64_t n_lists, void* temp, size_t& temp_bytes, cudaStream_t stream);
// rows (positions inside T) -> gmb_location {sequence, offset}
cudaError_t locate_convert(const uint32_t* rows, uint64_t n_rows, const uint32_t* seq_start, uint32_t n_seq, void* out,
                           cudaStream_t stream);

// --exclude-pseudo beyond 64 files: out[pos0 + j] = number of distinct FASTA files among the occurrences of both
// strands of position j (lists sorted by text position; file ids must not decrease along the sequences)
cudaError_t locate_distinct_files(const uint32_t* rows, const uint64_t* offsets, uint64_t n_pos, const uint32_t* seq_start,
                                  uint32_t n_seq, const uint32_t* seq_to_file, void* out, uint32_t value_bits, uint64_t pos0,
                                  cudaStream_t stream);

} // namespace gmb
