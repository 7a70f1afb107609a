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

// ---- the N pass of Dna5 calls whose searches skip the text's N (MapCtx::skip_n; capi.cu: NFix) ----
// window starts t of the concatenated text whose K characters lie inside one sequence and hold 1..E N -> out_pos[*counter++]
// (any order; entries beyond cap are counted, not written)
cudaError_t nfix_collect_windows(const uint64_t* nmask, uint64_t n_text, const uint32_t* seq_start, uint32_t n_seq, uint32_t K, uint32_t E,
                                 uint32_t* out_pos, unsigned long long* counter, uint64_t cap, cudaStream_t stream);
// from the sorted located lists of m such windows (offsets: 2m + 1): counts[i] = occurrences of window i on both strands;
// hits[*counter++] = concatenated-text position of every occurrence whose own window holds no N
cudaError_t nfix_collect_hits(const uint32_t* rows, const uint64_t* offsets, uint64_t m, uint64_t n_rows, const uint32_t* seq_start,
                              uint32_t n_seq, const uint64_t* nmask, uint32_t K, uint32_t* counts, uint32_t* hits,
                              unsigned long long* counter, cudaStream_t stream);
// counts != nullptr: out[pos[i] - text_begin] = min(counts[i], max value); counts == nullptr: saturating += 1 — for the
// positions inside one of the call's work ranges
cudaError_t nfix_apply(const uint32_t* pos, const uint32_t* counts, uint64_t n, uint64_t text_begin, const uint64_t* range_begin,
                       const uint64_t* range_end, uint32_t n_ranges, void* out, uint32_t value_bits, cudaStream_t stream);

} // namespace gmb
