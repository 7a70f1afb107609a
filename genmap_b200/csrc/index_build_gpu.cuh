// index_build_gpu.cuh — device-side index construction (suffix sorting by prefix doubling).
#pragma once
#include <cstdint>
#include <string>

#include "gmb_layout.h"

namespace gmb {

struct GpuBuildTimings {
    double h2d_ms = 0, sort_ms = 0, pack_ms = 0, total_ms = 0;
    uint32_t doubling_rounds[2] = {0, 0};
};

// Builds the index blob in a fresh cudaMalloc'ed buffer on `device` (caller frees with cudaFree).
// codes/limits are host pointers.  Returns 0 or a negative gmb_status with `err` set.
int build_index_gpu_device(const uint8_t* codes, const uint64_t* limits, uint32_t n_seq, bool with_sa, int device,
                           uint8_t** d_blob_out, IndexHeader* header_out, GpuBuildTimings* timings, std::string& err);

// Same, then copies the blob to malloc'ed host memory.
int build_index_gpu(const uint8_t* codes, const uint64_t* limits, uint32_t n_seq, bool with_sa, int device,
                    void** blob_out, uint64_t* bytes_out, std::string& err);

} // namespace gmb
