// map_kernel.cuh — launcher interface of the (K,E)-frequency kernel (map_kernel.cu).
#pragma once
#include <cuda_runtime.h>

#include "gmb_core.h"
#include "gmb_host.h"

namespace gmb {

struct MapLaunch {
    MapCtx cx;                 // device pointers into the index blob; cx.steps / cx.starts: device copies of the
                               // search tables (n_step_words words, then (B+1)*kMaxSearches SearchStart entries)
    uint32_t E;
    uint32_t n_step_words;
    uint32_t p1_off[kMaxBlockKmers + 1], fl_off[kMaxBlockKmers + 1];
    uint32_t chunk;            // positions per work chunk: a multiple of cx.B
    uint32_t sigma;            // 4 or 5: alphabet of the index (selects the rank-block layout)
    const uint64_t* text;      // 2-bit packed concatenated text (device)
    const uint64_t* nmask;     // sigma == 5: N mask of the text, one bit per position
    uint64_t text_begin;       // start of this FASTA file's text inside the concatenated text
    // work = chunks of <= `chunk` consecutive positions; a chunk never straddles two ranges
    const uint64_t* range_begin;  // device: n_ranges work ranges, file-local [begin, end)
    const uint64_t* range_end;
    const uint64_t* chunk_prefix; // device: n_ranges+1, number of chunks before range r
    uint32_t n_ranges;
    uint64_t n_chunks;
    uint64_t n_work;           // total k-mer starts to search (sizing only)
    unsigned long long* work_counter;  // device, zeroed before launch: next chunk id
    unsigned long long* fetch_counter; // device, 14 words (count_fetches): [0] rank-block fetches, [1] jump-table reads,
                                       // [2..9] fetches by interval size, [10] thin paths, [11] state-machine iterations
    void* out;                 // device, value_bits/8 bytes per file-local position
    uint32_t value_bits;
    bool count_fetches;
    bool exclude_pseudo;       // cx.sa / seq_start / seq_to_file / all_files are set
    // locate instantiation (launch_locate_kernel): `out` holds two uint32 list lengths per position of
    // [loc_pos0, ...) in the counting pass; the fill pass (cx.loc_rows != nullptr) reads the list starts from loc_off
    const uint64_t* loc_off;
    uint64_t loc_pos0;
    // locate instantiation, optional: "position" j stands for the text position loc_list[j] (the N pass of capi.cu
    // locates a list of scattered windows; the work ranges then cover list indices)
    const uint32_t* loc_list;
    // E = 0 on a Dna4 index entered through 16-byte table entries: the straight-line kernel of exact_kernel.cu
    // (nullptr: not applicable / switched off, the general kernel runs)
    const JtFull* e0_table;
    uint32_t e0_depth;
    // E >= 1 with every search entered through 16-byte entries (Dna4; Dna5 with cx.skip_n): the two-phase kernel of block_kernel.cu.
    // keylist: for every block size cnt the flat list of table keys of one strand, key_n[cnt] entries from key_off[cnt]:
    // x = XOR mask of the substituted characters on the key window, y = search | errors << 4 | (nothing substituted) << 8
    // (nullptr: not applicable / switched off).  `chunk` is then 32 * B: one block per lane.
    const uint2* keylist;
    uint32_t key_off[kMaxBlockKmers + 1], key_n[kMaxBlockKmers + 1];
    bool force_block_kernel;   // GMB_BLOCK_KERNEL=2: also for E >= 3 (measurements)
};

constexpr unsigned kChunk = 128; // positions handed out per global atomic (rounded down to a multiple of B)

// dynamic shared memory the kernel needs for these tables (the host shrinks B if this exceeds the SM's limit)
// (`blocked`: the instantiation that keeps per-window counters in the frame store: B > 1 or a Dna5 index)
size_t map_kernel_smem_bytes(uint32_t n_step_words, uint32_t E, uint32_t B, bool ep, uint32_t sigma, bool blocked);

// Enqueue the kernel on `stream`.  Returns cudaSuccess or the launch error.
cudaError_t launch_map_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream);

// E = 0 (exact_kernel.cu): does the launch qualify, and the launcher launch_map_kernel forwards to when it does
bool exact_kernel_applies(const MapLaunch& L);
cudaError_t launch_exact_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream);

// E >= 1 (block_kernel.cu): the same
bool block_kernel_applies(const MapLaunch& L);
size_t block_kernel_smem_bytes(uint32_t n_step_words, uint32_t E, uint32_t B, bool ep, uint32_t sigma = 4);
cudaError_t launch_block_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream);

// Locate variant (locate_kernel.cu): one k-mer per chain, every occurrence reported (csv output).
cudaError_t launch_locate_kernel(const MapLaunch& L, int sm_count, cudaStream_t stream);

} // namespace gmb
