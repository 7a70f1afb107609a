// jump_table.cuh — device builder of the per-depth jump tables (see SearchStart in gmb_core.h).
#pragma once
#include <cuda_runtime.h>

#include "gmb_core.h"

namespace gmb {

// Level d from level d-1 (d == 1: from the root): one thread per d-mer extends its (d-1)-mer parent by
// one character to the right on the bidirectional index.  out_uni / out_lof have 4^d entries; with out_full set
// the level is written as 16-byte entries holding both intervals instead (out_uni / out_lof unused).
cudaError_t build_jump_level(const MapCtx& cx, uint32_t sigma, uint32_t d, const JtEntry* prev_uni, const uint32_t* prev_lof,
                             JtEntry* out_uni, uint32_t* out_lof, JtFull* out_full, cudaStream_t stream);

// Rewrite the entries of a finished level of 16-byte entries whose key occurs exactly once as LOCATED entries
// (text position + context characters, gmb_core.h: JtFull); one pass over the packed text.  nmask: the N mask of a
// Dna5 index (nullptr for Dna4): keys whose window or context holds an N keep their intervals.
cudaError_t locate_jump_singletons(const uint64_t* text, const uint64_t* nmask, uint64_t n_text, const uint32_t* seq_start, uint32_t n_seq,
                                   uint32_t d, JtFull* full, cudaStream_t stream);

} // namespace gmb
