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

} // namespace gmb
