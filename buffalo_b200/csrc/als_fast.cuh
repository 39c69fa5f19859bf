// Tuned ALS kernels (team-per-row, TMA-staged gathers).  Placeholder until the first GPU
// validation of the generic path; see DESIGN.md.
#pragma once
#include "als_generic.cuh"
#include "bfl_common.cuh"

namespace bfl {
inline bool fast_als_applicable(int /*optimizer_code*/, int /*d*/, int /*vdim*/, int /*block_size*/) { return false; }
inline int fast_als_launch(const AlsArgs&, int, int, DevBuf<int32_t>&, DevBuf<int32_t>&, cudaStream_t) {
    BFL_FAIL(BFL_ERR_STATE, "tuned ALS kernels not built");
}
}  // namespace bfl
