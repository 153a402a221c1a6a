// kernels_common.cuh -- constants shared by the engine's kernels and its host side.
#pragma once

#include "common.cuh"

namespace bbx {


static constexpr uint32_t kNoJob = 0xFFFFFFFFu;
static constexpr uint32_t kSameJob = 0xFFFFFFFEu;  // crossfade a stream with itself (delay-only switch)
static constexpr int kNumSMs = 148;
__host__ __device__ __forceinline__ uint32_t ceil_div_dev(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace bbx
