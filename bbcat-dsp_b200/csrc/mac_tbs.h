// mac_tbs.h -- host entry of the shared-stream time-batched FDL MAC (mac_tbs.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mac_common.cuh"

namespace bbx {

struct MacTbsArgs {
  const MacSeg* segs;             // device: the plan's segments
  const uint32_t* cta_seg_begin;  // device: n_plan_ctas + 1 segment offsets
  uint32_t n_plan_ctas;
  const float2* fdl;              // [n_in][R][B]
  float2* ypart;                  // partial sums of block-step t0: [nt][slot_stride][B]
  float* nyq_part;                // Nyquist partial sums of block-step t0: [nt][slot_stride]
  uint32_t B, R, head, t0, nt, slot_stride;
  int* status;                    // device-visible word set non-zero when a barrier wait timed out (may be NULL)
};

// the Nyquist sums of column 0 (k_nyq_mac2, a few microseconds) and the MAC itself (k_fdl_mac_tbs<NTILE>) on `st`;
// *kernel_name = the MAC instantiation that ran
cudaError_t launch_nyq_mac2(const MacTbsArgs& a, cudaStream_t st);
cudaError_t launch_mac_tbs(const MacTbsArgs& a, cudaStream_t st, const char** kernel_name);

}  // namespace bbx
