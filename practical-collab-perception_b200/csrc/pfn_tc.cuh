// pfn_tc.cuh - interface between pfn.cu (C ABI entry points) and pfn_tc.cu (tensor-core PFN kernel)
#pragma once
#include "common.cuh"

namespace pcp {

struct TcArgs {
  const float* points;
  int64_t stride;
  pcp_grid g;
  int c_in, n_raw, raw_col0, with_distance, k0;   // k0 = c_in rounded up to a multiple of 8
  const float* params_simt;   // a0 | b0 | a1 | b1 live in the SIMT block
  const float* params_tc;     // w0h | w0l | w1h | w1l panels
  int a0_off, a1_off;         // float offsets of a0 and a1 inside params_simt
  const int32_t* hdr;
  const int32_t* seg_off;
  const int32_t* sorted_idx;
  const int32_t* tile_first;
  float* out;
  float* mean_out;
};

int launch_pfn_tc(const TcArgs& a, int64_t n_points, cudaStream_t stream);
// alpha1: the folded BN scale of layer 1 (device pointer into the SIMT parameter block, already written on `stream`)
int launch_pack_tc(int c_in, int k0, const float* w0, const float* w1, const float* alpha1, float* out,
                   cudaStream_t stream);

}  // namespace pcp
