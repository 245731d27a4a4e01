// pfn_tc.cuh - interface between pfn.cu (C ABI entry points, parameter packing) and pfn_tc.cu (tensor-core PFN kernel)
#pragma once
#include "common.cuh"

namespace pcp {

constexpr int kHidden = 32;   // layer-0 width of the two-layer PFN (NUM_FILTERS[0] / 2)
constexpr int kCout = 64;     // NUM_FILTERS[-1]
constexpr int kMaxCin = 24;   // input features, rounded up to a multiple of 8 (K of the first MMA)

__host__ __device__ inline int pfn_k0(int c_in) { return (c_in + 7) / 8 * 8; }
// layer-0 BN folded into the operand (scaled rows + a bias column fed with 1.0): two-layer PFN with a spare K column
__host__ __device__ inline bool pfn_fold0(int c_in, int num_layers) { return num_layers == 2 && pfn_k0(c_in) > c_in; }

// Packed parameter block (floats).  Operand panels are the K-major "interleaved" layout of umma.cuh:
// [K/4][rows][4], TF32 hi / lo (round-to-nearest split).  Rows of the LAST layer are multiplied by the sign of its
// folded BN scale so that BN + ReLU is non-decreasing in the accumulator and commutes with the per-pillar max.
//   two layers: w0h | w0l  [k0/4][32][4]      layer 0: Linear(c_in -> 32)
//               w1ah | w1al [8][64][4]        layer 1, columns 0..31  (the per-point half:  x . W1[:, :32]^T)
//               w1bh | w1bl [8][64][4]        layer 1, columns 32..63 (the per-pillar half: x_max . W1[:, 32:]^T)
//               w1sh | w1sl [8][64][4]        W1[:, :32] + W1[:, 32:]: one-point pillars (x_max == x) need a single product
//               a0[32] b0[32]                 folded BN of layer 0 (signed scale, shift); (1, 0) when pfn_fold0(): then w0 holds
//                                             alpha * W0 and column c_in holds beta (the producers feed it with 1.0)
//               a1[64] b1[64]                 |scale|, shift of layer 1
//               w1b_f32 [64][32]              sign-folded fp32 copy of W1[:, 32:] for the long-pillar finishing kernel
//   one layer:  w0h | w0l  [k0/4][64][4] (sign-folded) | a0[64] (|scale|) b0[64]
struct ParamLayout {
  int w0h, w0l, w1ah, w1al, w1bh, w1bl, w1sh, w1sl, a0, b0, a1, b1, w1b_f32, total;
};
__host__ __device__ inline ParamLayout param_layout(int c_in, int num_layers) {
  ParamLayout P{};
  const int k0 = pfn_k0(c_in);
  const int n0 = num_layers == 2 ? kHidden : kCout;
  int o = 0;
  P.w0h = o; o += k0 * n0;
  P.w0l = o; o += k0 * n0;
  if (num_layers == 2) {
    P.w1ah = o; o += kHidden * kCout;
    P.w1al = o; o += kHidden * kCout;
    P.w1bh = o; o += kHidden * kCout;
    P.w1bl = o; o += kHidden * kCout;
    P.w1sh = o; o += kHidden * kCout;
    P.w1sl = o; o += kHidden * kCout;
  }
  P.a0 = o; o += n0;
  P.b0 = o; o += n0;
  if (num_layers == 2) {
    P.a1 = o; o += kCout;
    P.b1 = o; o += kCout;
    P.w1b_f32 = o; o += kCout * kHidden;
  }
  P.total = o;
  return P;
}

struct TcArgs {
  const float* points;
  int64_t stride;
  pcp_grid g;
  int c_in, c_raw, n_raw, raw_col0, with_distance, k0, num_layers;
  const float* params;
  const int32_t* hdr;
  const int32_t* seg_off;
  const int32_t* sorted_idx;
  const unsigned long long* lists;
  ListOffsets lo;
  const float4* mean;
  const float4* long_mean;
  unsigned* long_acc;
  const int4* long_table;
  float* out;
  float* mean_out;
};

int launch_pfn_tc(const TcArgs& a, int64_t n_points, cudaStream_t stream);

}  // namespace pcp
