// pfn.cu - pillar feature network, fused (reference: dynamic_pillar_vfe.py:110-129 + PFNLayerV2 :35-46).
//
// One CTA owns every pillar whose first sorted point lies in its 128-point window, so a pillar is
// never split between CTAs: the per-pillar mean, the layer-0 max, the hoisted half of layer 1
// (W1[:, 32:] . x_max is computed once per pillar instead of once per point) and the final max all stay
// in shared memory; HBM sees one gather of the point rows and one coalesced store of pillar_features.
// Pillars with more points than a window are streamed through the same stages in 128-point chunks
// (three passes: sums, layer-0 max, layer 1), so any point distribution is handled by one launch.
#include "common.cuh"
#include "pfn_tc.cuh"

namespace pcp {

constexpr int kT = 128;        // points per chunk == max pillars per CTA
constexpr int kLd = kT + 1;    // smem row stride (conflict-free column access)
constexpr int kThreads = 256;
constexpr int kMaxCin = 24;
constexpr int kHidden = 32;
constexpr int kCout = 64;

struct PfnSmem {
  int off[kT + 2];
  int lp[kT];
  float xyz[3][kLd];
  float mean[3][kLd];
  float feat[kMaxCin][kLd];
  float x0[kHidden][kLd];
  float max0[kHidden][kLd];
  float h[kCout][kLd];
  float y[kCout][kLd];
  alignas(16) float w0[kMaxCin * kCout];
  alignas(16) float a0[kCout];
  alignas(16) float b0[kCout];
  alignas(16) float w1a[kHidden * kCout];
  alignas(16) float w1b[kHidden * kCout];
  alignas(16) float a1[kCout];
  alignas(16) float b1[kCout];
};

struct PfnArgs {
  const float* points;
  int64_t stride;
  pcp_grid g;
  int c_in, n_raw, raw_col0, with_distance;
  const float* params;
  const int32_t* hdr;
  const int32_t* seg_off;
  const int32_t* sorted_idx;
  const int32_t* tile_first;
  const int32_t* long_list;
  float* out;
  float* mean_out;
};

// C[m][n] (+)= sum_k A[k][m] * B[k][n];  A: smem [K][kLd], B: smem [K][N]; thread tile 4 x TN
template <int TN>
__device__ __forceinline__ void tile_gemm(const float (*A)[kLd], const float* B, int K, int N, int m0, int n0,
                                          float (&acc)[4][TN]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[4], b[TN];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = A[k][m0 + i];
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(B + k * N + n0 + j);
      b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// packed parameter block (floats):  w0t[c_in][H0] | a0[H0] | b0[H0] | w1a_t[32][64] | w1b_t[32][64] | a1[64] | b1[64]
__host__ __device__ inline int pfn_h0(int num_layers) { return num_layers == 2 ? kHidden : kCout; }
__host__ __device__ inline size_t pfn_param_floats_simt(int c_in, int num_layers) {
  const int h0 = pfn_h0(num_layers);
  size_t n = (size_t)c_in * h0 + 2 * h0;
  if (num_layers == 2) n += 2 * (size_t)kHidden * kCout + 2 * kCout;
  return (n + 3) / 4 * 4;   // keeps the tensor-core section 16-byte aligned
}
__host__ __device__ inline int pfn_k0(int c_in) { return (c_in + 7) / 8 * 8; }
// tensor-core section (two layers only): w0h | w0l ([k0/4][32][4]) | w1h | w1l ([16][64][4], sign-folded) | |alpha1|[64]
__host__ __device__ inline size_t pfn_param_floats(int c_in, int num_layers) {
  size_t n = pfn_param_floats_simt(c_in, num_layers);
  if (num_layers == 2) n += 2 * (size_t)pfn_k0(c_in) * 32 + 2 * 64 * 64 + 64;   // + |alpha1|
  return n;
}

// all PFN stages for the pillars [pa, pb) (at most kT pillars), streamed in chunks of kT rows
template <int kLayers>
__device__ __forceinline__ void pfn_process_range(PfnSmem& S, const PfnArgs& A, const int pa, const int pb) {
  constexpr int H0 = (kLayers == 2) ? kHidden : kCout;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npil = pb - pa;
  for (int i = tid; i <= npil; i += kThreads) S.off[i] = A.seg_off[pa + i];
  for (int i = tid; i < 3 * kLd; i += kThreads) (&S.mean[0][0])[i] = 0.f;
  __syncthreads();
  const int q_begin = S.off[0], q_end = S.off[npil];
  const int n_chunks = (q_end - q_begin + kT - 1) / kT;

  // GEMM thread mapping: 4 points x TN channels per thread
  const int ty = tid >> 3, tx = tid & 7;
  const int m0 = ty * 4;

  // ---- stage: gather the rows of sorted positions [cb, ce) ----
  auto stage_load = [&](int cb, int ce) {
    if (tid < kT) {
      const int pos = cb + tid;
      int lp = -1;
      float x = 0.f, y = 0.f, z = 0.f;
      if (pos < ce) {
        int lo = 0, hi = npil;  // last lp with off[lp] <= pos
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (S.off[mid] <= pos) lo = mid; else hi = mid;
        }
        lp = lo;
        const float* row = A.points + (int64_t)A.sorted_idx[pos] * A.stride;
        x = __ldg(row + 1); y = __ldg(row + 2); z = __ldg(row + 3);
        for (int f = 0; f < A.n_raw; ++f) S.feat[f][tid] = __ldg(row + A.raw_col0 + f);
      } else {
        for (int f = 0; f < A.c_in; ++f) S.feat[f][tid] = 0.f;
      }
      S.lp[tid] = lp;
      S.xyz[0][tid] = x; S.xyz[1][tid] = y; S.xyz[2][tid] = z;
    }
    __syncthreads();
  };
  // ---- stage: per-pillar running sums, sequential in sorted (ascending row) order ----
  auto stage_sums = [&](int cb, int ce) {
    if (tid < npil) {
      const int qs = max(S.off[tid], cb) - cb, qe = min(S.off[tid + 1], ce) - cb;
      float sx = S.mean[0][tid], sy = S.mean[1][tid], sz = S.mean[2][tid];
      for (int q = qs; q < qe; ++q) {
        sx = __fadd_rn(sx, S.xyz[0][q]); sy = __fadd_rn(sy, S.xyz[1][q]); sz = __fadd_rn(sz, S.xyz[2][q]);
      }
      S.mean[0][tid] = sx; S.mean[1][tid] = sy; S.mean[2][tid] = sz;
    }
    __syncthreads();
  };
  // ---- stage: scatter_mean = sum / clamp(count, 1), true division (dynamic_pillar_vfe.py:110) ----
  auto stage_means = [&]() {
    if (tid < npil) {
      const float cnt = (float)max(S.off[tid + 1] - S.off[tid], 1);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float m = __fdiv_rn(S.mean[d][tid], cnt);
        S.mean[d][tid] = m;
        if (A.mean_out) A.mean_out[(int64_t)(pa + tid) * 3 + d] = m;
      }
    }
    __syncthreads();
  };
  // ---- stage: f_cluster, f_center (:111-116), optional distance (:124) ----
  auto stage_features = [&]() {
    if (tid < kT && S.lp[tid] >= 0) {
      const int lp = S.lp[tid];
      const float x = S.xyz[0][tid], y = S.xyz[1][tid], z = S.xyz[2][tid];
      int f = A.n_raw;
      S.feat[f++][tid] = __fsub_rn(x, S.mean[0][lp]);
      S.feat[f++][tid] = __fsub_rn(y, S.mean[1][lp]);
      S.feat[f++][tid] = __fsub_rn(z, S.mean[2][lp]);
      const float cx = quantise(x, A.g.range_min_x, A.g.voxel_x);
      const float cy = quantise(y, A.g.range_min_y, A.g.voxel_y);
      S.feat[f++][tid] = __fsub_rn(x, __fadd_rn(__fmul_rn(cx, A.g.voxel_x), A.g.x_offset));
      S.feat[f++][tid] = __fsub_rn(y, __fadd_rn(__fmul_rn(cy, A.g.voxel_y), A.g.y_offset));
      S.feat[f++][tid] = __fsub_rn(z, A.g.z_offset);
      if (A.with_distance)
        S.feat[f++][tid] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    }
    __syncthreads();
  };
  // ---- stage: layer 0, Linear + BN(eval) + ReLU -> x0 (two layers) or y (single layer) ----
  auto stage_layer0 = [&]() {
    if (kLayers == 2) {
      const int n0 = tx * 4;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      tile_gemm<4>(S.feat, S.w0, A.c_in, H0, m0, n0, acc);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = S.a0[n0 + j], b = S.b0[n0 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i) S.x0[n0 + j][m0 + i] = fmaxf(fmaf(acc[i][j], a, b), 0.f);
      }
    } else {
      const int n0 = tx * 8;
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      tile_gemm<8>(S.feat, S.w0, A.c_in, H0, m0, n0, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = S.a0[n0 + j], b = S.b0[n0 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i) S.y[n0 + j][m0 + i] = fmaxf(fmaf(acc[i][j], a, b), 0.f);
      }
    }
    __syncthreads();
  };
  // ---- stage: layer-0 segment max (owner computes: warp -> pillar, lane -> channel) ----
  auto stage_max0 = [&](int cb, int ce) {
    for (int lp = warp; lp < npil; lp += kThreads / 32) {
      const int qs = max(S.off[lp], cb) - cb, qe = min(S.off[lp + 1], ce) - cb;
      if (qs >= qe) continue;
      float m = (S.off[lp] >= cb) ? 0.f : S.max0[lane][lp];
      for (int q = qs; q < qe; ++q) m = fmaxf(m, S.x0[lane][q]);
      S.max0[lane][lp] = m;
    }
    __syncthreads();
  };
  // ---- stage: hoisted half of layer 1, h[lp] = W1[:, 32:] . max0[lp] (once per pillar) ----
  auto stage_hoist = [&]() {
    const int n0 = tx * 8;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    tile_gemm<8>(S.max0, S.w1b, kHidden, kCout, m0, n0, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) S.h[n0 + j][m0 + i] = acc[i][j];
    __syncthreads();
  };
  // ---- stage: layer 1, x . W1[:, :32]^T + h[pillar], BN(eval), ReLU -> y ----
  auto stage_layer1 = [&]() {
    const int n0 = tx * 8;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lp = S.lp[m0 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = (lp >= 0) ? S.h[n0 + j][lp] : 0.f;
    }
    tile_gemm<8>(S.x0, S.w1a, kHidden, kCout, m0, n0, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = S.a1[n0 + j], b = S.b1[n0 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) S.y[n0 + j][m0 + i] = fmaxf(fmaf(acc[i][j], a, b), 0.f);
    }
    __syncthreads();
  };
  // ---- stage: final segment max -> pillar_features (coalesced 256-byte rows) ----
  auto stage_out = [&](int cb, int ce) {
    for (int lp = warp; lp < npil; lp += kThreads / 32) {
      const int qs = max(S.off[lp], cb) - cb, qe = min(S.off[lp + 1], ce) - cb;
      if (qs >= qe) continue;
      float* dst = A.out + (int64_t)(pa + lp) * kCout;
      float v0 = 0.f, v1 = 0.f;
      if (S.off[lp] < cb) { v0 = dst[lane]; v1 = dst[lane + 32]; }  // running max of a multi-chunk pillar
      for (int q = qs; q < qe; ++q) { v0 = fmaxf(v0, S.y[lane][q]); v1 = fmaxf(v1, S.y[lane + 32][q]); }
      dst[lane] = v0; dst[lane + 32] = v1;
    }
    __syncthreads();
  };

  if (n_chunks == 1) {
    // common case: the whole tile fits one chunk, every stage runs once out of shared memory
    stage_load(q_begin, q_end);
    stage_sums(q_begin, q_end);
    stage_means();
    stage_features();
    stage_layer0();
    if (kLayers == 2) {
      stage_max0(q_begin, q_end);
      stage_hoist();
      stage_layer1();
    }
    stage_out(q_begin, q_end);
  } else {
    // long pillars: stream the chunks three times (sums | layer-0 max | layer 1), rows re-gathered from L2
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int cb = q_begin + ch * kT, ce = min(cb + kT, q_end);
      stage_load(cb, ce);
      stage_sums(cb, ce);
    }
    stage_means();
    if (kLayers == 2) {
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int cb = q_begin + ch * kT, ce = min(cb + kT, q_end);
        stage_load(cb, ce);
        stage_features();
        stage_layer0();
        stage_max0(cb, ce);
      }
      stage_hoist();
    }
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int cb = q_begin + ch * kT, ce = min(cb + kT, q_end);
      stage_load(cb, ce);
      stage_features();
      stage_layer0();
      if (kLayers == 2) stage_layer1();
      stage_out(cb, ce);
    }
  }
}

// kFromList = false: one group (window of kWin sorted positions, tile_first[]) per loop iteration
// kFromList = true : one long pillar (more than kLongSeg points, long_list[]) per loop iteration
template <int kLayers, bool kFromList>
__global__ void __launch_bounds__(kThreads, 1)
pfn_kernel(const PfnArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PfnSmem& S = *reinterpret_cast<PfnSmem*>(smem_raw);
  constexpr int H0 = (kLayers == 2) ? kHidden : kCout;
  const int tid = threadIdx.x;
  const int n_items = kFromList ? A.hdr[kHdrLongCount] : (A.hdr[PCP_COUNT_KEPT] + kWin - 1) / kWin;
  if ((int)blockIdx.x >= n_items) return;
  {
    const float* p = A.params;
    for (int i = tid; i < A.c_in * H0; i += kThreads) S.w0[i] = p[i];
    p += A.c_in * H0;
    for (int i = tid; i < H0; i += kThreads) { S.a0[i] = p[i]; S.b0[i] = p[H0 + i]; }
    p += 2 * H0;
    if (kLayers == 2) {
      for (int i = tid; i < kHidden * kCout; i += kThreads) { S.w1a[i] = p[i]; S.w1b[i] = p[kHidden * kCout + i]; }
      p += 2 * kHidden * kCout;
      for (int i = tid; i < kCout; i += kThreads) { S.a1[i] = p[i]; S.b1[i] = p[kCout + i]; }
    }
  }
  __syncthreads();
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    int pa, pb;
    if (kFromList) { pa = A.long_list[it]; pb = pa + 1; }
    else { pa = A.tile_first[it]; pb = A.tile_first[it + 1]; }
    if (pb > pa) pfn_process_range<kLayers>(S, A, pa, pb);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// parameter packing: fold BN(eval) into scale / shift, transpose the weights
// ------------------------------------------------------------------------------------------------
__global__ void pack_params_kernel(int c_in, int num_layers, const float* w0, const float* lb0, const float* g0,
                                   const float* be0, const float* mu0, const float* var0, const float* w1,
                                   const float* lb1, const float* g1, const float* be1, const float* mu1,
                                   const float* var1, float eps, float* out) {
  const int h0 = pfn_h0(num_layers);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  float* w0t = out;
  float* a0 = w0t + c_in * h0;
  float* b0 = a0 + h0;
  for (int i = tid; i < c_in * h0; i += nth) {
    const int k = i / h0, c = i % h0;
    w0t[i] = w0[c * c_in + k];
  }
  for (int c = tid; c < h0; c += nth) {
    if (g0) {
      // ATen eval batch_norm: alpha = weight * rsqrt(var + eps), beta = bias - mean * alpha
      const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var0[c], eps)));
      const float al = __fmul_rn(g0[c], inv);
      a0[c] = al; b0[c] = __fsub_rn(be0[c], __fmul_rn(mu0[c], al));
    } else {
      a0[c] = 1.f; b0[c] = lb0 ? lb0[c] : 0.f;
    }
  }
  if (num_layers == 2) {
    float* w1a = b0 + h0;
    float* w1b = w1a + kHidden * kCout;
    float* a1 = w1b + kHidden * kCout;
    float* b1 = a1 + kCout;
    for (int i = tid; i < kHidden * kCout; i += nth) {
      const int k = i / kCout, c = i % kCout;
      w1a[i] = w1[c * (2 * kHidden) + k];
      w1b[i] = w1[c * (2 * kHidden) + kHidden + k];
    }
    for (int c = tid; c < kCout; c += nth) {
      if (g1) {
        const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var1[c], eps)));
        const float al = __fmul_rn(g1[c], inv);
        a1[c] = al; b1[c] = __fsub_rn(be1[c], __fmul_rn(mu1[c], al));
      } else {
        a1[c] = 1.f; b1[c] = lb1 ? lb1[c] : 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// standalone segmented mean / max over the pillars (torch_scatter.scatter_mean / scatter_max)
// one warp per pillar, lanes over channels; rows are visited in ascending row order
// ------------------------------------------------------------------------------------------------
template <int kMode>
__global__ void __launch_bounds__(256)
segment_reduce_kernel(const float* __restrict__ values, int64_t vstride, int channels,
                      const int32_t* __restrict__ hdr, const int32_t* __restrict__ seg_off,
                      const int32_t* __restrict__ sorted_idx, float* __restrict__ out) {
  const int P = hdr[PCP_COUNT_PILLARS];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < P; r += gridDim.x * wpb) {
    const int off = seg_off[r], n = seg_off[r + 1] - off;
    for (int c0 = 0; c0 < channels; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      bool first = true;
      for (int j0 = 0; j0 < n; j0 += 32) {
        const int my = (j0 + lane < n) ? sorted_idx[off + j0 + lane] : 0;
        const int cnt = min(32, n - j0);
        for (int j = 0; j < cnt; ++j) {
          const int idx = __shfl_sync(0xffffffffu, my, j);
          if (c < channels) {
            const float v = __ldg(values + (int64_t)idx * vstride + c);
            if (kMode == 0) acc = __fadd_rn(acc, v);
            else { acc = first ? v : fmaxf(acc, v); }
            first = false;
          }
        }
      }
      if (c < channels) {
        if (kMode == 0) acc = __fdiv_rn(acc, (float)max(n, 1));
        out[(int64_t)r * channels + c] = acc;
      }
    }
  }
}

}  // namespace pcp

using namespace pcp;

extern "C" size_t pcp_pfn_param_floats(const pcp_pfn_desc* d) {
  if (!d) return 0;
  const int c_in = d->c_raw + (d->use_absolute_xyz ? 6 : 3) + (d->with_distance ? 1 : 0);
  return pfn_param_floats(c_in, d->num_layers);
}

static int check_desc(const pcp_pfn_desc* d, int* c_in_out) {
  PCP_REQUIRE(d, PCP_E_INVALID, "pfn: null desc");
  PCP_REQUIRE(d->num_layers == 1 || d->num_layers == 2, PCP_E_UNSUPPORTED,
              "pfn: num_layers %d unsupported (NUM_FILTERS must have 1 or 2 entries)", d->num_layers);
  PCP_REQUIRE(d->c_out == kCout, PCP_E_UNSUPPORTED, "pfn: c_out %d unsupported (NUM_FILTERS[-1] must be 64)", d->c_out);
  PCP_REQUIRE(d->num_layers == 1 || d->hidden == kHidden, PCP_E_UNSUPPORTED,
              "pfn: hidden %d unsupported (NUM_FILTERS[0] must be 64)", d->hidden);
  PCP_REQUIRE(d->c_raw >= 3, PCP_E_INVALID, "pfn: c_raw < 3");
  const int c_in = d->c_raw + (d->use_absolute_xyz ? 6 : 3) + (d->with_distance ? 1 : 0);
  PCP_REQUIRE(c_in <= kMaxCin, PCP_E_UNSUPPORTED, "pfn: %d input features > %d supported", c_in, kMaxCin);
  *c_in_out = c_in;
  return 0;
}

extern "C" int pcp_pack_pfn_params(const pcp_pfn_desc* desc, const float* w0, const float* lin_bias0,
                                   const float* bn0_weight, const float* bn0_bias, const float* bn0_mean,
                                   const float* bn0_var, const float* w1, const float* lin_bias1,
                                   const float* bn1_weight, const float* bn1_bias, const float* bn1_mean,
                                   const float* bn1_var, float eps, float* packed_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int c_in = 0;
  if (int rc = check_desc(desc, &c_in)) return rc;
  PCP_REQUIRE(w0 && packed_out, PCP_E_INVALID, "pcp_pack_pfn_params: null argument");
  PCP_REQUIRE(desc->num_layers == 1 || w1, PCP_E_INVALID, "pcp_pack_pfn_params: null w1");
  PCP_REQUIRE(!bn0_weight || (bn0_bias && bn0_mean && bn0_var), PCP_E_INVALID, "pcp_pack_pfn_params: partial bn0");
  PCP_REQUIRE(!bn1_weight || (bn1_bias && bn1_mean && bn1_var), PCP_E_INVALID, "pcp_pack_pfn_params: partial bn1");
  pack_params_kernel<<<8, 256, 0, stream>>>(c_in, desc->num_layers, w0, lin_bias0, bn0_weight, bn0_bias, bn0_mean,
                                            bn0_var, w1, lin_bias1, bn1_weight, bn1_bias, bn1_mean, bn1_var, eps,
                                            packed_out);
  PCP_LAUNCH_CHECK("pack_params_kernel");
  if (desc->num_layers == 2)
    return launch_pack_tc(c_in, pfn_k0(c_in), w0, w1,
                          packed_out + c_in * kHidden + 2 * kHidden + 2 * kHidden * kCout /* a1 */,
                          packed_out + pfn_param_floats_simt(c_in, 2), stream);
  return 0;
}

extern "C" int pcp_pfn(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                       const pcp_grid* grid, const pcp_pfn_desc* desc, const float* packed_params,
                       const void* workspace, size_t workspace_bytes, float* pillar_features_out,
                       float* pillar_mean_out, int64_t pillar_capacity, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int c_in = 0;
  if (int rc = check_desc(desc, &c_in)) return rc;
  PCP_REQUIRE(grid && packed_params && workspace && pillar_features_out, PCP_E_INVALID, "pcp_pfn: null argument");
  PCP_REQUIRE(n_points >= 0 && (n_points == 0 || points), PCP_E_INVALID, "pcp_pfn: bad points");
  PCP_REQUIRE(row_stride >= 1 + desc->c_raw, PCP_E_INVALID, "pcp_pfn: row_stride %lld < 1 + c_raw", (long long)row_stride);
  const WsLayout L = ws_layout(n_points, max_frames, grid->nx, grid->ny);
  PCP_REQUIRE(workspace_bytes >= L.total, PCP_E_WORKSPACE, "pcp_pfn: workspace %zu < %zu bytes", workspace_bytes, L.total);
  PCP_REQUIRE(pillar_capacity >= L.cap, PCP_E_INVALID, "pcp_pfn: pillar_capacity too small");
  if (n_points == 0) return 0;
  const WsView W = ws_view(const_cast<void*>(workspace), L);
  PfnArgs a;
  a.points = points; a.stride = row_stride; a.g = *grid; a.c_in = c_in;
  a.raw_col0 = desc->use_absolute_xyz ? 1 : 4;
  a.n_raw = desc->use_absolute_xyz ? desc->c_raw : desc->c_raw - 3;
  a.with_distance = desc->with_distance;
  a.params = packed_params; a.hdr = W.hdr; a.seg_off = W.seg_off; a.sorted_idx = W.sorted_idx;
  a.out = pillar_features_out; a.mean_out = pillar_mean_out;
  a.tile_first = W.tile_first; a.long_list = W.long_list;
  const size_t smem = sizeof(PfnSmem);
  if (desc->num_layers == 2) {
    // tensor-core kernel for every pillar of up to kLongSeg points ...
    TcArgs t;
    t.points = points; t.stride = row_stride; t.g = *grid; t.c_in = c_in; t.n_raw = a.n_raw; t.raw_col0 = a.raw_col0;
    t.with_distance = a.with_distance; t.k0 = pfn_k0(c_in);
    t.params_simt = packed_params; t.params_tc = packed_params + pfn_param_floats_simt(c_in, 2);
    t.a0_off = c_in * kHidden; t.a1_off = c_in * kHidden + 2 * kHidden + 2 * kHidden * kCout;
    t.hdr = W.hdr; t.seg_off = W.seg_off; t.sorted_idx = W.sorted_idx; t.tile_first = W.tile_first;
    t.out = pillar_features_out; t.mean_out = pillar_mean_out;
    if (int rc = launch_pfn_tc(t, n_points, stream)) return rc;
    // ... and the chunk-streaming kernel for the (few) longer ones
    PCP_CUDA(cudaFuncSetAttribute(pfn_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pfn_kernel<2, true><<<148, kThreads, smem, stream>>>(a);
  } else {
    const int64_t groups = (n_points + kWin - 1) / kWin;
    const unsigned blocks = (unsigned)(groups < 148 * 8 ? groups : 148 * 8);
    PCP_CUDA(cudaFuncSetAttribute(pfn_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pfn_kernel<1, false><<<blocks, kThreads, smem, stream>>>(a);
  }
  PCP_LAUNCH_CHECK("pfn_kernel");
  return 0;
}

extern "C" int pcp_segment_reduce(const float* values, int64_t value_stride, int32_t channels, int32_t mode,
                                  int64_t n_points, int32_t max_frames, int32_t nx, int32_t ny,
                                  const void* workspace, size_t workspace_bytes, float* out,
                                  int64_t pillar_capacity, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(workspace && out, PCP_E_INVALID, "pcp_segment_reduce: null argument");
  PCP_REQUIRE(n_points >= 0 && (n_points == 0 || values), PCP_E_INVALID, "pcp_segment_reduce: bad values");
  PCP_REQUIRE(channels > 0 && value_stride >= channels, PCP_E_INVALID, "pcp_segment_reduce: bad channels/stride");
  PCP_REQUIRE(mode == 0 || mode == 1, PCP_E_INVALID, "pcp_segment_reduce: mode must be 0 (mean) or 1 (max)");
  const WsLayout L = ws_layout(n_points, max_frames, nx, ny);
  PCP_REQUIRE(workspace_bytes >= L.total, PCP_E_WORKSPACE, "pcp_segment_reduce: workspace too small");
  PCP_REQUIRE(pillar_capacity >= L.cap, PCP_E_INVALID, "pcp_segment_reduce: pillar_capacity too small");
  if (n_points == 0) return 0;
  const WsView W = ws_view(const_cast<void*>(workspace), L);
  const unsigned blocks = 148 * 8;
  if (mode == 0)
    segment_reduce_kernel<0><<<blocks, 256, 0, stream>>>(values, value_stride, channels, W.hdr, W.seg_off, W.sorted_idx, out);
  else
    segment_reduce_kernel<1><<<blocks, 256, 0, stream>>>(values, value_stride, channels, W.hdr, W.seg_off, W.sorted_idx, out);
  PCP_LAUNCH_CHECK("segment_reduce_kernel");
  return 0;
}
