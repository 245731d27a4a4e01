// pfn.cu - C-ABI entry points of the pillar feature network (reference: dynamic_pillar_vfe.py:110-129 +
// PFNLayerV2 :35-46): parameter packing (BatchNorm folding, TF32 hi/lo operand panels), the launch of the
// tensor-core kernel (pfn_tc.cu), the finishing kernel of long pillars, and the stand-alone segmented
// mean / max (torch_scatter.scatter_mean / scatter_max).
#include "internal.cuh"
#include "pfn_tc.cuh"
#include "umma.cuh"

namespace pcp {

// ------------------------------------------------------------------------------------------------
// parameter packing: fold BN(eval) into scale / shift, sign-fold the last layer, split into TF32 hi / lo panels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fold_bn(const float* g, const float* be, const float* mu, const float* var,
                                        const float* lin_bias, float eps, int c, float& a, float& b) {
  if (g) {
    // ATen eval batch_norm: alpha = weight * rsqrt(var + eps), beta = bias - mean * alpha
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var[c], eps)));
    a = __fmul_rn(g[c], inv);
    b = __fsub_rn(be[c], __fmul_rn(mu[c], a));
  } else {
    a = 1.f;
    b = lin_bias ? lin_bias[c] : 0.f;
  }
}

__global__ void pack_params_kernel(int c_in, int num_layers, const float* w0, const float* lb0, const float* g0,
                                   const float* be0, const float* mu0, const float* var0, const float* w1,
                                   const float* lb1, const float* g1, const float* be1, const float* mu1,
                                   const float* var1, float eps, float* out) {
  const ParamLayout P = param_layout(c_in, num_layers);
  const int k0 = pfn_k0(c_in);
  const int n0 = num_layers == 2 ? kHidden : kCout;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // layer 0 panels [k0/4][n0][4]; a single layer is the last layer: its rows carry the sign of its BN scale.
  // Two layers with a spare K column (k0 > c_in, pfn_fold0): the BN of layer 0 goes INTO the operand - rows scaled by
  // alpha, beta in column c_in, which the producers feed with 1.0 - so the layer-0 epilogue is a bare ReLU.
  const bool fold = pfn_fold0(c_in, num_layers);
  for (int i = tid; i < k0 * n0; i += nth) {
    const int kc = i / (n0 * 4), n = (i / 4) % n0, k = kc * 4 + (i & 3);
    float a, b;
    fold_bn(g0, be0, mu0, var0, lb0, eps, n, a, b);
    float w;
    if (fold) w = (k < c_in) ? __fmul_rn(a, w0[n * c_in + k]) : (k == c_in ? b : 0.f);
    else w = (k < c_in) ? ((num_layers == 1 && a < 0.f) ? -1.f : 1.f) * w0[n * c_in + k] : 0.f;
    float h, l;
    umma::split_tf32_rn(w, h, l);
    out[P.w0h + i] = h; out[P.w0l + i] = l;
  }
  for (int c = tid; c < n0; c += nth) {
    float a, b;
    fold_bn(g0, be0, mu0, var0, lb0, eps, c, a, b);
    out[P.a0 + c] = fold ? 1.f : ((num_layers == 1) ? fabsf(a) : a);
    out[P.b0 + c] = fold ? 0.f : b;
  }
  if (num_layers == 2) {
    // layer 1: columns 0..31 act on x (per point), columns 32..63 on x_max (per pillar); rows whose BN scale is
    // negative are negated (exact) so that BN+ReLU is non-decreasing in the accumulator for every channel
    for (int i = tid; i < kHidden * kCout; i += nth) {
      const int kc = i / (kCout * 4), n = (i / 4) % kCout, k = kc * 4 + (i & 3);
      float a, b;
      fold_bn(g1, be1, mu1, var1, lb1, eps, n, a, b);
      const float sgn = (a < 0.f) ? -1.f : 1.f;
      float h, l;
      umma::split_tf32_rn(sgn * w1[n * (2 * kHidden) + k], h, l);
      out[P.w1ah + i] = h; out[P.w1al + i] = l;
      umma::split_tf32_rn(sgn * w1[n * (2 * kHidden) + kHidden + k], h, l);
      out[P.w1bh + i] = h; out[P.w1bl + i] = l;
      umma::split_tf32_rn(sgn * __fadd_rn(w1[n * (2 * kHidden) + k], w1[n * (2 * kHidden) + kHidden + k]), h, l);
      out[P.w1sh + i] = h; out[P.w1sl + i] = l;
    }
    for (int i = tid; i < kCout * kHidden; i += nth) {
      const int n = i / kHidden, k = i % kHidden;
      float a, b;
      fold_bn(g1, be1, mu1, var1, lb1, eps, n, a, b);
      out[P.w1b_f32 + i] = ((a < 0.f) ? -1.f : 1.f) * w1[n * (2 * kHidden) + kHidden + k];
    }
    for (int c = tid; c < kCout; c += nth) {
      float a, b;
      fold_bn(g1, be1, mu1, var1, lb1, eps, c, a, b);
      out[P.a1 + c] = fabsf(a);
      out[P.b1 + c] = b;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// long pillars: the segments' partial maxima are in long_acc (ordered ints); one warp per pillar applies the
// per-pillar half of layer 1 (fp32 FMA chain) and BN + ReLU, then re-arms the accumulator for the next call
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pfn_finish_long_kernel(const int32_t* __restrict__ hdr, const int4* __restrict__ long_table, unsigned* __restrict__ long_acc,
                       const float4* __restrict__ long_mean, const float* __restrict__ params, int c_in, int num_layers,
                       float* __restrict__ out, float* __restrict__ mean_out) {
  __shared__ float s_w[kHidden][kCout + 1];     // W1[:, 32:] transposed: s_w[k][n], lanes read consecutive n
  const ParamLayout P = param_layout(c_in, num_layers);
  const int nlong = hdr[kHdrLongCount];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  if ((int)(blockIdx.x * wpb) >= nlong) return;
  if (num_layers == 2) {
    for (int i = threadIdx.x; i < kCout * kHidden; i += blockDim.x) s_w[i % kHidden][i / kHidden] = __ldg(params + P.w1b_f32 + i);
  }
  __syncthreads();
  const int pa = num_layers == 2 ? P.a1 : P.a0, pb = num_layers == 2 ? P.b1 : P.b0;
  const float sa0 = __ldg(params + pa + lane), sa1 = __ldg(params + pa + 32 + lane);
  const float sb0 = __ldg(params + pb + lane), sb1 = __ldg(params + pb + 32 + lane);
  for (int li = blockIdx.x * wpb + (threadIdx.x >> 5); li < nlong; li += gridDim.x * wpb) {
    const int r = long_table[li].x;
    if (mean_out && lane == 0) {
      const float4 m = long_mean[li];
      mean_out[(int64_t)r * 3 + 0] = m.x; mean_out[(int64_t)r * 3 + 1] = m.y; mean_out[(int64_t)r * 3 + 2] = m.z;
    }
    unsigned* acc = long_acc + (int64_t)li * 96;
    const float x0 = ord_dec(acc[lane]);
    const float ma = ord_dec(acc[32 + lane]), mb = ord_dec(acc[64 + lane]);
    __syncwarp();
    acc[lane] = kAccInit; acc[32 + lane] = kAccInit; acc[64 + lane] = kAccInit;
    float ha = 0.f, hb = 0.f;
    if (num_layers == 2) {
#pragma unroll
      for (int k = 0; k < kHidden; ++k) {
        const float xk = __shfl_sync(0xffffffffu, x0, k);
        ha = fmaf(xk, s_w[k][lane], ha);
        hb = fmaf(xk, s_w[k][lane + 32], hb);
      }
    }
    out[(int64_t)r * kCout + lane] = fmaxf(fmaf(__fadd_rn(ma, ha), sa0, sb0), 0.f);
    out[(int64_t)r * kCout + 32 + lane] = fmaxf(fmaf(__fadd_rn(mb, hb), sa1, sb1), 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// standalone segmented mean / max over the pillars (torch_scatter.scatter_mean / scatter_max).
// L lanes per pillar (a power of two), 32 / L pillars per warp side by side, kVec consecutive channels per lane; the rows of a
// pillar are visited in ascending row order, four row gathers in flight per lane (the mean is summed sequentially in
// that order: the CPU scatter_mean's rounding, bit for bit; the loads run ahead of the dependent adds).
// ------------------------------------------------------------------------------------------------
template <int kMode, int kVec>
__global__ void __launch_bounds__(256)
segment_reduce_kernel(const float* __restrict__ values, int64_t vstride, int channels, int L,
                      const int32_t* __restrict__ hdr, const int32_t* __restrict__ seg_off,
                      const int32_t* __restrict__ sorted_idx, float* __restrict__ out) {
  constexpr int kAhead = 4;
  const int P = hdr[PCP_COUNT_PILLARS];
  const int lane = threadIdx.x & 31;
  const int G = 32 / L, sub = lane / L, ln = lane - sub * L;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t base = warp_global * G; base < P; base += nwarps * G) {
    const int64_t r = base + sub;
    const bool valid = r < P;
    const int off = valid ? seg_off[r] : 0;
    const int n = valid ? seg_off[r + 1] - off : 0;
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    for (int c0 = 0; c0 < channels; c0 += L * kVec) {
      const int c = c0 + ln * kVec;
      const bool cok = c < channels;
      float acc[kVec];
#pragma unroll
      for (int t = 0; t < kVec; ++t) acc[t] = (kMode == 0) ? 0.f : -INFINITY;
      for (int j0 = 0; j0 < nmax; j0 += kAhead) {
        float v[kAhead][kVec];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
          const bool on = cok && (j0 + u < n);
          if (on) {
            const float* src = values + (int64_t)sorted_idx[off + j0 + u] * vstride + c;
            if (kVec == 4) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(src));
              v[u][0] = q.x; v[u][1 % kVec] = q.y; v[u][2 % kVec] = q.z; v[u][3 % kVec] = q.w;
            } else {
              v[u][0] = __ldg(src);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
          if (cok && (j0 + u < n)) {
#pragma unroll
            for (int t = 0; t < kVec; ++t) acc[t] = (kMode == 0) ? __fadd_rn(acc[t], v[u][t]) : fmaxf(acc[t], v[u][t]);
          }
        }
      }
      if (valid && cok) {
        float* dst = out + r * channels + c;
        if (kMode == 0) {
          const float d = (float)max(n, 1);
#pragma unroll
          for (int t = 0; t < kVec; ++t) acc[t] = __fdiv_rn(acc[t], d);
        }
        if (kVec == 4) *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1 % kVec], acc[2 % kVec], acc[3 % kVec]);
        else dst[0] = acc[0];
      }
    }
  }
}

static int pow2_at_least(int v) {
  int p = 1;
  while (p < v && p < 32) p <<= 1;
  return p;
}

int launch_segment_reduce(const float* values, int64_t value_stride, int32_t channels, int32_t mode, const WsView& W,
                          float* out, cudaStream_t stream) {
  const unsigned blocks = (unsigned)sm_count() * 8;
  const bool vec4 = (channels % 4 == 0) && (value_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(values) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec4) {
    const int L = pow2_at_least(channels / 4);
    if (mode == 0)
      segment_reduce_kernel<0, 4><<<blocks, 256, 0, stream>>>(values, value_stride, channels, L, W.hdr, W.seg_off, W.sorted_idx, out);
    else
      segment_reduce_kernel<1, 4><<<blocks, 256, 0, stream>>>(values, value_stride, channels, L, W.hdr, W.seg_off, W.sorted_idx, out);
  } else {
    const int L = pow2_at_least(channels);
    if (mode == 0)
      segment_reduce_kernel<0, 1><<<blocks, 256, 0, stream>>>(values, value_stride, channels, L, W.hdr, W.seg_off, W.sorted_idx, out);
    else
      segment_reduce_kernel<1, 1><<<blocks, 256, 0, stream>>>(values, value_stride, channels, L, W.hdr, W.seg_off, W.sorted_idx, out);
  }
  PCP_LAUNCH_CHECK("segment_reduce_kernel");
  return 0;
}

}  // namespace pcp

using namespace pcp;

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

extern "C" size_t pcp_pfn_param_floats(const pcp_pfn_desc* d) {
  int c_in = 0;
  if (!d || check_desc(d, &c_in)) return 0;
  return (size_t)param_layout(c_in, d->num_layers).total;
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
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(packed_out) & 15) == 0, PCP_E_INVALID, "pcp_pack_pfn_params: packed_out not 16-byte aligned");
  pack_params_kernel<<<8, 256, 0, stream>>>(c_in, desc->num_layers, w0, lin_bias0, bn0_weight, bn0_bias, bn0_mean,
                                            bn0_var, w1, lin_bias1, bn1_weight, bn1_bias, bn1_mean, bn1_var, eps,
                                            packed_out);
  PCP_LAUNCH_CHECK("pack_params_kernel");
  return 0;
}

extern "C" int pcp_pfn(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                       const pcp_grid* grid, const pcp_pfn_desc* desc, const float* packed_params,
                       const void* workspace, size_t workspace_bytes, float* pillar_features_out,
                       float* pillar_mean_out, int64_t pillar_capacity, void* stream_) {
  return pcp_pfn_stages(points, row_stride, n_points, max_frames, grid, desc, packed_params, workspace, workspace_bytes,
                        pillar_features_out, pillar_mean_out, pillar_capacity, PCP_PFN_STAGE_ALL, stream_);
}

extern "C" int pcp_pfn_stages(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                              const pcp_grid* grid, const pcp_pfn_desc* desc, const float* packed_params,
                              const void* workspace, size_t workspace_bytes, float* pillar_features_out,
                              float* pillar_mean_out, int64_t pillar_capacity, int32_t stages, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(stages != 0 && (stages & ~PCP_PFN_STAGE_ALL) == 0, PCP_E_INVALID, "pcp_pfn_stages: bad stage mask %d", stages);
  int c_in = 0;
  if (int rc = check_desc(desc, &c_in)) return rc;
  PCP_REQUIRE(grid && packed_params && workspace && pillar_features_out, PCP_E_INVALID, "pcp_pfn: null argument");
  PCP_REQUIRE(n_points >= 0 && (n_points == 0 || points), PCP_E_INVALID, "pcp_pfn: bad points");
  PCP_REQUIRE(row_stride >= 1 + desc->c_raw, PCP_E_INVALID, "pcp_pfn: row_stride %lld < 1 + c_raw", (long long)row_stride);
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(packed_params) & 15) == 0 && (reinterpret_cast<uintptr_t>(pillar_features_out) & 15) == 0,
              PCP_E_INVALID, "pcp_pfn: packed_params / pillar_features_out not 16-byte aligned");
  const WsLayout L = ws_layout(n_points, max_frames, grid->nx, grid->ny);
  PCP_REQUIRE(workspace_bytes >= L.total, PCP_E_WORKSPACE, "pcp_pfn: workspace %zu < %zu bytes", workspace_bytes, L.total);
  PCP_REQUIRE(pillar_capacity >= L.cap, PCP_E_INVALID, "pcp_pfn: pillar_capacity too small");
  if (n_points == 0) return 0;
  const WsView W = ws_view(const_cast<void*>(workspace), L);
  TcArgs t;
  t.points = points; t.stride = row_stride; t.g = *grid; t.c_in = c_in;
  t.raw_col0 = desc->use_absolute_xyz ? 1 : 4; t.c_raw = desc->c_raw;
  t.n_raw = desc->use_absolute_xyz ? desc->c_raw : desc->c_raw - 3;
  t.with_distance = desc->with_distance; t.k0 = pfn_k0(c_in); t.num_layers = desc->num_layers;
  t.params = packed_params;
  t.hdr = W.hdr; t.seg_off = W.seg_off; t.sorted_idx = W.sorted_idx; t.lists = W.lists; t.lo = L.lo;
  t.mean = W.mean; t.long_mean = W.long_mean; t.long_acc = W.long_acc; t.long_table = W.long_table;
  t.out = pillar_features_out; t.mean_out = pillar_mean_out;
  if (stages & PCP_PFN_STAGE_SLOTS)
    if (int rc = launch_pfn_tc(t, n_points, stream)) return rc;
  if ((stages & PCP_PFN_STAGE_LONG) && n_points > kSegRows) {
    const int64_t want = (n_points / (kSegRows + 1) + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 4;
    const unsigned blocks = (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
    pfn_finish_long_kernel<<<blocks, 256, 0, stream>>>(W.hdr, W.long_table, W.long_acc, W.long_mean, packed_params, c_in,
                                                       desc->num_layers, pillar_features_out, pillar_mean_out);
    PCP_LAUNCH_CHECK("pfn_finish_long_kernel");
  }
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
  return launch_segment_reduce(values, value_stride, channels, mode, W, out, stream);
}
