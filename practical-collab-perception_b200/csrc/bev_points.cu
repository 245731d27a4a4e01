// bev_points.cu - point <-> BEV image exchange of the HunterJr correction head
// (reference: pcdet/models/bev_layers/hunter_toolbox.py:8-41 bilinear_interpolate_torch, :65-96 bev_scatter,
//  :99-131 interpolate_points_feat_from_bev_img; called every forward by hunter_jr.py:268-279,300).
//
//   pcp_bev_interpolate   : per point, bilinear blend of the 4 neighbouring BEV pixels, all channels.  The image is
//                           NCHW in the reference (it gathers through a 'C H W -> H W C' VIEW, i.e. one 4-byte read
//                           per (point, corner, channel), each in a different 1 MB plane); here the image is first
//                           transposed to channels-last ONCE (a streaming pass over an L2-sized tensor), so that a
//                           warp reads each corner of its point as contiguous 128-byte lines.
//   pcp_bev_scatter_mean  : points -> pixels with the SAME machinery as the pillar path: dense per-pixel histogram,
//                           cell scan, counting-sort placement, ascending row order inside a pixel, then the
//                           per-pixel mean (sequential fp32 sum in row order / count == the CPU scatter_mean, bit for
//                           bit) and the dense canvas writer of scatter.cu (every output element written once).
#include "internal.cuh"

namespace pcp {

// ------------------------------------------------------------------------------------------------
// keying: hunter_toolbox.py:78-84
//   mask_in = (x > 0) & (x < width) & (y > 0) & (y < height)       (float compares, strict)
//   coord.long(): truncation; merge = batch * area + y * width + x
// ------------------------------------------------------------------------------------------------
constexpr int kKeyPts = 4;

__global__ void __launch_bounds__(256)
bev_key_count_kernel(const float* __restrict__ coord, int64_t cstride, const int64_t* __restrict__ batch, int64_t n,
                     int32_t frames, int32_t height, int32_t width, int32_t* __restrict__ cell,
                     int32_t* __restrict__ key, int32_t* __restrict__ within, int32_t* __restrict__ hdr) {
  const int64_t base = (int64_t)blockIdx.x * (256 * kKeyPts) + threadIdx.x;
  int32_t k[kKeyPts], w[kKeyPts];
#pragma unroll
  for (int u = 0; u < kKeyPts; ++u) {
    const int64_t i = base + u * 256;
    k[u] = -1; w[u] = 0;
    if (i < n) {
      const float x = __ldg(coord + i * cstride), y = __ldg(coord + i * cstride + 1);
      if (x > 0.f && x < (float)width && y > 0.f && y < (float)height) {
        const int64_t b = __ldg(batch + i);
        if (b < 0 || b >= frames) atomicAdd(&hdr[PCP_COUNT_BAD_FRAME], 1);
        else k[u] = (int32_t)b * (height * width) + (int32_t)y * width + (int32_t)x;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kKeyPts; ++u)
    if (k[u] >= 0) w[u] = atomicAdd(&cell[k[u]], 1);
#pragma unroll
  for (int u = 0; u < kKeyPts; ++u) {
    const int64_t i = base + u * 256;
    if (i < n) { key[i] = k[u]; within[i] = w[u]; }
  }
}

__global__ void __launch_bounds__(256)
max_index_kernel(const int64_t* __restrict__ idx, int64_t n, int32_t* __restrict__ out) {
  long long m = -1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (long long)__ldg(idx + i));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(out, (int32_t)min(m, (long long)0x7ffffffe) + 1);
}

// ------------------------------------------------------------------------------------------------
// NCHW -> NHWC (per frame: [C][HW] -> [HW][C]) through a 32 x 33 shared-memory tile
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* src = in + (int64_t)b * C * HW;
  float* dst = out + (int64_t)b * C * HW;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j;
    const int64_t p = p0 + tx;
    tile[ty + j][tx] = (c < C && p < HW) ? __ldg(src + (int64_t)c * HW + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int64_t p = p0 + ty + j;
    const int c = c0 + tx;
    if (c < C && p < HW) dst[p * C + c] = tile[tx][ty + j];
  }
}

// ------------------------------------------------------------------------------------------------
// bilinear gather: hunter_toolbox.py:18-40 evaluated per point, one warp per point, lanes over channels.
//   x0 = floor(x), x1 = x0 + 1, both clamped to [0, W-1] (same for y)
//   wa = (x1 - x) * (y1 - y); wb = (x1 - x) * (y - y0); wc = (x - x0) * (y1 - y); wd = (x - x0) * (y - y0)   (CLAMPED x0..y1)
//   ans = ((Ia * wa + Ib * wb) + Ic * wc) + Id * wd      Ia = im[y0, x0], Ib = im[y1, x0], Ic = im[y0, x1], Id = im[y1, x1]
// every product and sum is a separate fp32 rounding in the reference (separate torch kernels): no FMA contraction here.
// Points whose frame index is outside [0, B) keep a zero row (:116-123 only fills b in range(batch_size)).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bilinear_gather_kernel(const float* __restrict__ img /* (B, H, W, C) */, int B, int C, int H, int W,
                       const float* __restrict__ points, int64_t stride, int64_t n, float min_x, float min_y, float pix_x,
                       float pix_y, float* __restrict__ feat_out, float* __restrict__ coord_out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const float* row = points + i * stride;
  const float bf = __ldg(row), px = __ldg(row + 1), py = __ldg(row + 2);
  const float x = __fdiv_rn(__fsub_rn(px, min_x), pix_x);       // :114
  const float y = __fdiv_rn(__fsub_rn(py, min_y), pix_y);
  if (coord_out && lane == 0) { coord_out[2 * i] = x; coord_out[2 * i + 1] = y; }
  float* dst = feat_out + i * C;
  const bool in_batch = (bf > -1.f) && (bf < (float)B);          // points[:, 0].long() truncates toward zero
  if (!in_batch) {
    for (int c = lane; c < C; c += 32) dst[c] = 0.f;
    return;
  }
  const int b = (int)bf;
  const float fx0 = floorf(x), fy0 = floorf(y);
  const float wm = (float)(W - 1), hm = (float)(H - 1);
  // clamp in the float domain (the reference clamps the int64 cast; identical for every finite coordinate)
  const float x0 = fminf(fmaxf(fx0, 0.f), wm), x1 = fminf(fmaxf(__fadd_rn(fx0, 1.f), 0.f), wm);
  const float y0 = fminf(fmaxf(fy0, 0.f), hm), y1 = fminf(fmaxf(__fadd_rn(fy0, 1.f), 0.f), hm);
  const float wa = __fmul_rn(__fsub_rn(x1, x), __fsub_rn(y1, y));
  const float wb = __fmul_rn(__fsub_rn(x1, x), __fsub_rn(y, y0));
  const float wc = __fmul_rn(__fsub_rn(x, x0), __fsub_rn(y1, y));
  const float wd = __fmul_rn(__fsub_rn(x, x0), __fsub_rn(y, y0));
  const int ix0 = (int)x0, ix1 = (int)x1, iy0 = (int)y0, iy1 = (int)y1;
  const float* base = img + (int64_t)b * H * W * C;
  const float* pa = base + ((int64_t)iy0 * W + ix0) * C;
  const float* pb = base + ((int64_t)iy1 * W + ix0) * C;
  const float* pc = base + ((int64_t)iy0 * W + ix1) * C;
  const float* pd = base + ((int64_t)iy1 * W + ix1) * C;
  for (int c = lane; c < C; c += 32) {
    const float ia = __ldg(pa + c), ib = __ldg(pb + c), ic = __ldg(pc + c), id = __ldg(pd + c);
    float acc = __fadd_rn(__fmul_rn(ia, wa), __fmul_rn(ib, wb));
    acc = __fadd_rn(acc, __fmul_rn(ic, wc));
    acc = __fadd_rn(acc, __fmul_rn(id, wd));
    dst[c] = acc;
  }
}

}  // namespace pcp

using namespace pcp;

extern "C" int pcp_max_index_i64(const int64_t* index, int64_t n, int32_t* max_plus_one_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(max_plus_one_out && n >= 0 && (n == 0 || index), PCP_E_INVALID, "pcp_max_index_i64: bad argument");
  PCP_CUDA(cudaMemsetAsync(max_plus_one_out, 0, sizeof(int32_t), stream));
  if (n > 0) {
    max_index_kernel<<<sm_count(), 256, 0, stream>>>(index, n, max_plus_one_out);
    PCP_LAUNCH_CHECK("max_index_kernel");
  }
  return 0;
}

extern "C" int pcp_bev_scatter_mean(const float* bev_coord, int64_t coord_stride, const int64_t* batch_idx,
                                    const float* feat, int64_t feat_stride, int32_t channels, int64_t n_points,
                                    int32_t num_frames, int32_t height, int32_t width, void* workspace,
                                    size_t workspace_bytes, float* cell_mean_scratch, float* bev_out,
                                    int32_t* counts_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(workspace && bev_out && counts_out, PCP_E_INVALID, "pcp_bev_scatter_mean: null argument");
  PCP_REQUIRE(n_points >= 0 && n_points < (1ll << 29), PCP_E_INVALID, "pcp_bev_scatter_mean: n_points out of range (< 2^29)");
  PCP_REQUIRE(n_points == 0 || (bev_coord && batch_idx && feat && cell_mean_scratch), PCP_E_INVALID,
              "pcp_bev_scatter_mean: null input");
  PCP_REQUIRE(channels > 0 && feat_stride >= channels && coord_stride >= 2, PCP_E_INVALID, "pcp_bev_scatter_mean: bad strides");
  PCP_REQUIRE(num_frames > 0 && num_frames <= 65535 && height > 0 && width > 0 && height <= 65535 && width <= 65535,
              PCP_E_INVALID, "pcp_bev_scatter_mean: bad image shape");
  PCP_REQUIRE((int64_t)num_frames * height * width < (1ll << 31), PCP_E_UNSUPPORTED,
              "pcp_bev_scatter_mean: frames*height*width must fit int32");
  // cells are keyed (frame, y, x): the grouping stage sees a grid of `height` x `width` cells per frame
  const WsLayout L = ws_layout(n_points, num_frames, height, width);
  PCP_REQUIRE(workspace_bytes >= L.total, PCP_E_WORKSPACE, "pcp_bev_scatter_mean: workspace %zu < %zu bytes", workspace_bytes,
              L.total);
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PCP_E_INVALID, "pcp_bev_scatter_mean: workspace not 256-byte aligned");
  const WsView W = ws_view(workspace, L);
  PCP_CUDA(cudaMemsetAsync(workspace, 0, L.clear_bytes, stream));
  if (n_points > 0) {
    const unsigned blocks = (unsigned)((n_points + 256 * kKeyPts - 1) / (256 * kKeyPts));
    bev_key_count_kernel<<<blocks, 256, 0, stream>>>(bev_coord, coord_stride, batch_idx, n_points, num_frames, height, width,
                                                     W.cell, W.key, W.within, W.hdr);
    PCP_LAUNCH_CHECK("bev_key_count_kernel");
  }
  pcp_grid g{};
  g.voxel_x = 1.f; g.voxel_y = 1.f; g.nx = height; g.ny = width;
  int rc = finish_grouping(L, W, n_points, height, width, nullptr, 0, g, nullptr, nullptr, nullptr, counts_out, stream);
  if (rc) return rc;
  if (n_points > 0) {
    rc = launch_segment_reduce(feat, feat_stride, channels, 0, W, cell_mean_scratch, stream);   // :86 scatter_mean
    if (rc) return rc;
  }
  // :85-88: zeros everywhere else, '(B H W) C -> B C H W'; the rank map is already canvas-ordered (frame, y, x)
  return launch_canvas_from_map(cell_mean_scratch, W.cell_rank, channels, num_frames, width, height, bev_out, stream);
}

extern "C" int pcp_bev_interpolate(const float* bev_img, int32_t channels_last, int32_t num_frames, int32_t channels,
                                   int32_t height, int32_t width, const float* points, int64_t row_stride, int64_t n_points,
                                   float range_min_x, float range_min_y, float pixel_x, float pixel_y, float* nhwc_scratch,
                                   float* points_feat_out, float* bev_coord_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(bev_img && points_feat_out, PCP_E_INVALID, "pcp_bev_interpolate: null argument");
  PCP_REQUIRE(num_frames > 0 && num_frames <= 65535 && channels > 0 && height > 0 && width > 0, PCP_E_INVALID,
              "pcp_bev_interpolate: bad image shape");
  PCP_REQUIRE(n_points >= 0 && (n_points == 0 || points) && row_stride >= 3, PCP_E_INVALID, "pcp_bev_interpolate: bad points");
  PCP_REQUIRE(channels_last || nhwc_scratch, PCP_E_INVALID, "pcp_bev_interpolate: an NCHW image needs the channels-last scratch");
  if (n_points == 0) return 0;
  const float* img = bev_img;
  if (!channels_last) {
    const int64_t hw = (int64_t)height * width;
    const dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((channels + 31) / 32), (unsigned)num_frames);
    nchw_to_nhwc_kernel<<<grid, 256, 0, stream>>>(bev_img, nhwc_scratch, channels, hw);
    PCP_LAUNCH_CHECK("nchw_to_nhwc_kernel");
    img = nhwc_scratch;
  }
  const unsigned blocks = (unsigned)((n_points + 7) / 8);
  bilinear_gather_kernel<<<blocks, 256, 0, stream>>>(img, num_frames, channels, height, width, points, row_stride, n_points,
                                                     range_min_x, range_min_y, pixel_x, pixel_y, points_feat_out, bev_coord_out);
  PCP_LAUNCH_CHECK("bilinear_gather_kernel");
  return 0;
}
