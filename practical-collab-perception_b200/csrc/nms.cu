// nms.cu - late-fusion box NMS on the GPU, end to end.
// Reference: pcdet/models/detectors/v2x_late_fusion.py:21-35 -> model_nms_utils.class_agnostic_nms
// (pcdet/models/model_utils/model_nms_utils.py:6-27: score mask, top-k by score, NMS, post max size) ->
// iou3d_nms_utils.nms_gpu (pcdet/ops/iou3d_nms/iou3d_nms_utils.py:84-99: sort by score, suppression mask on the GPU
// (iou3d_nms_kernel.cu:267-312), the greedy scan over the mask on the HOST after a cudaMalloc + D2H copy,
// iou3d_nms.cpp:100-135).
//
// Here nothing leaves the device and nothing is allocated:
//   nms_order_kernel  : one CTA - score mask, bitonic sort by (score desc, index asc) in shared memory, pre-NMS top-k
//   nms_mask_kernel   : 64 x 64 tiles of the upper triangle: bit j of mask[i][j / 64] = IoU_bev(i, j) > thresh
//   nms_scan_kernel   : one warp - greedy scan in score order with the removal words in registers, post-NMS max size
// The BEV IoU is evaluated differently from the reference (which intersects all edge pairs, collects inside corners and
// sorts the points by angle): box A's corners are taken into box B's frame, where B is axis-aligned, clipped against B's
// four sides (Sutherland-Hodgman) and the polygon area is a shoelace sum.  Same quantity, fp32, agrees to ~1e-6.
#include "internal.cuh"

namespace pcp {

constexpr int kNmsMaxBoxes = 4096;      // boxes after the score mask that one CTA can order in shared memory

struct P2 { float x, y; };

// clip polygon (n <= 8 vertices) against the half-plane  s * coord <= lim  (coord = x if axis == 0 else y)
__device__ __forceinline__ int clip_axis(const P2* in, int n, P2* out, int axis, float s, float lim) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const P2 a = in[i], b = in[(i + 1 == n) ? 0 : i + 1];
    const float da = s * (axis ? a.y : a.x) - lim, db = s * (axis ? b.y : b.x) - lim;   // <= 0 : inside
    if (da <= 0.f) out[m++] = a;
    if ((da < 0.f && db > 0.f) || (da > 0.f && db < 0.f)) {
      const float t = da / (da - db);
      P2 c;
      c.x = a.x + t * (b.x - a.x);
      c.y = a.y + t * (b.y - a.y);
      out[m++] = c;
    }
  }
  return m;
}

// overlap area of two BEV boxes [x, y, z, dx, dy, dz, heading]
__device__ float bev_overlap(const float* __restrict__ a, const float* __restrict__ b) {
  const float ca = cosf(a[6]), sa = sinf(a[6]), cb = cosf(b[6]), sb = sinf(b[6]);
  const float hax = 0.5f * a[3], hay = 0.5f * a[4], hbx = 0.5f * b[3], hby = 0.5f * b[4];
  const float dx = a[0] - b[0], dy = a[1] - b[1];
  P2 p[10], q[10];
  const float lx[4] = {-hax, hax, hax, -hax}, ly[4] = {-hay, -hay, hay, hay};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // corner of A in the world frame (relative to B's centre), then rotated by -heading_b
    const float wx = lx[k] * ca - ly[k] * sa + dx, wy = lx[k] * sa + ly[k] * ca + dy;
    p[k].x = wx * cb + wy * sb;
    p[k].y = -wx * sb + wy * cb;
  }
  int n = 4;
  n = clip_axis(p, n, q, 0, 1.f, hbx);
  if (n < 3) return 0.f;
  n = clip_axis(q, n, p, 0, -1.f, hbx);
  if (n < 3) return 0.f;
  n = clip_axis(p, n, q, 1, 1.f, hby);
  if (n < 3) return 0.f;
  n = clip_axis(q, n, p, 1, -1.f, hby);
  if (n < 3) return 0.f;
  float area = 0.f;
  for (int i = 1; i + 1 < n; ++i)
    area += (p[i].x - p[0].x) * (p[i + 1].y - p[0].y) - (p[i].y - p[0].y) * (p[i + 1].x - p[0].x);
  return 0.5f * fabsf(area);
}

// iou3d_nms_kernel.cu:227-234
__device__ __forceinline__ float bev_iou(const float* a, const float* b) {
  const float sa = a[3] * a[4], sb = b[3] * b[4];
  const float so = bev_overlap(a, b);
  return so / fmaxf(sa + sb - so, 1e-8f);
}

__global__ void __launch_bounds__(256)
boxes_iou_bev_kernel(const float* __restrict__ boxes_a, int64_t na, const float* __restrict__ boxes_b, int64_t nb,
                     float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na * nb) return;
  out[i] = bev_iou(boxes_a + (i / nb) * 7, boxes_b + (i % nb) * 7);
}

// ---- 1. order: score mask + sort by (score desc, index asc) + top-k -------------------------------------------------
__global__ void __launch_bounds__(1024)
nms_order_kernel(const float* __restrict__ scores, int64_t n, float score_thresh, int apply_thresh, int pre_max,
                 int32_t* __restrict__ order, int32_t* __restrict__ hdr) {
  __shared__ float s_key[kNmsMaxBoxes];
  __shared__ int32_t s_idx[kNmsMaxBoxes];
  __shared__ int s_cnt;
  const int tid = threadIdx.x;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  // compaction in index order is not needed: the sort key carries the index
  for (int64_t i = tid; i < n; i += blockDim.x) {
    const float s = scores[i];
    if (!apply_thresh || s >= score_thresh) {                 // model_nms_utils.py:9
      const int pos = atomicAdd(&s_cnt, 1);
      if (pos < kNmsMaxBoxes) { s_key[pos] = s; s_idx[pos] = (int32_t)i; }
    }
  }
  __syncthreads();
  const int cnt = s_cnt;
  if (cnt > kNmsMaxBoxes) {                                   // reported to the host through hdr[1]
    if (tid == 0) { hdr[0] = 0; hdr[1] = cnt; }
    return;
  }
  int m = 1;
  while (m < cnt) m <<= 1;
  for (int i = cnt + tid; i < m; i += blockDim.x) { s_key[i] = -INFINITY; s_idx[i] = 0x7fffffff; }
  __syncthreads();
  // "a before b" = higher score first, ties by lower index
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < m; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const float ka = s_key[i], kb = s_key[p];
          const int32_t ia = s_idx[i], ib = s_idx[p];
          const bool a_first = (ka > kb) || (ka == kb && ia < ib);
          const bool up = (i & k) == 0;
          if (a_first != up) { s_key[i] = kb; s_key[p] = ka; s_idx[i] = ib; s_idx[p] = ia; }
        }
      }
      __syncthreads();
    }
  }
  const int keep = min(cnt, pre_max > 0 ? pre_max : cnt);     // torch.topk(k = min(NMS_PRE_MAXSIZE, n)), :15
  for (int i = tid; i < keep; i += blockDim.x) order[i] = s_idx[i];
  if (tid == 0) { hdr[0] = keep; hdr[1] = cnt; }
}

// ---- 2. suppression mask (upper triangle) -----------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float* __restrict__ boxes, int64_t box_stride, const int32_t* __restrict__ order,
                const int32_t* __restrict__ hdr, float thresh, int col_blocks, unsigned long long* __restrict__ mask) {
  const int n = hdr[0];
  const int rb = blockIdx.y, cbk = blockIdx.x;
  if (cbk < rb || rb * 64 >= n || cbk * 64 >= n) return;     // only the upper triangle is ever read
  __shared__ float s_col[64 * 7];
  const int t = threadIdx.x;
  const int cj = cbk * 64 + t;
  if (cj < n) {
    const float* src = boxes + (int64_t)order[cj] * box_stride;
#pragma unroll
    for (int k = 0; k < 7; ++k) s_col[t * 7 + k] = src[k];
  }
  __syncthreads();
  const int ri = rb * 64 + t;
  if (ri >= n) return;
  float a[7];
  const float* src = boxes + (int64_t)order[ri] * box_stride;
#pragma unroll
  for (int k = 0; k < 7; ++k) a[k] = src[k];
  unsigned long long bits = 0;
  const int ncol = min(64, n - cbk * 64);
  const int start = (rb == cbk) ? t + 1 : 0;
  for (int j = start; j < ncol; ++j)
    if (bev_iou(a, s_col + j * 7) > thresh) bits |= 1ull << j;
  mask[(int64_t)ri * col_blocks + cbk] = bits;
}

// ---- 3. greedy scan in score order (iou3d_nms.cpp:116-131), one warp, lanes own removal words -------------------------
__global__ void __launch_bounds__(32)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int32_t* __restrict__ order,
                const int32_t* __restrict__ hdr, int col_blocks, int post_max, int64_t* __restrict__ keep_out,
                int32_t* __restrict__ count_out) {
  const int n = hdr[0];
  const int lane = threadIdx.x;
  unsigned long long remv[2] = {0ull, 0ull};                 // words lane and lane + 32 (col_blocks <= 64)
  int kept = 0;
  const int limit = post_max > 0 ? post_max : n;
  const int nblk = (n + 63) >> 6;                            // column blocks nms_mask_kernel wrote
  for (int i = 0; i < n && kept < limit; ++i) {
    const int w = i >> 6;
    const unsigned long long word = __shfl_sync(0xffffffffu, w < 32 ? remv[0] : remv[1], w & 31);
    if (!((word >> (i & 63)) & 1ull)) {
      if (lane == 0) keep_out[kept] = order[i];
      ++kept;
      // only words at or after the diagonal block were written by nms_mask_kernel
      // (and only the column blocks that hold candidates: n may be far below the capacity col_blocks was sized for)
      if (lane >= w && lane < nblk) remv[0] |= mask[(int64_t)i * col_blocks + lane];
      if (lane + 32 >= w && lane + 32 < nblk) remv[1] |= mask[(int64_t)i * col_blocks + lane + 32];
    }
  }
  if (lane == 0) *count_out = (hdr[1] > kNmsMaxBoxes) ? -1 : kept;      // -1: more boxes pass the score mask than one CTA can order
}

}  // namespace pcp

using namespace pcp;

extern "C" int pcp_boxes_iou_bev(const float* boxes_a, int64_t num_a, const float* boxes_b, int64_t num_b, float* iou_out,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(num_a >= 0 && num_b >= 0, PCP_E_INVALID, "pcp_boxes_iou_bev: negative size");
  if (num_a == 0 || num_b == 0) return 0;
  PCP_REQUIRE(boxes_a && boxes_b && iou_out, PCP_E_INVALID, "pcp_boxes_iou_bev: null argument");
  const int64_t total = num_a * num_b;
  boxes_iou_bev_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(boxes_a, num_a, boxes_b, num_b, iou_out);
  PCP_LAUNCH_CHECK("boxes_iou_bev_kernel");
  return 0;
}

extern "C" size_t pcp_nms_scratch_bytes(int64_t num_boxes) {
  if (num_boxes < 0) return 0;
  const int64_t n = num_boxes < kNmsMaxBoxes ? num_boxes : kNmsMaxBoxes;
  const int64_t col_blocks = (n + 63) / 64;
  // hdr int32[4] | order int32[n] | mask u64[n * col_blocks]
  return 256 + align_up(4 * (size_t)(n + 1), 256) + 8 * (size_t)(n * col_blocks + 1);
}

extern "C" int pcp_nms_bev(const float* boxes, int64_t box_stride, const float* scores, int64_t num_boxes,
                           int32_t apply_score_thresh, float score_thresh, float iou_thresh, int32_t pre_max_size,
                           int32_t post_max_size, void* scratch, size_t scratch_bytes, int64_t* keep_out,
                           int32_t* count_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(count_out && num_boxes >= 0 && num_boxes < (1ll << 31), PCP_E_INVALID, "pcp_nms_bev: bad argument");
  if (num_boxes == 0) {
    PCP_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), stream));
    return 0;
  }
  PCP_REQUIRE(boxes && scores && keep_out && scratch && box_stride >= 7, PCP_E_INVALID, "pcp_nms_bev: null argument / stride < 7");
  PCP_REQUIRE(scratch_bytes >= pcp_nms_scratch_bytes(num_boxes), PCP_E_WORKSPACE, "pcp_nms_bev: scratch too small");
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, PCP_E_INVALID, "pcp_nms_bev: scratch not 256-byte aligned");
  const int64_t n = num_boxes < kNmsMaxBoxes ? num_boxes : kNmsMaxBoxes;
  const int col_blocks = (int)((n + 63) / 64);
  char* sp = static_cast<char*>(scratch);
  int32_t* hdr = reinterpret_cast<int32_t*>(sp);                          sp += 256;
  int32_t* order = reinterpret_cast<int32_t*>(sp);                        sp += align_up(4 * (size_t)(n + 1), 256);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(sp);
  nms_order_kernel<<<1, 1024, 0, stream>>>(scores, num_boxes, score_thresh, apply_score_thresh, pre_max_size, order, hdr);
  PCP_LAUNCH_CHECK("nms_order_kernel");
  nms_mask_kernel<<<dim3((unsigned)col_blocks, (unsigned)col_blocks), 64, 0, stream>>>(boxes, box_stride, order, hdr, iou_thresh,
                                                                                      col_blocks, mask);
  PCP_LAUNCH_CHECK("nms_mask_kernel");
  nms_scan_kernel<<<1, 32, 0, stream>>>(mask, order, hdr, col_blocks, post_max_size, keep_out, count_out);
  PCP_LAUNCH_CHECK("nms_scan_kernel");
  return 0;
}
