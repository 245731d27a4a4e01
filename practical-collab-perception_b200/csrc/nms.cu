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
// The BEV overlap RESTATES THE REFERENCE'S ARITHMETIC (iou3d_nms_kernel.cu:39-225), operation for operation: rotated corners,
// the 16 edge-pair intersections (bounding-rectangle rejection, strict straddle test, EPS = 1e-8 branch), the corner-in-box
// test with its MARGIN = 1e-2 m, the centroid, the ordering by atan2 and the fan area - because that kernel is only an
// approximation of the true overlap (errors up to a few 1e-3 in IoU), so a kept-index list identical to the reference's
// needs the same procedure, not a better area.  tests/test_gpu_nms.py holds the IoU matrix to 1e-5 of the reference kernel
// compiled from its source (> 99 % of the values bit-equal: mul/add fusion is ptxas's per-context choice) and the kept
// indices IDENTICAL to the reference's mask + host scan on unfiltered scenes.
#include "internal.cuh"

namespace pcp {

constexpr int kNmsMaxBoxes = 4096;      // candidates that one CTA can order in shared memory (after score mask + top-k)

struct P2 { float x, y; };

__device__ __forceinline__ P2 p2_add(const P2& a, const P2& b) { P2 r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
__device__ __forceinline__ P2 p2_sub(const P2& a, const P2& b) { P2 r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// :35-37
__device__ __forceinline__ float cross2(const P2& a, const P2& b) { return a.x * b.y - a.y * b.x; }
// :39-41
__device__ __forceinline__ float cross3(const P2& p1, const P2& p2, const P2& p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
// :43-49 bounding rectangles of the two segments overlap
__device__ __forceinline__ bool rect_cross(const P2& p1, const P2& p2, const P2& q1, const P2& q2) {
  return min(p1.x, p2.x) <= max(q1.x, q2.x) && min(q1.x, q2.x) <= max(p1.x, p2.x) &&
         min(p1.y, p2.y) <= max(q1.y, q2.y) && min(q1.y, q2.y) <= max(p1.y, p2.y);
}
// :51-62 point inside the rotated rectangle, 1e-2 m margin
__device__ __forceinline__ bool in_box2d(const float* box, const P2& p) {
  const float MARGIN = 1e-2;
  float center_x = box[0], center_y = box[1];
  float angle_cos = cos(-box[6]), angle_sin = sin(-box[6]);
  float rot_x = (p.x - center_x) * angle_cos + (p.y - center_y) * (-angle_sin);
  float rot_y = (p.x - center_x) * angle_sin + (p.y - center_y) * angle_cos;
  return (fabs(rot_x) < box[3] / 2 + MARGIN && fabs(rot_y) < box[4] / 2 + MARGIN);
}
// :64-95 intersection of segments p0-p1 and q0-q1
__device__ __forceinline__ bool seg_intersection(const P2& p1, const P2& p0, const P2& q1, const P2& q0, P2& ans) {
  const float EPS = 1e-8;
  if (!rect_cross(p0, p1, q0, q1)) return false;
  float s1 = cross3(q0, p1, p0);
  float s2 = cross3(p1, q1, p0);
  float s3 = cross3(p0, q1, q0);
  float s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
  float s5 = cross3(q1, p1, p0);
  if (fabs(s5 - s1) > EPS) {
    ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    float D = a0 * b1 - a1 * b0;
    ans.x = (b0 * c1 - b1 * c0) / D;
    ans.y = (a1 * c0 - a0 * c1) / D;
  }
  return true;
}
// :97-101
__device__ __forceinline__ void rotate_about(const P2& center, float angle_cos, float angle_sin, P2& p) {
  float new_x = (p.x - center.x) * angle_cos + (p.y - center.y) * (-angle_sin) + center.x;
  float new_y = (p.x - center.x) * angle_sin + (p.y - center.y) * angle_cos + center.y;
  p.x = new_x; p.y = new_y;
}

// overlap area of two BEV boxes [x, y, z, dx, dy, dz, heading]: :107-225
__device__ float bev_overlap(const float* box_a, const float* box_b) {
  float a_angle = box_a[6], b_angle = box_b[6];
  float a_dx_half = box_a[3] / 2, b_dx_half = box_b[3] / 2, a_dy_half = box_a[4] / 2, b_dy_half = box_b[4] / 2;
  float a_x1 = box_a[0] - a_dx_half, a_y1 = box_a[1] - a_dy_half;
  float a_x2 = box_a[0] + a_dx_half, a_y2 = box_a[1] + a_dy_half;
  float b_x1 = box_b[0] - b_dx_half, b_y1 = box_b[1] - b_dy_half;
  float b_x2 = box_b[0] + b_dx_half, b_y2 = box_b[1] + b_dy_half;
  P2 center_a, center_b;
  center_a.x = box_a[0]; center_a.y = box_a[1];
  center_b.x = box_b[0]; center_b.y = box_b[1];
  P2 ca[5], cb[5];
  ca[0].x = a_x1; ca[0].y = a_y1; ca[1].x = a_x2; ca[1].y = a_y1; ca[2].x = a_x2; ca[2].y = a_y2; ca[3].x = a_x1; ca[3].y = a_y2;
  cb[0].x = b_x1; cb[0].y = b_y1; cb[1].x = b_x2; cb[1].y = b_y1; cb[2].x = b_x2; cb[2].y = b_y2; cb[3].x = b_x1; cb[3].y = b_y2;
  float a_angle_cos = cos(a_angle), a_angle_sin = sin(a_angle);
  float b_angle_cos = cos(b_angle), b_angle_sin = sin(b_angle);
  for (int k = 0; k < 4; k++) {
    rotate_about(center_a, a_angle_cos, a_angle_sin, ca[k]);
    rotate_about(center_b, b_angle_cos, b_angle_sin, cb[k]);
  }
  ca[4] = ca[0];
  cb[4] = cb[0];
  // polygon vertices: edge intersections (a-edge major), then corners of b inside a / of a inside b, interleaved per k
  P2 pts[16];
  P2 centre;
  centre.x = 0; centre.y = 0;
  int cnt = 0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], pts[cnt])) {
        centre = p2_add(centre, pts[cnt]);
        cnt++;
      }
  for (int k = 0; k < 4; k++) {
    if (in_box2d(box_a, cb[k])) { centre = p2_add(centre, cb[k]); pts[cnt] = cb[k]; cnt++; }
    if (in_box2d(box_b, ca[k])) { centre = p2_add(centre, ca[k]); pts[cnt] = ca[k]; cnt++; }
  }
  centre.x /= cnt;
  centre.y /= cnt;
  // :198-207 bubble sort by the angle around the centroid (the angle of a vertex never changes: evaluated once here,
  // compared in the reference's order, so the permutation is the reference's)
  float ang[16];
  for (int i = 0; i < cnt; i++) ang[i] = atan2(pts[i].y - centre.y, pts[i].x - centre.x);
  for (int j = 0; j < cnt - 1; j++)
    for (int i = 0; i < cnt - j - 1; i++)
      if (ang[i] > ang[i + 1]) {
        const P2 tp = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = tp;
        const float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
      }
  float area = 0;
  for (int k = 0; k < cnt - 1; k++) area += cross2(p2_sub(pts[k], pts[0]), p2_sub(pts[k + 1], pts[0]));
  return fabs(area) / 2.0;
}

// :227-234
__device__ __forceinline__ float bev_iou(const float* box_a, const float* box_b) {
  const float EPS = 1e-8;
  float sa = box_a[3] * box_a[4];
  float sb = box_b[3] * box_b[4];
  float s_overlap = bev_overlap(box_a, box_b);
  return s_overlap / fmaxf(sa + sb - s_overlap, EPS);
}

// axis-aligned IoU of nms_normal_gpu: :316-327
__device__ __forceinline__ float normal_iou(const float* a, const float* b) {
  const float EPS = 1e-8;
  float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  float interS = width * height;
  float Sa = a[3] * a[4];
  float Sb = b[3] * b[4];
  return interS / fmaxf(Sa + Sb - interS, EPS);
}

__global__ void __launch_bounds__(256)
boxes_iou_bev_kernel(const float* __restrict__ boxes_a, int64_t na, const float* __restrict__ boxes_b, int64_t nb,
                     float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= na * nb) return;
  out[i] = bev_iou(boxes_a + (i / nb) * 7, boxes_b + (i % nb) * 7);
}

// ---- 1. order: score mask + (top-k pre-selection) + sort by (score desc, index asc) ------------------------------------
// Candidates = boxes that pass the score mask (model_nms_utils.py:9).  Up to kNmsMaxBoxes of them are ordered in shared
// memory.  When more pass and the caller asked for a pre-NMS top-k of at most kNmsMaxBoxes (torch.topk, :15 - every shipped
// config does: NMS_PRE_MAXSIZE 1000 .. 4096), the k best are selected first by a radix select on the order-preserving
// integer image of the score (4 passes of 8 bits over the candidates); ties at the k-th score keep the lower indices.
__global__ void __launch_bounds__(1024)
nms_order_kernel(const float* __restrict__ scores, int64_t n, float score_thresh, int apply_thresh, int pre_max,
                 int32_t* __restrict__ order, int32_t* __restrict__ hdr) {
  __shared__ float s_key[kNmsMaxBoxes];
  __shared__ int32_t s_idx[kNmsMaxBoxes];
  __shared__ int s_cnt, s_all;
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_prefix, s_need;
  __shared__ int s_warp[32];
  const int tid = threadIdx.x;
  if (tid == 0) { s_cnt = 0; s_all = 0; }
  __syncthreads();
  auto passes = [&](float s) { return !apply_thresh || s >= score_thresh; };
  {
    int mine = 0;
    for (int64_t i = tid; i < n; i += blockDim.x) mine += passes(scores[i]) ? 1 : 0;
    atomicAdd(&s_all, mine);
  }
  __syncthreads();
  const int all = s_all;
  unsigned cut = 0;            // candidates are the boxes with ord_enc(score) > cut, plus the first s_need with == cut
  bool select = false;
  if (all > kNmsMaxBoxes) {
    if (pre_max <= 0 || pre_max > kNmsMaxBoxes) {             // reported to the host through hdr[1]
      if (tid == 0) { hdr[0] = 0; hdr[1] = all; }
      return;
    }
    // radix select of the pre_max-th largest key
    if (tid == 0) { s_prefix = 0; s_need = (unsigned)pre_max; }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (tid < 256) s_hist[tid] = 0;
      __syncthreads();
      const unsigned prefix = s_prefix;
      const unsigned himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
      for (int64_t i = tid; i < n; i += blockDim.x) {
        const float s = scores[i];
        if (!passes(s)) continue;
        const unsigned k = ord_enc(s);
        if ((k & himask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned need = s_need, d = 255;
        for (;; --d) {                                         // walk the digits from the largest
          if (s_hist[d] >= need || d == 0) break;
          need -= s_hist[d];
        }
        s_prefix = prefix | (d << shift);
        s_need = need;                                         // how many of the keys with this prefix are still wanted
      }
      __syncthreads();
    }
    cut = s_prefix;
    select = true;
  }
  // gather the candidates (any order: the sort key carries the index)
  const unsigned need_eq = select ? s_need : 0u;
  for (int64_t i = tid; i < n; i += blockDim.x) {
    const float s = scores[i];
    if (passes(s) && (!select || ord_enc(s) > cut)) {
      const int pos = atomicAdd(&s_cnt, 1);
      s_key[pos] = s; s_idx[pos] = (int32_t)i;
    }
  }
  __syncthreads();
  if (select) {
    // the first need_eq boxes (in index order) whose key equals the cut: ordered block compaction
    int taken = 0;
    for (int64_t base = 0; base < n && taken < (int)need_eq; base += blockDim.x) {
      const int64_t i = base + tid;
      bool hit = false;
      float s = 0.f;
      if (i < n) { s = scores[i]; hit = passes(s) && ord_enc(s) == cut; }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if ((tid & 31) == 0) s_warp[tid >> 5] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < 32; ++w) { if (w < (tid >> 5)) before += s_warp[w]; total += s_warp[w]; }
      const int rank = taken + before + __popc(bal & ((1u << (tid & 31)) - 1u));
      if (hit && rank < (int)need_eq) {
        const int pos = atomicAdd(&s_cnt, 1);
        s_key[pos] = s; s_idx[pos] = (int32_t)i;
      }
      taken += total;
      __syncthreads();
    }
  }
  __syncthreads();
  const int cnt = s_cnt;
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
  if (tid == 0) { hdr[0] = keep; hdr[1] = min(all, kNmsMaxBoxes); }
}

// ---- 2. suppression mask (upper triangle) -----------------------------------------------------------------------------
template <bool kNormal>
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
    if ((kNormal ? normal_iou(a, s_col + j * 7) : bev_iou(a, s_col + j * 7)) > thresh) bits |= 1ull << j;   // a = the higher-scored box, as the reference
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

static int nms_launch(bool normal, const float* boxes, int64_t box_stride, const float* scores, int64_t num_boxes,
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
  const dim3 mg((unsigned)col_blocks, (unsigned)col_blocks);
  if (normal) nms_mask_kernel<true><<<mg, 64, 0, stream>>>(boxes, box_stride, order, hdr, iou_thresh, col_blocks, mask);
  else nms_mask_kernel<false><<<mg, 64, 0, stream>>>(boxes, box_stride, order, hdr, iou_thresh, col_blocks, mask);
  PCP_LAUNCH_CHECK("nms_mask_kernel");
  nms_scan_kernel<<<1, 32, 0, stream>>>(mask, order, hdr, col_blocks, post_max_size, keep_out, count_out);
  PCP_LAUNCH_CHECK("nms_scan_kernel");
  return 0;
}

extern "C" int pcp_nms_bev(const float* boxes, int64_t box_stride, const float* scores, int64_t num_boxes,
                           int32_t apply_score_thresh, float score_thresh, float iou_thresh, int32_t pre_max_size,
                           int32_t post_max_size, void* scratch, size_t scratch_bytes, int64_t* keep_out,
                           int32_t* count_out, void* stream) {
  return nms_launch(false, boxes, box_stride, scores, num_boxes, apply_score_thresh, score_thresh, iou_thresh, pre_max_size,
                    post_max_size, scratch, scratch_bytes, keep_out, count_out, stream);
}

extern "C" int pcp_nms_normal(const float* boxes, int64_t box_stride, const float* scores, int64_t num_boxes,
                              int32_t apply_score_thresh, float score_thresh, float iou_thresh, int32_t pre_max_size,
                              int32_t post_max_size, void* scratch, size_t scratch_bytes, int64_t* keep_out,
                              int32_t* count_out, void* stream) {
  return nms_launch(true, boxes, box_stride, scores, num_boxes, apply_score_thresh, score_thresh, iou_thresh, pre_max_size,
                    post_max_size, scratch, scratch_bytes, keep_out, count_out, stream);
}
