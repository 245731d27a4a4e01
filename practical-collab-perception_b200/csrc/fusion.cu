// fusion.cu - early-fusion input assembly on the GPU
// (reference: pcdet/datasets/v2x_sim/v2x_sim_dataset_ego_early.py:85-92 - every agent's sweep stack mapped into the ego
//  frame with apply_se3_ (nuscenes_temporal_utils.py:62-63: xyz @ R.T + t in fp64, stored back to fp32) and concatenated
//  after the ego points; pcdet/datasets/processor/data_processor.py:78-84 + common_utils.py:64-68 - the range mask
//  x,y,z in [min, max); pcdet/datasets/dataset.py:224-229 - the frame-index column collate_batch prepends).
// The reference then shuffles the rows with np.random.permutation (data_processor.py:95-104); every consumer on this path
// is order-independent, so the rows are kept in input order (deterministic).
//
// Three launches: per-block kept counts, one-block scan of the block counts, then the transform is evaluated again and the
// surviving rows are written at their final position (order-preserving compaction without a row-sized temporary).
#include "internal.cuh"

namespace pcp {

constexpr int kFuseThreads = 256;
constexpr int kFuseItems = 4;
constexpr int kFuseBlock = kFuseThreads * kFuseItems;   // rows per CTA

struct FuseArgs {
  const float* const* ptrs;  // per-agent clouds (device array of device pointers), or NULL: one concatenated block at `points`
  const float* points;
  int64_t in_stride;
  int32_t n_cols;            // columns of an input row (x, y, z, ...), all copied
  const int32_t* agent_off;  // [num_agents + 1]
  const double* se3;         // (num_agents, 12) rows 0..2 of target_se3_agent, row-major
  int32_t num_agents;
  float lo[3], hi[3];
  int32_t apply_mask;
};

__device__ __forceinline__ int agent_of(const int32_t* __restrict__ off, int na, int64_t i) {
  int a = 0;
  while (a + 1 < na && i >= off[a + 1]) ++a;           // a handful of agents: linear search
  return a;
}

// transformed xyz of row i (fp64 products and sums in the reference's matmul order, one rounding to fp32) and its mask
__device__ __forceinline__ const float* fuse_src(const FuseArgs& A, int64_t i, int a) {
  return A.ptrs ? A.ptrs[a] + (i - A.agent_off[a]) * A.in_stride : A.points + i * A.in_stride;
}

__device__ __forceinline__ bool fuse_row(const FuseArgs& A, int64_t i, float& x, float& y, float& z) {
  const int a = agent_of(A.agent_off, A.num_agents, i);
  const float* row = fuse_src(A, i, a);
  const double px = (double)__ldg(row), py = (double)__ldg(row + 1), pz = (double)__ldg(row + 2);
  const double* T = A.se3 + 12 * a;
  // numpy: (xyz @ R.T)[k] = x R[k,0] + y R[k,1] + z R[k,2] accumulated left to right, then + t[k]
  x = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, T[0]), __dmul_rn(py, T[1])), __dmul_rn(pz, T[2])), T[3]);
  y = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, T[4]), __dmul_rn(py, T[5])), __dmul_rn(pz, T[6])), T[7]);
  z = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, T[8]), __dmul_rn(py, T[9])), __dmul_rn(pz, T[10])), T[11]);
  if (!A.apply_mask) return true;
  return x >= A.lo[0] && x < A.hi[0] && y >= A.lo[1] && y < A.hi[1] && z >= A.lo[2] && z < A.hi[2];
}

__global__ void __launch_bounds__(kFuseThreads)
fuse_count_kernel(const FuseArgs A, int64_t n, int32_t* __restrict__ block_sum) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
#pragma unroll
  for (int u = 0; u < kFuseItems; ++u) {
    const int64_t i = (int64_t)blockIdx.x * kFuseBlock + (int64_t)threadIdx.x * kFuseItems + u;
    float x, y, z;
    if (i < n && fuse_row(A, i, x, y, z)) ++c;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) block_sum[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(1024)
fuse_scan_kernel(int32_t* __restrict__ block_sum, int32_t num_blocks, int32_t* __restrict__ count_out) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < num_blocks; b0 += 1024) {
    const int i = b0 + tid;
    const int v = (i < num_blocks) ? block_sum[i] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wex = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { if (w < warp) wex += s_warp[w]; tot += s_warp[w]; }
    const int carry = s_carry;
    if (i < num_blocks) block_sum[i] = carry + wex + incl - v;
    __syncthreads();
    if (tid == 0) s_carry = carry + tot;
    __syncthreads();
  }
  if (tid == 0) *count_out = s_carry;
}

__global__ void __launch_bounds__(kFuseThreads)
fuse_write_kernel(const FuseArgs A, int64_t n, const int32_t* __restrict__ block_excl, int32_t with_batch_col, float batch_idx,
                  float* __restrict__ out, int64_t out_stride) {
  __shared__ int s_warp[kFuseThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float x[kFuseItems], y[kFuseItems], z[kFuseItems];
  bool keep[kFuseItems];
  int c = 0;
#pragma unroll
  for (int u = 0; u < kFuseItems; ++u) {
    const int64_t i = (int64_t)blockIdx.x * kFuseBlock + (int64_t)tid * kFuseItems + u;
    keep[u] = (i < n) && fuse_row(A, i, x[u], y[u], z[u]);
    c += keep[u] ? 1 : 0;
  }
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int pos = block_excl[blockIdx.x] + incl - c;
  for (int w = 0; w < warp; ++w) pos += s_warp[w];
#pragma unroll
  for (int u = 0; u < kFuseItems; ++u) {
    if (!keep[u]) continue;
    const int64_t i = (int64_t)blockIdx.x * kFuseBlock + (int64_t)tid * kFuseItems + u;
    const float* row = fuse_src(A, i, agent_of(A.agent_off, A.num_agents, i));
    float* dst = out + (int64_t)pos * out_stride;
    if (with_batch_col) *dst++ = batch_idx;
    dst[0] = x[u]; dst[1] = y[u]; dst[2] = z[u];
    for (int k = 3; k < A.n_cols; ++k) dst[k] = __ldg(row + k);
    ++pos;
  }
}

}  // namespace pcp

namespace pcp {

// ------------------------------------------------------------------------------------------------
// device half of the points loader (pcdet/models/__init__.py:23-34 load_data_to_gpu for batch_dict['points'], after the
// collate step of pcdet/datasets/dataset.py:224-229): the host ships only the per-point columns a consumer reads, frames
// back to back, plus one row offset per frame; this kernel rebuilds the (N, 1 + C) rows collate_batch would have produced -
// frame index in column 0, the shipped columns at their original places, zeros elsewhere.
// ------------------------------------------------------------------------------------------------
struct UnpackCols { int32_t k; int32_t col[PCP_UNPACK_MAX_COLS]; };

__global__ void __launch_bounds__(256)
unpack_points_kernel(const float* __restrict__ packed, int64_t n, const int32_t* __restrict__ frame_off, int32_t frames,
                     const UnpackCols cols, int32_t n_out_cols, float* __restrict__ out, int64_t out_stride) {
  extern __shared__ int32_t s_off[];                    // [frames + 1]
  for (int i = threadIdx.x; i <= frames; i += blockDim.x) s_off[i] = __ldg(frame_off + i);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = frames;                              // frame f: s_off[f] <= i < s_off[f + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (i >= s_off[mid]) lo = mid; else hi = mid;
  }
  float v[PCP_UNPACK_MAX_COLS];
#pragma unroll
  for (int j = 0; j < PCP_UNPACK_MAX_COLS; ++j)
    if (j < cols.k) v[j] = __ldg(packed + i * cols.k + j);
  float* row = out + i * out_stride;
  if (n_out_cols == 8 && (out_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
    float r[8] = {(float)lo, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < PCP_UNPACK_MAX_COLS; ++j)
      if (j < cols.k) {
#pragma unroll
        for (int c = 1; c < 8; ++c)
          if (cols.col[j] + 1 == c) r[c] = v[j];
      }
    reinterpret_cast<float4*>(row)[0] = make_float4(r[0], r[1], r[2], r[3]);
    reinterpret_cast<float4*>(row)[1] = make_float4(r[4], r[5], r[6], r[7]);
  } else {
    row[0] = (float)lo;
    for (int c = 1; c < n_out_cols; ++c) row[c] = 0.f;
    for (int j = 0; j < cols.k; ++j) row[1 + cols.col[j]] = v[j];
  }
}

}  // namespace pcp

extern "C" int pcp_unpack_points(const float* packed, int32_t n_packed_cols, const int32_t* col_index_host, int64_t n_points,
                                 const int32_t* frame_offsets, int32_t num_frames, int32_t n_point_cols, float* rows_out,
                                 int64_t out_stride, void* stream_) {
  using namespace pcp;       // (defined ahead of the file-wide using-directive below)
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(n_points >= 0 && num_frames > 0 && num_frames <= 8192, PCP_E_INVALID, "pcp_unpack_points: bad n_points / num_frames");
  PCP_REQUIRE(n_packed_cols > 0 && n_packed_cols <= PCP_UNPACK_MAX_COLS && n_point_cols >= n_packed_cols, PCP_E_INVALID,
              "pcp_unpack_points: 1 <= shipped columns <= %d <= point columns", PCP_UNPACK_MAX_COLS);
  PCP_REQUIRE(out_stride >= 1 + n_point_cols, PCP_E_INVALID, "pcp_unpack_points: out_stride < 1 + n_point_cols");
  if (n_points == 0) return 0;
  PCP_REQUIRE(packed && col_index_host && frame_offsets && rows_out, PCP_E_INVALID, "pcp_unpack_points: null argument");
  UnpackCols cols{};
  cols.k = n_packed_cols;
  for (int j = 0; j < n_packed_cols; ++j) {
    PCP_REQUIRE(col_index_host[j] >= 0 && col_index_host[j] < n_point_cols, PCP_E_INVALID, "pcp_unpack_points: column index out of range");
    cols.col[j] = col_index_host[j];
  }
  const unsigned blocks = (unsigned)((n_points + 255) / 256);
  unpack_points_kernel<<<blocks, 256, sizeof(int32_t) * (size_t)(num_frames + 1), stream>>>(
      packed, n_points, frame_offsets, num_frames, cols, 1 + n_point_cols, rows_out, out_stride);
  PCP_LAUNCH_CHECK("unpack_points_kernel");
  return 0;
}

namespace pcp {

// ------------------------------------------------------------------------------------------------
// producer side of the exchange: foreground selection (reference: pcdet/models/bev_layers/hunter_jr.py:377-397)
//   prob = sigmoid(cls_logit);  send = prob[:, 0] < threshold;  row = [point columns 1.. | prob (3) | flow (3)];
//   per sample b: the rows with int(points[:, 0]) == b, in input order.
// Stable partition by sample: per-CTA kept counts per sample (sample-major table), one-block exclusive scan, write.
// ------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 256;
constexpr int kSelItems = 4;
constexpr int kSelBlock = kSelThreads * kSelItems;

struct SelArgs {
  const float* points; int64_t p_stride; int32_t n_pt_cols;      // columns 1 .. n_pt_cols are sent (column 0 = sample index)
  const float* logit; int64_t l_stride;
  const float* flow; int64_t f_stride;
  int32_t frames; float threshold;
};

__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// sample index of row i if it is sent, else -1
__device__ __forceinline__ int sel_frame(const SelArgs& A, int64_t i) {
  const float p0 = sigmoid_rn(__ldg(A.logit + i * A.l_stride));
  if (!(p0 < A.threshold)) return -1;
  const float bf = __ldg(A.points + i * A.p_stride);
  if (!(bf > -1.f) || !(bf < (float)A.frames)) return -1;
  return (int)bf;                                              // .long(): truncation toward zero
}

__global__ void __launch_bounds__(kSelThreads)
sel_count_kernel(const SelArgs A, int64_t n, int32_t nblk, int32_t* __restrict__ table) {
  extern __shared__ int32_t s_cnt[];                           // [frames]
  for (int f = threadIdx.x; f < A.frames; f += kSelThreads) s_cnt[f] = 0;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kSelItems; ++u) {
    const int64_t i = (int64_t)blockIdx.x * kSelBlock + (int64_t)threadIdx.x * kSelItems + u;
    if (i < n) {
      const int f = sel_frame(A, i);
      if (f >= 0) atomicAdd(&s_cnt[f], 1);
    }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < A.frames; f += kSelThreads) table[(int64_t)f * nblk + blockIdx.x] = s_cnt[f];
}

// frame_offsets[f] = first output row of sample f (the scanned table entry of its first CTA); [frames] = total
__global__ void sel_offsets_kernel(const int32_t* __restrict__ table, const int32_t* __restrict__ total, int32_t nblk,
                                   int32_t frames, int32_t* __restrict__ frame_offsets) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < frames) frame_offsets[f] = table[(int64_t)f * nblk];
  if (f == frames) frame_offsets[frames] = *total;
}

__global__ void __launch_bounds__(kSelThreads)
sel_write_kernel(const SelArgs A, int64_t n, int32_t nblk, const int32_t* __restrict__ table, float* __restrict__ out,
                 int64_t out_stride) {
  __shared__ int s_warp[kSelThreads / 32];
  __shared__ int s_lo, s_hi;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_lo = 0x7fffffff; s_hi = -1; }
  __syncthreads();
  int fr[kSelItems];
  int lo = 0x7fffffff, hi = -1;
#pragma unroll
  for (int u = 0; u < kSelItems; ++u) {
    const int64_t i = (int64_t)blockIdx.x * kSelBlock + (int64_t)tid * kSelItems + u;
    fr[u] = (i < n) ? sel_frame(A, i) : -1;
    if (fr[u] >= 0) { lo = min(lo, fr[u]); hi = max(hi, fr[u]); }
  }
  lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
  if (lane == 0 && hi >= 0) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
  __syncthreads();
  const int f_lo = s_lo, f_hi = s_hi;
  // one pass per sample present in this CTA (rows arrive grouped by sample: one pass, two at a sample boundary)
  for (int f = f_lo; f <= f_hi; ++f) {
    int c = 0;
#pragma unroll
    for (int u = 0; u < kSelItems; ++u) c += (fr[u] == f) ? 1 : 0;
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    __syncthreads();                       // s_warp of the previous pass has been read
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int pos = table[(int64_t)f * nblk + blockIdx.x] + incl - c;
    for (int w = 0; w < warp; ++w) pos += s_warp[w];
#pragma unroll
    for (int u = 0; u < kSelItems; ++u) {
      if (fr[u] != f) continue;
      const int64_t i = (int64_t)blockIdx.x * kSelBlock + (int64_t)tid * kSelItems + u;
      float* dst = out + (int64_t)pos * out_stride;
      const float* prow = A.points + i * A.p_stride + 1;
      for (int k = 0; k < A.n_pt_cols; ++k) dst[k] = __ldg(prow + k);
      dst += A.n_pt_cols;
      const float* lrow = A.logit + i * A.l_stride;
      dst[0] = sigmoid_rn(__ldg(lrow)); dst[1] = sigmoid_rn(__ldg(lrow + 1)); dst[2] = sigmoid_rn(__ldg(lrow + 2));
      const float* frow = A.flow + i * A.f_stride;
      dst[3] = __ldg(frow); dst[4] = __ldg(frow + 1); dst[5] = __ldg(frow + 2);
      ++pos;
    }
  }
}

}  // namespace pcp

using namespace pcp;

extern "C" size_t pcp_select_scratch_bytes(int64_t n_points, int32_t num_frames) {
  if (n_points < 0 || num_frames <= 0) return 0;
  return sizeof(int32_t) * ((size_t)((n_points + kSelBlock - 1) / kSelBlock) * (size_t)num_frames + 2);
}

extern "C" int pcp_select_foreground(const float* points, int64_t point_stride, int32_t n_point_cols, const float* cls_logit,
                                     int64_t logit_stride, const float* flow3d, int64_t flow_stride, int64_t n_points,
                                     int32_t num_frames, float threshold, int32_t* scratch, float* rows_out,
                                     int64_t out_stride, int32_t* frame_offsets_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(scratch && frame_offsets_out, PCP_E_INVALID, "pcp_select_foreground: null argument");
  PCP_REQUIRE(n_points >= 0 && n_points < (1ll << 31) - kSelBlock, PCP_E_INVALID, "pcp_select_foreground: n_points out of range");
  PCP_REQUIRE(num_frames >= 1 && num_frames <= 4096, PCP_E_INVALID, "pcp_select_foreground: 1 <= num_frames <= 4096");
  PCP_REQUIRE(n_points == 0 || (points && cls_logit && flow3d && rows_out), PCP_E_INVALID, "pcp_select_foreground: null input");
  PCP_REQUIRE(n_point_cols >= 1 && point_stride >= 1 + n_point_cols && logit_stride >= 3 && flow_stride >= 3 &&
                  out_stride >= n_point_cols + 6, PCP_E_INVALID, "pcp_select_foreground: bad column counts / strides");
  const int nblk = (int)((n_points + kSelBlock - 1) / kSelBlock);
  if (nblk == 0) {
    PCP_CUDA(cudaMemsetAsync(frame_offsets_out, 0, sizeof(int32_t) * (size_t)(num_frames + 1), stream));
    return 0;
  }
  SelArgs A{points, point_stride, n_point_cols, cls_logit, logit_stride, flow3d, flow_stride, num_frames, threshold};
  int32_t* total = scratch + (size_t)nblk * num_frames;
  sel_count_kernel<<<nblk, kSelThreads, sizeof(int32_t) * (size_t)num_frames, stream>>>(A, n_points, nblk, scratch);
  PCP_LAUNCH_CHECK("sel_count_kernel");
  fuse_scan_kernel<<<1, 1024, 0, stream>>>(scratch, nblk * num_frames, total);
  PCP_LAUNCH_CHECK("fuse_scan_kernel");
  sel_offsets_kernel<<<(num_frames + 256) / 256, 256, 0, stream>>>(scratch, total, nblk, num_frames, frame_offsets_out);
  PCP_LAUNCH_CHECK("sel_offsets_kernel");
  sel_write_kernel<<<nblk, kSelThreads, 0, stream>>>(A, n_points, nblk, scratch, rows_out, out_stride);
  PCP_LAUNCH_CHECK("sel_write_kernel");
  return 0;
}

extern "C" size_t pcp_fuse_scratch_bytes(int64_t n_points) {
  if (n_points < 0) return 0;
  return sizeof(int32_t) * (size_t)((n_points + kFuseBlock - 1) / kFuseBlock + 1);
}

static int fuse_launch(const float* const* cloud_ptrs, const float* points, int64_t in_stride, int32_t n_cols, int64_t n_points,
                       const int32_t* agent_offsets, const double* se3, int32_t num_agents, const float* range6_host,
                       int32_t with_batch_col, float batch_idx, int32_t* scratch, float* rows_out, int64_t out_stride,
                       int32_t* count_out, cudaStream_t stream) {
  PCP_REQUIRE(count_out && scratch, PCP_E_INVALID, "pcp_fuse_agent_points: null argument");
  PCP_REQUIRE(n_points >= 0 && n_points < (1ll << 31) - kFuseBlock, PCP_E_INVALID, "pcp_fuse_agent_points: n_points out of range");
  PCP_REQUIRE(n_points == 0 || ((points || cloud_ptrs) && rows_out && agent_offsets && se3), PCP_E_INVALID,
              "pcp_fuse_agent_points: null input");
  PCP_REQUIRE(n_cols >= 3 && in_stride >= n_cols && out_stride >= n_cols + (with_batch_col ? 1 : 0), PCP_E_INVALID,
              "pcp_fuse_agent_points: bad column counts / strides");
  PCP_REQUIRE(num_agents >= 1 && num_agents <= 64, PCP_E_INVALID, "pcp_fuse_agent_points: 1 <= num_agents <= 64");
  FuseArgs A{};
  A.ptrs = cloud_ptrs; A.points = points; A.in_stride = in_stride; A.n_cols = n_cols; A.agent_off = agent_offsets; A.se3 = se3; A.num_agents = num_agents;
  A.apply_mask = range6_host != nullptr;
  if (range6_host)
    for (int k = 0; k < 3; ++k) { A.lo[k] = range6_host[k]; A.hi[k] = range6_host[3 + k]; }
  const int nblk = (int)((n_points + kFuseBlock - 1) / kFuseBlock);
  if (nblk == 0) {
    PCP_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), stream));
    return 0;
  }
  fuse_count_kernel<<<nblk, kFuseThreads, 0, stream>>>(A, n_points, scratch);
  PCP_LAUNCH_CHECK("fuse_count_kernel");
  fuse_scan_kernel<<<1, 1024, 0, stream>>>(scratch, nblk, count_out);
  PCP_LAUNCH_CHECK("fuse_scan_kernel");
  fuse_write_kernel<<<nblk, kFuseThreads, 0, stream>>>(A, n_points, scratch, with_batch_col, batch_idx, rows_out, out_stride);
  PCP_LAUNCH_CHECK("fuse_write_kernel");
  return 0;
}

extern "C" int pcp_fuse_agent_points(const float* points, int64_t in_stride, int32_t n_cols, int64_t n_points,
                                     const int32_t* agent_offsets, const double* se3, int32_t num_agents,
                                     const float* range6_host, int32_t with_batch_col, float batch_idx, int32_t* scratch,
                                     float* rows_out, int64_t out_stride, int32_t* count_out, void* stream_) {
  return fuse_launch(nullptr, points, in_stride, n_cols, n_points, agent_offsets, se3, num_agents, range6_host, with_batch_col,
                     batch_idx, scratch, rows_out, out_stride, count_out, static_cast<cudaStream_t>(stream_));
}

extern "C" int pcp_fuse_agent_clouds(const float* const* cloud_ptrs, int64_t in_stride, int32_t n_cols, int64_t n_points,
                                     const int32_t* agent_offsets, const double* se3, int32_t num_agents,
                                     const float* range6_host, int32_t with_batch_col, float batch_idx, int32_t* scratch,
                                     float* rows_out, int64_t out_stride, int32_t* count_out, void* stream_) {
  PCP_REQUIRE(n_points == 0 || cloud_ptrs, PCP_E_INVALID, "pcp_fuse_agent_clouds: null cloud_ptrs");
  return fuse_launch(cloud_ptrs, nullptr, in_stride, n_cols, n_points, agent_offsets, se3, num_agents, range6_host, with_batch_col,
                     batch_idx, scratch, rows_out, out_stride, count_out, static_cast<cudaStream_t>(stream_));
}
