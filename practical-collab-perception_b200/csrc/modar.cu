// modar.cu - MoDAR point synthesis (reference: v2x_sim_dataset_ego.py:203-232, visualize_collab.py:118-142).
//
// The reference runs, per agent: a points-in-boxes CUDA op, boolean masks, torch.unique, a scatter-mean,
// an indexed add, a D2H copy, a numpy fp64 SE(3), and np.concatenate - ~15 launches and a device sync per agent.
// Here all agents of a frame are two launches (membership over point chunks x agents, rows over box chunks x agents) and
// nothing leaves the device:
//   1. boxes -> shared memory (with cos/sin of -heading evaluated once per box instead of once per pair)
//   2. box membership of every foreground point (first containing box wins, default -1)
//   3. per-box mean flow, summed in ascending point order (the CPU reference's order) by one warp per box
//   4. xyz += scale * mean, SE(3) in fp64, heading += yaw(R) wrapped by atan2(sin, cos), pack the row
#include "common.cuh"

namespace pcp {

constexpr int kMaxBoxes = 1024;  // boxes per agent held in shared memory (reference caps NMS output at 83)

struct BoxS { float cx, cy, cz, hx, hy, hz, cosa, msina; };  // hx/hy/hz are the FULL sizes dx, dy, dz

// pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-36 as nvcc compiles it (sm_100a SASS):
//   local_x = fma(sx, cosa, sy * (-sina));  local_y = fma(sy, cosa, -(sx * (-sina)))
// comparisons in double: |z - cz| > dz / 2.0 rejects;  |local| < d / 2.0 + (double)1e-5f accepts
__device__ __forceinline__ bool point_in_box(float x, float y, float z, const BoxS& b) {
  if ((double)fabsf(__fsub_rn(z, b.cz)) > (double)b.hz / 2.0) return false;
  const float sx = __fsub_rn(x, b.cx), sy = __fsub_rn(y, b.cy);
  const float lx = __fmaf_rn(sx, b.cosa, __fmul_rn(sy, b.msina));
  const float ly = __fmaf_rn(sy, b.cosa, -__fmul_rn(sx, b.msina));
  const double margin = (double)1e-5f;
  return ((double)fabsf(lx) < (double)b.hx / 2.0 + margin) & ((double)fabsf(ly) < (double)b.hy / 2.0 + margin);
}

// 1 + 2. membership of every foreground point: grid (point chunks, agents); the agent's boxes sit in shared memory
__global__ void __launch_bounds__(256)
modar_membership_kernel(const float* __restrict__ boxes_cat, const float* const* __restrict__ box_ptrs,
                        const int32_t* __restrict__ box_off, const float* __restrict__ fg_cat,
                        const float* const* __restrict__ fg_ptrs, const int32_t* __restrict__ fg_off,
                        int32_t* __restrict__ box_idx_out) {
  __shared__ BoxS s_box[kMaxBoxes];
  const int a = blockIdx.y;
  const int b0 = box_off[a], M = box_off[a + 1] - b0;
  const int f0 = fg_off[a], F = fg_off[a + 1] - f0;
  const int tid = threadIdx.x;
  if ((int64_t)blockIdx.x * blockDim.x >= F) return;
  const int i = blockIdx.x * blockDim.x + tid;
  // records either concatenated (one array + offsets) or per agent (a table of device pointers: nothing is copied together)
  const float* boxes = box_ptrs ? box_ptrs[a] : boxes_cat + (int64_t)b0 * 9;
  const float* fg = fg_ptrs ? fg_ptrs[a] : fg_cat + (int64_t)f0 * 13;
  float x = 0.f, y = 0.f, z = 0.f;
  if (i < F) {
    const float* p = fg + (int64_t)i * 13;
    x = p[0]; y = p[1]; z = p[2];
  }
  int hit = -1;
  for (int mb = 0; mb < M; mb += kMaxBoxes) {       // boxes in panels of kMaxBoxes (one panel in practice)
    const int mc = min(kMaxBoxes, M - mb);
    __syncthreads();
    for (int k = tid; k < mc; k += blockDim.x) {
      const float* bx = boxes + (int64_t)(mb + k) * 9;
      BoxS s;
      s.cx = bx[0]; s.cy = bx[1]; s.cz = bx[2]; s.hx = bx[3]; s.hy = bx[4]; s.hz = bx[5];
      const float rz = bx[6];
      s.cosa = cosf(-rz);
      s.msina = -sinf(-rz);
      s_box[k] = s;
    }
    __syncthreads();
    // first containing box wins (roiaware_pool3d_kernel.cu:329-335)
    if (i < F && hit < 0) {
      for (int k = 0; k < mc; ++k)
        if (point_in_box(x, y, z, s_box[k])) { hit = mb + k; break; }
    }
  }
  if (i < F) box_idx_out[f0 + i] = hit;
}

// 3 + 4. one warp per box: grid (box chunks, agents).  The agent's membership array is staged in shared memory in panels, so
// the scan over the foreground points never waits on global memory.  The box's own points are first COLLECTED (their numbers,
// in ascending order, kMemberCap at a time), then their flow vectors are gathered with every lane's loads in flight at once,
// then summed sequentially in point order by lanes 0 / 1 / 2 - the scan never stalls on a gather (it did: one exposed global
// round trip per 32-point step that held a member made this the longest kernel of the exchange, 79 us for 170 boxes).
constexpr int kIdxPanel = 4096;
constexpr int kMemberCap = 192;

__global__ void __launch_bounds__(256)
modar_rows_kernel(const float* __restrict__ boxes_cat, const float* const* __restrict__ box_ptrs,
                  const int32_t* __restrict__ box_off, const float* __restrict__ fg_cat,
                  const float* const* __restrict__ fg_ptrs, const int32_t* __restrict__ fg_off, int has_fg,
                  const double* __restrict__ se3, float scale, float max_sweep_idx_arg,
                  const float* __restrict__ max_sweep_idx_dev, int with_batch_col, float batch_idx,
                  float* __restrict__ rows_out, int64_t out_stride, const int32_t* __restrict__ box_idx) {
  __shared__ int32_t s_idx[kIdxPanel];
  const int a = blockIdx.y;
  const int b0 = box_off[a], M = box_off[a + 1] - b0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  if (blockIdx.x * nwarp >= M) return;
  const bool propagate = has_fg && (scale != 0.f);
  const int f0 = propagate ? fg_off[a] : 0, F = propagate ? fg_off[a + 1] - f0 : 0;
  const float* boxes = box_ptrs ? box_ptrs[a] : boxes_cat + (int64_t)b0 * 9;
  const float* fg = !propagate ? nullptr : (fg_ptrs ? fg_ptrs[a] : fg_cat + (int64_t)f0 * 13);
  const float max_sweep_idx = max_sweep_idx_dev ? __ldg(max_sweep_idx_dev) : max_sweep_idx_arg;
  const double* T = se3 + 12 * a;
  const double yaw_T = atan2(T[4], T[0]);  // rotation_matrix_to_yaw: arctan2(R10, R00), nuscenes_temporal_utils.py:28-29
  const int k = blockIdx.x * nwarp + warp;
  __shared__ int32_t s_mem[8][kMemberCap];            // per warp: point numbers of the box's members, ascending
  __shared__ float s_flow[8][3][kMemberCap];          // per warp: their flow vectors
  float acc = 0.f;                                    // lanes 0 / 1 / 2: running sums of flow x / y / z
  int cnt = 0, nm = 0;
  const unsigned lt = (1u << lane) - 1u;
  auto flush = [&](int n_m) {
    __syncwarp();
    for (int j = lane; j < n_m; j += 32) {
      const float* p = fg + (int64_t)s_mem[warp][j] * 13;
      s_flow[warp][0][j] = p[10]; s_flow[warp][1][j] = p[11]; s_flow[warp][2][j] = p[12];   // flow3 = last three columns (:213)
    }
    __syncwarp();
    if (lane < 3)
      for (int j = 0; j < n_m; ++j) acc = __fadd_rn(acc, s_flow[warp][lane][j]);           // sequential, ascending point order
    __syncwarp();
  };
  for (int p0 = 0; p0 < F; p0 += kIdxPanel) {
    const int pc = min(kIdxPanel, F - p0);
    __syncthreads();
    for (int i = tid; i < pc; i += blockDim.x) s_idx[i] = box_idx[f0 + p0 + i];
    __syncthreads();
    if (k < M) {
      for (int i0 = 0; i0 < pc; i0 += 32) {
        const int i = i0 + lane;
        const bool mine = (i < pc) && (s_idx[i] == k);
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        if (m == 0) continue;
        if (mine) s_mem[warp][nm + __popc(m & lt)] = p0 + i;
        const int c = __popc(m);
        nm += c; cnt += c;
        if (nm > kMemberCap - 32) { flush(nm); nm = 0; }
      }
    }
  }
  if (k < M && nm) flush(nm);
  const float sx = __shfl_sync(0xffffffffu, acc, 0), sy = __shfl_sync(0xffffffffu, acc, 1), sz = __shfl_sync(0xffffffffu, acc, 2);
  if (k >= M) return;
  const float* bx = boxes + (int64_t)k * 9;
  float ox = 0.f, oy = 0.f, oz = 0.f;
  if (cnt > 0) {
    const float c = (float)cnt;
    ox = __fmul_rn(__fdiv_rn(sx, c), scale);   // scatter(reduce='mean') * 2.  (:213)
    oy = __fmul_rn(__fdiv_rn(sy, c), scale);
    oz = __fmul_rn(__fdiv_rn(sz, c), scale);
  }
  if (lane == 0) {
    // modar[unq_box_idx, :3] += boxes_offset (:215); boxes without points are untouched
    const float x = (F > 0) ? __fadd_rn(bx[0], ox) : bx[0];
    const float y = (F > 0) ? __fadd_rn(bx[1], oy) : bx[1];
    const float z = (F > 0) ? __fadd_rn(bx[2], oz) : bx[2];
    // apply_se3_ boxes branch (nuscenes_temporal_utils.py:66-70): fp32 row @ fp64 R^T + t, stored as fp32
    const double xd = x, yd = y, zd = z;
    const float tx = (float)(xd * T[0] + yd * T[1] + zd * T[2] + T[3]);
    const float ty = (float)(xd * T[4] + yd * T[5] + zd * T[6] + T[7]);
    const float tz = (float)(xd * T[8] + yd * T[9] + zd * T[10] + T[11]);
    float yaw = (float)((double)bx[6] + yaw_T);
    yaw = atan2f(sinf(yaw), cosf(yaw));
    float* o = rows_out + (int64_t)(b0 + k) * out_stride;
    if (with_batch_col) *o++ = batch_idx;
    // [x, y, z, 0, 0, dx, dy, dz, heading, score, label, max_sweep_idx, -1]  (v2x_sim_dataset_ego.py:221-226)
    o[0] = tx; o[1] = ty; o[2] = tz; o[3] = 0.f; o[4] = 0.f;
    o[5] = bx[3]; o[6] = bx[4]; o[7] = bx[5]; o[8] = yaw; o[9] = bx[7]; o[10] = bx[8];
    o[11] = max_sweep_idx; o[12] = -1.f;
  }
}

// max of one column of a row-major fp32 matrix (ego_points[:, sweep column].max(), v2x_sim_dataset_ego.py:174), one CTA;
// an empty matrix gives 0 (the host wrapper's convention)
__global__ void __launch_bounds__(1024)
column_max_kernel(const float* __restrict__ rows, int64_t stride, int64_t n, int32_t col, float* __restrict__ out) {
  __shared__ float s_red[32];
  float m = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, __ldg(rows + i * stride + col));
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : -INFINITY;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (threadIdx.x == 0) *out = (n > 0) ? m : 0.f;
  }
}

}  // namespace pcp

using namespace pcp;

static int modar_launch(const float* boxes, const float* const* box_ptrs, const int32_t* box_offsets, const float* foreground,
                        const float* const* fg_ptrs, const int32_t* fg_offsets, const double* se3, int32_t num_agents,
                        int32_t max_boxes_per_agent, int32_t max_fg_per_agent, float scale, float max_sweep_idx,
                        const float* max_sweep_idx_dev, int32_t with_batch_col, float batch_idx, float* rows_out,
                        int64_t out_stride, int32_t* box_idx_out, cudaStream_t stream) {
  PCP_REQUIRE(num_agents >= 0, PCP_E_INVALID, "pcp_modar: num_agents < 0");
  if (num_agents == 0) return 0;
  PCP_REQUIRE((boxes || box_ptrs) && box_offsets && se3 && rows_out, PCP_E_INVALID, "pcp_modar: null argument");
  PCP_REQUIRE(out_stride >= 13 + (with_batch_col ? 1 : 0), PCP_E_INVALID, "pcp_modar: out_stride too small");
  const bool has_fg = foreground != nullptr || fg_ptrs != nullptr;
  PCP_REQUIRE(!has_fg || (fg_offsets && box_idx_out), PCP_E_INVALID,
              "pcp_modar: foreground given without fg_offsets / box_idx_out");
  PCP_REQUIRE(max_boxes_per_agent >= 0 && max_fg_per_agent >= 0 && num_agents <= 65535, PCP_E_INVALID, "pcp_modar: bad sizes");
  if (max_boxes_per_agent == 0) return 0;
  const bool propagate = has_fg && scale != 0.f && max_fg_per_agent > 0;
  if (propagate) {
    const dim3 grid((unsigned)((max_fg_per_agent + 255) / 256), (unsigned)num_agents);
    modar_membership_kernel<<<grid, 256, 0, stream>>>(boxes, box_ptrs, box_offsets, foreground, fg_ptrs, fg_offsets, box_idx_out);
    PCP_LAUNCH_CHECK("modar_membership_kernel");
  }
  const dim3 grid((unsigned)((max_boxes_per_agent + 7) / 8), (unsigned)num_agents);
  modar_rows_kernel<<<grid, 256, 0, stream>>>(boxes, box_ptrs, box_offsets, foreground, fg_ptrs, fg_offsets, propagate ? 1 : 0, se3,
                                              scale, max_sweep_idx, max_sweep_idx_dev, with_batch_col, batch_idx, rows_out,
                                              out_stride, box_idx_out);
  PCP_LAUNCH_CHECK("modar_rows_kernel");
  return 0;
}

extern "C" int pcp_modar(const float* boxes, const int32_t* box_offsets, const float* foreground,
                         const int32_t* fg_offsets, const double* se3, int32_t num_agents, int32_t max_boxes_per_agent,
                         int32_t max_fg_per_agent, float scale,
                         float max_sweep_idx, int32_t with_batch_col, float batch_idx, float* rows_out,
                         int64_t out_stride, int32_t* box_idx_out, void* stream_) {
  return modar_launch(boxes, nullptr, box_offsets, foreground, nullptr, fg_offsets, se3, num_agents, max_boxes_per_agent,
                      max_fg_per_agent, scale, max_sweep_idx, nullptr, with_batch_col, batch_idx, rows_out, out_stride,
                      box_idx_out, static_cast<cudaStream_t>(stream_));
}

extern "C" int pcp_modar_agents(const float* const* box_ptrs, const int32_t* box_offsets, const float* const* fg_ptrs,
                                const int32_t* fg_offsets, const double* se3, int32_t num_agents, int32_t max_boxes_per_agent,
                                int32_t max_fg_per_agent, float scale, const float* max_sweep_idx_dev, int32_t with_batch_col,
                                float batch_idx, float* rows_out, int64_t out_stride, int32_t* box_idx_out, void* stream_) {
  PCP_REQUIRE(num_agents <= 0 || (box_ptrs && max_sweep_idx_dev), PCP_E_INVALID, "pcp_modar_agents: null argument");
  return modar_launch(nullptr, box_ptrs, box_offsets, nullptr, fg_ptrs, fg_offsets, se3, num_agents, max_boxes_per_agent,
                      max_fg_per_agent, scale, 0.f, max_sweep_idx_dev, with_batch_col, batch_idx, rows_out, out_stride,
                      box_idx_out, static_cast<cudaStream_t>(stream_));
}

extern "C" int pcp_column_max(const float* rows, int64_t row_stride, int64_t n_rows, int32_t column, float* max_out,
                              void* stream_) {
  PCP_REQUIRE(max_out && n_rows >= 0 && column >= 0 && column < row_stride && (n_rows == 0 || rows), PCP_E_INVALID,
              "pcp_column_max: bad argument");
  column_max_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream_)>>>(rows, row_stride, n_rows, column, max_out);
  PCP_LAUNCH_CHECK("column_max_kernel");
  return 0;
}
