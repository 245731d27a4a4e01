// voxel3d.cu - dynamic 3-D voxelisation with per-voxel feature means
// (reference: pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:42-79, the VFE of the SECOND configs
//  tools/cfgs/v2x_sim_models/v2x_second_{car,rsu,ego}.yaml).
//
// The reference keys every point with b*nx*ny*nz + cx*ny*nz + cy*nz + cz and sorts (torch.unique).  The 3-D key space
// (1024 x 1024 x 40 cells per frame) is too sparse for a dense histogram, but its ORDER is "pillar (b, cx, cy) first, then
// cz", and a pillar has at most nz <= 64 voxels.  So the pillar stage of voxelize.cu is reused unchanged (dense 2-D
// histogram, cell scan, counting-sort placement, ascending rows inside a pillar) and the third axis costs one 64-bit
// occupancy mask per pillar:
//   voxel_mask_kernel  : per pillar, OR of (1 << cz) over its points, popcount = its number of voxels
//   voxel scan (2 launches): exclusive prefix of the popcounts = first voxel rank of every pillar
//   voxel_mean_kernel  : one warp per pillar, lane v owns the pillar's v-th set bit: sequential fp32 sum of the point
//                        features in ascending row order / count (== the CPU scatter_mean, bit for bit), voxel_coords
#include "internal.cuh"

namespace pcp {

constexpr int kQPts = 4;

// quantise + cull on all three axes (dynamic_mean_vfe.py:56-58), keyed by PILLAR
template <bool kVec4>
__global__ void __launch_bounds__(256)
quantise_count3d_kernel(const float* __restrict__ points, int64_t stride, int64_t n, int32_t frames, pcp_grid g,
                        float min_z, float voxel_z, int32_t nz, int32_t* __restrict__ cell, int32_t* __restrict__ key,
                        int32_t* __restrict__ within, int32_t* __restrict__ hdr) {
  const int64_t base = (int64_t)blockIdx.x * (256 * kQPts) + threadIdx.x;
  int32_t k[kQPts], w[kQPts];
#pragma unroll
  for (int u = 0; u < kQPts; ++u) {
    const int64_t i = base + u * 256;
    k[u] = -1; w[u] = 0;
    if (i < n) {
      const float* row = points + i * stride;
      float bf, x, y, z;
      if (kVec4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row));
        bf = v.x; x = v.y; y = v.z; z = v.w;
      } else {
        bf = __ldg(row); x = __ldg(row + 1); y = __ldg(row + 2); z = __ldg(row + 3);
      }
      const float qx = quantise(x, g.range_min_x, g.voxel_x);
      const float qy = quantise(y, g.range_min_y, g.voxel_y);
      const float qz = quantise(z, min_z, voxel_z);
      const bool keep = (qx >= 0.f) && (qx < (float)g.nx) && (qy >= 0.f) && (qy < (float)g.ny) && (qz >= 0.f) && (qz < (float)nz);
      if (keep) {
        if (!(bf > -1.f) || !(bf < (float)frames)) atomicAdd(&hdr[PCP_COUNT_BAD_FRAME], 1);
        else k[u] = (int32_t)bf * (g.nx * g.ny) + (int32_t)qx * g.ny + (int32_t)qy;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kQPts; ++u)
    if (k[u] >= 0) w[u] = atomicAdd(&cell[k[u]], 1);
#pragma unroll
  for (int u = 0; u < kQPts; ++u) {
    const int64_t i = base + u * 256;
    if (i < n) { key[i] = k[u]; within[i] = w[u]; }
  }
}

// per pillar: occupancy mask over cz; 8 lanes per pillar
__global__ void __launch_bounds__(256)
voxel_mask_kernel(const float* __restrict__ points, int64_t stride, const int32_t* __restrict__ hdr,
                  const int32_t* __restrict__ seg_off, const int32_t* __restrict__ sorted_idx, float min_z, float voxel_z,
                  unsigned long long* __restrict__ mask, int32_t* __restrict__ nvox) {
  const int P = hdr[PCP_COUNT_PILLARS];
  const int sub = threadIdx.x & 7;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  unsigned long long m = 0;
  if (r < P) {
    const int off = seg_off[r], cnt = seg_off[r + 1] - off;
    for (int j = sub; j < cnt; j += 8) {
      const float z = __ldg(points + (int64_t)sorted_idx[off + j] * stride + 3);
      m |= 1ull << (int)quantise(z, min_z, voxel_z);
    }
  }
#pragma unroll
  for (int d = 4; d > 0; d >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, d);
  if (r < P && sub == 0) { mask[r] = m; nvox[r] = __popcll(m); }
}

// exclusive scan of nvox[0, P) in two launches: per-block local prefixes + block sums, then one block scans the sums
constexpr int kVScanItems = 8;
constexpr int kVScanBlock = 256 * kVScanItems;   // 2048

__global__ void __launch_bounds__(256)
voxel_scan_local_kernel(const int32_t* __restrict__ hdr, const int32_t* __restrict__ nvox, int32_t* __restrict__ local,
                        int32_t* __restrict__ block_sum) {
  __shared__ int s_warp[8];
  const int P = hdr[PCP_COUNT_PILLARS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * kVScanBlock + (int64_t)tid * kVScanItems;
  int v[kVScanItems], sum = 0;
#pragma unroll
  for (int j = 0; j < kVScanItems; ++j) { v[j] = (base + j < P) ? nvox[base + j] : 0; sum += v[j]; }
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int wex = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { if (w < warp) wex += s_warp[w]; tot += s_warp[w]; }
  int run = wex + incl - sum;
#pragma unroll
  for (int j = 0; j < kVScanItems; ++j) {
    if (base + j < P) local[base + j] = run;
    run += v[j];
  }
  if (tid == 0) block_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024)
voxel_scan_blocks_kernel(int32_t* __restrict__ hdr, int32_t* __restrict__ block_sum, int32_t num_blocks, int32_t* __restrict__ counts_out) {
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
    if (i < num_blocks) block_sum[i] = carry + wex + incl - v;     // exclusive prefix of the block sums
    __syncthreads();
    if (tid == 0) s_carry = carry + tot;
    __syncthreads();
  }
  if (tid == 0) {
    hdr[PCP_COUNT_VOXELS] = s_carry;
    if (counts_out) counts_out[PCP_COUNT_VOXELS] = s_carry;
  }
}

// one warp per pillar; lane v (and v + 32) owns the pillar's v-th voxel in ascending cz
template <int kMaxC>
__global__ void __launch_bounds__(256)
voxel_mean_kernel(const float* __restrict__ points, int64_t stride, int channels, const int32_t* __restrict__ hdr,
                  const int32_t* __restrict__ seg_off, const int32_t* __restrict__ sorted_idx,
                  const unsigned long long* __restrict__ mask, const int32_t* __restrict__ local,
                  const int32_t* __restrict__ block_excl, const int32_t* __restrict__ voxel_coords2d, float min_z, float voxel_z,
                  float* __restrict__ vf_out, int32_t* __restrict__ vc_out, int32_t* __restrict__ point_voxel) {
  const int P = hdr[PCP_COUNT_PILLARS];
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= P) return;
  const int off = seg_off[r], cnt = seg_off[r + 1] - off;
  const unsigned long long m = mask[r];
  const int nv = __popcll(m);
  const int vbase = block_excl[r / kVScanBlock] + local[r];
  const int4 c2 = __ldg(reinterpret_cast<const int4*>(voxel_coords2d) + r);     // (frame, 0, cy, cx)
  for (int v0 = 0; v0 < nv; v0 += 32) {
    const int v = v0 + lane;
    // cz of this lane's voxel: position of the v-th set bit
    int cz = -1;
    if (v < nv) {
      unsigned long long t = m;
      for (int s = 0; s < v; ++s) t &= t - 1;
      cz = __ffsll((long long)t) - 1;
    }
    float acc[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) acc[c] = 0.f;
    int n_in = 0;
    for (int j = 0; j < cnt; ++j) {
      const int idx = sorted_idx[off + j];                                       // same address in every lane: one broadcast load
      const float* row = points + (int64_t)idx * stride;
      const int pz = (int)quantise(__ldg(row + 3), min_z, voxel_z);
      if (pz == cz) {
        ++n_in;
#pragma unroll
        for (int c = 0; c < kMaxC; ++c)
          if (c < channels) acc[c] = __fadd_rn(acc[c], __ldg(row + 1 + c));
        if (point_voxel) point_voxel[idx] = vbase + v;
      }
    }
    if (v < nv) {
      const float d = (float)max(n_in, 1);
      float* dst = vf_out + (int64_t)(vbase + v) * channels;
#pragma unroll
      for (int c = 0; c < kMaxC; ++c)
        if (c < channels) dst[c] = __fdiv_rn(acc[c], d);
      // (b, cz, cy, cx): dynamic_mean_vfe.py:71-76 after the [0, 3, 2, 1] reorder
      *reinterpret_cast<int4*>(vc_out + 4 * (int64_t)(vbase + v)) = make_int4(c2.x, cz, c2.z, c2.w);
    }
  }
}

}  // namespace pcp

using namespace pcp;

extern "C" size_t pcp_voxel3d_scratch_bytes(int64_t n_points, int32_t max_frames, int32_t nx, int32_t ny) {
  if (n_points < 0 || max_frames <= 0 || nx <= 0 || ny <= 0) return 0;
  const WsLayout L = ws_layout(n_points, max_frames, nx, ny);
  const size_t cap = (size_t)L.cap + 1;
  // mask u64[cap] | nvox i32[cap] | local i32[cap] | block sums i32[cap / 2048 + 1] | pillar coords i32[4 cap]
  return align_up(8 * cap, 256) + 2 * align_up(4 * cap, 256) + align_up(4 * (cap / kVScanBlock + 2), 256) + align_up(16 * cap, 256);
}

extern "C" int pcp_voxelize3d_mean(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                                   const pcp_grid* grid, float range_min_z, float voxel_z, int32_t nz, int32_t channels,
                                   void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                                   float* voxel_features_out, int32_t* voxel_coords_out, int32_t* point_voxel_out,
                                   int64_t voxel_capacity, int32_t* counts_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(grid && workspace && scratch && voxel_features_out && voxel_coords_out && counts_out, PCP_E_INVALID,
              "pcp_voxelize3d_mean: null argument");
  PCP_REQUIRE(n_points >= 0 && n_points < (1ll << 29), PCP_E_INVALID, "pcp_voxelize3d_mean: n_points out of range (< 2^29)");
  PCP_REQUIRE(n_points == 0 || points, PCP_E_INVALID, "pcp_voxelize3d_mean: null points");
  PCP_REQUIRE(channels >= 3 && channels <= 16 && row_stride >= 1 + channels, PCP_E_UNSUPPORTED,
              "pcp_voxelize3d_mean: 3 <= channels <= 16 and row_stride >= 1 + channels");
  PCP_REQUIRE(nz > 0 && nz <= 64, PCP_E_UNSUPPORTED, "pcp_voxelize3d_mean: nz must be in [1, 64] (one occupancy word per pillar)");
  PCP_REQUIRE(max_frames > 0 && grid->nx > 0 && grid->ny > 0 && grid->nx <= 65535 && grid->ny <= 65535, PCP_E_INVALID,
              "pcp_voxelize3d_mean: bad grid");
  PCP_REQUIRE((int64_t)max_frames * grid->nx * grid->ny * nz < (1ll << 31), PCP_E_UNSUPPORTED,
              "pcp_voxelize3d_mean: frames*nx*ny*nz must fit int32 (the reference's merge_coords is int32 too)");
  const WsLayout L = ws_layout(n_points, max_frames, grid->nx, grid->ny);
  PCP_REQUIRE(workspace_bytes >= L.total, PCP_E_WORKSPACE, "pcp_voxelize3d_mean: workspace %zu < %zu bytes", workspace_bytes, L.total);
  PCP_REQUIRE(scratch_bytes >= pcp_voxel3d_scratch_bytes(n_points, max_frames, grid->nx, grid->ny), PCP_E_WORKSPACE,
              "pcp_voxelize3d_mean: scratch too small");
  PCP_REQUIRE(voxel_capacity >= n_points || voxel_capacity >= (int64_t)L.cells * nz, PCP_E_INVALID,
              "pcp_voxelize3d_mean: voxel_capacity < min(N, cells)");
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 255) == 0 &&
              (reinterpret_cast<uintptr_t>(voxel_coords_out) & 15) == 0, PCP_E_INVALID, "pcp_voxelize3d_mean: misaligned buffer");
  const WsView W = ws_view(workspace, L);
  const size_t cap = (size_t)L.cap + 1;
  char* sp = static_cast<char*>(scratch);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(sp);   sp += align_up(8 * cap, 256);
  int32_t* nvox = reinterpret_cast<int32_t*>(sp);                           sp += align_up(4 * cap, 256);
  int32_t* local = reinterpret_cast<int32_t*>(sp);                          sp += align_up(4 * cap, 256);
  int32_t* bsum = reinterpret_cast<int32_t*>(sp);                           sp += align_up(4 * (cap / kVScanBlock + 2), 256);
  int32_t* coords2d = reinterpret_cast<int32_t*>(sp);

  PCP_CUDA(cudaMemsetAsync(workspace, 0, L.clear_bytes, stream));
  if (n_points > 0) {
    const bool vec4 = (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
    const unsigned blocks = (unsigned)((n_points + 256 * kQPts - 1) / (256 * kQPts));
    if (vec4)
      quantise_count3d_kernel<true><<<blocks, 256, 0, stream>>>(points, row_stride, n_points, max_frames, *grid, range_min_z, voxel_z,
                                                                nz, W.cell, W.key, W.within, W.hdr);
    else
      quantise_count3d_kernel<false><<<blocks, 256, 0, stream>>>(points, row_stride, n_points, max_frames, *grid, range_min_z, voxel_z,
                                                                 nz, W.cell, W.key, W.within, W.hdr);
    PCP_LAUNCH_CHECK("quantise_count3d_kernel");
  }
  // pillar stage; the rows inside a pillar come out in ascending order (no xyz mean needed: points = nullptr)
  int rc = finish_grouping(L, W, n_points, grid->nx, grid->ny, nullptr, 0, *grid, nullptr, coords2d, nullptr, counts_out, stream);
  if (rc) return rc;
  const int64_t pcap = L.cap > 0 ? L.cap : 1;
  const int nblk = (int)((pcap + kVScanBlock - 1) / kVScanBlock);
  if (n_points > 0) {
    voxel_mask_kernel<<<(unsigned)((pcap * 8 + 255) / 256), 256, 0, stream>>>(points, row_stride, W.hdr, W.seg_off, W.sorted_idx,
                                                                             range_min_z, voxel_z, mask, nvox);
    PCP_LAUNCH_CHECK("voxel_mask_kernel");
  }
  voxel_scan_local_kernel<<<nblk, 256, 0, stream>>>(W.hdr, nvox, local, bsum);
  PCP_LAUNCH_CHECK("voxel_scan_local_kernel");
  voxel_scan_blocks_kernel<<<1, 1024, 0, stream>>>(W.hdr, bsum, nblk, counts_out);
  PCP_LAUNCH_CHECK("voxel_scan_blocks_kernel");
  if (n_points > 0) {
    const unsigned blocks = (unsigned)((pcap * 32 + 255) / 256);
    if (channels <= 8)
      voxel_mean_kernel<8><<<blocks, 256, 0, stream>>>(points, row_stride, channels, W.hdr, W.seg_off, W.sorted_idx, mask, local, bsum,
                                                       coords2d, range_min_z, voxel_z, voxel_features_out, voxel_coords_out, point_voxel_out);
    else
      voxel_mean_kernel<16><<<blocks, 256, 0, stream>>>(points, row_stride, channels, W.hdr, W.seg_off, W.sorted_idx, mask, local, bsum,
                                                        coords2d, range_min_z, voxel_z, voxel_features_out, voxel_coords_out, point_voxel_out);
    PCP_LAUNCH_CHECK("voxel_mean_kernel");
  }
  return 0;
}
