// scatter.cu - dense BEV canvas (reference: pointpillar_scatter.py:14-37).
//
// The reference zero-fills a canvas per frame, boolean-masks the pillars of that frame, index-assigns the
// transposed features and finally torch.stack()s the frames (a second full copy).  Here every canvas
// element is written exactly ONCE - zero or feature - by a kernel that is a pure streaming write:
// a CTA owns a (8 rows x 128 columns) patch of one frame, looks the pillar rank of its 1024 cells up
// once, and for every group of 4 channels gathers 16-byte pieces of the pillar rows, transposes them in
// registers and issues 128-bit streaming stores, 512 contiguous bytes per warp per (channel, row).
#include "internal.cuh"

namespace pcp {

STAGE_TABLE(g_stage_canvas);  // 0 canvas_v8_kernel
constexpr int kTileX = 128;   // columns per CTA (32 lanes x 4)
constexpr int kTileY = 8;     // rows per CTA (one warp per row)

__device__ __forceinline__ void st_stream_f4(float* p, float a, float b, float c, float d) {
  // canvas lines are never re-read by this library: keep them from displacing pillar rows in L2
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// kXMajor = true : rank map is the voxelize workspace, laid out like the reference's linear key
//                  (frame, cx, cy) -> b*nx*ny + cx*ny + cy
// kXMajor = false: rank map is canvas-ordered (frame, y, x) (generic path)
template <bool kXMajor, int kUnroll, int kMinBlocks>
__global__ void __launch_bounds__(kTileY * 32, kMinBlocks)
canvas_kernel(const float* __restrict__ pf, const int32_t* __restrict__ rank_map, int channels, int nx, int ny,
              float* __restrict__ canvas) {
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int y = blockIdx.y * kTileY + wy;
  const int x0 = blockIdx.x * kTileX + lane * 4;
  if (y >= ny || x0 >= nx) return;
  const int64_t nxy = (int64_t)nx * ny;
  int r[4];
  if (kXMajor) {
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = (x0 + i < nx) ? __ldg(rank_map + b * nxy + (int64_t)(x0 + i) * ny + y) : -1;
  } else {
    const int32_t* m = rank_map + b * nxy + (int64_t)y * nx + x0;
    if (x0 + 3 < nx && ((reinterpret_cast<uintptr_t>(m) & 15) == 0)) {
      const int4 t = __ldg(reinterpret_cast<const int4*>(m));
      r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = (x0 + i < nx) ? __ldg(m + i) : -1;
    }
  }
  float* dst = canvas + ((int64_t)b * channels) * nxy + (int64_t)y * nx + x0;
  const bool vec = (x0 + 3 < nx) && ((nx & 3) == 0) && ((reinterpret_cast<uintptr_t>(canvas) & 15) == 0);
  const bool any = (r[0] >= 0) | (r[1] >= 0) | (r[2] >= 0) | (r[3] >= 0);
  if (vec && (channels & 3) == 0) {
    if (!__any_sync(0xffffffffu, any)) {
      // whole 512-byte row piece is empty for every channel: pure zero stream
#pragma unroll 8
      for (int c = 0; c < channels; ++c) st_stream_f4(dst + (int64_t)c * nxy, 0.f, 0.f, 0.f, 0.f);
      return;
    }
    for (int c0 = 0; c0 < channels; c0 += 4 * kUnroll) {
      float4 v[kUnroll][4];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          v[u][i] = (r[i] >= 0 && c0 + 4 * u < channels)
                        ? __ldg(reinterpret_cast<const float4*>(pf + (int64_t)r[i] * channels + c0 + 4 * u))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int c = c0 + 4 * u;
        if (c < channels) {
          st_stream_f4(dst + (int64_t)(c + 0) * nxy, v[u][0].x, v[u][1].x, v[u][2].x, v[u][3].x);
          st_stream_f4(dst + (int64_t)(c + 1) * nxy, v[u][0].y, v[u][1].y, v[u][2].y, v[u][3].y);
          st_stream_f4(dst + (int64_t)(c + 2) * nxy, v[u][0].z, v[u][1].z, v[u][2].z, v[u][3].z);
          st_stream_f4(dst + (int64_t)(c + 3) * nxy, v[u][0].w, v[u][1].w, v[u][2].w, v[u][3].w);
        }
      }
    }
  } else {
    for (int c = 0; c < channels; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (x0 + i < nx) dst[(int64_t)c * nxy + i] = (r[i] >= 0) ? __ldg(pf + (int64_t)r[i] * channels + c) : 0.f;
  }
}

// 256-bit gathers (sm_100: ld.global.v8.f32): one 32-byte sector of a pillar row per lane and instruction
struct float8 { float v[8]; };
__device__ __forceinline__ float8 ldg_f8(const float* p) {
  float8 r;
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}

// Fast path of the workspace variant: channels % 8 == 0, nx % 4 == 0, 32-byte aligned pillar rows, 16-byte aligned canvas.
// Same tiling as canvas_kernel (a warp = one canvas row piece of 128 cells, a lane = 4 consecutive cells); per step a
// lane gathers 8 channels of each of its pillars with one 256-bit load and issues eight 128-bit streaming stores.
template <int kMinBlocks>
__global__ void __launch_bounds__(kTileY * 32, kMinBlocks)
canvas_v8_kernel(const float* __restrict__ pf, const int32_t* __restrict__ rank_map, int channels, int nx, int ny,
                 float* __restrict__ canvas) {
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  STAGE_BEGIN(g_stage_canvas, 0);
  const int b = blockIdx.z;
  const int y = blockIdx.y * kTileY + wy;
  const int x0 = blockIdx.x * kTileX + lane * 4;
  if (y >= ny || x0 >= nx) return;
  const int64_t nxy = (int64_t)nx * ny;
  int r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = __ldg(rank_map + b * nxy + (int64_t)(x0 + i) * ny + y);
  float* dst = canvas + ((int64_t)b * channels) * nxy + (int64_t)y * nx + x0;
  const bool any = (r[0] >= 0) | (r[1] >= 0) | (r[2] >= 0) | (r[3] >= 0);
  if (!__any_sync(0xffffffffu, any)) {
#pragma unroll 8
    for (int c = 0; c < channels; ++c) st_stream_f4(dst + (int64_t)c * nxy, 0.f, 0.f, 0.f, 0.f);
    return;
  }
  for (int c = 0; c < channels; c += 8) {
    float8 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (r[i] >= 0) v[i] = ldg_f8(pf + (int64_t)r[i] * channels + c);
      else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[i].v[q] = 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) st_stream_f4(dst + (int64_t)(c + q) * nxy, v[0].v[q], v[1].v[q], v[2].v[q], v[3].v[q]);
  }
  __syncthreads();            // trace only: warps that streamed zeros left earlier
  STAGE_END(g_stage_canvas, 0);
}

// TMA variant of canvas_v8_kernel (additionally nx % 128 == 0, ny % 8 == 0: every CTA owns a full 8 x 128 patch).  Per step
// of 8 channels the patch's (8 channels x 8 rows x 128 columns) block is assembled in shared memory - every lane stores the
// same transposed 16-byte pieces it would have sent to global memory - and leaves as 64 bulk copies (cp.async.bulk, the
// TMA unit) of 512 contiguous bytes, issued by two warps, while all warps already gather the next 8 channels.  Two staging
// buffers: a buffer is rewritten once the bulk copies issued from it two steps ago have finished READING shared memory.
constexpr int kTmaStageFloats = 8 * kTileY * kTileX;          // one buffer: 32 KB

__device__ __forceinline__ void bulk_store_g(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}

template <int kMinBlocks>
__global__ void __launch_bounds__(kTileY * 32, kMinBlocks)
canvas_tma_kernel(const float* __restrict__ pf, const int32_t* __restrict__ rank_map, int channels, int nx, int ny,
                  float* __restrict__ canvas) {
  extern __shared__ __align__(128) float s_stage[];            // [2][8 channels][8 rows][128 columns]
  const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * kTileY, xt = blockIdx.x * kTileX;
  const int y = y0 + wy, x0 = xt + lane * 4;
  const int64_t nxy = (int64_t)nx * ny;
  int r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = __ldg(rank_map + b * nxy + (int64_t)(x0 + i) * ny + y);
  // issuing threads (two warps): thread t sends channel t >> 3, row t & 7 of the step
  const int iq = tid >> 3, irow = tid & 7;
  float* const gdst = canvas + ((int64_t)b * channels + iq) * nxy + (int64_t)(y0 + irow) * nx + xt;
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_stage);
  int step = 0;
  for (int c = 0; c < channels; c += 8, ++step) {
    float8 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (r[i] >= 0) v[i] = ldg_f8(pf + (int64_t)r[i] * channels + c);
      else {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[i].v[q] = 0.f;
      }
    }
    const int buf = step & 1;
    if (step >= 2) {
      if (tid < 64) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
    }
    float* sb = s_stage + buf * kTmaStageFloats + wy * kTileX + lane * 4;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<float4*>(sb + q * (kTileY * kTileX)) = make_float4(v[0].v[q], v[1].v[q], v[2].v[q], v[3].v[q]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < 64) {
      bulk_store_g(gdst + (int64_t)c * nxy, s_base + 4u * (uint32_t)(buf * kTmaStageFloats + iq * (kTileY * kTileX) + irow * kTileX),
                   kTileX * 4);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid < 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// generic path: canvas-ordered rank map from arbitrary voxel_coords rows (frame, z, y, x)
__global__ void __launch_bounds__(256)
coords_to_map_kernel(const int32_t* __restrict__ coords, int64_t P, int frames, int nx, int ny,
                     int32_t* __restrict__ map) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P) return;
  const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + r);
  // reference: indices = z + y * nx + x (pointpillar_scatter.py:27); nz == 1 so z is 0
  const int64_t idx = (int64_t)c.y + (int64_t)c.z * nx + c.w;
  if (c.x < 0 || c.x >= frames || idx < 0 || idx >= (int64_t)nx * ny) return;
  // duplicate coordinates: the highest row wins, as the CPU reference's sequential index_put does
  atomicMax(&map[(int64_t)c.x * nx * ny + idx], (int32_t)r);
}

__global__ void __launch_bounds__(256)
num_frames_kernel(const int32_t* __restrict__ coords, int64_t P, int32_t* __restrict__ out) {
  int m = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P; r += (int64_t)gridDim.x * blockDim.x)
    m = max(m, __ldg(coords + 4 * r) + 1);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

}  // namespace pcp

using namespace pcp;

STAGE_EXPORT(pcp_debug_stage_canvas, g_stage_canvas)

static dim3 canvas_grid(int nx, int ny, int frames) {
  return dim3((unsigned)((nx + kTileX - 1) / kTileX), (unsigned)((ny + kTileY - 1) / kTileY), (unsigned)frames);
}

extern "C" int pcp_bev_scatter_ws(const float* pillar_features, int32_t channels, int32_t num_frames,
                                  int64_t n_points, int32_t max_frames, const pcp_grid* grid,
                                  const void* workspace, size_t workspace_bytes, float* canvas_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(grid && workspace && canvas_out, PCP_E_INVALID, "pcp_bev_scatter_ws: null argument");
  PCP_REQUIRE(channels > 0 && num_frames >= 0 && num_frames <= max_frames && num_frames <= 65535, PCP_E_INVALID,
              "pcp_bev_scatter_ws: bad channels/num_frames");
  const WsLayout L = ws_layout(n_points, max_frames, grid->nx, grid->ny);
  PCP_REQUIRE(workspace_bytes >= L.total, PCP_E_WORKSPACE, "pcp_bev_scatter_ws: workspace too small");
  if (num_frames == 0) return 0;
  PCP_REQUIRE(pillar_features, PCP_E_INVALID, "pcp_bev_scatter_ws: null pillar_features");
  const WsView W = ws_view(const_cast<void*>(workspace), L);
  {
    const dim3 cg = canvas_grid(grid->nx, grid->ny, num_frames);
    const int th = kTileY * 32;
    const bool fast8 = (channels % 8 == 0) && (grid->nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(pillar_features) & 31) == 0) &&
                       ((reinterpret_cast<uintptr_t>(canvas_out) & 15) == 0);
#ifdef PCP_CANVAS_TMA
    if (fast8 && grid->nx % kTileX == 0 && grid->ny % kTileY == 0) {
      const int smem = 2 * kTmaStageFloats * (int)sizeof(float);
      PCP_CUDA(cudaFuncSetAttribute(canvas_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      canvas_tma_kernel<3><<<cg, th, smem, stream>>>(pillar_features, W.cell_rank, channels, grid->nx, grid->ny, canvas_out);
    } else
#endif
    if (fast8)
      canvas_v8_kernel<3><<<cg, th, 0, stream>>>(pillar_features, W.cell_rank, channels, grid->nx, grid->ny, canvas_out);
    else
      canvas_kernel<true, 2, 3><<<cg, th, 0, stream>>>(pillar_features, W.cell_rank, channels, grid->nx, grid->ny, canvas_out);
  }
  PCP_LAUNCH_CHECK("canvas_kernel<ws>");
  return 0;
}

namespace pcp {
int launch_canvas_from_map(const float* rows, const int32_t* rank_map, int32_t channels, int32_t frames, int32_t nx,
                           int32_t ny, float* canvas, cudaStream_t stream) {
  canvas_kernel<false, 2, 3><<<canvas_grid(nx, ny, frames), kTileY * 32, 0, stream>>>(rows, rank_map, channels, nx, ny, canvas);
  PCP_LAUNCH_CHECK("canvas_kernel<map>");
  return 0;
}
}  // namespace pcp

extern "C" int pcp_bev_scatter(const float* pillar_features, const int32_t* voxel_coords, int64_t num_pillars,
                               int32_t channels, int32_t num_frames, int32_t nx, int32_t ny,
                               int32_t* cell_map_scratch, float* canvas_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(cell_map_scratch && canvas_out, PCP_E_INVALID, "pcp_bev_scatter: null argument");
  PCP_REQUIRE(channels > 0 && nx > 0 && ny > 0 && num_frames >= 0 && num_frames <= 65535 && num_pillars >= 0,
              PCP_E_INVALID, "pcp_bev_scatter: bad shape");
  PCP_REQUIRE(num_pillars == 0 || (pillar_features && voxel_coords), PCP_E_INVALID, "pcp_bev_scatter: null input");
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(voxel_coords) & 15) == 0, PCP_E_INVALID, "pcp_bev_scatter: voxel_coords not 16-byte aligned");
  if (num_frames == 0) return 0;
  PCP_CUDA(cudaMemsetAsync(cell_map_scratch, 0xff, sizeof(int32_t) * (size_t)num_frames * nx * ny, stream));
  if (num_pillars > 0) {
    coords_to_map_kernel<<<(unsigned)((num_pillars + 255) / 256), 256, 0, stream>>>(voxel_coords, num_pillars, num_frames,
                                                                                  nx, ny, cell_map_scratch);
    PCP_LAUNCH_CHECK("coords_to_map_kernel");
  }
  canvas_kernel<false, 2, 3><<<canvas_grid(nx, ny, num_frames), kTileY * 32, 0, stream>>>(pillar_features, cell_map_scratch,
                                                                                     channels, nx, ny, canvas_out);
  PCP_LAUNCH_CHECK("canvas_kernel<map>");
  return 0;
}

extern "C" int pcp_num_frames(const int32_t* voxel_coords, int64_t num_pillars, int32_t* num_frames_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(num_frames_out && num_pillars >= 0 && (num_pillars == 0 || voxel_coords), PCP_E_INVALID,
              "pcp_num_frames: bad argument");
  PCP_CUDA(cudaMemsetAsync(num_frames_out, 0, sizeof(int32_t), stream));
  if (num_pillars > 0) {
    num_frames_kernel<<<sm_count(), 256, 0, stream>>>(voxel_coords, num_pillars, num_frames_out);
    PCP_LAUNCH_CHECK("num_frames_kernel");
  }
  return 0;
}
