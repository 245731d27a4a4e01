// internal.cuh - launchers shared between the translation units of libpcp_b200.so (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace pcp {

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The kernels of the voxelize -> PFN chain run back to back on one stream; launched
// with launch_pdl() each of them may be scheduled while its predecessor is still draining: it says so itself by calling
// pdl_launch_dependents() first thing, and it calls pdl_wait() before it touches anything its predecessor wrote (the wait
// returns when the prerequisite grid has completed and its writes are visible).  The hand-over of the SMs from one kernel of
// the chain to the next then never passes through an idle scheduler.  Without the launch attribute both calls are no-ops.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Which edges of the chain are programmatic (make variant NAME=pdl15 SRC="voxelize pfn_tc" DEFS=-DPCP_PDL_MASK=15).
// MEASURED (profiles/r02_pdl_matrix.jsonl): with every edge programmatic the chain alone gains 2 % (serial step 457 -> 449 us,
// voxelize 152 -> 149 us) and voxelize beside the canvas stream drops from 254 to 176 us - but only by starving the canvas
// (its CTAs never find a free slot between two kernels of the chain any more: 228 -> 417 us), and the PFN, queued early,
// shuts it out completely: the pipelined step goes from 386 to 499 us (434 us with the PFN edge alone).  The steady state is
// what the library is built for, so the shipped default is 0: no programmatic edges.
#ifndef PCP_PDL_MASK
#define PCP_PDL_MASK 0
#endif
constexpr int kPdlScan = 1, kPdlPlace = 2, kPdlPrep = 4, kPdlPfn = 8;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int edge, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (PCP_PDL_MASK & edge) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- optional wall-clock trace of the kernels of a step (debug build only: make dbg; tools/stage_trace.py): every kernel
//      records the earliest CTA start and the latest CTA end (globaltimer, ns) in a per-translation-unit table ----
#if defined(PCP_STAGE_TRACE) && defined(__CUDACC__)
#define STAGE_TABLE(name) static __device__ unsigned long long name[8][2]
#define STAGE_BEGIN(tab, id)                                                                          \
  do {                                                                                                \
    if (threadIdx.x == 0) {                                                                           \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      atomicMin(&tab[id][0], t_);                                                                     \
    }                                                                                                 \
  } while (0)
#define STAGE_END(tab, id)                                                                            \
  do {                                                                                                \
    if (threadIdx.x == 0) {                                                                           \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      atomicMax(&tab[id][1], t_);                                                                     \
    }                                                                                                 \
  } while (0)
#define STAGE_EXPORT(fn, tab)                                                                         \
  extern "C" int fn(unsigned long long* host16, int reset) {                                          \
    if (reset) {                                                                                      \
      unsigned long long init[8][2];                                                                  \
      for (int i = 0; i < 8; ++i) { init[i][0] = ~0ull; init[i][1] = 0ull; }                          \
      return (int)cudaMemcpyToSymbol(tab, init, sizeof(init));                                        \
    }                                                                                                 \
    cudaDeviceSynchronize();                                                                          \
    return (int)cudaMemcpyFromSymbol(host16, tab, sizeof(unsigned long long) * 16);                   \
  }
#else
#define STAGE_TABLE(name)
#define STAGE_BEGIN(tab, id)
#define STAGE_END(tab, id)
#define STAGE_EXPORT(fn, tab)
#endif

// voxelize.cu: everything after the keying kernel (cell scan, placement, ascending row order inside every cell;
// the xyz mean only when `points` is given).
int finish_grouping(const WsLayout& L, const WsView& W, int64_t n, int32_t nx, int32_t ny, const float* points, int64_t stride,
                    const pcp_grid& grid, int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                    int32_t* counts_out, cudaStream_t stream);

// voxelize_radix.cu: the whole of pcp_voxelize() as a stable two-digit radix sort on the key (see the file header)
int voxelize_radix(const WsLayout& L, const WsView& W, const RadixPlan& rp, const float* points, int64_t stride, int64_t n,
                   int32_t frames, const pcp_grid& grid, int32_t* point_pillar_out, int32_t* voxel_coords_out,
                   int32_t* pillar_count_out, int32_t* counts_out, cudaStream_t stream);

// voxelize_binned.cu: the whole of pcp_voxelize() as one coarse partition pass + a per-tile finish (see the file header)
int voxelize_binned(const WsLayout& L, const WsView& W, const float* points, int64_t stride, int64_t n, int32_t frames,
                    const pcp_grid& grid, int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                    int32_t* counts_out, cudaStream_t stream);
// voxelize.cu: pillar_prep_kernel for the pillars above 8 rows, reading the binned path's placed records
int launch_pillar_prep_rec(const WsLayout& L, const WsView& W, const float* points, int64_t stride, int64_t n,
                            const pcp_grid& grid, cudaStream_t stream);

// api.cu: multiprocessors of the current device (queried once per device)
int sm_count();

// pfn.cu: per-cell mean (mode 0) / max (mode 1) of `channels` columns of `values`, rows visited in ascending row order.
int launch_segment_reduce(const float* values, int64_t value_stride, int32_t channels, int32_t mode, const WsView& W,
                          float* out, cudaStream_t stream);

// scatter.cu: dense (frames, channels, ny, nx) canvas from a canvas-ordered (frame, y, x) rank map.
int launch_canvas_from_map(const float* rows, const int32_t* rank_map, int32_t channels, int32_t frames, int32_t nx,
                           int32_t ny, float* canvas, cudaStream_t stream);

}  // namespace pcp
