// internal.cuh - launchers shared between the translation units of libpcp_b200.so (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace pcp {

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
