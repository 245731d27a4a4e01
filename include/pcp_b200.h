/*
 * pcp_b200.h - C ABI of libpcp_b200.so: the B200 (sm_100a) point->BEV front end.
 *
 * This is the drop-in boundary for the reference's hot path (quan-dao/practical-collab-perception,
 * an OpenPCDet fork).  Each entry point names the reference code it replaces (paths relative to the
 * reference root).  The reference reaches this arithmetic through torch / torch_scatter /
 * roiaware_pool3d_cuda calls inside three Python call sites; a maintainer binds this library with
 * ctypes - no torch types cross the boundary.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the parameter name ends in _host;
 *    the library never allocates, frees or retains caller memory;
 *  - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); no entry point
 *    synchronises the device or the stream; results that size later tensors (`counts`) are read back
 *    by the caller (one 32-byte D2H copy);
 *  - scratch lives in a caller-provided workspace of pcp_workspace_bytes() bytes; the workspace filled
 *    by pcp_voxelize() is consumed by pcp_pfn(), pcp_segment_reduce() and pcp_bev_scatter_ws(), which
 *    must be given the same (n_points, max_frames, nx, ny) so that they find the same layout;
 *  - return value: 0 = ok, < 0 = invalid argument (PCP_E_*), > 0 = a cudaError_t from a launch.
 *    pcp_last_error_string() describes the last failure on the calling thread.  Never exit()s
 *    (contrast pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:350-354);
 *  - stateless and re-entrant: no globals, no static device buffers; safe for one process per GPU or
 *    several streams in one process; every call is CUDA-graph capturable.
 */
#ifndef PCP_B200_H_
#define PCP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCP_ABI_VERSION 1

#define PCP_E_INVALID   (-1)   /* null pointer, negative size, unsupported shape */
#define PCP_E_WORKSPACE (-2)   /* workspace smaller than pcp_workspace_bytes() */
#define PCP_E_UNSUPPORTED (-3) /* configuration outside what the kernels implement */

/* indices into the int32 `counts[PCP_COUNTS_LEN]` block written by pcp_voxelize() */
#define PCP_COUNTS_LEN        8
#define PCP_COUNT_PILLARS     0   /* P  = number of non-empty pillars                      */
#define PCP_COUNT_KEPT        1   /* N' = points that survived the range cull              */
#define PCP_COUNT_FRAMES      2   /* max frame index among pillars + 1 (0 if P == 0)        */
#define PCP_COUNT_BAD_FRAME   3   /* points whose frame index was < 0 or >= max_frames      */
#define PCP_COUNT_MAX_PER_PILLAR 4 /* largest number of points in one pillar               */
#define PCP_COUNT_VOXELS      5   /* V  = non-empty 3-D voxels (pcp_voxelize3d_mean only)     */

/* Constants DynamicPillarVFE.__init__ derives: pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py:77-89.
 * x/y/z_offset are computed by the HOST exactly as the reference does (voxel/2 + range_min, :80-82) and
 * passed as fp32, because their rounding depends on the caller's scalar types. */
typedef struct pcp_grid {
  float range_min_x, range_min_y;  /* point_cloud_range[0], [1]            */
  float voxel_x, voxel_y;          /* voxel_size[0], [1]                   */
  float x_offset, y_offset, z_offset;
  int32_t nx, ny;                  /* grid_size[0], grid_size[1]; nz == 1  */
} pcp_grid;

/* Feature assembly + PFN stack description: dynamic_pillar_vfe.py:53-75,118-126. */
typedef struct pcp_pfn_desc {
  int32_t c_raw;             /* NUM_RAW_POINT_FEATURES: columns 1..c_raw of a point row          */
  int32_t use_absolute_xyz;  /* USE_ABSLOTE_XYZ: 1 -> features start at column 1, 0 -> column 4 */
  int32_t with_distance;     /* WITH_DISTANCE: append ||xyz||                                  */
  int32_t num_layers;        /* 1 or 2 PFNLayerV2                                                */
  int32_t hidden;            /* layer-0 output width when num_layers == 2 (NUM_FILTERS[0]/2 = 32)*/
  int32_t c_out;             /* NUM_FILTERS[-1] (64)                                             */
} pcp_pfn_desc;

int  pcp_abi_version(void);
const char* pcp_last_error_string(void);

/* Bytes of scratch for a batch of n_points rows over max_frames frames on an nx x ny grid. */
size_t pcp_workspace_bytes(int64_t n_points, int32_t max_frames, int32_t nx, int32_t ny);

/* Number of float32 elements of the packed PFN parameter block pcp_pack_pfn_params() writes. */
size_t pcp_pfn_param_floats(const pcp_pfn_desc* desc);

/*
 * Fold BatchNorm1d(eval) into per-channel scale/shift and lay the PFN weights out for pcp_pfn().
 * Replaces the per-forward nn.Linear / nn.BatchNorm1d parameter reads of PFNLayerV2
 * (dynamic_pillar_vfe.py:24-39).  alpha = gamma / sqrt(var + eps), beta = bias - mean * alpha
 * (ATen's eval-mode batch_norm transform).  For use_norm == 0 pass bn_* = NULL and lin_bias != NULL.
 * w0: (w0_out, c_in) row-major, w1: (c_out, 2 * hidden) row-major or NULL when num_layers == 1.
 */
int pcp_pack_pfn_params(const pcp_pfn_desc* desc,
                        const float* w0, const float* lin_bias0,
                        const float* bn0_weight, const float* bn0_bias, const float* bn0_mean, const float* bn0_var,
                        const float* w1, const float* lin_bias1,
                        const float* bn1_weight, const float* bn1_bias, const float* bn1_mean, const float* bn1_var,
                        float eps, float* packed_out, void* stream);

/*
 * Quantise + cull + linearise + compact.  Replaces dynamic_pillar_vfe.py:98-108 (floor((xy-min)/voxel),
 * range mask, merge_coords, torch.unique(return_inverse, return_counts)) and :137-143 (voxel_coords).
 *   points            (n_points, row_stride) fp32, column 0 = frame index, 1..3 = x,y,z
 *   point_pillar_out  (n_points) int32: pillar rank of each input row, -1 for culled rows.  The
 *                     reference's unq_inv is this array with the -1 entries removed.  May be NULL.
 *   voxel_coords_out  (pillar_capacity, 4) int32 rows (frame, 0, y, x), ascending linear key, first P valid
 *   pillar_count_out  (pillar_capacity) int32 points per pillar (unq_cnt).  May be NULL.
 *   counts_out        int32[PCP_COUNTS_LEN]
 * Points whose frame index is outside [0, max_frames) are dropped and counted in PCP_COUNT_BAD_FRAME.
 * Non-finite x or y are culled (the reference's float->int cast of NaN is undefined behaviour).
 */
int pcp_voxelize(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                 const pcp_grid* grid, void* workspace, size_t workspace_bytes,
                 int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                 int64_t pillar_capacity, int32_t* counts_out, void* stream);

/*
 * pcp_voxelize() with the compaction algorithm chosen by the caller (pillars, maps, counts identical bit for bit; means
 * identical too - sequential sums in ascending row order for pillars of any length):
 *   PCP_VOXELIZE_HISTOGRAM  dense per-cell histogram with one L2 atomic per point + cell scan + counting-sort placement + an
 *                           ordering pass per pillar (csrc/voxelize.cu); any size.  The faster one on the B200 at every size
 *                           measured (profiles/r02_voxelize_time_*.json), hence:
 *   PCP_VOXELIZE_AUTO       = HISTOGRAM (what pcp_voxelize() does).
 *   PCP_VOXELIZE_RADIX      stable two-digit MSD radix sort on the linear key: partition into bins of 512 / 1024 consecutive
 *                           cells, then a counting sort + run-length per bin in shared memory (csrc/voxelize_radix.cu).  No
 *                           global atomics, no gathers; rows ascend inside every pillar by construction, so the per-pillar
 *                           mean is the sequential sum in row order for pillars of ANY length.  Covers key spaces of at most 4 M cells (16 frames of
 *                           512 x 512) and 1 .. 16.6 M rows; PCP_E_UNSUPPORTED outside.
 *   PCP_VOXELIZE_BINNED     one coarse partition pass into bins of 2048 consecutive cells ({x, y, z, row} records, unordered
 *                           inside a bin), then one CTA per bin: rows per cell, scan, placement, pillars of up to 8 rows put in
 *                           row order and averaged from the contiguous records, global ranks from a look-back over 64-byte
 *                           tile records; longer pillars go through the work lists (csrc/voxelize_binned.cu).  One returning
 *                           global atomic per (CTA, bin) instead of one per point, 20 KB cleared instead of 8 MB, 4 launches
 *                           instead of 5 - and still slower than HISTOGRAM (169 vs 152 us on the bench batch, DESIGN.md
 *                           section 3).  Covers n_points >= 1 and at most 4 M cells; PCP_E_UNSUPPORTED outside.
 */
#define PCP_VOXELIZE_AUTO      0
#define PCP_VOXELIZE_HISTOGRAM 1
#define PCP_VOXELIZE_RADIX     2
#define PCP_VOXELIZE_BINNED    3
int pcp_voxelize_method(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                        const pcp_grid* grid, void* workspace, size_t workspace_bytes,
                        int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                        int64_t pillar_capacity, int32_t* counts_out, int32_t method, void* stream);

/*
 * Pillar feature network.  Replaces dynamic_pillar_vfe.py:110-129 (scatter_mean, f_cluster, f_center,
 * concat) and PFNLayerV2.forward :35-46 for every layer, fused: no per-point intermediate reaches HBM.
 *   pillar_features_out (pillar_capacity, c_out) fp32, first P rows valid.
 *   pillar_mean_out     (pillar_capacity, 3) fp32 per-pillar mean xyz (points_mean, :110).  May be NULL.
 */
int pcp_pfn(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames, const pcp_grid* grid,
            const pcp_pfn_desc* desc, const float* packed_params,
            const void* workspace, size_t workspace_bytes,
            float* pillar_features_out, float* pillar_mean_out, int64_t pillar_capacity, void* stream);

/*
 * pcp_pfn() in two separately schedulable stages (same arguments plus a stage mask): PCP_PFN_STAGE_SLOTS is the tensor-core
 * kernel (every pillar of at most 32 points complete, partial maxima of the longer ones), PCP_PFN_STAGE_LONG finishes the
 * pillars above 32 points (a few thousand per batch, latency bound).  A pipelined caller enqueues the second stage on the
 * stream that writes the canvas, so that the next batch's voxelize kernels start right behind the tensor-core kernel.
 */
#define PCP_PFN_STAGE_SLOTS 1
#define PCP_PFN_STAGE_LONG  2
#define PCP_PFN_STAGE_ALL   3
int pcp_pfn_stages(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames, const pcp_grid* grid,
                   const pcp_pfn_desc* desc, const float* packed_params, const void* workspace, size_t workspace_bytes,
                   float* pillar_features_out, float* pillar_mean_out, int64_t pillar_capacity, int32_t stages, void* stream);

/*
 * Segmented reduction of arbitrary per-point values over the pillars found by pcp_voxelize():
 * torch_scatter.scatter_mean / scatter_max (call sites dynamic_pillar_vfe.py:40,110).
 *   values (n_points, value_stride) fp32 indexed by ORIGINAL row number, first `channels` columns reduced
 *   mode   0 = mean (sum in ascending row order / count), 1 = max
 *   out    (pillar_capacity, channels)
 */
int pcp_segment_reduce(const float* values, int64_t value_stride, int32_t channels, int32_t mode,
                       int64_t n_points, int32_t max_frames, int32_t nx, int32_t ny,
                       const void* workspace, size_t workspace_bytes,
                       float* out, int64_t pillar_capacity, void* stream);

/*
 * Dense BEV canvas from the workspace left by pcp_voxelize() (fast path: no search, no memset; every
 * canvas element is written exactly once).  Replaces PointPillarScatter.forward,
 * pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:14-37.
 *   canvas_out (num_frames, channels, ny, nx) fp32.
 */
int pcp_bev_scatter_ws(const float* pillar_features, int32_t channels, int32_t num_frames,
                       int64_t n_points, int32_t max_frames, const pcp_grid* grid,
                       const void* workspace, size_t workspace_bytes,
                       float* canvas_out, void* stream);

/*
 * Dense BEV canvas from arbitrary (pillar_features, voxel_coords) - the generic PointPillarScatter for
 * callers that did not run pcp_voxelize().  cell_map_scratch: (num_frames * ny * nx) int32 scratch.
 * Rows with coordinates outside the canvas are ignored; duplicate coordinates: the highest row wins
 * (the CPU reference's sequential index_put order).
 */
int pcp_bev_scatter(const float* pillar_features, const int32_t* voxel_coords, int64_t num_pillars,
                    int32_t channels, int32_t num_frames, int32_t nx, int32_t ny,
                    int32_t* cell_map_scratch, float* canvas_out, void* stream);

/* max(voxel_coords[:,0]) + 1 into *num_frames_out (device int32): pointpillar_scatter.py:17. */
int pcp_num_frames(const int32_t* voxel_coords, int64_t num_pillars, int32_t* num_frames_out, void* stream);

/*
 * MoDAR point synthesis for all agents of one frame in one launch.  Replaces
 * pcdet/datasets/v2x_sim/v2x_sim_dataset_ego.py:203-232 (== workspace/visualize_collab.py:118-142,253-262):
 * points_in_boxes_gpu (roiaware_pool3d_kernel.cu:16-36,313-336) + unique + scatter(mean) * scale +
 * modar[:, :3] += offset, then apply_se3_ (nuscenes_temporal_utils.py:66-70), then the row packing.
 *   boxes        (total_boxes, 9) fp32: box7 | score | label, agents concatenated
 *   box_offsets  int32[num_agents + 1] prefix offsets into boxes
 *   foreground   (total_fg, 13) fp32: pt5 | sweep | inst | cls3 | flow3; may be NULL (no propagation)
 *   fg_offsets   int32[num_agents + 1]
 *   se3          (num_agents, 12) fp64: rows 0..2 of target_se3_agent, row-major
 *   max_boxes_per_agent, max_fg_per_agent: largest per-agent counts (the host built the offsets, so it knows them): grid sizes
 *   scale        flow multiplier: 2.0 * latency / sweep interval; 2.0 in the reference (:213); 0 = EXCHANGE_NOW
 *   rows_out     (total_boxes, out_stride) fp32; with_batch_col != 0 prepends the frame index column
 *                (collate_batch layout), i.e. 14 columns instead of 13
 *   box_idx_out  (total_fg) int32 per-agent box index of each foreground point, -1 = none.  May be NULL.
 */
int pcp_modar(const float* boxes, const int32_t* box_offsets, const float* foreground, const int32_t* fg_offsets,
              const double* se3, int32_t num_agents, int32_t max_boxes_per_agent, int32_t max_fg_per_agent,
              float scale, float max_sweep_idx,
              int32_t with_batch_col, float batch_idx, float* rows_out, int64_t out_stride,
              int32_t* box_idx_out, void* stream);

/*
 * pcp_modar() over per-agent record arrays, with nothing gathered on the host side: box_ptrs / fg_ptrs are DEVICE arrays of
 * num_agents device pointers (agent a's (M_a, 9) boxes and (F_a, 13) foreground records; fg_ptrs may be NULL = no
 * propagation, an agent without records has F_a = 0), box_offsets / fg_offsets their prefix sums as above (rows_out and
 * box_idx_out are indexed through them), and max_sweep_idx is read from device memory (pcp_column_max of the ego cloud's
 * sweep column, v2x_sim_dataset_ego.py:174) - no device->host synchronisation anywhere on the exchange path.
 */
int pcp_modar_agents(const float* const* box_ptrs, const int32_t* box_offsets, const float* const* fg_ptrs,
                     const int32_t* fg_offsets, const double* se3, int32_t num_agents, int32_t max_boxes_per_agent,
                     int32_t max_fg_per_agent, float scale, const float* max_sweep_idx_dev, int32_t with_batch_col,
                     float batch_idx, float* rows_out, int64_t out_stride, int32_t* box_idx_out, void* stream);

/* max over the rows of one column of a row-major fp32 matrix into *max_out (device); 0 for an empty matrix. */
int pcp_column_max(const float* rows, int64_t row_stride, int64_t n_rows, int32_t column, float* max_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * the callers and data formats either side of the pillar path.
 * ------------------------------------------------------------------------------------------------------------ */

/* max(index) + 1 into *max_plus_one_out (device int32, 0 for an empty array): the `torch.max(points_batch_idx).item() + 1`
 * of pcdet/models/bev_layers/hunter_toolbox.py:74. */
int pcp_max_index_i64(const int64_t* index, int64_t n, int32_t* max_plus_one_out, void* stream);

/*
 * Points -> BEV image by per-pixel mean.  Replaces bev_scatter(), pcdet/models/bev_layers/hunter_toolbox.py:65-96
 * (called every forward by pcdet/models/bev_layers/hunter_jr.py:279): strict in-image mask (:78-79), .long() truncation
 * (:81), merge = b * area + y * width + x (:82), torch.unique + torch_scatter.scatter_mean (:84-86), zero fill and
 * '(B H W) C -> B C H W' (:85,88).
 *   bev_coord   (n_points, coord_stride >= 2) fp32: bev_x, bev_y (pixels)
 *   batch_idx   (n_points) int64
 *   feat        (n_points, feat_stride >= channels) fp32
 *   workspace   pcp_workspace_bytes(n_points, num_frames, height, width) bytes (the pixel grid plays the pillar grid)
 *   cell_mean_scratch (min(n_points, num_frames * height * width), channels) fp32
 *   bev_out     (num_frames, channels, height, width) fp32, every element written
 *   counts_out  int32[PCP_COUNTS_LEN]: [PCP_COUNT_PILLARS] = occupied pixels, [PCP_COUNT_KEPT] = points inside the image
 * The per-pixel sum runs in ascending point order (the CPU scatter_mean's order): deterministic, bit-reproducible.
 */
int pcp_bev_scatter_mean(const float* bev_coord, int64_t coord_stride, const int64_t* batch_idx,
                         const float* feat, int64_t feat_stride, int32_t channels, int64_t n_points,
                         int32_t num_frames, int32_t height, int32_t width, void* workspace, size_t workspace_bytes,
                         float* cell_mean_scratch, float* bev_out, int32_t* counts_out, void* stream);

/*
 * BEV image -> points by bilinear interpolation.  Replaces interpolate_points_feat_from_bev_img() and
 * bilinear_interpolate_torch(), hunter_toolbox.py:8-41,99-131 (hunter_jr.py:268,300).
 *   bev_img     (num_frames, channels, height, width) fp32, or (num_frames, height, width, channels) when channels_last != 0
 *   points      (n_points, row_stride) fp32: column 0 = frame index, 1..2 = x, y (metres)
 *   bev coordinate = (xy - range_min) / pixel, fp32 subtract and IEEE divide (:114)
 *   nhwc_scratch     num_frames * channels * height * width floats (unused when channels_last != 0)
 *   points_feat_out  (n_points, channels); rows whose frame index is outside [0, num_frames) are zero (:116-123)
 *   bev_coord_out    (n_points, 2) or NULL (return_bev_coord)
 * Every product / sum is rounded separately, in the reference's order: bit-exact against the CPU reference.
 */
int pcp_bev_interpolate(const float* bev_img, int32_t channels_last, int32_t num_frames, int32_t channels,
                        int32_t height, int32_t width, const float* points, int64_t row_stride, int64_t n_points,
                        float range_min_x, float range_min_y, float pixel_x, float pixel_y, float* nhwc_scratch,
                        float* points_feat_out, float* bev_coord_out, void* stream);

/*
 * Dynamic 3-D voxelisation with per-voxel feature means.  Replaces DynamicMeanVFE.forward,
 * pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:42-79 (the VFE of tools/cfgs/v2x_sim_models/v2x_second_*.yaml).
 *   points   (n_points, row_stride) fp32: column 0 = frame index, columns 1 .. channels are averaged (x, y, z first)
 *   grid     x / y constants as for pcp_voxelize(); range_min_z, voxel_z, nz: the third axis (nz <= 64)
 *   workspace pcp_workspace_bytes(n, max_frames, nx, ny); scratch pcp_voxel3d_scratch_bytes(n, max_frames, nx, ny)
 *   voxel_features_out (voxel_capacity, channels) fp32; voxel_coords_out (voxel_capacity, 4) int32 rows (frame, z, y, x) in
 *   ascending key order b*nx*ny*nz + cx*ny*nz + cy*nz + cz (= torch.unique order); voxel_capacity >= min(n, cells * nz)
 *   point_voxel_out (n_points) int32 voxel rank per input row (-1 = culled), may be NULL
 *   counts_out[PCP_COUNT_VOXELS] = V, [PCP_COUNT_KEPT] = N'
 */
size_t pcp_voxel3d_scratch_bytes(int64_t n_points, int32_t max_frames, int32_t nx, int32_t ny);
int pcp_voxelize3d_mean(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                        const pcp_grid* grid, float range_min_z, float voxel_z, int32_t nz, int32_t channels,
                        void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                        float* voxel_features_out, int32_t* voxel_coords_out, int32_t* point_voxel_out,
                        int64_t voxel_capacity, int32_t* counts_out, void* stream);

/*
 * Early-fusion input assembly.  Replaces pcdet/datasets/v2x_sim/v2x_sim_dataset_ego_early.py:85-92 (per-agent apply_se3_,
 * nuscenes_temporal_utils.py:62-63, + np.concatenate), the range mask of pcdet/datasets/processor/data_processor.py:78-84
 * (common_utils.py:64-68) and the frame-index column of collate_batch (pcdet/datasets/dataset.py:224-229).
 *   points        (n_points, in_stride) fp32 rows x, y, z, ... of all agents, ego first (identity transform)
 *   agent_offsets int32[num_agents + 1] device prefix offsets; se3 (num_agents, 12) fp64 device, rows 0..2 of target_se3_agent
 *   range6_host   HOST float[6] point_cloud_range, or NULL for no mask
 *   scratch       pcp_fuse_scratch_bytes(n_points) bytes
 *   rows_out      (n_points, out_stride): [frame index if with_batch_col] x' y' z' then columns 3 .. n_cols-1 unchanged;
 *                 surviving rows in input order; *count_out (device int32) = their number
 */
size_t pcp_fuse_scratch_bytes(int64_t n_points);
int pcp_fuse_agent_points(const float* points, int64_t in_stride, int32_t n_cols, int64_t n_points,
                          const int32_t* agent_offsets, const double* se3, int32_t num_agents,
                          const float* range6_host, int32_t with_batch_col, float batch_idx, int32_t* scratch,
                          float* rows_out, int64_t out_stride, int32_t* count_out, void* stream);
/* pcp_fuse_agent_points() over per-agent clouds left where they are: cloud_ptrs is a DEVICE array of num_agents device pointers
 * (agent a's (N_a, in_stride) rows; agent_offsets their prefix sums as above) - no concatenation of the inputs. */
int pcp_fuse_agent_clouds(const float* const* cloud_ptrs, int64_t in_stride, int32_t n_cols, int64_t n_points,
                          const int32_t* agent_offsets, const double* se3, int32_t num_agents,
                          const float* range6_host, int32_t with_batch_col, float batch_idx, int32_t* scratch,
                          float* rows_out, int64_t out_stride, int32_t* count_out, void* stream);

/*
 * Producer side of the exchange: the foreground records one agent broadcasts.  Replaces, in the test-time branch of
 * pcdet/models/bev_layers/hunter_jr.py:377-397, torch.sigmoid(points_cls_logit), the mask prob[:, 0] < 0.3, the torch.cat of
 * [points[mask, 1:], prob[mask], points_flow3d[mask]] and the per-sample boolean split: one stable partition on the GPU.
 *   points       (n_points, point_stride) fp32: column 0 = sample index, columns 1 .. n_point_cols are sent (7 in the reference)
 *   cls_logit    (n_points, logit_stride >= 3), flow3d (n_points, flow_stride >= 3)
 *   rows_out     (>= n_points, out_stride >= n_point_cols + 6): [point columns | prob3 | flow3], grouped by sample, input order
 *   frame_offsets_out  device int32[num_frames + 1]: first row of every sample, [num_frames] = rows sent
 *   scratch      pcp_select_scratch_bytes(n_points, num_frames) bytes
 * Rows whose sample index is outside [0, num_frames) are not sent (the reference's loop over metadata never reaches them).
 */
size_t pcp_select_scratch_bytes(int64_t n_points, int32_t num_frames);
int pcp_select_foreground(const float* points, int64_t point_stride, int32_t n_point_cols, const float* cls_logit,
                          int64_t logit_stride, const float* flow3d, int64_t flow_stride, int64_t n_points,
                          int32_t num_frames, float threshold, int32_t* scratch, float* rows_out, int64_t out_stride,
                          int32_t* frame_offsets_out, void* stream);

/*
 * Device half of the points loader.  Replaces, for batch_dict['points'], the frame-index padding + concatenation of
 * collate_batch (pcdet/datasets/dataset.py:224-229) and the .float().cuda() of load_data_to_gpu (pcdet/models/__init__.py:
 * 23-34): the host ships only the per-point columns a consumer reads (x, y, z, intensity, time = 20 of the 28 bytes of an
 * early-fusion row; the frame index travels as one row offset per frame), this entry rebuilds the (N, 1 + C) rows.
 *   packed          (n_points, n_packed_cols) fp32 device: the shipped columns, frames back to back
 *   col_index_host  HOST int32[n_packed_cols]: which per-point column (0-based, frame-index column not counted) each is
 *   frame_offsets   device int32[num_frames + 1]: first row of every frame, [num_frames] = n_points
 *   rows_out        (n_points, out_stride >= 1 + n_point_cols): column 0 = frame index, shipped columns in place, others 0
 */
#define PCP_UNPACK_MAX_COLS 16
int pcp_unpack_points(const float* packed, int32_t n_packed_cols, const int32_t* col_index_host, int64_t n_points,
                      const int32_t* frame_offsets, int32_t num_frames, int32_t n_point_cols, float* rows_out,
                      int64_t out_stride, void* stream);

/*
 * Pairwise BEV IoU of rotated boxes [x, y, z, dx, dy, dz, heading]: iou3d_nms_utils.boxes_iou_bev,
 * pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu:227-265.  iou_out (num_a, num_b) fp32.
 * The overlap follows the reference kernel's arithmetic (box_overlap, :107-225: edge-pair intersections, corner test with
 * its 1e-2 m margin, ordering by atan2, fan area): results agree with the reference kernel's to 1e-5 (> 99 % bit-equal; FMA
 * fusion is ptxas's choice), NMS keep lists are identical.
 */
int pcp_boxes_iou_bev(const float* boxes_a, int64_t num_a, const float* boxes_b, int64_t num_b, float* iou_out, void* stream);

/*
 * Class-agnostic rotated-box NMS, entirely on the device.  Replaces the late-fusion box merge
 * pcdet/models/detectors/v2x_late_fusion.py:21-35 -> model_nms_utils.class_agnostic_nms
 * (pcdet/models/model_utils/model_nms_utils.py:6-27) -> iou3d_nms_utils.nms_gpu (iou3d_nms_utils.py:84-99,
 * iou3d_nms_kernel.cu:267-312, and the HOST scan of iou3d_nms.cpp:100-135 with its cudaMalloc + D2H copy).
 *   boxes (num_boxes, box_stride >= 7) fp32, scores (num_boxes) fp32
 *   apply_score_thresh / score_thresh: keep scores >= thresh (:9); pre_max_size: top-k before NMS (:15, <= 0: all);
 *   iou_thresh: suppress when IoU_bev > thresh; post_max_size: at most this many survivors (:20, <= 0: all)
 *   scratch  pcp_nms_scratch_bytes(num_boxes) bytes, 256-byte aligned
 *   keep_out int64[num_boxes]: indices into `boxes` of the survivors, highest score first (ties: lower index first)
 *   count_out device int32: number of survivors; -1 if more than 4096 boxes pass the score mask AND no pre_max_size
 *            in [1, 4096] narrows them (with one, the top pre_max_size are selected first by a radix select: any num_boxes)
 * pcp_nms_normal() is the same pipeline with the axis-aligned IoU of nms_normal_gpu (iou3d_nms_utils.py:102-116,
 * iou3d_nms_kernel.cu:316-327).
 */
size_t pcp_nms_scratch_bytes(int64_t num_boxes);
int pcp_nms_bev(const float* boxes, int64_t box_stride, const float* scores, int64_t num_boxes,
                int32_t apply_score_thresh, float score_thresh, float iou_thresh, int32_t pre_max_size,
                int32_t post_max_size, void* scratch, size_t scratch_bytes, int64_t* keep_out, int32_t* count_out,
                void* stream);
int pcp_nms_normal(const float* boxes, int64_t box_stride, const float* scores, int64_t num_boxes,
                   int32_t apply_score_thresh, float score_thresh, float iou_thresh, int32_t pre_max_size,
                   int32_t post_max_size, void* scratch, size_t scratch_bytes, int64_t* keep_out, int32_t* count_out,
                   void* stream);

/*
 * Diagnostic: C[128 x n] = A[128 x k] . B[n x k]^T through exactly the tensor-core path pcp_pfn() uses
 * (shared-memory operand panels, tcgen05.mma kind::tf32 with the 3xTF32 split, TMEM accumulator,
 * tcgen05.ld).  k multiple of 8 up to 64, n = 32 or 64, all row-major fp32 device pointers.
 */
int pcp_selftest_umma(const float* a, const float* b, int32_t k, int32_t n, float* c, void* stream);

/* Same product with the A operand staged in TENSOR MEMORY (tcgen05.st, then tcgen05.mma with a TMEM A address):
 * the path layer 1 of pcp_pfn() uses for the activations it reads back from layer 0.  k multiple of 8 up to 32. */
int pcp_selftest_umma_ts(const float* a, const float* b, int32_t k, int32_t n, float* c, void* stream);

/* Diagnostic micro-benchmark: SM cycles (clock64) of `reps` back-to-back 3xTF32 groups (3 * ksteps tcgen05.mma each,
 * M = 128, N = n; mode 0: A from shared memory, 1: A from tensor memory), issue -> commit -> mbarrier wait.
 * out[0] = total cycles, out[1] = cycles of an empty commit + wait, out[2] = cycles spent issuing. */
int pcp_selftest_umma_cycles(int32_t mode, int32_t n, int32_t ksteps, int32_t reps, long long* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* PCP_B200_H_ */
