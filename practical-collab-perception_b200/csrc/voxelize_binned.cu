// voxelize_binned.cu - pillar compaction by ONE coarse partition pass + a per-tile finish in shared memory
// (reference: dynamic_pillar_vfe.py:98-108 torch.unique(return_inverse, return_counts), :110 scatter_mean, :137-143 voxel_coords).
//
// The key b*nx*ny + cx*ny + cy is split at bit 11: a BIN is one scan tile of kScanTileCells = 2048 consecutive cells (for a
// 512-wide grid: four x columns of one frame - a strip across the whole y range, which averages the radial lidar density).
//   K1 bin_key      one CTA per 2048 consecutive rows: quantise + cull (bit-exact fp32), key[], rows per bin counted in
//                   shared memory and added to the global per-bin counts (one reduction per touched bin and CTA); the last
//                   CTA to finish turns the counts into the first row of every bin.
//   K2 bin_scatter  same chunks: rows per bin in shared memory again (the returning shared-memory atomic is the row's slot
//                   inside the CTA's share of the bin), one global atomic per touched bin reserves the CTA's range, every
//                   row is written as a 16-byte {x, y, z, row number} record + its key, bin-contiguous, in no particular order.
//   K3 bin_finish   one CTA per bin, everything per-cell in shared memory: rows per cell, the scan that yields pillar ranks
//                   (ascending key = torch.unique order) and first sorted positions, placement of the records by cell,
//                   then per pillar: rows put in ascending order (sorting network / rank by counting on the 16-byte
//                   records, which are contiguous and cache-hot - no gather from the point rows), the sequential fp32 sum of
//                   x, y, z in that order (the CPU scatter_mean, bit for bit), voxel_coords, counts, per-length-class work
//                   lists, long-pillar tables.  Global ranks: every tile publishes a 64-byte record of its counters and sums
//                   the records of all tiles before it; a tile publishes before it waits and tiles are numbered by a
//                   ticket in start order, so the wait cannot deadlock.
//   K4 pillar_prep  pillars above 8 rows (few, concentrated in the dense tiles) are ordered and averaged by pillar_prep_kernel
//                   (voxelize.cu) through the work lists, spread over the whole GPU; it reads the placed records, not the rows.
// Compared with the histogram path (voxelize.cu): no returning L2 atomic per point, no scan launches over the dense cell
// array, no 8 MB memset, no random placement pass over the whole batch, no three-deep gather chain per pillar.
#include "internal.cuh"

namespace pcp {

constexpr int kBnThreads = 256;
constexpr int kBnRows = 8;                                // rows per thread of K1 / K2
constexpr int kBnChunk = kBnThreads * kBnRows;            // 2048 rows per CTA
constexpr int kBnShift = 11;                              // log2(kScanTileCells)
static_assert((1 << kBnShift) == kScanTileCells, "bin = scan tile");
static_assert(kScanTileCells == kBnThreads * 8, "eight cells per thread in the tile scan");

__device__ __forceinline__ int ld_volatile(const int32_t* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int4 ld_volatile4(const int32_t* p) {
  int4 v;
  asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// ---- optional per-tile phase trace of the finish kernel (debug build only: make dbg; tools/bin_timing.py) ----
#ifdef PCP_BIN_TIMING
constexpr int kBinTraceStamps = 12;
__device__ unsigned long long g_bin_trace[kBnMaxBins][kBinTraceStamps];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define BTRACE(tile, i) do { if (threadIdx.x == 0) g_bin_trace[tile][i] = gtime(); } while (0)
#define BTRACE_V(tile, i, v) do { if (threadIdx.x == 0) g_bin_trace[tile][i] = (unsigned long long)(v); } while (0)
#else
#define BTRACE(tile, i)
#define BTRACE_V(tile, i, v)
#endif

template <bool kVec4>
__device__ __forceinline__ void load_row_head(const float* __restrict__ points, int64_t stride, int64_t i, float& bf, float& x,
                                              float& y, float& z) {
  const float* row = points + i * stride;
  if (kVec4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row));
    bf = v.x; x = v.y; y = v.z; z = v.w;
  } else {
    bf = __ldg(row); x = __ldg(row + 1); y = __ldg(row + 2); z = __ldg(row + 3);
  }
}

// ------------------------------------------------------------------------------------------------
// K1: keys + rows per bin
// ------------------------------------------------------------------------------------------------
template <bool kVec4>
__global__ void __launch_bounds__(kBnThreads)
bin_key_kernel(const float* __restrict__ points, int64_t stride, int64_t n, int32_t frames, pcp_grid g, int32_t nbins,
               int32_t* __restrict__ key, int32_t* __restrict__ bin_count, int32_t* __restrict__ hdr) {
  extern __shared__ int32_t s_hist[];                      // [nbins]
  const int tid = threadIdx.x;
  for (int b = tid; b < nbins; b += kBnThreads) s_hist[b] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kBnChunk + tid;
  float bf[kBnRows], x[kBnRows], y[kBnRows];
#pragma unroll
  for (int u = 0; u < kBnRows; ++u) {
    const int64_t i = base + u * kBnThreads;
    bf[u] = 0.f; x[u] = 0.f; y[u] = 0.f;
    float z;
    if (i < n) load_row_head<kVec4>(points, stride, i, bf[u], x[u], y[u], z);
  }
  int bad = 0;
#pragma unroll
  for (int u = 0; u < kBnRows; ++u) {
    const int64_t i = base + u * kBnThreads;
    if (i < n) {
      bool bad_frame;
      const int32_t k = point_key(bf[u], x[u], y[u], frames, g, bad_frame);
      bad += bad_frame ? 1 : 0;
      key[i] = k;
      if (k >= 0) atomicAdd(&s_hist[k >> kBnShift], 1);
    }
  }
  if (bad) atomicAdd(&hdr[PCP_COUNT_BAD_FRAME], bad);
  __syncthreads();
  for (int b = tid; b < nbins; b += kBnThreads) {
    const int32_t c = s_hist[b];
    if (c) atomicAdd(&bin_count[b], c);
  }
}

// ------------------------------------------------------------------------------------------------
// K2: partition by bin (unordered inside a bin).  The chunk's rows are first grouped by bin in shared memory, so that the
// global writes of a warp are a few contiguous runs (a 2048-row chunk of one 512 x 512 frame touches 128 bins: runs of
// ~16 records = 256 bytes) instead of 32 separate sectors.
// Dynamic shared memory: s_hist[nbins] | s_delta[nbins] | s_rec[kBnChunk] float4 | s_key[kBnChunk]
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t bin_scatter_smem(int nbins) {
  return sizeof(int32_t) * 2 * (size_t)nbins + (sizeof(float4) + sizeof(int32_t)) * (size_t)kBnChunk;
}

template <bool kVec4>
__global__ void __launch_bounds__(kBnThreads)
bin_scatter_kernel(const float* __restrict__ points, int64_t stride, int64_t n, int32_t nbins,
                   const int32_t* __restrict__ key, const int32_t* __restrict__ bin_count, int32_t* __restrict__ bin_start_out,
                   int32_t* __restrict__ bin_cursor, float4* __restrict__ rec, int32_t* __restrict__ rkey) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float4* s_rec = reinterpret_cast<float4*>(s_raw);                                  // [kBnChunk] rows grouped by bin
  int32_t* s_key = reinterpret_cast<int32_t*>(s_raw + sizeof(float4) * kBnChunk);   // [kBnChunk]
  int32_t* s_hist = s_key + kBnChunk;      // [nbins] rows of this chunk per bin, then their first position in s_rec
  int32_t* s_delta = s_hist + nbins;       // [nbins] global position of the bin's first row of this chunk - its position in s_rec
  __shared__ unsigned long long s_wsum[kBnThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int b = tid; b < nbins; b += kBnThreads) s_hist[b] = 0;
  const int64_t base = (int64_t)blockIdx.x * kBnChunk + tid;
  int32_t k[kBnRows];
  float x[kBnRows], y[kBnRows], z[kBnRows];
#pragma unroll
  for (int u = 0; u < kBnRows; ++u) {
    const int64_t i = base + u * kBnThreads;
    k[u] = (i < n) ? __ldg(key + i) : -1;
  }
#pragma unroll
  for (int u = 0; u < kBnRows; ++u) {
    const int64_t i = base + u * kBnThreads;
    x[u] = 0.f; y[u] = 0.f; z[u] = 0.f;
    float bf;
    if (i < n) load_row_head<kVec4>(points, stride, i, bf, x[u], y[u], z[u]);     // not waiting for the key
  }
  __syncthreads();
  int32_t slot[kBnRows];
#pragma unroll
  for (int u = 0; u < kBnRows; ++u) slot[u] = (k[u] >= 0) ? atomicAdd(&s_hist[k[u] >> kBnShift], 1) : 0;
  __syncthreads();
  // Two exclusive scans over the bins at once (thread t owns `per` consecutive bins; packed 64-bit sums): the chunk's own
  // counts -> first position of the bin's rows in s_rec, and the batch's counts (K1) -> first row of the bin in the output
  // (every CTA repeats this small scan: the counts are 4 bytes per bin in L2).  Then ONE global atomic per touched bin
  // reserves the chunk's range behind the rows other chunks already claimed.
  int32_t kept = 0;
  {
    const int per = (nbins + kBnThreads - 1) / kBnThreads;        // <= 32
    const int b0 = tid * per;
    unsigned long long mine = 0;
    for (int j = 0; j < per; ++j)
      if (b0 + j < nbins) mine += ((unsigned long long)(uint32_t)__ldg(bin_count + b0 + j) << 32) | (uint32_t)s_hist[b0 + j];
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    unsigned long long run = incl - mine;
#pragma unroll
    for (int w = 0; w < kBnThreads / 32; ++w) {
      if (w < warp) run += s_wsum[w];
      kept += (int32_t)(s_wsum[w] & 0xffffffffull);
    }
    for (int j = 0; j < per; ++j) {
      const int b = b0 + j;
      if (b < nbins) {
        const int32_t c = s_hist[b];
        s_hist[b] = (int32_t)(run & 0xffffffffull);                            // first position of the bin's rows in s_rec
        s_delta[b] = (int32_t)(run >> 32) - (int32_t)(run & 0xffffffffull);    // global first row of the bin - local first position
        if (blockIdx.x == 0) bin_start_out[b] = (int32_t)(run >> 32);
        run += ((unsigned long long)(uint32_t)__ldg(bin_count + b) << 32) | (uint32_t)c;
      }
    }
    if (blockIdx.x == 0 && tid == kBnThreads - 1) bin_start_out[nbins] = (int32_t)(run >> 32);
  }
  __syncthreads();
  // strided over the bins: the iterations (one returning atomic each) are independent of each other; the chunk's count of
  // a bin is the difference of two neighbouring first positions
  for (int b = tid; b < nbins; b += kBnThreads) {
    const int32_t c = ((b + 1 < nbins) ? s_hist[b + 1] : kept) - s_hist[b];
    if (c) s_delta[b] += atomicAdd(&bin_cursor[b], c);
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kBnRows; ++u) {
    if (k[u] >= 0) {
      const int32_t j = s_hist[k[u] >> kBnShift] + slot[u];
      s_rec[j] = make_float4(x[u], y[u], z[u], __int_as_float((int32_t)(base + u * kBnThreads)));
      s_key[j] = k[u];
    }
  }
  __syncthreads();
  for (int j = tid; j < kept; j += kBnThreads) {
    const int32_t kk = s_key[j];
    const int32_t pos = s_delta[kk >> kBnShift] + j;
    rec[pos] = s_rec[j];
    rkey[pos] = kk;
  }
}

// ------------------------------------------------------------------------------------------------
// K3: per-tile finish
// ------------------------------------------------------------------------------------------------
// tile record (16 ints), the same fields as the histogram path's tile_info
constexpr int kTrPoints = 0, kTrPillars = 1, kTrClass = 2, kTrLong = 12, kTrSegs = 13, kTrBig = 14, kTrMax = 15, kTrInts = 16;

__device__ __forceinline__ void cswap_rec(float4& a, float4& b) {
  if (__float_as_int(a.w) > __float_as_int(b.w)) { const float4 t = a; a = b; b = t; }
}

// one pillar of at most NMAX rows per thread: records -> ascending row order -> sorted_idx, sequential sums, mean
template <int NMAX>
__device__ __forceinline__ void finish_short(const float4* __restrict__ srec, int off, int n, int r, uint32_t cxy,
                                             int32_t* __restrict__ sorted_idx, float4* __restrict__ mean) {
  float4 v[NMAX];
#pragma unroll
  for (int j = 0; j < NMAX; ++j) v[j] = (j < n) ? __ldcg(srec + off + j) : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
  if (NMAX > 1) {
#pragma unroll
    for (int i = 1; i < NMAX; ++i)
#pragma unroll
      for (int j = i; j > 0; --j) cswap_rec(v[j - 1], v[j]);
  }
  float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
  for (int j = 0; j < NMAX; ++j)
    if (j < n) {
      sorted_idx[off + j] = __float_as_int(v[j].w);
      ax = __fadd_rn(ax, v[j].x); ay = __fadd_rn(ay, v[j].y); az = __fadd_rn(az, v[j].z);
    }
  const float cnt = (float)n;
  mean[r] = make_float4(__fdiv_rn(ax, cnt), __fdiv_rn(ay, cnt), __fdiv_rn(az, cnt), __uint_as_float(cxy));
}

__global__ void __launch_bounds__(kBnThreads, 4)
bin_finish_kernel2(const float4* __restrict__ rec, float4* __restrict__ srec, const int32_t* __restrict__ rkey,
                   const int32_t* __restrict__ bin_start, int32_t nbins, int64_t total_cells, int32_t nx, int32_t ny,
                   int32_t* __restrict__ tile_rec, int32_t* __restrict__ tile_frames, int32_t* __restrict__ ctrl,
                   int32_t* __restrict__ hdr, int32_t* __restrict__ cell_rank, int32_t* __restrict__ seg_off,
                   int32_t* __restrict__ sorted_idx, int32_t* __restrict__ voxel_coords, int32_t* __restrict__ pillar_count,
                   unsigned long long* __restrict__ lists, const ListOffsets lo, int4* __restrict__ long_table,
                   int32_t* __restrict__ big_list, float4* __restrict__ mean, int32_t* __restrict__ counts_out) {
  __shared__ __align__(16) int32_t s_cnt[kScanTileCells];     // rows per cell, then the running placement offset of the cell
  __shared__ int32_t s_pcnt[kScanTileCells];                  // per pillar of the tile (local rank): rows
  __shared__ int32_t s_poff[kScanTileCells];                  //                                      first position inside the bin
  __shared__ uint16_t s_pidx[kScanTileCells];                 //                                      cell inside the tile
  __shared__ uint32_t s_pxy[kScanTileCells];                  //                                      cx | cy << 16
  __shared__ uint16_t s_order[kScanTileCells];                // local ranks ordered by length class
  __shared__ int s_rec[kTrInts];                              // this tile's record
  __shared__ int s_before[kTrInts];                           // sums over the tiles before this one (kTrMax: max)
  __shared__ int s_part[kBnThreads / 32][kTrInts];
  __shared__ unsigned long long s_warp[kBnThreads / 32];
  __shared__ int s_cls[kTrInts];                              // running in-tile counters
  __shared__ int s_cbase[kNumClasses + 2];                    // first position of every class (10 = long) in s_order
  __shared__ long long s_lo[kNumLists];
  __shared__ int s_frw[kBnThreads / 32];
  __shared__ int s_tile, s_fr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- phase 0: tile ticket (tiles are numbered in start order: every lower-numbered tile is already running) ----
#ifdef PCP_BIN_TIMING
  const unsigned long long t_start = gtime();
#endif
  if (tid == 0) { s_tile = atomicAdd(&ctrl[1], 1); s_fr = 0; }
  if (tid < kTrInts) { s_rec[tid] = 0; s_cls[tid] = 0; }
  if (tid < kNumLists) s_lo[tid] = lo.off[tid];
  {
    int4* z = reinterpret_cast<int4*>(s_cnt);
    z[tid] = make_int4(0, 0, 0, 0);
    z[tid + kBnThreads] = make_int4(0, 0, 0, 0);
  }
  __syncthreads();
  const int tile = s_tile;
  const int32_t beg = __ldg(bin_start + tile), end = __ldg(bin_start + tile + 1);
  const int32_t base_key = tile << kBnShift;
  BTRACE_V(tile, 0, t_start);
  BTRACE(tile, 1);
  BTRACE_V(tile, 10, end - beg);
#ifdef PCP_BIN_TIMING
  { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); BTRACE_V(tile, 11, smid); }
#endif

  // ---- phase 1: rows per cell ----
  for (int32_t p0 = beg + tid; p0 < end; p0 += kBnThreads * 16) {
    int32_t c[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int32_t p = p0 + u * kBnThreads;
      c[u] = (p < end) ? __ldg(rkey + p) - base_key : -1;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u)
      if (c[u] >= 0) atomicAdd(&s_cnt[c[u]], 1);
  }
  __syncthreads();
  BTRACE(tile, 2);

  // ---- phase 2: tile statistics + block scan over the cells (thread t owns cells 8 t .. 8 t + 7) ----
  int32_t c[8];
  {
    const int4 a = *reinterpret_cast<const int4*>(s_cnt + tid * 8);
    const int4 b = *reinterpret_cast<const int4*>(s_cnt + tid * 8 + 4);
    c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
  }
  unsigned long long mine = 0;
  {
    int pts = 0, pil = 0, cmax = 0, last_nonempty = -1, nlong = 0, nseg = 0, nbig = 0;
    unsigned long long cls = 0;           // 10 x 6-bit counters: pillars of this thread per class (at most 8 each)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c[j] > 0) {
        pts += c[j]; ++pil; cmax = max(cmax, c[j]); last_nonempty = j;
        if (c[j] <= kSegRows) cls += 1ull << (6 * class_of(c[j]));
        else { ++nlong; nseg += (c[j] + kSegRows - 1) / kSegRows; nbig += (c[j] > kWarpLongMax) ? 1 : 0; }
      }
    }
    mine = ((unsigned long long)(uint32_t)pts << 32) | (unsigned long long)(uint32_t)pil;
    const int wpts = __reduce_add_sync(0xffffffffu, pts), wpil = __reduce_add_sync(0xffffffffu, pil);
    const int wmax = __reduce_max_sync(0xffffffffu, cmax);
    if (lane == 0 && wpil) { atomicAdd(&s_rec[kTrPoints], wpts); atomicAdd(&s_rec[kTrPillars], wpil); atomicMax(&s_rec[kTrMax], wmax); }
    if (wpil) {
#pragma unroll
      for (int k = 0; k < kNumClasses; ++k) {
        const int v = __reduce_add_sync(0xffffffffu, (int)((cls >> (6 * k)) & 63ull));
        if (lane == 0 && v) atomicAdd(&s_rec[kTrClass + k], v);
      }
    }
    if (__any_sync(0xffffffffu, nlong > 0)) {
      const int a = __reduce_add_sync(0xffffffffu, nlong), b = __reduce_add_sync(0xffffffffu, nseg);
      const int d = __reduce_add_sync(0xffffffffu, nbig);
      if (lane == 0) { atomicAdd(&s_rec[kTrLong], a); atomicAdd(&s_rec[kTrSegs], b); atomicAdd(&s_rec[kTrBig], d); }
    }
    // frame of the last non-empty cell of the tile + 1 (0: empty tile)
    int fr = 0;
    if (last_nonempty >= 0)
      fr = (int)(((uint32_t)base_key + (uint32_t)(tid * 8 + last_nonempty)) / ((uint32_t)nx * (uint32_t)ny)) + 1;
    fr = __reduce_max_sync(0xffffffffu, fr);
    if (lane == 0 && fr) atomicMax(&s_fr, fr);
  }
  unsigned long long incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  // ---- publish the tile record (before anything waits).  Every field is stored + 1 into zeroed memory and the record
  //      goes out as four 16-byte stores: a reader takes a quarter as published when none of its four words is zero, so
  //      there is no separate flag and no fence ----
  if (tid < 4) {
    const int4 v = make_int4(s_rec[4 * tid] + 1, s_rec[4 * tid + 1] + 1, s_rec[4 * tid + 2] + 1, s_rec[4 * tid + 3] + 1);
    *reinterpret_cast<int4*>(tile_rec + (int64_t)tile * kTrInts + 4 * tid) = v;
  }
  if (tid == 4) tile_frames[tile] = s_fr + 1;
  unsigned long long warp_excl = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < kBnThreads / 32; ++w) {
    if (w < warp) warp_excl += s_warp[w];
    tile_total += s_warp[w];
  }
  const int npil = (int)(tile_total & 0xffffffffull);
  // per cell: local rank (kept in registers until the tile's first global rank is known) + first position inside the bin,
  // which replaces the count as the cell's running placement offset; non-empty cells -> compact per-pillar arrays
  int32_t lr[8];
  {
    const unsigned long long excl0 = warp_excl + (incl - mine);
    int rl = (int)(excl0 & 0xffffffffull);
    int ol = (int)(excl0 >> 32);
    // (cx, cy) of this thread's first cell: two divisions per thread, then a step per cell
    const uint32_t nxy_ = (uint32_t)nx * (uint32_t)ny;
    const uint32_t rem0 = ((uint32_t)base_key + (uint32_t)(tid * 8)) % nxy_;
    uint32_t cx = rem0 / (uint32_t)ny, cy = rem0 - cx * (uint32_t)ny;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      lr[j] = -1;
      if (c[j] > 0) {
        lr[j] = rl;
        s_pcnt[rl] = c[j]; s_poff[rl] = ol; s_pidx[rl] = (uint16_t)(tid * 8 + j); s_pxy[rl] = (cx & 0xffffu) | (cy << 16);
        s_cnt[tid * 8 + j] = ol;
        ++rl; ol += c[j];
      }
      if (++cy == (uint32_t)ny) { cy = 0; if (++cx == (uint32_t)nx) cx = 0; }
    }
  }
  __syncthreads();
  BTRACE(tile, 3);

  // ---- phase 3: placement by cell (arrival order inside a cell; put in row order per pillar below) ----
  for (int32_t p0 = beg + tid; p0 < end; p0 += kBnThreads * 8) {
    int32_t k[8];
    float4 r[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int32_t p = p0 + u * kBnThreads;
      k[u] = -1;
      if (p < end) { k[u] = __ldg(rkey + p) - base_key; r[u] = __ldg(rec + p); }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (k[u] >= 0) srec[beg + atomicAdd(&s_cnt[k[u]], 1)] = r[u];
  }

  BTRACE(tile, 4);
  // ---- phase 4: sum of the records of ALL tiles before this one (at most kBnMaxBins x 64 bytes in L2; two records in
  //      flight per thread) ----
  {
    int acc[kTrInts];
#pragma unroll
    for (int i = 0; i < kTrInts; ++i) acc[i] = 0;
    for (int t0 = tid; t0 < tile; t0 += kBnThreads * 2) {
      int4 v[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + u * kBnThreads;
        if (t < tile) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[u][q] = ld_volatile4(tile_rec + (int64_t)t * kTrInts + 4 * q);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + u * kBnThreads;
        if (t < tile) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            while (v[u][q].x == 0 || v[u][q].y == 0 || v[u][q].z == 0 || v[u][q].w == 0) {
              __nanosleep(40);
              v[u][q] = ld_volatile4(tile_rec + (int64_t)t * kTrInts + 4 * q);
            }
            acc[4 * q + 0] += v[u][q].x - 1; acc[4 * q + 1] += v[u][q].y - 1; acc[4 * q + 2] += v[u][q].z - 1;
            if (q < 3) acc[4 * q + 3] += v[u][q].w - 1; else acc[kTrMax] = max(acc[kTrMax], v[u][q].w - 1);
          }
        }
      }
    }
    int frames = 0;
    if (tile == nbins - 1) {
      for (int t = tid; t < tile; t += kBnThreads) {
        int f = ld_volatile(tile_frames + t);
        while (f == 0) { __nanosleep(40); f = ld_volatile(tile_frames + t); }
        frames = max(frames, f - 1);
      }
      frames = __reduce_max_sync(0xffffffffu, frames);
    }
#pragma unroll
    for (int i = 0; i < kTrInts; ++i)
      acc[i] = (i == kTrMax) ? __reduce_max_sync(0xffffffffu, acc[i]) : __reduce_add_sync(0xffffffffu, acc[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kTrInts; ++i) s_part[warp][i] = acc[i];
      s_frw[warp] = frames;
    }
  }
  __syncthreads();
  if (tid < kTrInts) {
    int v = 0;
#pragma unroll
    for (int w = 0; w < kBnThreads / 32; ++w) v = (tid == kTrMax) ? max(v, s_part[w][tid]) : v + s_part[w][tid];
    s_before[tid] = v;
  }
  if (tid == 32) {
    // first position of every length class in the tile's class-ordered pillar list
    int run = 0;
    for (int k = 0; k < kNumClasses; ++k) { s_cbase[k] = run; run += s_rec[kTrClass + k]; }
    s_cbase[kNumClasses] = run;
    s_cbase[kNumClasses + 1] = run + s_rec[kTrLong];
  }
  __syncthreads();
  const int32_t r_tile = s_before[kTrPillars];
  BTRACE(tile, 5);
  if (tile == nbins - 1 && tid == 0) {
    const int32_t P = r_tile + s_rec[kTrPillars];
    const int32_t Nk = s_before[kTrPoints] + s_rec[kTrPoints];
    int frames = s_fr;
#pragma unroll
    for (int w = 0; w < kBnThreads / 32; ++w) frames = max(frames, s_frw[w]);
    const int cmax = max(s_before[kTrMax], s_rec[kTrMax]);
    const int bad = __ldcg(hdr + PCP_COUNT_BAD_FRAME);
    hdr[PCP_COUNT_PILLARS] = P; hdr[PCP_COUNT_KEPT] = Nk; hdr[PCP_COUNT_FRAMES] = frames; hdr[PCP_COUNT_MAX_PER_PILLAR] = cmax;
    seg_off[P] = Nk;
    hdr[kHdrLongCount] = s_before[kTrLong] + s_rec[kTrLong];
    hdr[kHdrListCount + kSegList] = s_before[kTrSegs] + s_rec[kTrSegs];
    hdr[kHdrBigCount] = s_before[kTrBig] + s_rec[kTrBig];
    for (int k = 0; k < kNumClasses; ++k) hdr[kHdrListCount + k] = s_before[kTrClass + k] + s_rec[kTrClass + k];
    if (counts_out) {
      counts_out[PCP_COUNT_PILLARS] = P; counts_out[PCP_COUNT_KEPT] = Nk; counts_out[PCP_COUNT_FRAMES] = frames;
      counts_out[PCP_COUNT_BAD_FRAME] = bad; counts_out[PCP_COUNT_MAX_PER_PILLAR] = cmax;
      for (int j = 5; j < PCP_COUNTS_LEN; ++j) counts_out[j] = 0;
    }
  }
  // ---- cell -> pillar rank map of the tile (every cell written: -1 = empty) ----
  {
    int32_t vr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) vr[j] = lr[j] >= 0 ? r_tile + lr[j] : -1;
    const int64_t cb = (int64_t)base_key + tid * 8;
    if (cb + 8 <= total_cells) {
      *reinterpret_cast<int4*>(cell_rank + cb) = make_int4(vr[0], vr[1], vr[2], vr[3]);
      *reinterpret_cast<int4*>(cell_rank + cb + 4) = make_int4(vr[4], vr[5], vr[6], vr[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (cb + j < total_cells) cell_rank[cb + j] = vr[j];
    }
  }
  // ---- phase 5a: per pillar in rank order (coalesced): seg_off, voxel_coords, counts; position in the class-ordered list ----
  const uint32_t nxy = (uint32_t)nx * (uint32_t)ny;
  for (int q = tid; q < npil; q += kBnThreads) {
    const int32_t cnt = s_pcnt[q];
    const int32_t r = r_tile + q;
    const uint32_t b = ((uint32_t)base_key + s_pidx[q]) / nxy;
    const uint32_t pxy = s_pxy[q];
    const uint32_t cx = pxy & 0xffffu, cy = pxy >> 16;
    seg_off[r] = beg + s_poff[q];
    // (frame, z = 0, y, x): dynamic_pillar_vfe.py:138-143 after the [0, 3, 2, 1] reorder
    if (voxel_coords) *reinterpret_cast<int4*>(voxel_coords + 4 * (int64_t)r) = make_int4((int)b, 0, (int)cy, (int)cx);
    if (pillar_count) pillar_count[r] = cnt;
    const int k = cnt <= kSegRows ? class_of(cnt) : kNumClasses;
    s_order[s_cbase[k] + atomicAdd(&s_cls[k], 1)] = (uint16_t)q;
  }
  __syncthreads();        // also: every record of the tile has been placed (phase 3)
  BTRACE(tile, 6);
  // ---- phase 5b: per pillar in class order: work-list entry; short pillars (1 .. 8 rows) are finished here, one thread
  //      each; longer ones (few, concentrated in the dense tiles) are left to pillar_prep_kernel, which spreads them over
  //      the whole GPU through its work lists ----
  {
    const int n_list = s_cbase[kNumClasses], l1 = s_cbase[kNumClasses + 1];
    for (int i = tid; i < l1; i += kBnThreads) {
      const int q = s_order[i];
      const int n = s_pcnt[q], off = beg + s_poff[q], r = r_tile + q;
      if (i < n_list) {
        int k = 0;
#pragma unroll
        for (int t = 1; t < kNumClasses; ++t) k += (i >= s_cbase[t]) ? 1 : 0;
        lists[s_lo[k] + s_before[kTrClass + k] + (i - s_cbase[k])] = pack_entry(r, off, n);
        const uint32_t cxy = s_pxy[q];
        switch (k) {
          case 0: finish_short<1>(srec, off, n, r, cxy, sorted_idx, mean); break;
          case 1: finish_short<2>(srec, off, n, r, cxy, sorted_idx, mean); break;
          case 2: finish_short<3>(srec, off, n, r, cxy, sorted_idx, mean); break;
          case 3: finish_short<4>(srec, off, n, r, cxy, sorted_idx, mean); break;
          case 4: finish_short<6>(srec, off, n, r, cxy, sorted_idx, mean); break;
          case 5: finish_short<8>(srec, off, n, r, cxy, sorted_idx, mean); break;
          default: break;
        }
      } else {
        const int nseg = (n + kSegRows - 1) / kSegRows;
        const int li = s_before[kTrLong] + (i - n_list);
        const int sb = s_before[kTrSegs] + atomicAdd(&s_cls[kTrSegs], nseg);
        long_table[li] = make_int4(r, off, n, sb);
        if (n > kWarpLongMax) big_list[s_before[kTrBig] + atomicAdd(&s_cls[kTrBig], 1)] = li;
      }
    }
  }
  BTRACE(tile, 7);
}

// point -> pillar map (unq_inv with -1 at culled rows) from the row keys and the cell -> rank map
__global__ void __launch_bounds__(256)
point_pillar_kernel(const int32_t* __restrict__ key, const int32_t* __restrict__ cell_rank, int64_t n,
                    int32_t* __restrict__ point_pillar) {
  const int64_t base = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  int32_t k[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = base + u * 256;
    k[u] = (i < n) ? __ldg(key + i) : -1;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = base + u * 256;
    if (i < n) point_pillar[i] = (k[u] >= 0) ? __ldg(cell_rank + k[u]) : -1;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int voxelize_binned(const WsLayout& L, const WsView& W, const float* points, int64_t stride, int64_t n, int32_t frames,
                    const pcp_grid& grid, int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                    int32_t* counts_out, cudaStream_t stream) {
  const int32_t nbins = (int32_t)L.scan_tiles;
  const bool vec4 = (stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
  const unsigned chunks = (unsigned)((n + kBnChunk - 1) / kBnChunk);
  const size_t hist_bytes = sizeof(int32_t) * (size_t)nbins;
  PCP_CUDA(cudaMemsetAsync(W.hdr, 0, L.bz_clear_bytes, stream));           // hdr | control words | counts | cursors | tile records
  if (vec4)
    bin_key_kernel<true><<<chunks, kBnThreads, hist_bytes, stream>>>(points, stride, n, frames, grid, nbins, W.key, W.bz_count,
                                                                     W.hdr);
  else
    bin_key_kernel<false><<<chunks, kBnThreads, hist_bytes, stream>>>(points, stride, n, frames, grid, nbins, W.key, W.bz_count,
                                                                      W.hdr);
  PCP_LAUNCH_CHECK("bin_key_kernel");
  const size_t sc_bytes = bin_scatter_smem(nbins);
  if (vec4) {
    if (sc_bytes > 40 * 1024)   // static shared memory counts too
      PCP_CUDA(cudaFuncSetAttribute(bin_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_bytes));
    bin_scatter_kernel<true><<<chunks, kBnThreads, sc_bytes, stream>>>(points, stride, n, nbins, W.key, W.bz_count, W.bz_start,
                                                                       W.bz_cursor, W.rrec, W.rkey);
  } else {
    if (sc_bytes > 40 * 1024)
      PCP_CUDA(cudaFuncSetAttribute(bin_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_bytes));
    bin_scatter_kernel<false><<<chunks, kBnThreads, sc_bytes, stream>>>(points, stride, n, nbins, W.key, W.bz_count, W.bz_start,
                                                                        W.bz_cursor, W.rrec, W.rkey);
  }
  PCP_LAUNCH_CHECK("bin_scatter_kernel");
  bin_finish_kernel2<<<(unsigned)nbins, kBnThreads, 0, stream>>>(
      W.rrec, W.rsrec, W.rkey, W.bz_start, nbins, L.cells, grid.nx, grid.ny, W.bz_rec, W.bz_frames, W.bz_ctrl, W.hdr, W.cell_rank, W.seg_off, W.sorted_idx, voxel_coords_out, pillar_count_out,
      W.lists, L.lo, W.long_table, W.big_list, W.mean, counts_out);
  PCP_LAUNCH_CHECK("bin_finish_kernel2");
  if (int rc = launch_pillar_prep_rec(L, W, points, stride, n, grid, stream)) return rc;
  if (point_pillar_out) {
    point_pillar_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(W.key, W.cell_rank, n, point_pillar_out);
    PCP_LAUNCH_CHECK("point_pillar_kernel");
  }
  return 0;
}

}  // namespace pcp

#ifdef PCP_BIN_TIMING
// debug: per tile of the last bin_finish_kernel2 launch: globaltimer stamps [0..7], rows [10], SM [11]
extern "C" int pcp_debug_read_bin_timing(unsigned long long* trace_host, int tiles) {
  PCP_CUDA(cudaDeviceSynchronize());
  PCP_CUDA(cudaMemcpyFromSymbol(trace_host, pcp::g_bin_trace, sizeof(unsigned long long) * pcp::kBinTraceStamps * tiles));
  return 0;
}
#endif
