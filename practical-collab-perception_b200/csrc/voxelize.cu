// voxelize.cu - quantise + cull + linearise + compact (reference: dynamic_pillar_vfe.py:98-108,137-143).
//
// B200-first design: the linear key space B*nx*ny is dense and small (2 M cells for 8 frames of
// 512 x 512), so instead of the reference's sort-based torch.unique the pillars are found with a
// dense per-cell histogram that lives in L2:
//   1. quantise_count : one thread per point, bit-exact fp32 quantisation, atomicAdd on the cell's
//                       counter returns the point's arrival slot inside its cell;
//   2. scan_cells     : ONE pass decoupled look-back scan over the cells produces, for every
//                       non-empty cell, its pillar rank (ascending key order == torch.unique order),
//                       the first sorted position of its points, voxel_coords and the counts;
//                       and bins the pillars by length into the work lists the PFN kernel consumes;
//   3. place          : counting-sort scatter of the row numbers + the point->pillar map (unq_inv);
//   4. pillar_prep    : row numbers inside each pillar are put in ascending order so that every
//                       per-pillar sum runs in the reference CPU path's order (deterministic), and the
//                       pillar mean (scatter_mean, :110) is evaluated in that order; one launch over the
//                       dense per-class lists (thread / warp / CTA per pillar); long pillars also get
//                       their segment entries here.
// No kernel synchronises with the host; the only data-dependent size (P) is read back by the caller.
#include "internal.cuh"

namespace pcp {

STAGE_TABLE(g_stage_vox);      // 0 quantise_count, 1 cell scan, 2 place, 3 pillar_prep

// ------------------------------------------------------------------------------------------------
// 1. quantise + count
// ------------------------------------------------------------------------------------------------
constexpr int kPtsPerThread = 4;   // independent rows per thread: 4 loads, then 4 atomics, then 4 stores in flight

template <bool kVec4>
__global__ void __launch_bounds__(256)
quantise_count_kernel(const float* __restrict__ points, int64_t stride, int64_t n, int32_t frames,
                      pcp_grid g, int32_t* __restrict__ cell, int32_t* __restrict__ key,
                      int32_t* __restrict__ within, int32_t* __restrict__ hdr) {
  STAGE_BEGIN(g_stage_vox, 0);
  pdl_launch_dependents();
  const int64_t base = (int64_t)blockIdx.x * (256 * kPtsPerThread) + threadIdx.x;
  float bf[kPtsPerThread], x[kPtsPerThread], y[kPtsPerThread];
#pragma unroll
  for (int u = 0; u < kPtsPerThread; ++u) {
    const int64_t i = base + u * 256;
    bf[u] = 0.f; x[u] = 0.f; y[u] = 0.f;
    if (i < n) {
      const float* row = points + i * stride;
      if (kVec4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row));
        bf[u] = v.x; x[u] = v.y; y[u] = v.z;
      } else {
        bf[u] = __ldg(row); x[u] = __ldg(row + 1); y[u] = __ldg(row + 2);
      }
    }
  }
  int32_t k[kPtsPerThread], w[kPtsPerThread];
#pragma unroll
  for (int u = 0; u < kPtsPerThread; ++u) {
    k[u] = -1; w[u] = 0;
    if (base + u * 256 < n) {
      bool bad_frame;
      k[u] = point_key(bf[u], x[u], y[u], frames, g, bad_frame);
      if (bad_frame) atomicAdd(&hdr[PCP_COUNT_BAD_FRAME], 1);
    }
  }
  pdl_wait();                  // everything above read the caller's rows only
#pragma unroll
  for (int u = 0; u < kPtsPerThread; ++u)
    if (k[u] >= 0) w[u] = atomicAdd(&cell[k[u]], 1);
#pragma unroll
  for (int u = 0; u < kPtsPerThread; ++u) {
    const int64_t i = base + u * 256;
    if (i < n) { key[i] = k[u]; within[i] = w[u]; }
  }
  STAGE_END(g_stage_vox, 0);
}

// ------------------------------------------------------------------------------------------------
// 2. scan over the cells in two launches: per-tile sums, then every tile adds up the sums of the tiles before it
//    (the whole array of tile sums is a few kB in L2; no spinning, no tile-to-tile dependency chain)
//    packed 64-bit partial: [61:32] sum of counts (points), [31:0] sum of flags (pillars)
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = kScanTileCells / kScanThreads;  // 8

__device__ __forceinline__ void load_tile_counts(const int32_t* __restrict__ cell, int64_t base, int64_t cells,
                                                 int32_t (&c)[kScanItems]) {
  if (base + kScanItems <= cells) {
    const int4 a = *reinterpret_cast<const int4*>(cell + base);
    const int4 b = *reinterpret_cast<const int4*>(cell + base + 4);
    c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) c[j] = (base + j < cells) ? cell[base + j] : 0;
  }
}

// per-tile record written by tile_sums_kernel (16 ints)
constexpr int kTiPoints = 0;     // points in the tile
constexpr int kTiPillars = 1;    // non-empty cells
constexpr int kTiClass = 2;      // [2, 12): pillars per length class
constexpr int kTiLong = 12;      // long pillars
constexpr int kTiSegs = 13;      // segments of the long pillars
constexpr int kTiBig = 14;       // long pillars above kWarpLongMax rows
constexpr int kTiMax = 15;       // largest pillar (max, not a sum)
constexpr int kTiInts = 16;

__global__ void __launch_bounds__(kScanThreads)
tile_sums_kernel(const int32_t* __restrict__ cell, int64_t cells, int32_t nx, int32_t ny, int32_t* __restrict__ tile_info,
                 int32_t* __restrict__ tile_frames) {
  __shared__ int s_acc[kTiInts];
  __shared__ int s_fr;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < kTiInts) s_acc[tid] = 0;
  if (tid == 0) s_fr = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kScanTileCells + (int64_t)tid * kScanItems;
  int32_t c[kScanItems];
  load_tile_counts(cell, base, cells, c);
  int pts = 0, pil = 0, cmax = 0, last_nonempty = -1;
  unsigned long long cls = 0;           // 10 x 6-bit counters: pillars of this thread per class (at most 8 each)
  int nlong = 0, nseg = 0, nbig = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (c[j] > 0) {
      pts += c[j]; ++pil; cmax = max(cmax, c[j]); last_nonempty = j;
      if (c[j] <= kSegRows) cls += 1ull << (6 * class_of(c[j]));
      else { ++nlong; nseg += (c[j] + kSegRows - 1) / kSegRows; nbig += (c[j] > kWarpLongMax) ? 1 : 0; }
    }
  }
  // warp reductions, then one shared-memory atomic per warp and counter
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    pts += __shfl_xor_sync(0xffffffffu, pts, d);
    pil += __shfl_xor_sync(0xffffffffu, pil, d);
    cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
  }
  if (lane == 0) { atomicAdd(&s_acc[kTiPoints], pts); atomicAdd(&s_acc[kTiPillars], pil); atomicMax(&s_acc[kTiMax], cmax); }
#pragma unroll
  for (int k = 0; k < kNumClasses; ++k) {
    int v = (int)((cls >> (6 * k)) & 63ull);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0 && v) atomicAdd(&s_acc[kTiClass + k], v);
  }
  if (__any_sync(0xffffffffu, nlong > 0)) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      nlong += __shfl_xor_sync(0xffffffffu, nlong, d);
      nseg += __shfl_xor_sync(0xffffffffu, nseg, d);
      nbig += __shfl_xor_sync(0xffffffffu, nbig, d);
    }
    if (lane == 0) { atomicAdd(&s_acc[kTiLong], nlong); atomicAdd(&s_acc[kTiSegs], nseg); atomicAdd(&s_acc[kTiBig], nbig); }
  }
  // frame of the last non-empty cell of the tile + 1 (0: empty tile); cells < 2^31 (checked by the host)
  int fr = 0;
  if (last_nonempty >= 0) fr = (int)(((uint32_t)base + (uint32_t)last_nonempty) / ((uint32_t)nx * (uint32_t)ny)) + 1;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) fr = max(fr, __shfl_xor_sync(0xffffffffu, fr, d));
  if (lane == 0 && fr) atomicMax(&s_fr, fr);
  __syncthreads();
  if (tid < kTiInts) tile_info[(int64_t)blockIdx.x * kTiInts + tid] = s_acc[tid];
  if (tid == 0) tile_frames[blockIdx.x] = s_fr;
}

// Second launch: every tile sums the records of the tiles before it (its exclusive prefix of every counter), scans its own
// cells and writes, for every non-empty cell, the pillar rank, the first sorted position, voxel_coords and the pillar's
// work-list entry.  No global atomics: list positions, long-pillar indices and segment ranges come from the prefixes.
__global__ void __launch_bounds__(kScanThreads, 4)
scan_cells_kernel(int32_t* __restrict__ cell, int32_t* __restrict__ cell_rank, int64_t cells, int32_t nx, int32_t ny,
                  const int32_t* __restrict__ tile_info, const int32_t* __restrict__ tile_frames, int32_t* __restrict__ hdr,
                  int32_t* __restrict__ seg_off, int32_t* __restrict__ voxel_coords,
                  int32_t* __restrict__ pillar_count, unsigned long long* __restrict__ lists, const ListOffsets lo,
                  int4* __restrict__ long_table, int32_t* __restrict__ big_list, int64_t scan_tiles) {
  __shared__ int s_cls[kTiInts];                 // running in-tile counters (classes, long, segs, big)
  __shared__ int s_before[kTiInts];              // sums over the tiles before this one (kTiMax: max)
  __shared__ int s_part[kScanThreads / 32][kTiInts];
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ long long s_lo[kNumClasses];
  __shared__ int s_frames[kScanThreads / 32];
  // the tile's non-empty cells, compacted in rank order (phase A fills, phase B walks them with all threads busy)
  __shared__ int32_t s_pcnt[kScanTileCells];     // rows of the pillar
  __shared__ int32_t s_poff[kScanTileCells];     // its first sorted position
  __shared__ uint16_t s_pidx[kScanTileCells];    // its cell inside the tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kTiInts) s_cls[tid] = 0;
  if (tid < kNumClasses) s_lo[tid] = lo.off[tid];
  const int64_t tile = blockIdx.x;
  const int64_t base = tile * kScanTileCells + (int64_t)tid * kScanItems;
  // ---- exclusive prefix of the tile records ----
  {
    int acc[kTiInts];
#pragma unroll
    for (int i = 0; i < kTiInts; ++i) acc[i] = 0;
    int frames = 0;
    if (tile == scan_tiles - 1)
      for (int64_t t = tid; t <= tile; t += kScanThreads) frames = max(frames, __ldg(tile_frames + t));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) frames = max(frames, __shfl_xor_sync(0xffffffffu, frames, d));
    if (lane == 0) s_frames[warp] = frames;
    for (int64_t t = tid; t < tile; t += kScanThreads) {
      const int4* rec = reinterpret_cast<const int4*>(tile_info + t * kTiInts);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 v = __ldg(rec + q);
        acc[4 * q + 0] += v.x; acc[4 * q + 1] += v.y; acc[4 * q + 2] += v.z;
        if (q < 3) acc[4 * q + 3] += v.w; else acc[kTiMax] = max(acc[kTiMax], v.w);
      }
    }
#pragma unroll
    for (int i = 0; i < kTiInts; ++i) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const int o = __shfl_xor_sync(0xffffffffu, acc[i], d);
        acc[i] = (i == kTiMax) ? max(acc[i], o) : acc[i] + o;
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kTiInts; ++i) s_part[warp][i] = acc[i];
    }
  }

  int32_t c[kScanItems];
  load_tile_counts(cell, base, cells, c);
  unsigned long long mine = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) mine += ((unsigned long long)(uint32_t)c[j] << 32) | (c[j] > 0 ? 1ull : 0ull);
  // block exclusive scan of `mine`
  unsigned long long incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (tid < kTiInts) {
    int v = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) v = (tid == kTiMax) ? max(v, s_part[w][tid]) : v + s_part[w][tid];
    s_before[tid] = v;
  }
  unsigned long long warp_excl = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    if (w < warp) warp_excl += s_warp[w];
    tile_total += s_warp[w];
  }
  __syncthreads();
  const int32_t r_tile = s_before[kTiPillars], off_tile = s_before[kTiPoints];     // first rank / sorted position of the tile
  if (tile == scan_tiles - 1) {
    // the last tile knows every total: counts block, list lengths, long-pillar counters
    const int32_t* own = tile_info + tile * kTiInts;
    if (tid == 0) {
      const int32_t P = r_tile + (int32_t)(tile_total & 0xffffffffull);
      const int32_t Nk = off_tile + (int32_t)((tile_total >> 32) & 0x3fffffffull);
      hdr[PCP_COUNT_PILLARS] = P;
      hdr[PCP_COUNT_KEPT] = Nk;
      hdr[PCP_COUNT_MAX_PER_PILLAR] = max(s_before[kTiMax], own[kTiMax]);
      seg_off[P] = Nk;
      hdr[kHdrLongCount] = s_before[kTiLong] + own[kTiLong];
      hdr[kHdrListCount + kSegList] = s_before[kTiSegs] + own[kTiSegs];
      hdr[kHdrBigCount] = s_before[kTiBig] + own[kTiBig];
      int frames = 0;
#pragma unroll
      for (int w = 0; w < kScanThreads / 32; ++w) frames = max(frames, s_frames[w]);
      hdr[PCP_COUNT_FRAMES] = frames;     // max frame index among pillars + 1
    }
    if (tid < kNumClasses) hdr[kHdrListCount + tid] = s_before[kTiClass + tid] + own[kTiClass + tid];
  }
  // ---- phase A: per cell - first sorted position / rank (or -1) as vector stores; non-empty cells -> compact list ----
  {
    const unsigned long long excl0 = warp_excl + (incl - mine);          // inside the tile
    int rl = (int)(excl0 & 0xffffffffull);                               // local rank
    int ol = (int)((excl0 >> 32) & 0x3fffffffull);                       // local sorted position
    int32_t vo[kScanItems], vr[kScanItems];
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      vo[j] = -1; vr[j] = -1;
      if (c[j] > 0) {
        vo[j] = off_tile + ol; vr[j] = r_tile + rl;
        s_pcnt[rl] = c[j]; s_poff[rl] = off_tile + ol; s_pidx[rl] = (uint16_t)(tid * kScanItems + j);
        ++rl; ol += c[j];
      }
    }
    if (base + kScanItems <= cells) {
      *reinterpret_cast<int4*>(cell + base) = make_int4(vo[0], vo[1], vo[2], vo[3]);
      *reinterpret_cast<int4*>(cell + base + 4) = make_int4(vo[4], vo[5], vo[6], vo[7]);
      *reinterpret_cast<int4*>(cell_rank + base) = make_int4(vr[0], vr[1], vr[2], vr[3]);
      *reinterpret_cast<int4*>(cell_rank + base + 4) = make_int4(vr[4], vr[5], vr[6], vr[7]);
    } else {
#pragma unroll
      for (int j = 0; j < kScanItems; ++j)
        if (base + j < cells) { cell[base + j] = vo[j]; cell_rank[base + j] = vr[j]; }
    }
  }
  __syncthreads();
  // ---- phase B: per pillar, consecutive threads = consecutive ranks ----
  const int npil = (int)(tile_total & 0xffffffffull);
  const uint32_t nxy = (uint32_t)nx * (uint32_t)ny;                      // cells < 2^31 (checked by the host)
  const uint32_t tile_cell0 = (uint32_t)(tile * kScanTileCells);
  for (int q = tid; q < npil; q += kScanThreads) {
    const int32_t cnt = s_pcnt[q], off = s_poff[q];
    const int32_t r = r_tile + q;
    const uint32_t idx = tile_cell0 + s_pidx[q];
    const uint32_t b = idx / nxy, rem = idx - b * nxy;
    const uint32_t cx = rem / (uint32_t)ny, cy = rem - cx * (uint32_t)ny;
    seg_off[r] = off;
    // (frame, z = 0, y, x): dynamic_pillar_vfe.py:138-143 after the [0, 3, 2, 1] reorder
    if (voxel_coords) *reinterpret_cast<int4*>(voxel_coords + 4 * (int64_t)r) = make_int4((int)b, 0, (int)cy, (int)cx);
    if (pillar_count) pillar_count[r] = cnt;
    if (cnt <= kSegRows) {
      // the tile's first position in the class list is the sum of the class counters of the tiles before it
      const int k = class_of(cnt);
      const int slot = s_before[kTiClass + k] + atomicAdd(&s_cls[kTiClass + k], 1);
      lists[s_lo[k] + slot] = pack_entry(r, off, cnt);
    } else {
      // long pillar: its index, its segment range (pillar_prep_kernel writes the segment entries)
      const int nseg = (cnt + kSegRows - 1) / kSegRows;
      const int li = s_before[kTiLong] + atomicAdd(&s_cls[kTiLong], 1);
      const int sb = s_before[kTiSegs] + atomicAdd(&s_cls[kTiSegs], nseg);
      long_table[li] = make_int4(r, off, cnt, sb);
      if (cnt > kWarpLongMax) big_list[s_before[kTiBig] + atomicAdd(&s_cls[kTiBig], 1)] = li;
    }
  }
}

// Single-launch form of the two kernels above for key spaces of at most kScanFusedMaxTiles tiles (4 M cells): every tile
// computes its own record, PUBLISHES it - every field stored + 1 into memory the prologue memset zeroed, as four 16-byte
// stores, so a quarter counts as published when none of its four words is zero: no flag, no fence - and then sums the
// records of all tiles before it.  A tile publishes before it waits and tiles are numbered by a ticket in start order
// (every lower-numbered tile is already running), so the wait cannot deadlock.  Saves one launch and the second read of
// the cell array's counts.
__device__ __forceinline__ int4 ld_volatile_int4(const int32_t* p) {
  int4 v;
  asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_int(const int32_t* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kScanThreads, 4)
scan_cells_fused_kernel(int32_t* __restrict__ cell, int32_t* __restrict__ cell_rank, int64_t cells, int32_t nx, int32_t ny,
                        int32_t* __restrict__ tile_info, int32_t* __restrict__ tile_frames, int32_t* __restrict__ hdr,
                        int32_t* __restrict__ seg_off, int32_t* __restrict__ voxel_coords,
                        int32_t* __restrict__ pillar_count, unsigned long long* __restrict__ lists, const ListOffsets lo,
                        int4* __restrict__ long_table, int32_t* __restrict__ big_list, int32_t scan_tiles) {
  __shared__ int s_cls[kTiInts];                 // running in-tile counters (classes, long, segs, big)
  __shared__ int s_rec[kTiInts];                 // this tile's record
  __shared__ int s_before[kTiInts];              // sums over the tiles before this one (kTiMax: max)
  __shared__ int s_part[kScanThreads / 32][kTiInts];
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ long long s_lo[kNumClasses];
  __shared__ int s_frw[kScanThreads / 32];
  __shared__ int s_tile, s_fr;
  __shared__ int32_t s_pcnt[kScanTileCells];     // rows of the pillar
  __shared__ int32_t s_poff[kScanTileCells];     // its first sorted position (local to the tile until the prefix is known)
  __shared__ uint16_t s_pidx[kScanTileCells];    // its cell inside the tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  STAGE_BEGIN(g_stage_vox, 1);
  pdl_launch_dependents();
  pdl_wait();
  if (tid == 0) { s_tile = atomicAdd(&hdr[kHdrScanTicket], 1); s_fr = 0; }
  if (tid < kTiInts) { s_cls[tid] = 0; s_rec[tid] = 0; }
  if (tid < kNumClasses) s_lo[tid] = lo.off[tid];
  __syncthreads();
  const int tile = s_tile;
  const int64_t base = (int64_t)tile * kScanTileCells + (int64_t)tid * kScanItems;
  int32_t c[kScanItems];
  load_tile_counts(cell, base, cells, c);
  // ---- this tile's record (as tile_sums_kernel) ----
  unsigned long long mine = 0;
  {
    int pts = 0, pil = 0, cmax = 0, last_nonempty = -1, nlong = 0, nseg = 0, nbig = 0;
    unsigned long long cls = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      if (c[j] > 0) {
        pts += c[j]; ++pil; cmax = max(cmax, c[j]); last_nonempty = j;
        if (c[j] <= kSegRows) cls += 1ull << (6 * class_of(c[j]));
        else { ++nlong; nseg += (c[j] + kSegRows - 1) / kSegRows; nbig += (c[j] > kWarpLongMax) ? 1 : 0; }
      }
    }
    mine = ((unsigned long long)(uint32_t)pts << 32) | (unsigned long long)(uint32_t)pil;
    const int wpts = __reduce_add_sync(0xffffffffu, pts), wpil = __reduce_add_sync(0xffffffffu, pil);
    const int wmax = __reduce_max_sync(0xffffffffu, cmax);
    if (wpil) {
      if (lane == 0) { atomicAdd(&s_rec[kTiPoints], wpts); atomicAdd(&s_rec[kTiPillars], wpil); atomicMax(&s_rec[kTiMax], wmax); }
#pragma unroll
      for (int k = 0; k < kNumClasses; ++k) {
        const int v = __reduce_add_sync(0xffffffffu, (int)((cls >> (6 * k)) & 63ull));
        if (lane == 0 && v) atomicAdd(&s_rec[kTiClass + k], v);
      }
    }
    if (__any_sync(0xffffffffu, nlong > 0)) {
      const int a = __reduce_add_sync(0xffffffffu, nlong), b = __reduce_add_sync(0xffffffffu, nseg);
      const int d = __reduce_add_sync(0xffffffffu, nbig);
      if (lane == 0) { atomicAdd(&s_rec[kTiLong], a); atomicAdd(&s_rec[kTiSegs], b); atomicAdd(&s_rec[kTiBig], d); }
    }
    int fr = 0;
    if (last_nonempty >= 0) fr = (int)(((uint32_t)base + (uint32_t)last_nonempty) / ((uint32_t)nx * (uint32_t)ny)) + 1;
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
  // ---- publish ----
  if (tid < 4) {
    const int4 v = make_int4(s_rec[4 * tid] + 1, s_rec[4 * tid + 1] + 1, s_rec[4 * tid + 2] + 1, s_rec[4 * tid + 3] + 1);
    *reinterpret_cast<int4*>(tile_info + (int64_t)tile * kTiInts + 4 * tid) = v;
  }
  if (tid == 4) tile_frames[tile] = s_fr + 1;
  unsigned long long warp_excl = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    if (w < warp) warp_excl += s_warp[w];
    tile_total += s_warp[w];
  }
  // ---- local ranks / positions; non-empty cells -> compact per-pillar arrays ----
  const unsigned long long excl0 = warp_excl + (incl - mine);
  const int rl0 = (int)(excl0 & 0xffffffffull), ol0 = (int)(excl0 >> 32);     // local rank / position of this thread's first pillar
  {
    int rl = rl0, ol = ol0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      if (c[j] > 0) {
        s_pcnt[rl] = c[j]; s_poff[rl] = ol; s_pidx[rl] = (uint16_t)(tid * kScanItems + j);
        ++rl; ol += c[j];
      }
    }
  }
  // ---- sum of the records of all tiles before this one ----
  {
    int acc[kTiInts];
#pragma unroll
    for (int i = 0; i < kTiInts; ++i) acc[i] = 0;
    for (int t0 = tid; t0 < tile; t0 += kScanThreads * 2) {
      int4 v[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + u * kScanThreads;
        if (t < tile) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[u][q] = ld_volatile_int4(tile_info + (int64_t)t * kTiInts + 4 * q);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = t0 + u * kScanThreads;
        if (t < tile) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            while (v[u][q].x == 0 || v[u][q].y == 0 || v[u][q].z == 0 || v[u][q].w == 0) {
              __nanosleep(40);
              v[u][q] = ld_volatile_int4(tile_info + (int64_t)t * kTiInts + 4 * q);
            }
            acc[4 * q + 0] += v[u][q].x - 1; acc[4 * q + 1] += v[u][q].y - 1; acc[4 * q + 2] += v[u][q].z - 1;
            if (q < 3) acc[4 * q + 3] += v[u][q].w - 1; else acc[kTiMax] = max(acc[kTiMax], v[u][q].w - 1);
          }
        }
      }
    }
    int frames = 0;
    if (tile == scan_tiles - 1) {
      for (int t = tid; t < tile; t += kScanThreads) {
        int f = ld_volatile_int(tile_frames + t);
        while (f == 0) { __nanosleep(40); f = ld_volatile_int(tile_frames + t); }
        frames = max(frames, f - 1);
      }
      frames = __reduce_max_sync(0xffffffffu, frames);
    }
#pragma unroll
    for (int i = 0; i < kTiInts; ++i)
      acc[i] = (i == kTiMax) ? __reduce_max_sync(0xffffffffu, acc[i]) : __reduce_add_sync(0xffffffffu, acc[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kTiInts; ++i) s_part[warp][i] = acc[i];
      s_frw[warp] = frames;
    }
  }
  __syncthreads();
  if (tid < kTiInts) {
    int v = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) v = (tid == kTiMax) ? max(v, s_part[w][tid]) : v + s_part[w][tid];
    s_before[tid] = v;
  }
  __syncthreads();
  const int32_t r_tile = s_before[kTiPillars], off_tile = s_before[kTiPoints];
  if (tile == scan_tiles - 1) {
    if (tid == 0) {
      const int32_t P = r_tile + s_rec[kTiPillars];
      const int32_t Nk = off_tile + s_rec[kTiPoints];
      hdr[PCP_COUNT_PILLARS] = P;
      hdr[PCP_COUNT_KEPT] = Nk;
      hdr[PCP_COUNT_MAX_PER_PILLAR] = max(s_before[kTiMax], s_rec[kTiMax]);
      seg_off[P] = Nk;
      hdr[kHdrLongCount] = s_before[kTiLong] + s_rec[kTiLong];
      hdr[kHdrListCount + kSegList] = s_before[kTiSegs] + s_rec[kTiSegs];
      hdr[kHdrBigCount] = s_before[kTiBig] + s_rec[kTiBig];
      int frames = s_fr;
#pragma unroll
      for (int w = 0; w < kScanThreads / 32; ++w) frames = max(frames, s_frw[w]);
      hdr[PCP_COUNT_FRAMES] = frames;
    }
    if (tid < kNumClasses) hdr[kHdrListCount + tid] = s_before[kTiClass + tid] + s_rec[kTiClass + tid];
  }
  // ---- per cell: first sorted position / rank (or -1) as vector stores ----
  {
    int32_t vo[kScanItems], vr[kScanItems];
    int rl = r_tile + rl0, ol = off_tile + ol0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      vo[j] = -1; vr[j] = -1;
      if (c[j] > 0) { vo[j] = ol; vr[j] = rl; ++rl; ol += c[j]; }
    }
    if (base + kScanItems <= cells) {
      *reinterpret_cast<int4*>(cell + base) = make_int4(vo[0], vo[1], vo[2], vo[3]);
      *reinterpret_cast<int4*>(cell + base + 4) = make_int4(vo[4], vo[5], vo[6], vo[7]);
      *reinterpret_cast<int4*>(cell_rank + base) = make_int4(vr[0], vr[1], vr[2], vr[3]);
      *reinterpret_cast<int4*>(cell_rank + base + 4) = make_int4(vr[4], vr[5], vr[6], vr[7]);
    } else {
#pragma unroll
      for (int j = 0; j < kScanItems; ++j)
        if (base + j < cells) { cell[base + j] = vo[j]; cell_rank[base + j] = vr[j]; }
    }
  }
  // ---- per pillar, consecutive threads = consecutive ranks (as scan_cells_kernel phase B) ----
  const int npil = (int)(tile_total & 0xffffffffull);
  const uint32_t nxy = (uint32_t)nx * (uint32_t)ny;
  const uint32_t tile_cell0 = (uint32_t)tile * (uint32_t)kScanTileCells;
  for (int q = tid; q < npil; q += kScanThreads) {
    const int32_t cnt = s_pcnt[q], off = off_tile + s_poff[q];
    const int32_t r = r_tile + q;
    const uint32_t idx = tile_cell0 + s_pidx[q];
    const uint32_t b = idx / nxy, rem = idx - b * nxy;
    const uint32_t cx = rem / (uint32_t)ny, cy = rem - cx * (uint32_t)ny;
    seg_off[r] = off;
    if (voxel_coords) *reinterpret_cast<int4*>(voxel_coords + 4 * (int64_t)r) = make_int4((int)b, 0, (int)cy, (int)cx);
    if (pillar_count) pillar_count[r] = cnt;
    if (cnt <= kSegRows) {
      const int k = class_of(cnt);
      const int slot = s_before[kTiClass + k] + atomicAdd(&s_cls[kTiClass + k], 1);
      lists[s_lo[k] + slot] = pack_entry(r, off, cnt);
    } else {
      const int nseg = (cnt + kSegRows - 1) / kSegRows;
      const int li = s_before[kTiLong] + atomicAdd(&s_cls[kTiLong], 1);
      const int sb = s_before[kTiSegs] + atomicAdd(&s_cls[kTiSegs], nseg);
      long_table[li] = make_int4(r, off, cnt, sb);
      if (cnt > kWarpLongMax) big_list[s_before[kTiBig] + atomicAdd(&s_cls[kTiBig], 1)] = li;
    }
  }
  STAGE_END(g_stage_vox, 1);
}

// ------------------------------------------------------------------------------------------------
// 3. counting-sort scatter + point -> pillar map
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
place_kernel(const int32_t* __restrict__ key, const int32_t* __restrict__ within,
             const int32_t* __restrict__ cell, const int32_t* __restrict__ cell_rank, int64_t n,
             int32_t* __restrict__ sorted_idx, int32_t* __restrict__ point_pillar,
             const int32_t* __restrict__ hdr, int32_t* __restrict__ counts_out) {
  STAGE_BEGIN(g_stage_vox, 2);
  pdl_launch_dependents();
  pdl_wait();
  const int64_t base = (int64_t)blockIdx.x * (256 * kPtsPerThread) + threadIdx.x;
  if (base == 0 && counts_out) {
#pragma unroll
    for (int j = 0; j < PCP_COUNTS_LEN; ++j) counts_out[j] = hdr[j];
  }
  int32_t k[kPtsPerThread], w[kPtsPerThread], o[kPtsPerThread];
#pragma unroll
  for (int u = 0; u < kPtsPerThread; ++u) {
    const int64_t i = base + u * 256;
    k[u] = (i < n) ? key[i] : -1;
    w[u] = (i < n) ? within[i] : 0;
  }
  // one random 4-byte read per point: the cell's first sorted position (the scan left it in the histogram array)
#pragma unroll
  for (int u = 0; u < kPtsPerThread; ++u) o[u] = (k[u] >= 0) ? __ldg(cell + k[u]) : -1;
#pragma unroll
  for (int u = 0; u < kPtsPerThread; ++u) {
    const int64_t i = base + u * 256;
    if (k[u] >= 0) sorted_idx[o[u] + w[u]] = (int32_t)i;
    if (point_pillar && i < n) point_pillar[i] = (k[u] >= 0) ? __ldg(cell_rank + k[u]) : -1;
  }
  STAGE_END(g_stage_vox, 2);
}

// ------------------------------------------------------------------------------------------------
// 4. pillar preparation: ascending row order inside every pillar + the pillar mean, ONE launch.
//    Row numbers inside a pillar arrive in atomic order; they are put in ascending order so that every
//    per-pillar sum runs in the reference CPU path's order (index_add_ walks the rows sequentially), which
//    makes scatter_mean (dynamic_pillar_vfe.py:110) bit-reproducible.  The mean is evaluated here - one
//    sequential fp32 sum per pillar, divided by the count - so that the PFN kernel streams each row once.
//    Every CTA walks three work lists in turn (longest items first):
//      long  pillars (> kSegRows rows)  : one CTA per pillar, rank-by-counting / bitonic sort in shared memory;
//                                         also emits the pillar's segment entries and arms its max accumulator
//      mid   pillars (9 .. 32 rows)     : one warp per pillar, rank by counting through shuffles
//      short pillars (1 .. 8 rows)      : one thread per pillar, 19-comparator network in registers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(int32_t& a, int32_t& b) {
  const int32_t lo = min(a, b), hi = max(a, b);
  a = lo; b = hi;
}

template <bool kVec4>
__device__ __forceinline__ void load_xyz(const float* __restrict__ points, int64_t stride, int32_t idx, float& x, float& y,
                                         float& z) {
  if (points == nullptr) { x = 0.f; y = 0.f; z = 0.f; return; }   // callers that only need the row order (no xyz mean)
  const float* row = points + (int64_t)idx * stride;
  if (kVec4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row));
    x = v.y; y = v.z; z = v.w;
  } else {
    x = __ldg(row + 1); y = __ldg(row + 2); z = __ldg(row + 3);
  }
}

// the pillar's cell, packed next to its mean: cx | cy << 16 (every point of a pillar quantises to the same cell)
__device__ __forceinline__ float pack_cell(float x, float y, const pcp_grid& g) {
  const unsigned cx = (unsigned)(int)quantise(x, g.range_min_x, g.voxel_x);
  const unsigned cy = (unsigned)(int)quantise(y, g.range_min_y, g.voxel_y);
  return __uint_as_float((cx & 0xffffu) | (cy << 16));
}

// one pillar of at most NMAX rows (one length class) per thread: order the row numbers (a sum of two values does not
// depend on the order, so classes 0 and 1 are left as they are), gather xyz, sum in ascending row order, divide
template <int NMAX, bool kVec4>
__device__ __forceinline__ void prep_short_one(const float* __restrict__ points, int64_t stride, const pcp_grid& g,
                                               unsigned long long entry, int32_t* __restrict__ sorted_idx,
                                               float4* __restrict__ mean) {
  int r, off, n;
  unpack_entry(entry, r, off, n);
  int32_t v[NMAX];
#pragma unroll
  for (int j = 0; j < NMAX; ++j) v[j] = (j < n) ? sorted_idx[off + j] : 0x7fffffff;
  if (NMAX > 2) {
#pragma unroll
    for (int i = 1; i < NMAX; ++i)
#pragma unroll
      for (int j = i; j > 0; --j) cswap(v[j - 1], v[j]);
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
      if (j < n) sorted_idx[off + j] = v[j];
  }
  float x[NMAX], y[NMAX], z[NMAX];
#pragma unroll
  for (int j = 0; j < NMAX; ++j)
    if (j < n) load_xyz<kVec4>(points, stride, v[j], x[j], y[j], z[j]);
  float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
  for (int j = 0; j < NMAX; ++j)
    if (j < n) { ax = __fadd_rn(ax, x[j]); ay = __fadd_rn(ay, y[j]); az = __fadd_rn(az, z[j]); }
  const float cnt = (float)n;
  mean[r] = make_float4(__fdiv_rn(ax, cnt), __fdiv_rn(ay, cnt), __fdiv_rn(az, cnt), pack_cell(x[0], y[0], g));
}

// W lanes per pillar (two length classes of at most W rows, walked back to back), items [w_begin, w_end) of the
// concatenated lists: rank by counting against the warp's shared-memory copy, lanes 0, 1, 2 of the group run the
// sequential sums of x, y, z.  warp_smem: 4 x 32 words of this warp.
// kRec: the pillar's rows are {x, y, z, row number} records at srec[off ..] (binned path) instead of row numbers in sorted_idx
template <int W, bool kVec4, bool kRec>
__device__ __forceinline__ void prep_mid(const float* __restrict__ points, int64_t stride, const float4* __restrict__ srec,
                                         const pcp_grid& g,
                                         const unsigned long long* __restrict__ list_a, int count_a,
                                         const unsigned long long* __restrict__ list_b, int count_b, int w_begin, int w_end,
                                         int32_t* __restrict__ sorted_idx, float4* __restrict__ mean, int32_t* warp_smem) {
  constexpr int kPer = 32 / W;                          // pillars per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / W, ln = lane & (W - 1);
  const int total = min(count_a + count_b, w_end);
  int32_t* sv = warp_smem + sub * W;                    // [32] row numbers
  float* sx = reinterpret_cast<float*>(warp_smem) + 32 + sub * W;   // [3][32] x | y | z in ascending row order
  for (int w0 = w_begin + warp * kPer; w0 < total; w0 += (blockDim.x >> 5) * kPer) {
    const int w = w0 + sub;
    int r = 0, off = 0, n = 0;
    if (w < total) unpack_entry(__ldg(w < count_a ? list_a + w : list_b + (w - count_a)), r, off, n);
    float4 rc = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
    if (kRec && ln < n) rc = __ldcg(srec + off + ln);
    const int32_t v = kRec ? __float_as_int(rc.w) : ((ln < n) ? sorted_idx[off + ln] : 0x7fffffff);
    sv[ln] = v;
    __syncwarp();
    int rank = 0;
#pragma unroll
    for (int i = 0; i < W; i += 4) {
      const int4 t = *reinterpret_cast<const int4*>(sv + i);       // padding lanes hold INT_MAX: never smaller
      rank += (t.x < v) + (t.y < v) + (t.z < v) + (t.w < v);
    }
    float x = 0.f, y = 0.f, z = 0.f;
    if (ln < n) {
      sorted_idx[off + rank] = v;
      if (kRec) { x = rc.x; y = rc.y; z = rc.z; }
      else load_xyz<kVec4>(points, stride, v, x, y, z);
      sx[rank] = x; sx[32 + rank] = y; sx[64 + rank] = z;
    }
    __syncwarp();
    float acc = 0.f;
    if (ln < 3) {
      const float* src = sx + 32 * ln;
      for (int i = 0; i < n; ++i) acc = __fadd_rn(acc, src[i]);
      acc = __fdiv_rn(acc, (float)max(n, 1));
    }
    const float my = __shfl_sync(0xffffffffu, acc, 1, W), mz = __shfl_sync(0xffffffffu, acc, 2, W);
    if (ln == 0 && w < total) mean[r] = make_float4(acc, my, mz, pack_cell(sx[0], sx[32], g));
    __syncwarp();
  }
}

constexpr int kPrepThreads = 256;
constexpr int kPrepWarps = kPrepThreads / 32;

constexpr int kCountSortMax = 1024;   // CTA path: ranked by counting up to this size, bitonic network above
constexpr int kSumChunk = 256;        // CTA path: rows staged per step of the sequential sum

// Shared memory of a CTA (24 KB, static; measured: 12 KB is no faster, 41 KB - which drops the L1 to its 92 KB configuration with
// four resident CTAs - costs 16 us on the bench batch): either 8 per-warp regions of 3 x 256 words (warp path), or the CTA path's
// row-number array [kBigSegMax] followed by 2048 words of scratch (counting-sort output / double-buffered xyz chunks).
struct PrepSmem {
  union {
    int32_t warp_words[kPrepWarps][3][kWarpLongMax];
    struct { int32_t s[kBigSegMax]; int32_t scratch[2048]; } cta;
  };
};

template <bool kVec4, bool kRec>
__global__ void __launch_bounds__(kPrepThreads, 4)
pillar_prep_kernel(const float* __restrict__ points, int64_t stride, const float4* __restrict__ srec, int32_t* __restrict__ hdr,
                   unsigned long long* __restrict__ lists, const ListOffsets lo, const int4* __restrict__ long_table,
                   const int32_t* __restrict__ big_list, int32_t* __restrict__ sorted_idx, float4* __restrict__ mean,
                   float4* __restrict__ long_mean, unsigned* __restrict__ long_acc, const pcp_grid g, int phase_mask,
                   int32_t* __restrict__ tmp) {
  __shared__ __align__(16) PrepSmem sm;
  __shared__ float s_red[3][8];
  __shared__ int s_red_i[kPrepWarps];
  __shared__ int s_flag;
  __shared__ int s_chunk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  STAGE_BEGIN(g_stage_vox, 3);
  pdl_launch_dependents();
  pdl_wait();
  const int nlong = hdr[kHdrLongCount];
  // Work is handed out in CTA-sized chunks through tickets: CTAs that spent time on a big pillar take fewer chunks.
  auto grab = [&](int which, int chunk) {
    __syncthreads();
    if (tid == 0) s_chunk = atomicAdd(&hdr[kHdrPrepTicket + which], chunk);
    __syncthreads();
    return s_chunk;
  };

  // ---------------- long pillars, CTA path (> kWarpLongMax rows) ----------------
  const int nbig = (phase_mask & 1) ? hdr[kHdrBigCount] : 0;
  for (int bi = blockIdx.x; bi < nbig; bi += gridDim.x) {
    const int li = big_list[bi];
    const int4 e = long_table[li];
    const int32_t r = e.x, off = e.y, n = e.z, sb = e.w;
    const int nseg = (n + kSegRows - 1) / kSegRows;
    for (int i = tid; i < nseg; i += kPrepThreads)
      lists[lo.off[kSegList] + sb + i] = pack_entry(li, off + i * kSegRows, min(kSegRows, n - i * kSegRows));
    for (int i = tid; i < 96; i += kPrepThreads) long_acc[(int64_t)li * 96 + i] = kAccInit;
    int32_t* s = sm.cta.s;
    // sort `cnt` (<= kBigSegMax) row numbers, fetched by `src(i)`, ascending in s[]
    auto sort_rows = [&](int cnt, auto src) {
      if (cnt <= kCountSortMax) {
        int32_t* s2 = sm.cta.scratch;
        const int n4 = (cnt + 3) & ~3;
        for (int i = tid; i < n4; i += kPrepThreads) s[i] = (i < cnt) ? src(i) : 0x7fffffff;
        __syncthreads();
        for (int i = tid; i < cnt; i += kPrepThreads) {
          const int32_t v = s[i];
          int rank = 0;
          for (int j = 0; j < n4; j += 4) {
            const int4 t = *reinterpret_cast<const int4*>(s + j);
            rank += (t.x < v) + (t.y < v) + (t.z < v) + (t.w < v);
          }
          s2[rank] = v;
        }
        __syncthreads();
        for (int i = tid; i < cnt; i += kPrepThreads) s[i] = s2[i];
        __syncthreads();
      } else {
        int m = 2048;
        while (m < cnt) m <<= 1;
        for (int i = tid; i < m; i += kPrepThreads) s[i] = (i < cnt) ? src(i) : 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1) {
          for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += kPrepThreads) {
              const int p2 = i ^ j;
              if (p2 > i) {
                const int32_t va = s[i], vb = s[p2];
                const bool up = (i & k) == 0;
                if ((va > vb) == up) { s[i] = vb; s[p2] = va; }
              }
            }
            __syncthreads();
          }
        }
      }
    };
    // the sorted rows s[0 .. cnt) go to sorted_idx[out_off ..); their x, y, z are added to the running sums IN ORDER (warps 0, 1, 2
    // lane 0 own the three accumulators), rows staged kSumChunk at a time (double buffered: one barrier per chunk)
    float acc = 0.f;
    float cellw = 0.f;
    auto sum_rows = [&](int cnt, int out_off, bool first) {
      float* ch = reinterpret_cast<float*>(sm.cta.scratch);      // [2][3][kSumChunk]
      for (int c0 = 0; c0 < cnt; c0 += kSumChunk) {
        float* buf = ch + ((c0 / kSumChunk) & 1) * (3 * kSumChunk);
        const int i = c0 + tid;
        if (i < cnt) {
          const int32_t idx = s[i];
          sorted_idx[out_off + i] = idx;
          float x, y, z;
          load_xyz<kVec4>(points, stride, idx, x, y, z);
          if (first && i == 0) cellw = pack_cell(x, y, g);
          buf[tid] = x; buf[kSumChunk + tid] = y; buf[2 * kSumChunk + tid] = z;
        }
        __syncthreads();
        if (warp < 3 && lane == 0) {
          const float* srcv = buf + warp * kSumChunk;
          const int m = min(kSumChunk, cnt - c0);
          for (int q = 0; q < m; ++q) acc = __fadd_rn(acc, srcv[q]);
        }
      }
      __syncthreads();                                           // s[] and the staging buffers are free again
    };
    auto row_at = [&](int i) { return kRec ? __float_as_int(__ldcg(srec + off + i).w) : sorted_idx[off + i]; };
    bool exact = true;
    if (n <= kBigSegMax) {
      sort_rows(n, row_at);
      sum_rows(n, off, true);
    } else {
      // Giant pillar.  Row numbers are distinct integers, so a bucket of 4096 consecutive row numbers holds at most 4096 of the
      // pillar's rows: counting sort by bucket (row >> 12) into the scratch array `tmp`, then whole buckets are taken
      // together while they fit, ordered in shared memory like a <= 4096-row pillar, and summed - one running sum over
      // the whole pillar, in ascending row order, as the reference's CPU scatter_mean does.  Covers row numbers below
      // 4096 * 4096 = 16.7 M; above that the sum falls back to arrival order (tolerance).
      for (int i = tid; i < kBigSegMax; i += kPrepThreads) s[i] = 0;
      if (tid == 0) s_flag = 0;
      __syncthreads();
      for (int i = tid; i < n; i += kPrepThreads) {
        const int bkt = row_at(i) >> 12;
        if (bkt < kBigSegMax) atomicAdd(&s[bkt], 1); else s_flag = 1;
      }
      __syncthreads();
      exact = (s_flag == 0) && (tmp != nullptr);
      if (exact) {
        // in-place exclusive scan of the 4096 bucket counts (16 per thread): s[b] = first position of bucket b = its cursor
        {
          int v[16], run = 0;
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = s[tid * 16 + j]; run += v[j]; }
          int incl = run;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
          }
          if (lane == 31) s_red_i[warp] = incl;
          __syncthreads();
          int base = incl - run;
          for (int w = 0; w < warp; ++w) base += s_red_i[w];
#pragma unroll
          for (int j = 0; j < 16; ++j) { s[tid * 16 + j] = base; base += v[j]; }
        }
        __syncthreads();
        for (int i = tid; i < n; i += kPrepThreads) {
          const int32_t row = row_at(i);
          tmp[off + atomicAdd(&s[row >> 12], 1)] = row;
        }
        __syncthreads();
        // tmp[off ..] now holds the rows grouped by bucket, buckets ascending.  Windows of up to 4096 rows, cut back to the last
        // bucket boundary inside them (the rows of the bucket that continues behind the window wait for the next one)
        bool first = true;
        for (int p0 = 0; p0 < n;) {
          int cnt = min(kBigSegMax, n - p0);
          if (p0 + cnt < n) {
            const int next_bkt = __ldcg(tmp + off + p0 + cnt) >> 12;
            if (tid == 0) s_flag = 0;
            __syncthreads();
            int mine = 0;
            for (int i = tid; i < cnt; i += kPrepThreads) mine += ((__ldcg(tmp + off + p0 + i) >> 12) == next_bkt) ? 1 : 0;
            mine = __reduce_add_sync(0xffffffffu, mine);
            if (lane == 0 && mine) atomicAdd(&s_flag, mine);
            __syncthreads();
            cnt -= s_flag;            // > 0: a bucket holds at most 4096 rows and one of them lies behind the window
            __syncthreads();
          }
          sort_rows(cnt, [&](int i) { return __ldcg(tmp + off + p0 + i); });
          sum_rows(cnt, off + p0, first);
          first = false;
          p0 += cnt;
        }
      } else {
        // arrival order, strided partial sums + tree (within tolerance, not order-canonical)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int i = tid; i < n; i += kPrepThreads) {
          float x, y, z;
          if (kRec) {
            const float4 rc = __ldcg(srec + off + i);
            x = rc.x; y = rc.y; z = rc.z;
            sorted_idx[off + i] = __float_as_int(rc.w);
          } else {
            load_xyz<kVec4>(points, stride, sorted_idx[off + i], x, y, z);
          }
          a0 += x; a1 += y; a2 += z;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          a0 += __shfl_xor_sync(0xffffffffu, a0, d); a1 += __shfl_xor_sync(0xffffffffu, a1, d); a2 += __shfl_xor_sync(0xffffffffu, a2, d);
        }
        if (lane == 0) { s_red[0][warp] = a0; s_red[1][warp] = a1; s_red[2][warp] = a2; }
        __syncthreads();
        if (tid < 3) {
          float t = 0.f;
          for (int w = 0; w < kPrepWarps; ++w) t += s_red[tid][w];
          s_red[tid][0] = __fdiv_rn(t, (float)n);
        }
        if (tid == 0) {
          float x, y, z;
          if (kRec) { const float4 rc = __ldcg(srec + off); x = rc.x; y = rc.y; }
          else load_xyz<kVec4>(points, stride, sorted_idx[off], x, y, z);
          s_red[0][1] = pack_cell(x, y, g);
        }
        __syncthreads();
      }
    }
    float mx, my, mz;
    if (exact) {
      if (warp < 3 && lane == 0) s_red[warp][0] = __fdiv_rn(acc, (float)n);
      if (tid == 0) s_red[0][1] = cellw;
      __syncthreads();
    }
    mx = s_red[0][0]; my = s_red[1][0]; mz = s_red[2][0];
    if (tid == 0) {
      const float cw = s_red[0][1];
      long_mean[li] = make_float4(mx, my, mz, cw);
      mean[r] = make_float4(mx, my, mz, cw);
    }
    __syncthreads();
  }

  // ---------------- long pillars, warp path (33 .. kWarpLongMax rows): one warp per pillar ----------------
  for (int l0 = (phase_mask & 2) ? grab(0, kPrepWarps) : nlong; l0 < nlong; l0 = grab(0, kPrepWarps)) {
    const int li = l0 + warp;
    if (li >= nlong) continue;
    const int4 e = long_table[li];
    const int32_t r = e.x, off = e.y, n = e.z, sb = e.w;
    if (n > kWarpLongMax) continue;
    const int nseg = (n + kSegRows - 1) / kSegRows;
    for (int i = lane; i < nseg; i += 32)
      lists[lo.off[kSegList] + sb + i] = pack_entry(li, off + i * kSegRows, min(kSegRows, n - i * kSegRows));
    for (int i = lane; i < 96; i += 32) long_acc[(int64_t)li * 96 + i] = kAccInit;
    int32_t* s0 = sm.warp_words[warp][0];
    int32_t* s1 = sm.warp_words[warp][1];
    const int n4 = (n + 3) & ~3;
    int32_t v[kWarpLongMax / 32];
    float x[kWarpLongMax / 32], y[kWarpLongMax / 32], z[kWarpLongMax / 32];
#pragma unroll
    for (int k = 0; k < kWarpLongMax / 32; ++k) {
      const int i = k * 32 + lane;
      if (kRec) {
        float4 rc = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fffffff));
        if (i < n) rc = __ldcg(srec + off + i);
        v[k] = __float_as_int(rc.w); x[k] = rc.x; y[k] = rc.y; z[k] = rc.z;
      } else {
        v[k] = (i < n) ? sorted_idx[off + i] : 0x7fffffff;
      }
      if (i < n4) s0[i] = v[k];
    }
    __syncwarp();
    int rank[kWarpLongMax / 32];
#pragma unroll
    for (int k = 0; k < kWarpLongMax / 32; ++k) rank[k] = 0;
    for (int j = 0; j < n4; j += 4) {
      const int4 t = *reinterpret_cast<const int4*>(s0 + j);
#pragma unroll
      for (int k = 0; k < kWarpLongMax / 32; ++k)
        rank[k] += (t.x < v[k]) + (t.y < v[k]) + (t.z < v[k]) + (t.w < v[k]);
    }
    float* fx = reinterpret_cast<float*>(sm.warp_words[warp][0]);
    float* fy = reinterpret_cast<float*>(sm.warp_words[warp][2]);
    float* fz = reinterpret_cast<float*>(sm.warp_words[warp][1]);
    if (kRec) {
      // every lane holds its own rows' coordinates: they go straight to their ranks
      __syncwarp();                     // every lane has finished reading s0 (= fx)
#pragma unroll
      for (int k = 0; k < kWarpLongMax / 32; ++k)
        if (k * 32 + lane < n) {
          sorted_idx[off + rank[k]] = v[k];
          fx[rank[k]] = x[k]; fy[rank[k]] = y[k]; fz[rank[k]] = z[k];
        }
      __syncwarp();
    } else {
#pragma unroll
      for (int k = 0; k < kWarpLongMax / 32; ++k)
        if (k * 32 + lane < n) s1[rank[k]] = v[k];
      __syncwarp();
#pragma unroll
      for (int k = 0; k < kWarpLongMax / 32; ++k) {
        const int i = k * 32 + lane;
        if (i < n) {
          const int32_t idx = s1[i];
          sorted_idx[off + i] = idx;
          load_xyz<kVec4>(points, stride, idx, x[k], y[k], z[k]);
        }
      }
      __syncwarp();                     // every lane has read its sorted row numbers: s1 can be reused for z
#pragma unroll
      for (int k = 0; k < kWarpLongMax / 32; ++k) {
        const int i = k * 32 + lane;
        if (i < n) { fx[i] = x[k]; fy[i] = y[k]; fz[i] = z[k]; }
      }
      __syncwarp();
    }
    float acc = 0.f;
    if (lane < 3) {
      const float* src = lane == 0 ? fx : (lane == 1 ? fy : fz);
      for (int i = 0; i < n; ++i) acc = __fadd_rn(acc, src[i]);
      acc = __fdiv_rn(acc, (float)n);
    }
    const float my = __shfl_sync(0xffffffffu, acc, 1), mz = __shfl_sync(0xffffffffu, acc, 2);
    if (lane == 0) {
      const float cw = pack_cell(fx[0], fy[0], g);
      long_mean[li] = make_float4(acc, my, mz, cw);
      mean[r] = make_float4(acc, my, mz, cw);
    }
    __syncwarp();
  }


  // ---------------- mid pillars: classes 6, 7 (9..16 rows): half a warp per pillar; 8, 9 (17..32 rows): a warp ----------------
  if (phase_mask & 4) {
    {
      const int ca = hdr[kHdrListCount + 6], cb = hdr[kHdrListCount + 7];
      constexpr int kChunk = kPrepWarps * 2 * 4;          // 4 rounds of the CTA
      for (int w0 = grab(1, kChunk); w0 < ca + cb; w0 = grab(1, kChunk))
        prep_mid<16, kVec4, kRec>(points, stride, srec, g, lists + lo.off[6], ca, lists + lo.off[7], cb, w0, w0 + kChunk, sorted_idx, mean,
                            &sm.warp_words[warp][0][0]);
    }
    {
      const int ca = hdr[kHdrListCount + 8], cb = hdr[kHdrListCount + 9];
      constexpr int kChunk = kPrepWarps * 4;
      for (int w0 = grab(2, kChunk); w0 < ca + cb; w0 = grab(2, kChunk))
        prep_mid<32, kVec4, kRec>(points, stride, srec, g, lists + lo.off[8], ca, lists + lo.off[9], cb, w0, w0 + kChunk, sorted_idx, mean,
                            &sm.warp_words[warp][0][0]);
    }
  }

  // ---------------- short pillars: classes 0..5 (1..8 rows), one thread per pillar, all classes in one index space ----------------
  if (phase_mask & 8) {
    int pre[7];
    pre[0] = 0;
#pragma unroll
    for (int k = 0; k <= 5; ++k) pre[k + 1] = pre[k] + hdr[kHdrListCount + k];
    constexpr int kChunk = kPrepThreads * 2;
    for (int w0 = grab(3, kChunk); w0 < pre[6]; w0 = grab(3, kChunk)) {
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        const int w = w0 + u * kPrepThreads + tid;
        if (w >= pre[6]) continue;
        if (w < pre[1]) prep_short_one<1, kVec4>(points, stride, g, __ldg(lists + lo.off[0] + w), sorted_idx, mean);
        else if (w < pre[2]) prep_short_one<2, kVec4>(points, stride, g, __ldg(lists + lo.off[1] + (w - pre[1])), sorted_idx, mean);
        else if (w < pre[3]) prep_short_one<3, kVec4>(points, stride, g, __ldg(lists + lo.off[2] + (w - pre[2])), sorted_idx, mean);
        else if (w < pre[4]) prep_short_one<4, kVec4>(points, stride, g, __ldg(lists + lo.off[3] + (w - pre[3])), sorted_idx, mean);
        else if (w < pre[5]) prep_short_one<6, kVec4>(points, stride, g, __ldg(lists + lo.off[4] + (w - pre[4])), sorted_idx, mean);
        else prep_short_one<8, kVec4>(points, stride, g, __ldg(lists + lo.off[5] + (w - pre[5])), sorted_idx, mean);
      }
    }
  }
  STAGE_END(g_stage_vox, 3);
}

}  // namespace pcp

namespace pcp {
// Everything after the keying kernel (which filled cell[] with the per-cell counts, key[] and within[]): cell scan, counting-sort
// placement, ascending row order inside every pillar (+ the xyz mean when `points` is given).  Shared by pcp_voxelize(),
// pcp_voxelize3d() and pcp_bev_scatter_mean().
int finish_grouping(const WsLayout& L, const WsView& W, int64_t n, int32_t nx, int32_t ny, const float* points, int64_t stride,
                    const pcp_grid& grid, int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                    int32_t* counts_out, cudaStream_t stream) {
  if (L.scan_tiles <= kScanFusedMaxTiles) {
    // one launch: tile records are published and summed inside the kernel (tile_info / hdr were zeroed by the prologue memset)
    PCP_CUDA(launch_pdl(kPdlScan, scan_cells_fused_kernel, dim3((unsigned)L.scan_tiles), dim3(kScanThreads), 0, stream,
                        W.cell, W.cell_rank, L.cells, nx, ny, W.tile_info, W.tile_info + 16 * (L.scan_tiles + 1), W.hdr, W.seg_off,
                        voxel_coords_out, pillar_count_out, W.lists, L.lo, W.long_table, W.big_list, (int32_t)L.scan_tiles));
  } else {
    tile_sums_kernel<<<(unsigned)L.scan_tiles, kScanThreads, 0, stream>>>(W.cell, L.cells, nx, ny, W.tile_info,
                                                                        W.tile_info + 16 * (L.scan_tiles + 1));
    PCP_LAUNCH_CHECK("tile_sums_kernel");
    scan_cells_kernel<<<(unsigned)L.scan_tiles, kScanThreads, 0, stream>>>(
        W.cell, W.cell_rank, L.cells, nx, ny, W.tile_info, W.tile_info + 16 * (L.scan_tiles + 1), W.hdr, W.seg_off, voxel_coords_out, pillar_count_out,
        W.lists, L.lo, W.long_table, W.big_list, L.scan_tiles);
    PCP_LAUNCH_CHECK("scan_cells_kernel");
  }
  {
    const unsigned blocks = (unsigned)((n + 256 * kPtsPerThread - 1) / (256 * kPtsPerThread));
    PCP_CUDA(launch_pdl(kPdlPlace, place_kernel, dim3(blocks ? blocks : 1), dim3(256), 0, stream, W.key, W.within, W.cell, W.cell_rank, n,
                        W.sorted_idx, point_pillar_out, W.hdr, counts_out));
  }
  if (n > 0) {
    const int64_t want = (n + kPrepThreads - 1) / kPrepThreads;
    const int64_t cap = (int64_t)sm_count() * 4;
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    const bool vec4 = (stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
    const float4* no_rec = nullptr;
    if (vec4)
      PCP_CUDA(launch_pdl(kPdlPrep, pillar_prep_kernel<true, false>, dim3(blocks), dim3(kPrepThreads), 0, stream, points, stride, no_rec, W.hdr,
                          W.lists, L.lo, W.long_table, W.big_list, W.sorted_idx, W.mean, W.long_mean, W.long_acc, grid, 15, W.within));
    else
      PCP_CUDA(launch_pdl(kPdlPrep, pillar_prep_kernel<false, false>, dim3(blocks), dim3(kPrepThreads), 0, stream, points, stride, no_rec, W.hdr,
                          W.lists, L.lo, W.long_table, W.big_list, W.sorted_idx, W.mean, W.long_mean, W.long_acc, grid, 15, W.within));
  }
  return 0;
}

// Pillars above 8 rows (phases 1 | 2 | 4 of pillar_prep_kernel: long pillars and the 9 .. 32-row classes) from the placed
// {x, y, z, row} records of the binned path (voxelize_binned.cu, which finishes every shorter pillar inside its per-tile
// kernel): ascending row order, mean, segment entries, max accumulators.
int launch_pillar_prep_rec(const WsLayout& L, const WsView& W, const float* points, int64_t stride, int64_t n,
                           const pcp_grid& grid, cudaStream_t stream) {
  if (n <= 8) return 0;
  const int64_t want = n / (9 * kPrepWarps) + 1;                       // at most n / 9 pillars, one (half) warp each
  const int64_t cap = (int64_t)sm_count() * 4;
  const unsigned blocks = (unsigned)(want < cap ? want : cap);
  const bool vec4 = (stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
  if (vec4)
    pillar_prep_kernel<true, true><<<blocks, kPrepThreads, 0, stream>>>(points, stride, W.rsrec, W.hdr, W.lists, L.lo, W.long_table,
                                                                       W.big_list, W.sorted_idx, W.mean, W.long_mean, W.long_acc, grid, 7, W.within);
  else
    pillar_prep_kernel<false, true><<<blocks, kPrepThreads, 0, stream>>>(points, stride, W.rsrec, W.hdr, W.lists, L.lo, W.long_table,
                                                                        W.big_list, W.sorted_idx, W.mean, W.long_mean, W.long_acc, grid, 7, W.within);
  PCP_LAUNCH_CHECK("pillar_prep_kernel(records)");
  return 0;
}
}  // namespace pcp

using namespace pcp;

STAGE_EXPORT(pcp_debug_stage_voxelize, g_stage_vox)

extern "C" size_t pcp_workspace_bytes(int64_t n_points, int32_t max_frames, int32_t nx, int32_t ny) {
  if (n_points < 0 || max_frames <= 0 || nx <= 0 || ny <= 0) return 0;
  return ws_layout(n_points, max_frames, nx, ny).total;
}

extern "C" int pcp_voxelize_method(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                                   const pcp_grid* grid, void* workspace, size_t workspace_bytes,
                                   int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                                   int64_t pillar_capacity, int32_t* counts_out, int32_t method, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(grid && workspace && voxel_coords_out && counts_out, PCP_E_INVALID, "pcp_voxelize: null argument");
  PCP_REQUIRE(n_points >= 0 && n_points < (1ll << 29), PCP_E_INVALID, "pcp_voxelize: n_points out of range (< 2^29)");
  PCP_REQUIRE(n_points == 0 || points, PCP_E_INVALID, "pcp_voxelize: null points");
  PCP_REQUIRE(row_stride >= 3, PCP_E_INVALID, "pcp_voxelize: row_stride < 3");
  PCP_REQUIRE(max_frames > 0 && grid->nx > 0 && grid->ny > 0, PCP_E_INVALID, "pcp_voxelize: bad grid");
  PCP_REQUIRE(grid->nx <= 65535 && grid->ny <= 65535, PCP_E_UNSUPPORTED, "pcp_voxelize: nx, ny must be <= 65535");
  PCP_REQUIRE((int64_t)max_frames * grid->nx * grid->ny < (1ll << 31), PCP_E_UNSUPPORTED,
              "pcp_voxelize: frames*nx*ny must fit int32 (the reference's merge_coords is int32 too)");
  PCP_REQUIRE(method == PCP_VOXELIZE_AUTO || method == PCP_VOXELIZE_HISTOGRAM || method == PCP_VOXELIZE_RADIX ||
                  method == PCP_VOXELIZE_BINNED, PCP_E_INVALID,
              "pcp_voxelize: unknown method %d", method);
  const WsLayout L = ws_layout(n_points, max_frames, grid->nx, grid->ny);
  PCP_REQUIRE(workspace_bytes >= L.total, PCP_E_WORKSPACE, "pcp_voxelize: workspace %zu < %zu bytes",
              workspace_bytes, L.total);
  PCP_REQUIRE(pillar_capacity >= L.cap, PCP_E_INVALID, "pcp_voxelize: pillar_capacity %lld < min(N, cells) = %lld",
              (long long)pillar_capacity, (long long)L.cap);
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PCP_E_INVALID, "pcp_voxelize: workspace not 256-byte aligned");
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(voxel_coords_out) & 15) == 0, PCP_E_INVALID, "pcp_voxelize: voxel_coords_out not 16-byte aligned");
  const WsView W = ws_view(workspace, L);

  // the radix path covers key spaces of at most 4 M cells and batches of at most 16.6 M rows (RadixPlan)
  // AUTO is the histogram path: measured on the B200 it is the faster of the two at every size (profiles/r02_voxelize_*:
  // 184 vs 392 us on the bench batch - the radix path's one-warp-per-bin stage is bound by its step-to-step latency chain
  // and by the few dense bins at the frame centres).  The radix path stays selectable: it is the one whose means are the
  // sequential row-order sums for pillars of ANY length.
  if (method == PCP_VOXELIZE_RADIX) {
    const RadixPlan rp = radix_plan(n_points, L.cells, sm_count());
    PCP_REQUIRE(rp.ok, PCP_E_UNSUPPORTED,
                "pcp_voxelize: the radix method covers 1 <= n_points <= %lld rows and at most %d cells",
                (long long)kRxMaxChunks * kRxMaxChunkPts, kRxMaxBins << kRxMaxShift);
    if (rp.ok)
      return voxelize_radix(L, W, rp, points, row_stride, n_points, max_frames, *grid, point_pillar_out, voxel_coords_out,
                            pillar_count_out, counts_out, stream);
  }

  if (method == PCP_VOXELIZE_BINNED) {
    PCP_REQUIRE(binned_applies(n_points, L.cells), PCP_E_UNSUPPORTED,
                "pcp_voxelize: the binned method covers n_points >= 1 and at most %lld cells",
                (long long)kBnMaxBins * kScanTileCells);
    return voxelize_binned(L, W, points, row_stride, n_points, max_frames, *grid, point_pillar_out, voxel_coords_out,
                           pillar_count_out, counts_out, stream);
  }

  PCP_CUDA(cudaMemsetAsync(workspace, 0, L.clear_bytes, stream));
  if (n_points > 0) {
    const bool vec4 = (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
    const unsigned blocks = (unsigned)((n_points + 256 * kPtsPerThread - 1) / (256 * kPtsPerThread));
    if (vec4)
      quantise_count_kernel<true><<<blocks, 256, 0, stream>>>(points, row_stride, n_points, max_frames, *grid,
                                                              W.cell, W.key, W.within, W.hdr);
    else
      quantise_count_kernel<false><<<blocks, 256, 0, stream>>>(points, row_stride, n_points, max_frames, *grid,
                                                               W.cell, W.key, W.within, W.hdr);
    PCP_LAUNCH_CHECK("quantise_count_kernel");
  }
  return finish_grouping(L, W, n_points, grid->nx, grid->ny, points, row_stride, *grid, point_pillar_out, voxel_coords_out,
                         pillar_count_out, counts_out, stream);
}

extern "C" int pcp_voxelize(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                            const pcp_grid* grid, void* workspace, size_t workspace_bytes,
                            int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                            int64_t pillar_capacity, int32_t* counts_out, void* stream) {
  return pcp_voxelize_method(points, row_stride, n_points, max_frames, grid, workspace, workspace_bytes, point_pillar_out,
                             voxel_coords_out, pillar_count_out, pillar_capacity, counts_out, PCP_VOXELIZE_AUTO, stream);
}
