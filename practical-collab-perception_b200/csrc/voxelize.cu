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
#include <stdlib.h>

#include "common.cuh"

namespace pcp {

// ------------------------------------------------------------------------------------------------
// 1. quantise + count
// ------------------------------------------------------------------------------------------------
template <bool kVec4>
__global__ void __launch_bounds__(256)
quantise_count_kernel(const float* __restrict__ points, int64_t stride, int64_t n, int32_t frames,
                      pcp_grid g, int32_t* __restrict__ cell, int32_t* __restrict__ key,
                      int32_t* __restrict__ within, int32_t* __restrict__ hdr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* row = points + i * stride;
  float bf, x, y;
  if (kVec4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row));
    bf = v.x; x = v.y; y = v.z;
  } else {
    bf = __ldg(row); x = __ldg(row + 1); y = __ldg(row + 2);
  }
  const float qx = quantise(x, g.range_min_x, g.voxel_x);
  const float qy = quantise(y, g.range_min_y, g.voxel_y);
  // reference: (coords >= 0) & (coords < grid) on the int-cast floor (dynamic_pillar_vfe.py:99);
  // qx, qy are integral floats, the float compare is the same predicate and is false for NaN.
  bool keep = (qx >= 0.f) && (qx < (float)g.nx) && (qy >= 0.f) && (qy < (float)g.ny);
  int32_t k = -1, w = 0;
  if (keep) {
    // points[:, 0].int() truncates toward zero (:104)
    if (!(bf > -1.f) || !(bf < (float)frames)) {
      atomicAdd(&hdr[PCP_COUNT_BAD_FRAME], 1);
    } else {
      const int32_t b = (int32_t)bf;
      k = b * (g.nx * g.ny) + (int32_t)qx * g.ny + (int32_t)qy;
      w = atomicAdd(&cell[k], 1);
    }
  }
  key[i] = k;
  within[i] = w;
}

// ------------------------------------------------------------------------------------------------
// 2. single-pass scan over the cells (decoupled look-back)
//    packed 64-bit partial: [63:62] status, [61:32] sum of counts (points), [31:0] sum of flags (pillars)
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kStatusAgg = 1ull << 62;
constexpr unsigned long long kStatusPrefix = 2ull << 62;
constexpr unsigned long long kPayloadMask = (1ull << 62) - 1;
constexpr int kScanThreads = 256;
constexpr int kScanItems = kScanTileCells / kScanThreads;  // 8

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_cells_kernel(int32_t* __restrict__ cell, int64_t cells, int32_t nx, int32_t ny,
                  unsigned long long* __restrict__ state, int32_t* __restrict__ hdr,
                  int32_t* __restrict__ seg_off, int32_t* __restrict__ voxel_coords,
                  int32_t* __restrict__ pillar_count, unsigned long long* __restrict__ lists, const ListOffsets lo,
                  int4* __restrict__ long_table, int32_t* __restrict__ big_list, int64_t scan_tiles) {
  __shared__ int s_tile;
  __shared__ int s_cls[kNumClasses], s_cls_base[kNumClasses];
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_red[2][kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(&hdr[kHdrScanTicket], 1);
  if (tid < kNumClasses) s_cls[tid] = 0;
  __syncthreads();
  const int64_t tile = s_tile;
  const int64_t base = tile * kScanTileCells + (int64_t)tid * kScanItems;

  int32_t c[kScanItems];
  if (base + kScanItems <= cells) {
    const int4 a = *reinterpret_cast<const int4*>(cell + base);
    const int4 b = *reinterpret_cast<const int4*>(cell + base + 4);
    c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) c[j] = (base + j < cells) ? cell[base + j] : 0;
  }
  unsigned long long mine = 0;
  int cmax = 0, last_nonempty = -1;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    mine += ((unsigned long long)(uint32_t)c[j] << 32) | (c[j] > 0 ? 1ull : 0ull);
    cmax = max(cmax, c[j]);
    if (c[j] > 0) last_nonempty = j;
  }
  // block exclusive scan of `mine`
  unsigned long long incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  // per-tile reductions for the header (max points per pillar, last frame that owns a pillar)
  const int64_t nxy = (int64_t)nx * ny;
  int fr = last_nonempty >= 0 ? (int)((base + last_nonempty) / nxy) + 1 : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
    fr = max(fr, __shfl_xor_sync(0xffffffffu, fr, d));
  }
  if (lane == 0) { s_red[0][warp] = cmax; s_red[1][warp] = fr; }
  __syncthreads();
  unsigned long long warp_excl = 0, tile_total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    if (w < warp) warp_excl += s_warp[w];
    tile_total += s_warp[w];
  }
  if (warp == 0) {
    unsigned long long running = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(&state[0], kStatusPrefix | tile_total);
    } else {
      if (lane == 0) atomicExch(&state[tile], kStatusAgg | tile_total);
      int64_t j = tile - 1;
      while (true) {
        const int64_t idx = j - lane;
        unsigned long long v = kStatusPrefix;  // virtual predecessor of tile 0: prefix 0
        if (idx >= 0) {
          v = ld_volatile_u64(&state[idx]);
          while ((v >> 62) == 0) v = ld_volatile_u64(&state[idx]);
        }
        const unsigned pref = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first = pref ? (__ffs(pref) - 1) : 32;
        unsigned long long contrib = (lane <= first) ? (v & kPayloadMask) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
        running += contrib;
        if (pref) break;
        j -= 32;
      }
      if (lane == 0) atomicExch(&state[tile], kStatusPrefix | ((running + tile_total) & kPayloadMask));
    }
    if (lane == 0) {
      s_prefix = running;
      int m = 0, f = 0;
#pragma unroll
      for (int w = 0; w < kScanThreads / 32; ++w) { m = max(m, s_red[0][w]); f = max(f, s_red[1][w]); }
      if (m > 0) {
        atomicMax(&hdr[PCP_COUNT_MAX_PER_PILLAR], m);
        atomicMax(&hdr[PCP_COUNT_FRAMES], f);
      }
      if (tile == scan_tiles - 1) {
        const unsigned long long tot = running + tile_total;
        const int32_t P = (int32_t)(tot & 0xffffffffull);
        const int32_t Nk = (int32_t)((tot >> 32) & 0x3fffffffull);
        hdr[PCP_COUNT_PILLARS] = P;
        hdr[PCP_COUNT_KEPT] = Nk;
        seg_off[P] = Nk;
      }
    }
  }
  __syncthreads();
  unsigned long long excl = s_prefix + warp_excl + (incl - mine);
  unsigned long long my_ent[kScanItems];          // packed list entry
  int my_slot[kScanItems];                        // class << 16 | slot inside this tile's class batch
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    const int64_t idx = base + j;
    my_ent[j] = 0ull; my_slot[j] = -1;
    if (idx < cells) {
      if (c[j] > 0) {
        const int32_t r = (int32_t)(excl & 0xffffffffull);
        const int32_t off = (int32_t)((excl >> 32) & 0x3fffffffull);
        seg_off[r] = off;
        const int32_t b = (int32_t)(idx / nxy);
        const int32_t rem = (int32_t)(idx - (int64_t)b * nxy);
        const int32_t cx = rem / ny, cy = rem - cx * ny;
        // (frame, z = 0, y, x): dynamic_pillar_vfe.py:138-143 after the [0, 3, 2, 1] reorder
        *reinterpret_cast<int4*>(voxel_coords + 4 * (int64_t)r) = make_int4(b, 0, cy, cx);
        if (pillar_count) pillar_count[r] = c[j];
        cell[idx] = r;
        if (c[j] <= kSegRows) {
          const int k = class_of(c[j]);
          my_ent[j] = pack_entry(r, off, c[j]);
          my_slot[j] = (k << 16) | atomicAdd(&s_cls[k], 1);
        } else {
          // long pillar: reserve its segments; sort_long_kernel fills the segment table
          const int nseg = (c[j] + kSegRows - 1) / kSegRows;
          const int li = atomicAdd(&hdr[kHdrLongCount], 1);
          const int sb = atomicAdd(&hdr[kHdrListCount + kSegList], nseg);
          long_table[li] = make_int4(r, off, c[j], sb);
          if (c[j] > kWarpLongMax) big_list[atomicAdd(&hdr[kHdrBigCount], 1)] = li;
        }
        excl += ((unsigned long long)(uint32_t)c[j] << 32) | 1ull;
      } else {
        cell[idx] = -1;
      }
    }
  }
  // work lists: one global atomic per (tile, class)
  __syncthreads();
  if (tid < kNumClasses) s_cls_base[tid] = s_cls[tid] > 0 ? atomicAdd(&hdr[kHdrListCount + tid], s_cls[tid]) : 0;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (my_slot[j] >= 0) {
      const int k = my_slot[j] >> 16;
      lists[lo.off[k] + s_cls_base[k] + (my_slot[j] & 0xffff)] = my_ent[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 3. counting-sort scatter + point -> pillar map
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
place_kernel(const int32_t* __restrict__ key, const int32_t* __restrict__ within,
             const int32_t* __restrict__ cell, const int32_t* __restrict__ seg_off, int64_t n,
             int32_t* __restrict__ sorted_idx, int32_t* __restrict__ point_pillar,
             const int32_t* __restrict__ hdr, int32_t* __restrict__ counts_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && counts_out) {
#pragma unroll
    for (int j = 0; j < PCP_COUNTS_LEN; ++j) counts_out[j] = hdr[j];
  }
  if (i >= n) return;
  const int32_t k = key[i];
  int32_t r = -1;
  if (k >= 0) {
    r = __ldg(cell + k);
    sorted_idx[__ldg(seg_off + r) + within[i]] = (int32_t)i;
  }
  if (point_pillar) point_pillar[i] = r;
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

constexpr int kPrepThreads = 256;
constexpr int kPrepWarps = kPrepThreads / 32;

constexpr int kCountSortMax = 1024;   // CTA path: ranked by counting up to this size, bitonic network above
constexpr int kSumChunk = 256;        // CTA path: rows staged per step of the sequential sum

// Shared memory of a CTA (24 KB, static): either 8 per-warp regions of 3 x 256 words (warp path), or the CTA path's
// row-number array [kBigSegMax] followed by 2048 words of scratch (counting-sort output / double-buffered xyz chunks).
struct PrepSmem {
  union {
    int32_t warp_words[kPrepWarps][3][kWarpLongMax];
    struct { int32_t s[kBigSegMax]; int32_t scratch[2048]; } cta;
  };
};

template <bool kVec4>
__global__ void __launch_bounds__(kPrepThreads, 5)
pillar_prep_kernel(const float* __restrict__ points, int64_t stride, const int32_t* __restrict__ hdr,
                   unsigned long long* __restrict__ lists, const ListOffsets lo, const int4* __restrict__ long_table,
                   const int32_t* __restrict__ big_list, int32_t* __restrict__ sorted_idx, float4* __restrict__ mean,
                   float4* __restrict__ long_mean, unsigned* __restrict__ long_acc, const pcp_grid g, int phase_mask) {
  __shared__ __align__(16) PrepSmem sm;
  __shared__ float s_red[3][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nlong = hdr[kHdrLongCount];

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
    float mx, my, mz;
    if (n <= kBigSegMax) {
      if (n <= kCountSortMax) {
        int32_t* s2 = sm.cta.scratch;
        const int n4 = (n + 3) & ~3;
        for (int i = tid; i < n4; i += kPrepThreads) s[i] = (i < n) ? sorted_idx[off + i] : 0x7fffffff;
        __syncthreads();
        for (int i = tid; i < n; i += kPrepThreads) {
          const int32_t v = s[i];
          int rank = 0;
          for (int j = 0; j < n4; j += 4) {
            const int4 t = *reinterpret_cast<const int4*>(s + j);
            rank += (t.x < v) + (t.y < v) + (t.z < v) + (t.w < v);
          }
          s2[rank] = v;
        }
        __syncthreads();
        for (int i = tid; i < n; i += kPrepThreads) s[i] = s2[i];
        __syncthreads();
      } else {
        int m = 2048;
        while (m < n) m <<= 1;
        for (int i = tid; i < m; i += kPrepThreads) s[i] = (i < n) ? sorted_idx[off + i] : 0x7fffffff;
        __syncthreads();
        for (int k = 2; k <= m; k <<= 1) {
          for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += kPrepThreads) {
              const int p = i ^ j;
              if (p > i) {
                const int32_t a = s[i], b = s[p];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { s[i] = b; s[p] = a; }
              }
            }
            __syncthreads();
          }
        }
      }
      // sequential sums in ascending row order, rows staged kSumChunk at a time (double buffered: one barrier per chunk)
      float* ch = reinterpret_cast<float*>(sm.cta.scratch);      // [2][3][kSumChunk]
      float acc = 0.f;
      float cellw = 0.f;
      for (int c0 = 0; c0 < n; c0 += kSumChunk) {
        float* buf = ch + ((c0 / kSumChunk) & 1) * (3 * kSumChunk);
        const int i = c0 + tid;
        if (i < n) {
          const int32_t idx = s[i];
          sorted_idx[off + i] = idx;
          float x, y, z;
          load_xyz<kVec4>(points, stride, idx, x, y, z);
          if (i == 0) cellw = pack_cell(x, y, g);
          buf[tid] = x; buf[kSumChunk + tid] = y; buf[2 * kSumChunk + tid] = z;
        }
        __syncthreads();
        if (warp < 3 && lane == 0) {
          const float* src = buf + warp * kSumChunk;
          const int m = min(kSumChunk, n - c0);
          for (int q = 0; q < m; ++q) acc = __fadd_rn(acc, src[q]);
        }
      }
      if (warp < 3 && lane == 0) s_red[warp][0] = __fdiv_rn(acc, (float)n);
      if (tid == 0) s_red[0][1] = cellw;
      __syncthreads();
      mx = s_red[0][0]; my = s_red[1][0]; mz = s_red[2][0];
    } else {
      // giant pillar: arrival order, strided partial sums + tree (within tolerance, not order-canonical)
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      for (int i = tid; i < n; i += kPrepThreads) {
        float x, y, z;
        load_xyz<kVec4>(points, stride, sorted_idx[off + i], x, y, z);
        a0 += x; a1 += y; a2 += z;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, d); a1 += __shfl_xor_sync(0xffffffffu, a1, d); a2 += __shfl_xor_sync(0xffffffffu, a2, d);
      }
      if (lane == 0) { s_red[0][warp] = a0; s_red[1][warp] = a1; s_red[2][warp] = a2; }
      __syncthreads();
      if (tid < 3) {
        float acc = 0.f;
        for (int w = 0; w < kPrepWarps; ++w) acc += s_red[tid][w];
        s_red[tid][0] = __fdiv_rn(acc, (float)n);
      }
      __syncthreads();
      mx = s_red[0][0]; my = s_red[1][0]; mz = s_red[2][0];
      if (tid == 0) {
        float x, y, z;
        load_xyz<kVec4>(points, stride, sorted_idx[off], x, y, z);
        s_red[0][1] = pack_cell(x, y, g);
      }
    }
    if (tid == 0) {
      const float cw = s_red[0][1];
      long_mean[li] = make_float4(mx, my, mz, cw);
      mean[r] = make_float4(mx, my, mz, cw);
    }
    __syncthreads();
  }

  // ---------------- long pillars, warp path (33 .. kWarpLongMax rows): one warp per pillar ----------------
  for (int li = blockIdx.x * kPrepWarps + warp; (phase_mask & 2) && li < nlong; li += gridDim.x * kPrepWarps) {
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
#pragma unroll
    for (int k = 0; k < kWarpLongMax / 32; ++k) {
      const int i = k * 32 + lane;
      v[k] = (i < n) ? sorted_idx[off + i] : 0x7fffffff;
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
#pragma unroll
    for (int k = 0; k < kWarpLongMax / 32; ++k)
      if (k * 32 + lane < n) s1[rank[k]] = v[k];
    __syncwarp();
    float* fx = reinterpret_cast<float*>(sm.warp_words[warp][0]);
    float* fy = reinterpret_cast<float*>(sm.warp_words[warp][2]);
    float x[kWarpLongMax / 32], y[kWarpLongMax / 32], z[kWarpLongMax / 32];
#pragma unroll
    for (int k = 0; k < kWarpLongMax / 32; ++k) {
      const int i = k * 32 + lane;
      if (i < n) {
        const int32_t idx = s1[i];
        sorted_idx[off + i] = idx;
        load_xyz<kVec4>(points, stride, idx, x[k], y[k], z[k]);
      }
    }
    __syncwarp();                       // every lane has read its sorted row numbers: s1 can be reused for z
    float* fz = reinterpret_cast<float*>(sm.warp_words[warp][1]);
#pragma unroll
    for (int k = 0; k < kWarpLongMax / 32; ++k) {
      const int i = k * 32 + lane;
      if (i < n) { fx[i] = x[k]; fy[i] = y[k]; fz[i] = z[k]; }
    }
    __syncwarp();
    float acc = 0.f;
    if (lane < 3) {
      const float* src = lane == 0 ? fx : (lane == 1 ? fy : fz);
      for (int i = 0; i < n; ++i) acc = __fadd_rn(acc, src[i]);
      acc = __fdiv_rn(acc, (float)n);
    }
    const float my = __shfl_sync(0xffffffffu, acc, 1), mz = __shfl_sync(0xffffffffu, acc, 2);
    if (lane == 0) {
      const float cw = pack_cell(x[0], y[0], g);
      long_mean[li] = make_float4(acc, my, mz, cw);
      mean[r] = make_float4(acc, my, mz, cw);
    }
    __syncwarp();
  }


  // ---------------- mid pillars: classes 6..9 (9..32 rows), one warp per pillar ----------------
  {
    int pre[5];
    pre[0] = 0;
#pragma unroll
    for (int k = 6; k <= 9; ++k) pre[k - 5] = pre[k - 6] + hdr[kHdrListCount + k];
    constexpr int wpb = kPrepWarps;
    for (int w = blockIdx.x * wpb + warp; (phase_mask & 4) && w < pre[4]; w += gridDim.x * wpb) {
      int q = 0;
#pragma unroll
      for (int t = 1; t < 4; ++t) q += (w >= pre[t]) ? 1 : 0;
      int r, off, n;
      unpack_entry(__ldg(lists + lo.off[6 + q] + (w - pre[q])), r, off, n);
      const int32_t v = (lane < n) ? sorted_idx[off + lane] : 0x7fffffff;
      int rank = 0;
      for (int i = 0; i < n; ++i) rank += (__shfl_sync(0xffffffffu, v, i) < v) ? 1 : 0;
      int32_t* sw = sm.warp_words[warp][0];
      if (lane < n) { sorted_idx[off + rank] = v; sw[rank] = v; }
      __syncwarp();
      float x = 0.f, y = 0.f, z = 0.f;
      if (lane < n) load_xyz<kVec4>(points, stride, sw[lane], x, y, z);
      // lanes 0, 1, 2 run the sequential sums of x, y, z
      float mine = 0.f;
      for (int i = 0; i < n; ++i) {
        const float vx = __shfl_sync(0xffffffffu, x, i), vy = __shfl_sync(0xffffffffu, y, i), vz = __shfl_sync(0xffffffffu, z, i);
        mine = __fadd_rn(mine, lane == 0 ? vx : (lane == 1 ? vy : vz));
      }
      mine = __fdiv_rn(mine, (float)n);
      const float my = __shfl_sync(0xffffffffu, mine, 1), mz = __shfl_sync(0xffffffffu, mine, 2);
      if (lane == 0) mean[r] = make_float4(mine, my, mz, pack_cell(x, y, g));
      __syncwarp();
    }
  }

  // ---------------- short pillars: classes 0..5 (1..8 rows), one thread per pillar ----------------
  {
    int pre[7];
    pre[0] = 0;
#pragma unroll
    for (int k = 0; k <= 5; ++k) pre[k + 1] = pre[k] + hdr[kHdrListCount + k];
    for (int w = blockIdx.x * kPrepThreads + tid; (phase_mask & 8) && w < pre[6]; w += gridDim.x * kPrepThreads) {
      int k = 0;
#pragma unroll
      for (int q = 1; q <= 5; ++q) k += (w >= pre[q]) ? 1 : 0;
      int r, off, n;
      unpack_entry(__ldg(lists + lo.off[k] + (w - pre[k])), r, off, n);
      int32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (j < n) ? sorted_idx[off + j] : 0x7fffffff;
      if (n > 1) {
        // optimal 8-input sorting network (19 compare-exchanges)
        cswap(v[0], v[1]); cswap(v[2], v[3]); cswap(v[4], v[5]); cswap(v[6], v[7]);
        cswap(v[0], v[2]); cswap(v[1], v[3]); cswap(v[4], v[6]); cswap(v[5], v[7]);
        cswap(v[1], v[2]); cswap(v[5], v[6]); cswap(v[0], v[4]); cswap(v[3], v[7]);
        cswap(v[1], v[5]); cswap(v[2], v[6]);
        cswap(v[1], v[4]); cswap(v[3], v[6]);
        cswap(v[2], v[4]); cswap(v[3], v[5]);
        cswap(v[3], v[4]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < n) sorted_idx[off + j] = v[j];
      }
      float x[8], y[8], z[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < n) load_xyz<kVec4>(points, stride, v[j], x[j], y[j], z[j]);
      float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < n) { ax = __fadd_rn(ax, x[j]); ay = __fadd_rn(ay, y[j]); az = __fadd_rn(az, z[j]); }
      const float cnt = (float)n;
      mean[r] = make_float4(__fdiv_rn(ax, cnt), __fdiv_rn(ay, cnt), __fdiv_rn(az, cnt), pack_cell(x[0], y[0], g));
    }
  }
}

}  // namespace pcp

using namespace pcp;

extern "C" size_t pcp_workspace_bytes(int64_t n_points, int32_t max_frames, int32_t nx, int32_t ny) {
  if (n_points < 0 || max_frames <= 0 || nx <= 0 || ny <= 0) return 0;
  return ws_layout(n_points, max_frames, nx, ny).total;
}

extern "C" int pcp_voxelize(const float* points, int64_t row_stride, int64_t n_points, int32_t max_frames,
                            const pcp_grid* grid, void* workspace, size_t workspace_bytes,
                            int32_t* point_pillar_out, int32_t* voxel_coords_out, int32_t* pillar_count_out,
                            int64_t pillar_capacity, int32_t* counts_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(grid && workspace && voxel_coords_out && counts_out, PCP_E_INVALID, "pcp_voxelize: null argument");
  PCP_REQUIRE(n_points >= 0 && n_points < (1ll << 29), PCP_E_INVALID, "pcp_voxelize: n_points out of range (< 2^29)");
  PCP_REQUIRE(n_points == 0 || points, PCP_E_INVALID, "pcp_voxelize: null points");
  PCP_REQUIRE(row_stride >= 3, PCP_E_INVALID, "pcp_voxelize: row_stride < 3");
  PCP_REQUIRE(max_frames > 0 && grid->nx > 0 && grid->ny > 0, PCP_E_INVALID, "pcp_voxelize: bad grid");
  PCP_REQUIRE(grid->nx <= 65535 && grid->ny <= 65535, PCP_E_UNSUPPORTED, "pcp_voxelize: nx, ny must be <= 65535");
  PCP_REQUIRE((int64_t)max_frames * grid->nx * grid->ny < (1ll << 31), PCP_E_UNSUPPORTED,
              "pcp_voxelize: frames*nx*ny must fit int32 (the reference's merge_coords is int32 too)");
  const WsLayout L = ws_layout(n_points, max_frames, grid->nx, grid->ny);
  PCP_REQUIRE(workspace_bytes >= L.total, PCP_E_WORKSPACE, "pcp_voxelize: workspace %zu < %zu bytes",
              workspace_bytes, L.total);
  PCP_REQUIRE(pillar_capacity >= L.cap, PCP_E_INVALID, "pcp_voxelize: pillar_capacity %lld < min(N, cells) = %lld",
              (long long)pillar_capacity, (long long)L.cap);
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PCP_E_INVALID, "pcp_voxelize: workspace not 256-byte aligned");
  PCP_REQUIRE((reinterpret_cast<uintptr_t>(voxel_coords_out) & 15) == 0, PCP_E_INVALID, "pcp_voxelize: voxel_coords_out not 16-byte aligned");
  const WsView W = ws_view(workspace, L);

  PCP_CUDA(cudaMemsetAsync(workspace, 0, L.clear_bytes, stream));
  if (n_points > 0) {
    const bool vec4 = (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
    const unsigned blocks = (unsigned)((n_points + 255) / 256);
    if (vec4)
      quantise_count_kernel<true><<<blocks, 256, 0, stream>>>(points, row_stride, n_points, max_frames, *grid,
                                                              W.cell, W.key, W.within, W.hdr);
    else
      quantise_count_kernel<false><<<blocks, 256, 0, stream>>>(points, row_stride, n_points, max_frames, *grid,
                                                               W.cell, W.key, W.within, W.hdr);
    PCP_LAUNCH_CHECK("quantise_count_kernel");
  }
  scan_cells_kernel<<<(unsigned)L.scan_tiles, kScanThreads, 0, stream>>>(
      W.cell, L.cells, grid->nx, grid->ny, W.scan_state, W.hdr, W.seg_off, voxel_coords_out, pillar_count_out,
      W.lists, L.lo, W.long_table, W.big_list, L.scan_tiles);
  PCP_LAUNCH_CHECK("scan_cells_kernel");
  {
    const unsigned blocks = (unsigned)((n_points + 255) / 256);
    place_kernel<<<blocks ? blocks : 1, 256, 0, stream>>>(W.key, W.within, W.cell, W.seg_off, n_points, W.sorted_idx,
                                                          point_pillar_out, W.hdr, counts_out);
    PCP_LAUNCH_CHECK("place_kernel");
  }
  if (n_points > 0) {
    const int64_t want = (n_points + kPrepThreads - 1) / kPrepThreads;
    const unsigned blocks = (unsigned)(want < 148 * 5 ? want : 148 * 5);
    const bool vec4 = (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
    // PCP_PREP_SPLIT=1 (diagnostic): one launch per phase so that a launch list shows each phase's time
    static const bool split = getenv("PCP_PREP_SPLIT") != nullptr;
    for (int ph = 0; ph < (split ? 4 : 1); ++ph) {
      const int mask = split ? (1 << ph) : 15;
      if (vec4)
        pillar_prep_kernel<true><<<blocks, kPrepThreads, 0, stream>>>(points, row_stride, W.hdr, W.lists, L.lo, W.long_table, W.big_list,
                                                                     W.sorted_idx, W.mean, W.long_mean, W.long_acc, *grid, mask);
      else
        pillar_prep_kernel<false><<<blocks, kPrepThreads, 0, stream>>>(points, row_stride, W.hdr, W.lists, L.lo, W.long_table, W.big_list,
                                                                      W.sorted_idx, W.mean, W.long_mean, W.long_acc, *grid, mask);
    }
    PCP_LAUNCH_CHECK("pillar_prep_kernel");
  }
  return 0;
}
