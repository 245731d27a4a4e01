// voxelize_radix.cu - pillar compaction as a STABLE two-digit MSD radix sort on the linear BEV key
// (reference: dynamic_pillar_vfe.py:98-108 torch.unique(return_inverse, return_counts) and :137-143 voxel_coords).
//
// The key b*nx*ny + cx*ny + cy is split into a HIGH digit (the "bin": key >> shift, at most 4096 bins of 512 or 1024
// consecutive cells - for a 512-wide grid a bin is one x column of one frame) and a LOW digit (the cell inside the bin).
//   K1 radix_count    one CTA per chunk of consecutive input rows: quantise + cull (bit-exact fp32, as the histogram
//                     path), key[] and the chunk's bin histogram (shared-memory counters -> table[chunk][bin]).
//   K2 radix_offsets  per bin, exclusive prefix of the chunk counts (table becomes "rows of this bin before chunk c") and
//                     the bin totals.
//   K3 radix_scatter  stable partition by the high digit.  A warp owns a contiguous sub-range of its chunk: it counts its
//                     rows per bin (16-bit counters, one row of counters per warp), the CTA turns the 16 rows into exclusive
//                     prefixes, and every warp walks its sub-range again IN ORDER, 32 rows per step: __match_any_sync groups
//                     the lanes of a bin, the group's first lane bumps the warp's counter, and each row lands at
//                     bin_start + rows of earlier chunks + rows of earlier warps + rows of earlier steps + rank in the group.
//                     Written per row: {x, y, z, row number} and the key - 20 bytes, bin-contiguous.
//   K4 bin_info / bin_finish  the low digit, one WARP per bin, everything in shared memory: per-cell counts, the scan that
//                     yields pillar ranks (ascending key = torch.unique order) and first sorted positions, a second in-order
//                     walk that places the row numbers (stable: rows ascend inside every pillar with no sorting) and sums
//                     x, y, z per cell in exactly that order - the CPU scatter_mean's sequential fp32 sum, for pillars of ANY
//                     length.  Emits everything the PFN and the canvas writer consume: cell_rank, seg_off, sorted_idx,
//                     voxel_coords, counts, per-length-class work lists, long-pillar tables, the per-pillar mean.
// No global atomics on the data path, no random gathers: after K3 every access is bin-local.  Compared with the dense
// histogram path (voxelize.cu: one returning L2 atomic per point, a random placement pass, three dependent gathers per
// pillar to restore row order) the point rows are read twice, sequentially, and never gathered.
#include "internal.cuh"

namespace pcp {

constexpr int kRxThreads = 512;
constexpr int kRxWarps = kRxThreads / 32;

// Lanes of the warp whose `val` equals this lane's (the low `nbits` bits decide; lanes with !valid match nobody and get 0).
// One ballot per bit: ~3 instructions per bit whatever the values.  The hardware MATCH.ANY walks the distinct values one
// after the other - several hundred cycles when most of the 32 values differ, which is the common case here.
__device__ __forceinline__ unsigned warp_match(unsigned val, int nbits, bool valid) {
  unsigned m = __ballot_sync(0xffffffffu, valid);
  for (int bit = 0; bit < nbits; ++bit) {
    const bool one = (val >> bit) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, one);
    m &= one ? bal : ~bal;
  }
  return valid ? m : 0u;
}

// ------------------------------------------------------------------------------------------------
// K1: keys + per-chunk bin histogram
// ------------------------------------------------------------------------------------------------
template <bool kVec4>
__global__ void __launch_bounds__(kRxThreads)
radix_count_kernel(const float* __restrict__ points, int64_t stride, int64_t n, int32_t frames, pcp_grid g,
                   const RadixPlan rp, int32_t* __restrict__ key, int32_t* __restrict__ point_pillar,
                   int32_t* __restrict__ table, int32_t* __restrict__ hdr) {
  extern __shared__ int32_t s_hist[];                      // [nbins]
  const int tid = threadIdx.x;
  for (int b = tid; b < rp.nbins; b += kRxThreads) s_hist[b] = 0;
  __syncthreads();
  const int64_t beg = (int64_t)blockIdx.x * rp.chunk_pts;
  const int64_t end = min(n, beg + (int64_t)rp.chunk_pts);
  constexpr int U = 8;
  int bad = 0;
  for (int64_t i0 = beg + tid; i0 < end; i0 += kRxThreads * U) {
    float bf[U], x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + (int64_t)u * kRxThreads;
      bf[u] = 0.f; x[u] = 0.f; y[u] = 0.f;
      if (i < end) {
        const float* row = points + i * stride;
        if (kVec4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row));
          bf[u] = v.x; x[u] = v.y; y[u] = v.z;
        } else {
          bf[u] = __ldg(row); x[u] = __ldg(row + 1); y[u] = __ldg(row + 2);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + (int64_t)u * kRxThreads;
      if (i < end) {
        bool bad_frame;
        const int32_t k = point_key(bf[u], x[u], y[u], frames, g, bad_frame);
        bad += bad_frame ? 1 : 0;
        key[i] = k;
        if (k >= 0) atomicAdd(&s_hist[k >> rp.shift], 1);
        else if (point_pillar) point_pillar[i] = -1;
      }
    }
  }
  if (bad) atomicAdd(&hdr[PCP_COUNT_BAD_FRAME], bad);
  __syncthreads();
  int32_t* dst = table + (int64_t)blockIdx.x * rp.nbins;
  for (int b = tid; b < rp.nbins; b += kRxThreads) dst[b] = s_hist[b];
}

// ------------------------------------------------------------------------------------------------
// K2: per bin, exclusive prefix over the chunks; bin totals
//     CTA = 64 bins x 8 chunk ranges: every thread loads its range of chunk counts at once (one latency round)
// ------------------------------------------------------------------------------------------------
constexpr int kOffBins = 64, kOffParts = 8, kOffMaxPer = (kRxMaxChunks + kOffParts - 1) / kOffParts;   // 32

__global__ void __launch_bounds__(kOffBins * kOffParts)
radix_offsets_kernel(int32_t* __restrict__ table, const RadixPlan rp, int32_t* __restrict__ bin_total) {
  __shared__ int32_t s_part[kOffParts][kOffBins];
  const int bl = threadIdx.x % kOffBins, part = threadIdx.x / kOffBins;
  const int b = blockIdx.x * kOffBins + bl;
  const int per = (rp.chunks + kOffParts - 1) / kOffParts;
  const int c0 = part * per, c1 = min(rp.chunks, c0 + per);
  int32_t v[kOffMaxPer];
  int32_t sum = 0;
#pragma unroll
  for (int j = 0; j < kOffMaxPer; ++j) {
    const int c = c0 + j;
    v[j] = (b < rp.nbins && c < c1) ? table[(int64_t)c * rp.nbins + b] : 0;
  }
#pragma unroll
  for (int j = 0; j < kOffMaxPer; ++j) sum += v[j];
  s_part[part][bl] = sum;
  __syncthreads();
  int32_t run = 0, total = 0;
#pragma unroll
  for (int p = 0; p < kOffParts; ++p) {
    const int32_t t = s_part[p][bl];
    if (p < part) run += t;
    total += t;
  }
  if (b < rp.nbins) {
#pragma unroll
    for (int j = 0; j < kOffMaxPer; ++j) {
      const int c = c0 + j;
      if (c < c1) { table[(int64_t)c * rp.nbins + b] = run; run += v[j]; }
    }
    if (part == 0) bin_total[b] = total;
  }
}

// ------------------------------------------------------------------------------------------------
// K3: stable partition by bin
// ------------------------------------------------------------------------------------------------
template <bool kVec4>
__global__ void __launch_bounds__(kRxThreads, 1)
radix_scatter_kernel(const float* __restrict__ points, int64_t stride, int64_t n, const RadixPlan rp,
                     const int32_t* __restrict__ key, const int32_t* __restrict__ table,
                     const int32_t* __restrict__ bin_total, int32_t* __restrict__ bin_start_out,
                     int32_t* __restrict__ dest, float4* __restrict__ rec, int32_t* __restrict__ rkey) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  int32_t* s_base = reinterpret_cast<int32_t*>(s_raw);                                       // [nbins_pad]
  uint16_t* s_cnt = reinterpret_cast<uint16_t*>(s_raw + sizeof(int32_t) * rp.nbins_pad);     // [16][nbins_pad]
  __shared__ int32_t s_wsum[kRxWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = rp.nbins, row = rp.nbins_pad;
  {
    uint4* z = reinterpret_cast<uint4*>(s_cnt);
    const int nz = kRxWarps * row * 2 / 16;
    for (int i = tid; i < nz; i += kRxThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  // ---- first row of every bin (exclusive scan of the bin totals), plus this chunk's rows-before inside the bin ----
  {
    const int per = (nb + kRxThreads - 1) / kRxThreads;           // <= 8 consecutive bins per thread
    const int b0 = tid * per;
    int32_t t[8];
    int32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      t[j] = (j < per && b0 + j < nb) ? __ldg(bin_total + b0 + j) : 0;
      mine += t[j];
    }
    int32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kRxWarps; ++w) {
      const int32_t s = s_wsum[w];
      if (w < warp) before += s;
      total += s;
    }
    int32_t run = before + incl - mine;
    const int32_t* trow = table + (int64_t)blockIdx.x * nb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < per && b0 + j < nb) {
        s_base[b0 + j] = run + __ldg(trow + b0 + j);
        if (blockIdx.x == 0) bin_start_out[b0 + j] = run;
        run += t[j];
      }
    }
    if (blockIdx.x == 0 && tid == 0) bin_start_out[nb] = total;
  }
  __syncthreads();
  // ---- phase 1: rows per bin of every warp's sub-range (two 16-bit counters per word) ----
  const int64_t cbeg = (int64_t)blockIdx.x * rp.chunk_pts;
  const int64_t cend = min(n, cbeg + (int64_t)rp.chunk_pts);
  const int64_t wbeg = cbeg + (int64_t)warp * rp.wpts;
  const int64_t wend = min(cend, wbeg + (int64_t)rp.wpts);
  uint16_t* my_cnt = s_cnt + warp * row;
  {
    unsigned* c32 = reinterpret_cast<unsigned*>(my_cnt);
    constexpr int U = 8;
    for (int64_t i0 = wbeg + lane; i0 < wend; i0 += 32 * U) {
      int32_t k[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + 32 * u;
        k[u] = (i < wend) ? __ldg(key + i) : -1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (k[u] >= 0) {
          const int b = k[u] >> rp.shift;
          atomicAdd(&c32[b >> 1], 1u << (16 * (b & 1)));
        }
    }
  }
  __syncthreads();
  // ---- counts -> exclusive prefixes over the warps (a chunk holds < 65536 rows: they fit 16 bits) ----
  for (int b = tid; b < nb; b += kRxThreads) {
    unsigned run = 0;
#pragma unroll
    for (int w = 0; w < kRxWarps; ++w) {
      const unsigned c = s_cnt[w * row + b];
      s_cnt[w * row + b] = (uint16_t)run;
      run += c;
    }
  }
  __syncthreads();
  // ---- phase 3: the sub-range again, in order; one step = 32 consecutive rows.  Only keys are touched (eight steps of them
  //      in flight) and only the destination of every row is produced: the chain from one step to the next runs through
  //      shared memory alone ----
  int bin_bits = 1;
  while ((1 << bin_bits) < nb) ++bin_bits;
  {
    constexpr int kAhead = 8;
    int32_t kq[kAhead];
#pragma unroll
    for (int d = 0; d < kAhead; ++d) {
      const int64_t i = wbeg + 32 * d + lane;
      kq[d] = (i < wend) ? __ldg(key + i) : -1;
    }
    for (int64_t i0 = wbeg; i0 < wend; i0 += 32 * kAhead) {
#pragma unroll
      for (int d = 0; d < kAhead; ++d) {
        const int64_t j0 = i0 + 32 * d;
        if (j0 >= wend) break;
        const int32_t k = kq[d];
        const int64_t i = j0 + lane;
        {
          const int64_t in = i + 32 * kAhead;
          kq[d] = (in < wend) ? __ldg(key + in) : -1;
        }
        const unsigned b = (k >= 0) ? (unsigned)(k >> rp.shift) : 0u;
        const unsigned m = warp_match(b, bin_bits, k >= 0);
        const int leader = __ffs(m) - 1;
        const int rank = __popc(m & ((1u << lane) - 1u));
        unsigned off = 0;
        if (k >= 0 && lane == leader) {
          off = my_cnt[b];
          my_cnt[b] = (uint16_t)(off + __popc(m));
        }
        off = __shfl_sync(0xffffffffu, off, leader);
        if (i < wend) dest[i] = (k >= 0) ? s_base[b] + (int32_t)off + rank : -1;
        __syncwarp();
      }
    }
  }
  __syncthreads();
  // ---- phase 4: the scatter itself, every row independent of every other (as many loads in flight as registers allow) ----
  auto load_row = [&](int64_t i, float& x, float& y, float& z) {
    const float* r = points + i * stride;
    if (kVec4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(r));
      x = v.y; y = v.z; z = v.w;
    } else {
      x = __ldg(r + 1); y = __ldg(r + 2); z = __ldg(r + 3);
    }
  };
  {
    constexpr int U = 4;
    for (int64_t i0 = cbeg + tid; i0 < cend; i0 += kRxThreads * U) {
      int32_t pos[U], kk[U];
      float x[U], y[U], z[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * kRxThreads;
        pos[u] = -1; kk[u] = -1; x[u] = 0.f; y[u] = 0.f; z[u] = 0.f;
        if (i < cend) {
          pos[u] = __ldcg(dest + i);                       // written by another warp of this CTA a moment ago
          kk[u] = __ldg(key + i);
          load_row(i, x[u], y[u], z[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * kRxThreads;
        if (pos[u] >= 0) {
          rec[pos[u]] = make_float4(x[u], y[u], z[u], __int_as_float((int32_t)i));
          rkey[pos[u]] = kk[u];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4: the low digit, one warp per bin.
// Per-warp shared memory (cells = rp.bin_cells):
//   a_off  int32 [cells]      point count of the cell, then its running placement offset inside the bin
//   p_off  int32 [cells + 32] first offset of the bin's q-th pillar (p_off[npil] = rows of the bin)
//   a_rank uint16[cells]      rank of the cell among the bin's pillars
//   p_cell uint16[cells]      cell of the bin's q-th pillar
//   s_xyz  float4[32]         staging of 32 records of a long pillar (x | y | z columns) for its in-order sums
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t bin_warp_smem(int cells, bool full) {
  return full ? (size_t)cells * (4 + 4 + 2 + 2) + 128 + 512 : (size_t)cells * 4;
}

// record of a bin / a group of bins (16 ints): sums, except the last two (maxima)
constexpr int kBiPoints = 0, kBiPillars = 1, kBiClass = 2, kBiLong = 12, kBiSegs = 13, kBiFrames = 14, kBiMax = 15, kBiInts = 16;

struct BinStats { int pts, pil, nlong, nseg, frames, cmax, cls_lane; };   // cls_lane: pillars of class `lane` (lanes 0 .. 9)

// counts the rows of the bin per cell into a_off (zeroed here), then walks the cells in key order: statistics for the
// record and - kFill - the scan results.  Cell c0 + lane belongs to `lane` in the step of c0 (conflict-free).
template <bool kFill>
__device__ __forceinline__ BinStats bin_scan(const int32_t* __restrict__ rkey, int32_t beg, int32_t end, int32_t base_key,
                                             int cells, int32_t nxy, int32_t* a_off, int32_t* p_off, uint16_t* a_rank,
                                             uint16_t* p_cell, int lane) {
  for (int c = lane; c < cells; c += 32) a_off[c] = 0;
  __syncwarp();
  // rows per cell (eight key loads in flight per lane; a dense cell serialises its atomics in the shared-memory unit)
  for (int32_t p0 = beg; p0 < end; p0 += 32 * 8) {
    int32_t c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int32_t p = p0 + 32 * u + lane;
      c[u] = (p < end) ? __ldg(rkey + p) - base_key : -1;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (c[u] >= 0) atomicAdd(&a_off[c[u]], 1);
  }
  __syncwarp();
  int pts = 0, cmax = 0, nlong = 0, nseg = 0;
  unsigned long long cls = 0;                       // 10 x 6-bit counters: a lane sees at most 32 cells
  int run_off = 0, run_rank = 0, last_cell = -1;
  const unsigned lt = (1u << lane) - 1u;
  for (int c0 = 0; c0 < cells; c0 += 32) {
    const int cell = c0 + lane;
    const int c = a_off[cell];
    const bool flag = c > 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const int row_total = __shfl_sync(0xffffffffu, incl, 31);
    if (flag) {
      pts += c; cmax = max(cmax, c);
      if (c <= kSegRows) cls += 1ull << (6 * class_of(c));
      else { nlong += 1; nseg += (c + kSegRows - 1) / kSegRows; }
    }
    if (kFill) {
      const int r = run_rank + __popc(bal & lt);
      a_off[cell] = run_off + incl - c;
      a_rank[cell] = flag ? (uint16_t)r : (uint16_t)0xffffu;
      if (flag) { p_cell[r] = (uint16_t)cell; p_off[r] = run_off + incl - c; }
    }
    if (bal) last_cell = c0 + 31 - __clz(bal);
    run_off += row_total;
    run_rank += __popc(bal);
  }
  if (kFill && lane == 0) p_off[run_rank] = run_off;
  BinStats s;
  s.pts = __reduce_add_sync(0xffffffffu, pts);
  s.pil = run_rank;
  s.cmax = __reduce_max_sync(0xffffffffu, cmax);
  s.nlong = __reduce_add_sync(0xffffffffu, nlong);
  s.nseg = __reduce_add_sync(0xffffffffu, nseg);
  s.cls_lane = 0;
#pragma unroll
  for (int k = 0; k < kNumClasses; ++k) {
    const int v = __reduce_add_sync(0xffffffffu, (int)((cls >> (6 * k)) & 63ull));
    if (lane == k) s.cls_lane = v;
  }
  s.frames = (last_cell >= 0) ? (int)(((int64_t)base_key + last_cell) / nxy) + 1 : 0;
  return s;
}

// adds a bin's record into the CTA accumulator (shared memory, 16 ints; lanes 0..9 carry the class counts)
__device__ __forceinline__ void acc_record(int* s_acc, const BinStats& s, int lane) {
  if (lane < kNumClasses && s.cls_lane) atomicAdd(&s_acc[kBiClass + lane], s.cls_lane);
  if (lane == 0) {
    atomicAdd(&s_acc[kBiPoints], s.pts); atomicAdd(&s_acc[kBiPillars], s.pil);
    if (s.nlong) { atomicAdd(&s_acc[kBiLong], s.nlong); atomicAdd(&s_acc[kBiSegs], s.nseg); }
    atomicMax(&s_acc[kBiFrames], s.frames); atomicMax(&s_acc[kBiMax], s.cmax);
  }
}

__global__ void __launch_bounds__(256)
bin_info_kernel(const int32_t* __restrict__ rkey, const int32_t* __restrict__ bin_start, const RadixPlan rp,
                int32_t nxy, int32_t* __restrict__ group_info) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_acc[kBiInts];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kBiInts) s_acc[tid] = 0;
  __syncthreads();
  const int bin = blockIdx.x * rp.gw + warp;
  if (warp < rp.gw && bin < rp.nbins) {
    int32_t* a_off = reinterpret_cast<int32_t*>(s_raw + (size_t)warp * bin_warp_smem(rp.bin_cells, false));
    const int32_t beg = __ldg(bin_start + bin), end = __ldg(bin_start + bin + 1);
    if (end > beg) {
      const BinStats s = bin_scan<false>(rkey, beg, end, bin << rp.shift, rp.bin_cells, nxy, a_off, nullptr, nullptr, nullptr, lane);
      acc_record(s_acc, s, lane);
    }
  }
  __syncthreads();
  if (tid < kBiInts) group_info[(int64_t)blockIdx.x * kBiInts + tid] = s_acc[tid];
}

__global__ void __launch_bounds__(256)
bin_finish_kernel(const float4* __restrict__ rec, float4* __restrict__ srec, const int32_t* __restrict__ rkey,
                  const int32_t* __restrict__ bin_start,
                  const RadixPlan rp, int64_t total_cells, int32_t nx, int32_t ny, const int32_t* __restrict__ group_info,
                  int32_t* __restrict__ hdr, int32_t* __restrict__ cell_rank, int32_t* __restrict__ seg_off,
                  int32_t* __restrict__ sorted_idx, int32_t* __restrict__ voxel_coords, int32_t* __restrict__ pillar_count,
                  int32_t* __restrict__ point_pillar, unsigned long long* __restrict__ lists, const ListOffsets lo,
                  int4* __restrict__ long_table, float4* __restrict__ mean, float4* __restrict__ long_mean,
                  unsigned* __restrict__ long_acc, int32_t* __restrict__ counts_out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_before[kBiInts];          // records of the groups before this one
  __shared__ int s_acc[kBiInts];             // this group's record
  __shared__ int s_cls[kBiInts];             // running list slots inside the group (classes, long pillars, segments)
  __shared__ int s_wpil[8];                  // pillars per warp (bin) of the group
  __shared__ int s_red[8][kBiInts];
  __shared__ long long s_lo[kNumLists];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int g = blockIdx.x;
  if (tid < kBiInts) { s_acc[tid] = 0; s_cls[tid] = 0; }
  if (tid < kNumLists) s_lo[tid] = lo.off[tid];
  // ---- records of the groups before this one ----
  {
    int acc[kBiInts];
#pragma unroll
    for (int i = 0; i < kBiInts; ++i) acc[i] = 0;
    for (int t = tid; t < g; t += blockDim.x) {
      const int4* r = reinterpret_cast<const int4*>(group_info + (int64_t)t * kBiInts);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 v = __ldg(r + q);
        acc[4 * q + 0] += v.x; acc[4 * q + 1] += v.y;
        if (q < 3) { acc[4 * q + 2] += v.z; acc[4 * q + 3] += v.w; }
        else { acc[kBiFrames] = max(acc[kBiFrames], v.z); acc[kBiMax] = max(acc[kBiMax], v.w); }
      }
    }
#pragma unroll
    for (int i = 0; i < kBiInts; ++i)
      acc[i] = (i >= kBiFrames) ? __reduce_max_sync(0xffffffffu, acc[i]) : __reduce_add_sync(0xffffffffu, acc[i]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kBiInts; ++i) s_red[warp][i] = acc[i];
    }
  }
  __syncthreads();
  if (tid < kBiInts) {
    int v = 0;
    for (int w = 0; w < nwarps; ++w) v = (tid >= kBiFrames) ? max(v, s_red[w][tid]) : v + s_red[w][tid];
    s_before[tid] = v;
  }
  // ---- per warp: count + scan of its bin ----
  const int bin = g * rp.gw + warp;
  const bool have_bin = warp < rp.gw && bin < rp.nbins;
  const int cells = rp.bin_cells;
  unsigned char* wbase = s_raw + (size_t)warp * bin_warp_smem(cells, true);
  float4* s_xyz = reinterpret_cast<float4*>(wbase);
  wbase += 512;
  int32_t* a_off = reinterpret_cast<int32_t*>(wbase);
  int32_t* p_off = a_off + cells;
  uint16_t* a_rank = reinterpret_cast<uint16_t*>(p_off + cells + 32);
  uint16_t* p_cell = a_rank + cells;
  int32_t beg = 0, end = 0;
  const int32_t base_key = have_bin ? (bin << rp.shift) : 0;
  const int32_t nxy = nx * ny;
  BinStats st{0, 0, 0, 0, 0, 0, 0};
  if (have_bin) {
    beg = __ldg(bin_start + bin); end = __ldg(bin_start + bin + 1);
    if (end > beg) {
      st = bin_scan<true>(rkey, beg, end, base_key, cells, nxy, a_off, p_off, a_rank, p_cell, lane);
      acc_record(s_acc, st, lane);
    }
  }
  if (lane == 0 && warp < 8) s_wpil[warp] = st.pil;
  __syncthreads();
  int rank0 = s_before[kBiPillars];          // global rank of this bin's first pillar
  for (int w = 0; w < warp; ++w) rank0 += s_wpil[w];
  // ---- totals: the last group knows every record ----
  if (g == (int)gridDim.x - 1 && tid == 0) {
    const int P = s_before[kBiPillars] + s_acc[kBiPillars];
    const int Nk = s_before[kBiPoints] + s_acc[kBiPoints];
    const int frames = max(s_before[kBiFrames], s_acc[kBiFrames]);
    const int cmax = max(s_before[kBiMax], s_acc[kBiMax]);
    hdr[PCP_COUNT_PILLARS] = P; hdr[PCP_COUNT_KEPT] = Nk; hdr[PCP_COUNT_FRAMES] = frames; hdr[PCP_COUNT_MAX_PER_PILLAR] = cmax;
    seg_off[P] = Nk;
    hdr[kHdrLongCount] = s_before[kBiLong] + s_acc[kBiLong];
    hdr[kHdrListCount + kSegList] = s_before[kBiSegs] + s_acc[kBiSegs];
    hdr[kHdrBigCount] = 0;
    for (int k = 0; k < kNumClasses; ++k) hdr[kHdrListCount + k] = s_before[kBiClass + k] + s_acc[kBiClass + k];
    if (counts_out) {
      counts_out[PCP_COUNT_PILLARS] = P; counts_out[PCP_COUNT_KEPT] = Nk; counts_out[PCP_COUNT_FRAMES] = frames;
      counts_out[PCP_COUNT_BAD_FRAME] = hdr[PCP_COUNT_BAD_FRAME]; counts_out[PCP_COUNT_MAX_PER_PILLAR] = cmax;
      for (int j = 5; j < PCP_COUNTS_LEN; ++j) counts_out[j] = 0;
    }
  }
  if (!have_bin) return;                      // no block-wide barrier below this line
  // ---- cell -> pillar rank map of the bin (every cell written: -1 = empty) ----
  for (int c = lane; c < cells; c += 32) {
    const int64_t cell = (int64_t)base_key + c;
    if (cell < total_cells) {
      int v = -1;
      if (end > beg) { const int q = a_rank[c]; if (q != 0xffff) v = rank0 + q; }
      cell_rank[cell] = v;
    }
  }
  if (end <= beg) return;
  // ---- second walk over the bin's rows, in order: every row goes to its sorted position (stable: rows ascend inside every
  //      pillar), as a row number for the PFN and as an {x, y, z, row} record for the means below ----
  const int cell_bits = 31 - __clz(cells);
  {
    constexpr int kAhead = 4;
    int32_t kq[kAhead];
    float4 rq[kAhead];
#pragma unroll
    for (int d = 0; d < kAhead; ++d) {
      const int32_t p = beg + 32 * d + lane;
      kq[d] = 0; rq[d] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < end) { kq[d] = __ldg(rkey + p); rq[d] = __ldg(rec + p); }
    }
    for (int32_t p0 = beg; p0 < end; p0 += 32 * kAhead) {
#pragma unroll
      for (int d = 0; d < kAhead; ++d) {
        const int32_t q0 = p0 + 32 * d;
        if (q0 >= end) break;
        const bool valid = q0 + lane < end;
        const int32_t k = kq[d];
        const float4 r = rq[d];
        {
          const int32_t pn = q0 + 32 * kAhead + lane;
          if (pn < end) { kq[d] = __ldg(rkey + pn); rq[d] = __ldg(rec + pn); }
        }
        const unsigned c = valid ? (unsigned)(k - base_key) : 0u;
        const unsigned m = warp_match(c, cell_bits, valid);
        const int leader = __ffs(m) - 1;
        const int rank = __popc(m & ((1u << lane) - 1u));
        int off = 0;
        if (valid && lane == leader) { off = a_off[c]; a_off[c] = off + __popc(m); }
        off = __shfl_sync(0xffffffffu, off, leader);
        if (valid) {
          const int32_t row = __float_as_int(r.w);
          const int32_t pos = beg + off + rank;
          sorted_idx[pos] = row;
          srec[pos] = r;
          if (point_pillar) point_pillar[row] = rank0 + a_rank[c];
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
  // ---- per pillar: coordinates, counts, work-list entry, and the mean - the sequential fp32 sum of its rows in row order
  //      (what index_add_ does on the CPU), read back from the sorted records ----
  const int npil = st.pil;
  for (int q0 = 0; q0 < npil; q0 += 32) {
    const int q = q0 + lane;
    bool is_long = false;
    int r = 0, off = 0, cnt = 0, li = 0, sb = 0;
    uint32_t cxy = 0;
    if (q < npil) {
      const int c = p_cell[q];
      off = beg + p_off[q];
      cnt = p_off[q + 1] - p_off[q];
      r = rank0 + q;
      const uint32_t idx = (uint32_t)base_key + (uint32_t)c;
      const uint32_t b = idx / (uint32_t)nxy, rem = idx - b * (uint32_t)nxy;
      const uint32_t cx = rem / (uint32_t)ny, cy = rem - cx * (uint32_t)ny;
      cxy = (cx & 0xffffu) | (cy << 16);
      seg_off[r] = off;
      // (frame, z = 0, y, x): dynamic_pillar_vfe.py:138-143 after the [0, 3, 2, 1] reorder
      if (voxel_coords) *reinterpret_cast<int4*>(voxel_coords + 4 * (int64_t)r) = make_int4((int)b, 0, (int)cy, (int)cx);
      if (pillar_count) pillar_count[r] = cnt;
      if (cnt <= kSegRows) {
        const int k = class_of(cnt);
        const int slot = s_before[kBiClass + k] + atomicAdd(&s_cls[kBiClass + k], 1);
        lists[s_lo[k] + slot] = pack_entry(r, off, cnt);
        // short pillar: this lane sums its rows (up to 32), four records in flight
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (int j = 0; j < cnt; j += 4) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j + u < cnt) v[u] = __ldcg(srec + off + j + u);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j + u < cnt) { ax = __fadd_rn(ax, v[u].x); ay = __fadd_rn(ay, v[u].y); az = __fadd_rn(az, v[u].z); }
        }
        const float fc = (float)cnt;
        mean[r] = make_float4(__fdiv_rn(ax, fc), __fdiv_rn(ay, fc), __fdiv_rn(az, fc), __uint_as_float(cxy));
      } else {
        is_long = true;
        const int nseg = (cnt + kSegRows - 1) / kSegRows;
        li = s_before[kBiLong] + atomicAdd(&s_cls[kBiLong], 1);
        sb = s_before[kBiSegs] + atomicAdd(&s_cls[kBiSegs], nseg);
        long_table[li] = make_int4(r, off, cnt, sb);
      }
    }
    // long pillars: the whole warp writes the segment entries and the max accumulators, and sums the rows 32 at a time:
    // the records go to shared memory, lanes 0 / 1 / 2 add x / y / z in order
    unsigned lm = __ballot_sync(0xffffffffu, is_long);
    while (lm) {
      const int src = __ffs(lm) - 1;
      lm &= lm - 1;
      const int off_s = __shfl_sync(0xffffffffu, off, src), cnt_s = __shfl_sync(0xffffffffu, cnt, src);
      const int li_s = __shfl_sync(0xffffffffu, li, src), sb_s = __shfl_sync(0xffffffffu, sb, src);
      const int r_s = __shfl_sync(0xffffffffu, r, src);
      const uint32_t cxy_s = __shfl_sync(0xffffffffu, cxy, src);
      const int nseg = (cnt_s + kSegRows - 1) / kSegRows;
      for (int i = lane; i < nseg; i += 32)
        lists[s_lo[kSegList] + sb_s + i] = pack_entry(li_s, off_s + i * kSegRows, min(kSegRows, cnt_s - i * kSegRows));
      for (int i = lane; i < 96; i += 32) long_acc[(int64_t)li_s * 96 + i] = kAccInit;
      float acc = 0.f;
      float* sx = reinterpret_cast<float*>(s_xyz);                 // [3][32]
      float4 nxt = (lane < cnt_s) ? __ldcg(srec + off_s + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j0 = 0; j0 < cnt_s; j0 += 32) {
        const float4 v = nxt;
        if (j0 + 32 + lane < cnt_s) nxt = __ldcg(srec + off_s + j0 + 32 + lane);
        sx[lane] = v.x; sx[32 + lane] = v.y; sx[64 + lane] = v.z;
        __syncwarp();
        if (lane < 3) {
          const float* col = sx + 32 * lane;
          const int n = min(32, cnt_s - j0);
          for (int t = 0; t < n; ++t) acc = __fadd_rn(acc, col[t]);
        }
        __syncwarp();
      }
      acc = __fdiv_rn(acc, (float)cnt_s);
      const float my = __shfl_sync(0xffffffffu, acc, 1), mz = __shfl_sync(0xffffffffu, acc, 2);
      if (lane == 0) {
        const float4 mv = make_float4(acc, my, mz, __uint_as_float(cxy_s));
        mean[r_s] = mv;
        long_mean[li_s] = mv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) PCP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

int voxelize_radix(const WsLayout& L, const WsView& W, const RadixPlan& rp, const float* points, int64_t stride, int64_t n,
                   int32_t frames, const pcp_grid& grid, int32_t* point_pillar_out, int32_t* voxel_coords_out,
                   int32_t* pillar_count_out, int32_t* counts_out, cudaStream_t stream) {
  const bool vec4 = (stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
  PCP_CUDA(cudaMemsetAsync(W.hdr, 0, sizeof(int32_t) * kHdrInts, stream));
  {
    const size_t smem = sizeof(int32_t) * (size_t)rp.nbins;
    if (vec4)
      radix_count_kernel<true><<<rp.chunks, kRxThreads, smem, stream>>>(points, stride, n, frames, grid, rp, W.key,
                                                                         point_pillar_out, W.rtable, W.hdr);
    else
      radix_count_kernel<false><<<rp.chunks, kRxThreads, smem, stream>>>(points, stride, n, frames, grid, rp, W.key,
                                                                          point_pillar_out, W.rtable, W.hdr);
    PCP_LAUNCH_CHECK("radix_count_kernel");
  }
  radix_offsets_kernel<<<(rp.nbins + kOffBins - 1) / kOffBins, kOffBins * kOffParts, 0, stream>>>(W.rtable, rp, W.rbin_total);
  PCP_LAUNCH_CHECK("radix_offsets_kernel");
  {
    const size_t smem = sizeof(int32_t) * (size_t)rp.nbins_pad + sizeof(uint16_t) * (size_t)kRxWarps * rp.nbins_pad;
    if (vec4) {
      if (int rc = set_smem(radix_scatter_kernel<true>, smem)) return rc;
      radix_scatter_kernel<true><<<rp.chunks, kRxThreads, smem, stream>>>(points, stride, n, rp, W.key, W.rtable, W.rbin_total,
                                                                           W.rbin_start, W.within, W.rrec, W.rkey);
    } else {
      if (int rc = set_smem(radix_scatter_kernel<false>, smem)) return rc;
      radix_scatter_kernel<false><<<rp.chunks, kRxThreads, smem, stream>>>(points, stride, n, rp, W.key, W.rtable, W.rbin_total,
                                                                            W.rbin_start, W.within, W.rrec, W.rkey);
    }
    PCP_LAUNCH_CHECK("radix_scatter_kernel");
  }
  {
    const size_t smem = (size_t)rp.gw * bin_warp_smem(rp.bin_cells, false);
    bin_info_kernel<<<rp.ngroups, rp.gw * 32, smem, stream>>>(W.rkey, W.rbin_start, rp, grid.nx * grid.ny, W.rgroup_info);
    PCP_LAUNCH_CHECK("bin_info_kernel");
  }
  {
    const size_t smem = (size_t)rp.gw * bin_warp_smem(rp.bin_cells, true);
    if (int rc = set_smem(bin_finish_kernel, smem)) return rc;
    bin_finish_kernel<<<rp.ngroups, rp.gw * 32, smem, stream>>>(
        W.rrec, W.rsrec, W.rkey, W.rbin_start, rp, L.cells, grid.nx, grid.ny, W.rgroup_info, W.hdr, W.cell_rank, W.seg_off, W.sorted_idx,
        voxel_coords_out, pillar_count_out, point_pillar_out, W.lists, L.lo, W.long_table, W.mean, W.long_mean, W.long_acc,
        counts_out);
    PCP_LAUNCH_CHECK("bin_finish_kernel");
  }
  return 0;
}

}  // namespace pcp
