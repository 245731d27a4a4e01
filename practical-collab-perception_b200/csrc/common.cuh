// common.cuh - shared definitions of libpcp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#include "../../include/pcp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpcp_b200 is written for sm_100a (B200) only"
#endif

namespace pcp {

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local string; defined in api.cu)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define PCP_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::pcp::set_error(__VA_ARGS__);            \
      return (code);                            \
    }                                           \
  } while (0)

#define PCP_LAUNCH_CHECK(what)                                  \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return ::pcp::cuda_fail(e__, what); \
  } while (0)

#define PCP_CUDA(call)                                            \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return ::pcp::cuda_fail(e__, #call);  \
  } while (0)

// ---------------------------------------------------------------------------------------------
// pillar length classes.  The PFN kernel puts ONE PILLAR PER TENSOR-MEMORY LANE and walks the pillar's
// points as "slots" (slot j = j-th point of each of the 128 pillars of a group), so the per-pillar max is an
// elementwise max between accumulators inside a thread.  Pillars are therefore binned by length: class k
// holds pillars of kClassMinLen[k] .. kClassSlots[k] points and is processed with kClassSlots[k] slots
// (shorter pillars repeat their last point - a max is idempotent).  Pillars above kSegRows points are cut
// into SEGMENTS of kSegRows rows that are processed like a 32-slot class; their partial maxima meet in a
// per-pillar accumulator (ordered-int atomicMax) that a finishing kernel turns into the output row.
// ---------------------------------------------------------------------------------------------
constexpr int kNumClasses = 10;
constexpr int kSegList = kNumClasses;       // list index of the long-pillar segments
constexpr int kNumLists = kNumClasses + 1;
constexpr int kSegRows = 32;                // rows per segment == largest class
constexpr int kGroup = 128;                 // pillars (TMEM lanes) per group
constexpr int kBigSegMax = 4096;            // long pillars up to this size are index-sorted by one CTA in smem
__host__ __device__ constexpr int class_slots(int k) {
  return k < 4 ? k + 1 : (k == 4 ? 6 : k == 5 ? 8 : k == 6 ? 12 : k == 7 ? 16 : k == 8 ? 24 : 32);
}
__host__ __device__ constexpr int class_min_len(int k) {
  return k < 4 ? k + 1 : (k == 4 ? 5 : k == 5 ? 7 : k == 6 ? 9 : k == 7 ? 13 : k == 8 ? 17 : k == 9 ? 25 : 33);
}
// class of a pillar of n points, 1 <= n <= kSegRows
__host__ __device__ inline int class_of(int n) {
  return n <= 4 ? n - 1 : (n <= 6 ? 4 : n <= 8 ? 5 : n <= 12 ? 6 : n <= 16 ? 7 : n <= 24 ? 8 : 9);
}

// ---------------------------------------------------------------------------------------------
// workspace layout.  One contiguous caller-owned block:
//   hdr        int32[64]        counts (PCP_COUNT_*), scan ticket, list lengths, long-pillar count
//   tile_info  int32[tiles][16] per-tile counters of the cell scan (points, pillars, per-class pillars, long pillars, ...)
//   cell       int32[cells]     per-cell point count -> (after the scan) first sorted position of the cell's points, or -1;
//                               x-major like the reference's linear key: b*nx*ny + cx*ny + cy
//   cell_rank  int32[cells]     pillar rank of the cell or -1 (same order)
//   key        int32[N]         linear key of every input row, -1 = culled
//   within     int32[N]         arrival slot of the row inside its cell
//   seg_off    int32[cap+1]     first sorted position of each pillar (exclusive scan of counts)
//   sorted_idx int32[N]         input row numbers grouped by pillar, ascending inside a pillar
//   lists      uint64[...]      per length class: packed pillar descriptors (list k at list_off[k], capacity
//                               N / min_len + 1): bits [0,29) pillar rank, [29,58) first sorted position, [58,63) rows - 1;
//                               the last list holds the long-pillar SEGMENTS (same packing, long index instead of rank)
//   mean       float4[cap+1]    per-pillar mean xyz (sequential fp32 sum in ascending row order / count), by pillar rank
//   long_table int4[N/33+1]     long pillars {pillar rank, first sorted position, rows, first segment}
//   long_mean  float4[N/33+1]   mean xyz of each long pillar (sequential fp32 sum in ascending row order)
//   long_acc   uint32[96*(N/33+1)] per long pillar: ordered-int running max of layer-0 (32) and layer-1 (64) values
// [hdr | tile_info | cell] is cleared by one memset at the start of pcp_voxelize().
// ---------------------------------------------------------------------------------------------
constexpr int kHdrInts = 64;
constexpr int kHdrPrepTicket = 16;   // [16, 20): work tickets of pillar_prep_kernel (long, mid-16, mid-32, short)
constexpr int kHdrScanTicket = 20;   // tile ticket of the single-launch cell scan
constexpr int kScanFusedMaxTiles = 2048;   // above this the scan runs as two launches (every tile sums ALL records before it)
constexpr int kHdrListCount = 32;    // [32, 32 + kNumLists): entries in each list
constexpr int kHdrLongCount = 48;    // long pillars
constexpr int kHdrBigCount = 49;     // long pillars above kWarpLongMax rows (handled by a whole CTA in pillar_prep_kernel)
constexpr int kWarpLongMax = 128;    // long pillars up to this size are ordered by ONE warp
constexpr int kScanTileCells = 2048; // cells per scan tile (256 threads x 8)
constexpr unsigned kAccInit = 0x007fffffu;   // ordered-int encoding of -inf

struct ListOffsets { int64_t off[kNumLists]; };   // entry offsets of every list inside `lists`

// packed work-list entry: one 8-byte load tells a PFN thread everything about its pillar
__host__ __device__ inline unsigned long long pack_entry(int r, int off, int len) {
  return (unsigned long long)(unsigned)r | ((unsigned long long)(unsigned)off << 29) | ((unsigned long long)(len - 1) << 58);
}
__host__ __device__ inline void unpack_entry(unsigned long long e, int& r, int& off, int& len) {
  r = (int)(e & 0x1fffffffull);
  off = (int)((e >> 29) & 0x1fffffffull);
  len = (int)((e >> 58) & 31ull) + 1;
}

// ---------------------------------------------------------------------------------------------
// radix path of pcp_voxelize (voxelize_radix.cu): key = bin << shift | cell-in-bin
// ---------------------------------------------------------------------------------------------
constexpr int kRxMaxBins = 4096;        // bins (high digit); 16-bit counters per (warp, bin) in shared memory
constexpr int kRxMinShift = 9;          // 512 cells per bin ...
constexpr int kRxMaxShift = 10;         // ... or 1024: at most 4 M cells (16 frames of 512 x 512, 4 of 1024 x 1024)
constexpr int kRxMaxChunks = 256;       // chunks of consecutive input rows (one CTA each; B200: one per SM)
constexpr int kRxChunkWarps = 16;
constexpr int kRxMaxChunkPts = 65024;   // rows per chunk: a multiple of 512 below 65536, so that prefixes inside a chunk fit 16 bits

struct RadixPlan {
  int32_t ok;                           // 0: this problem runs on the histogram path
  int32_t shift, bin_cells, nbins, nbins_pad;
  int32_t chunks, chunk_pts, wpts;      // K1 / K3: chunks of chunk_pts rows, wpts rows per warp
  int32_t gw, ngroups;                  // K4: bins (= warps) per CTA, CTAs
};

// bins of a key space of `cells` cells; false when the radix path does not cover it
__host__ __device__ inline bool radix_bins(int64_t cells, int32_t& shift, int32_t& nbins) {
  shift = kRxMinShift;
  while (((cells + (1ll << shift) - 1) >> shift) > kRxMaxBins && shift <= kRxMaxShift) ++shift;
  nbins = (int32_t)((cells + (1ll << shift) - 1) >> shift);
  return shift <= kRxMaxShift && cells > 0;
}

__host__ inline RadixPlan radix_plan(int64_t n, int64_t cells, int sms) {
  RadixPlan p{};
  if (n <= 0 || !radix_bins(cells, p.shift, p.nbins)) return p;
  p.bin_cells = 1 << p.shift;
  p.nbins_pad = (p.nbins + 7) & ~7;
  int64_t chunks = sms < kRxMaxChunks ? sms : kRxMaxChunks;
  if (chunks < 1) chunks = 1;
  const int64_t unit = 32 * kRxChunkWarps;                      // 512 rows: 32 per warp
  if (chunks > (n + unit - 1) / unit) chunks = (n + unit - 1) / unit;
  int64_t chunk_pts = ((n + chunks - 1) / chunks + unit - 1) / unit * unit;
  if (chunk_pts > kRxMaxChunkPts) chunk_pts = kRxMaxChunkPts;
  chunks = (n + chunk_pts - 1) / chunk_pts;
  if (chunks > kRxMaxChunks) return p;                          // more than 16.6 M rows: histogram path
  p.chunks = (int32_t)chunks;
  p.chunk_pts = (int32_t)chunk_pts;
  p.wpts = (int32_t)(chunk_pts / kRxChunkWarps);
  p.gw = (p.shift == kRxMinShift && p.nbins >= 2048) ? 8 : 4;
  p.ngroups = (p.nbins + p.gw - 1) / p.gw;
  p.ok = 1;
  return p;
}

// ---------------------------------------------------------------------------------------------
// binned path of pcp_voxelize (voxelize_binned.cu): bin = one scan tile of kScanTileCells consecutive cells
// ---------------------------------------------------------------------------------------------
constexpr int kBnMaxBins = 2048;        // every tile of the finish kernel sums the records of all tiles before it: keep that small
constexpr int kBnCtrlInts = 64;         // [1] tile ticket of the finish kernel

// the binned method covers 1 .. 2^29 - 1 rows and key spaces of at most kBnMaxBins * kScanTileCells = 4 M cells
__host__ __device__ inline bool binned_applies(int64_t n, int64_t cells) {
  return n > 0 && cells > 0 && (cells + kScanTileCells - 1) / kScanTileCells <= kBnMaxBins;
}

struct WsLayout {
  size_t bz_ctrl, bz_count, bz_cursor, bz_frames, bz_rec, bz_start, bz_clear_bytes;   // binned path
  size_t hdr, tile_info, cell, cell_rank, key, within, seg_off, sorted_idx, lists, mean, long_table, big_list, long_mean, long_acc, total;
  size_t rtable, rbin_total, rbin_start, rgroup_info, rrec, rsrec, rkey;   // radix path scratch (zero-sized when it does not apply)
  size_t clear_bytes;  // bytes from hdr that the prologue memset clears
  int64_t cells, cap, scan_tiles, seg_cap, long_cap;
  ListOffsets lo;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ inline WsLayout ws_layout(int64_t n, int32_t frames, int32_t nx, int32_t ny) {
  WsLayout L;
  L.cells = (int64_t)frames * nx * ny;
  L.cap = n < L.cells ? n : L.cells;
  L.scan_tiles = (L.cells + kScanTileCells - 1) / kScanTileCells;
  L.seg_cap = n / 16 + 2;
  L.long_cap = n / (kSegRows + 1) + 1;
  size_t o = 0;
  L.hdr = o;         o = align_up(o + sizeof(int32_t) * kHdrInts, 256);
  {
    // binned path: control words, per-bin row counts / fill cursors, per-tile frame count + 1 and per-tile records (16
    // ints, every field stored + 1: zero = not published yet) - cleared together with hdr by one small memset
    const size_t nb = binned_applies(n, L.cells) ? (size_t)L.scan_tiles : 0;
    L.bz_ctrl = o;   o += sizeof(int32_t) * kBnCtrlInts;
    L.bz_count = o;  o += sizeof(int32_t) * nb;
    L.bz_cursor = o; o += sizeof(int32_t) * nb;
    L.bz_frames = o; o += sizeof(int32_t) * nb;
    o = align_up(o, 64);
    L.bz_rec = o;    o = align_up(o + sizeof(int32_t) * 16 * nb, 256);
    L.bz_clear_bytes = o;
    L.bz_start = o;  o = align_up(o + sizeof(int32_t) * (nb + 1), 256);
  }
  L.tile_info = o;   o = align_up(o + sizeof(int32_t) * 17 * (size_t)(L.scan_tiles + 1), 256);   // records + per-tile frame count
  L.cell = o;        o = align_up(o + sizeof(int32_t) * (size_t)L.cells, 256);
  L.clear_bytes = o;
  L.cell_rank = o;   o = align_up(o + sizeof(int32_t) * (size_t)L.cells, 256);
  L.key = o;         o = align_up(o + sizeof(int32_t) * (size_t)(n + 1), 256);
  L.within = o;      o = align_up(o + sizeof(int32_t) * (size_t)(n + 1), 256);
  L.seg_off = o;     o = align_up(o + sizeof(int32_t) * (size_t)(L.cap + 2), 256);
  L.sorted_idx = o;  o = align_up(o + sizeof(int32_t) * (size_t)(n + 1), 256);
  L.lists = o;
  int64_t lo = 0;
  for (int k = 0; k < kNumClasses; ++k) {
    L.lo.off[k] = lo;
    int64_t c = n / class_min_len(k) + 1;
    if (c > L.cap + 1) c = L.cap + 1;
    lo += (c + 63) / 64 * 64;
  }
  L.lo.off[kSegList] = lo;   // long-pillar segments, written by pillar_prep_kernel
  lo += (L.seg_cap + 63) / 64 * 64;
  o = align_up(o + sizeof(unsigned long long) * (size_t)(lo + 64), 256);
  L.mean = o;        o = align_up(o + 16 * (size_t)(L.cap + 1), 256);
  L.long_table = o;  o = align_up(o + 16 * (size_t)L.long_cap, 256);
  L.big_list = o;    o = align_up(o + 4 * (size_t)L.long_cap, 256);
  L.long_mean = o;   o = align_up(o + 16 * (size_t)L.long_cap, 256);
  L.long_acc = o;    o = align_up(o + sizeof(uint32_t) * 96 * (size_t)L.long_cap, 256);
  {
    // radix path: table[chunk][bin] | bin totals | first row of every bin | K4 group records | {x, y, z, row} records in bin
    // order, the same records in pillar order, keys in bin order.  Capacities depend on (n, cells) only, never on the device.
    int32_t shift = 0, nbins = 0;
    const bool rx = radix_bins(L.cells, shift, nbins) && n > 0 && n <= (int64_t)kRxMaxChunks * kRxMaxChunkPts;
    const size_t nb = rx ? (size_t)nbins : 0, np = (rx || binned_applies(n, L.cells)) ? (size_t)n : 0;
    L.rtable = o;      o = align_up(o + sizeof(int32_t) * kRxMaxChunks * nb, 256);
    L.rbin_total = o;  o = align_up(o + sizeof(int32_t) * (nb + 1), 256);
    L.rbin_start = o;  o = align_up(o + sizeof(int32_t) * (nb + 1), 256);
    L.rgroup_info = o; o = align_up(o + sizeof(int32_t) * 16 * (nb / 4 + 1), 256);
    L.rrec = o;        o = align_up(o + 16 * (np + 1), 256);
    L.rsrec = o;       o = align_up(o + 16 * (np + 1), 256);
    L.rkey = o;        o = align_up(o + sizeof(int32_t) * (np + 1), 256);
  }
  L.total = o;
  return L;
}

struct WsView {
  int32_t* hdr;
  int32_t* tile_info;
  int32_t* cell;
  int32_t* cell_rank;
  int32_t* key;
  int32_t* within;
  int32_t* seg_off;
  int32_t* sorted_idx;
  unsigned long long* lists;
  float4* mean;
  int4* long_table;
  int32_t* big_list;
  float4* long_mean;
  unsigned* long_acc;
  int32_t* bz_ctrl;
  int32_t* bz_count;
  int32_t* bz_cursor;
  int32_t* bz_frames;
  int32_t* bz_rec;
  int32_t* bz_start;
  int32_t* rtable;
  int32_t* rbin_total;
  int32_t* rbin_start;
  int32_t* rgroup_info;
  float4* rrec;
  float4* rsrec;
  int32_t* rkey;
};

__host__ inline WsView ws_view(void* base, const WsLayout& L) {
  char* p = static_cast<char*>(base);
  WsView v;
  v.hdr = reinterpret_cast<int32_t*>(p + L.hdr);
  v.tile_info = reinterpret_cast<int32_t*>(p + L.tile_info);
  v.cell = reinterpret_cast<int32_t*>(p + L.cell);
  v.cell_rank = reinterpret_cast<int32_t*>(p + L.cell_rank);
  v.key = reinterpret_cast<int32_t*>(p + L.key);
  v.within = reinterpret_cast<int32_t*>(p + L.within);
  v.seg_off = reinterpret_cast<int32_t*>(p + L.seg_off);
  v.sorted_idx = reinterpret_cast<int32_t*>(p + L.sorted_idx);
  v.lists = reinterpret_cast<unsigned long long*>(p + L.lists);
  v.mean = reinterpret_cast<float4*>(p + L.mean);
  v.long_table = reinterpret_cast<int4*>(p + L.long_table);
  v.big_list = reinterpret_cast<int32_t*>(p + L.big_list);
  v.long_mean = reinterpret_cast<float4*>(p + L.long_mean);
  v.long_acc = reinterpret_cast<unsigned*>(p + L.long_acc);
  v.bz_ctrl = reinterpret_cast<int32_t*>(p + L.bz_ctrl);
  v.bz_count = reinterpret_cast<int32_t*>(p + L.bz_count);
  v.bz_cursor = reinterpret_cast<int32_t*>(p + L.bz_cursor);
  v.bz_frames = reinterpret_cast<int32_t*>(p + L.bz_frames);
  v.bz_rec = reinterpret_cast<int32_t*>(p + L.bz_rec);
  v.bz_start = reinterpret_cast<int32_t*>(p + L.bz_start);
  v.rtable = reinterpret_cast<int32_t*>(p + L.rtable);
  v.rbin_total = reinterpret_cast<int32_t*>(p + L.rbin_total);
  v.rbin_start = reinterpret_cast<int32_t*>(p + L.rbin_start);
  v.rgroup_info = reinterpret_cast<int32_t*>(p + L.rgroup_info);
  v.rrec = reinterpret_cast<float4*>(p + L.rrec);
  v.rsrec = reinterpret_cast<float4*>(p + L.rsrec);
  v.rkey = reinterpret_cast<int32_t*>(p + L.rkey);
  return v;
}


// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
// Pillar cell of a coordinate, bit-exact with the reference's
//   torch.floor((x - range_min) / voxel)            (dynamic_pillar_vfe.py:98)
// IEEE fp32 subtract, IEEE fp32 divide (not a reciprocal multiply), floor.  Returned as float so that
// NaN / inf fail the range test the way the reference's int compare does on CPU.
__device__ __forceinline__ float quantise(float v, float vmin, float vsize) {
  return floorf(__fdiv_rn(__fsub_rn(v, vmin), vsize));
}

// Linear key of a point row, or -1 when it is culled (dynamic_pillar_vfe.py:98-106).  The range test runs on the integral
// floats (the same predicate as the reference's compare on the int cast, and false for NaN); points[:, 0].int() truncates
// toward zero (:104); a frame index outside [0, frames) drops the row and is reported through `bad_frame`.
__device__ __forceinline__ int32_t point_key(float bf, float x, float y, int32_t frames, const pcp_grid& g, bool& bad_frame) {
  bad_frame = false;
  const float qx = quantise(x, g.range_min_x, g.voxel_x);
  const float qy = quantise(y, g.range_min_y, g.voxel_y);
  if (!((qx >= 0.f) && (qx < (float)g.nx) && (qy >= 0.f) && (qy < (float)g.ny))) return -1;
  if (!(bf > -1.f) || !(bf < (float)frames)) { bad_frame = true; return -1; }
  return (int32_t)bf * (g.nx * g.ny) + (int32_t)qx * g.ny + (int32_t)qy;
}

__device__ __forceinline__ float ld_stream(const float* p) { return __ldg(p); }

// order-preserving float <-> uint32 map (atomicMax on the encoded value == max on the float)
__device__ __forceinline__ unsigned ord_enc(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_dec(unsigned u) {
  return (u & 0x80000000u) ? __uint_as_float(u & 0x7fffffffu) : __uint_as_float(~u);
}

}  // namespace pcp
