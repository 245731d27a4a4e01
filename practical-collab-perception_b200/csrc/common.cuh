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
// workspace layout.  One contiguous caller-owned block:
//   hdr        int32[64]        counts (PCP_COUNT_*), scan ticket, big-segment list length
//   scan_state uint64[tiles]    decoupled look-back tile descriptors
//   cell       int32[cells]     per-cell point count -> (after the scan) pillar rank or -1;
//                               x-major like the reference's linear key: b*nx*ny + cx*ny + cy
//   key        int32[N]         linear key of every input row, -1 = culled
//   within     int32[N]         arrival slot of the row inside its cell
//   seg_off    int32[cap+1]     first sorted position of each pillar (exclusive scan of counts)
//   sorted_idx int32[N]         input row numbers grouped by pillar, ascending inside a pillar
//   big_list   int32[N/32+1]    pillars with more than kSmallSeg points (sorted by a CTA each)
//   tile_first int32[N/kWin+3]  first pillar whose first sorted point lies in PFN window t (ascending)
//   long_list  int32[N/128+1]   pillars with more than kLongSeg points (streamed by the SIMT PFN kernel)
// [hdr | scan_state | cell] is cleared by one memset at the start of pcp_voxelize().
// ---------------------------------------------------------------------------------------------
constexpr int kHdrInts = 64;
constexpr int kHdrScanTicket = 16;   // dynamic tile id of the scan
constexpr int kHdrBigCount = 17;     // entries in big_list
constexpr int kScanTileCells = 2048; // cells per scan tile (256 threads x 8)
constexpr int kSmallSeg = 32;        // segments up to this size are index-sorted by one warp
constexpr int kBigSegMax = 4096;     // segments up to this size are index-sorted by one CTA in smem
constexpr int kHdrLongCount = 18;    // entries in long_list
constexpr int kWin = 112;            // PFN group window: pillars whose first sorted point lies in a window of kWin positions
constexpr int kLongSeg = 128;        // pillars with more points than one PFN sub-tile (handled by the streaming kernel)

struct WsLayout {
  size_t hdr, scan_state, cell, key, within, seg_off, sorted_idx, big_list, tile_first, long_list, total;
  size_t clear_bytes;  // bytes from hdr that the prologue memset clears
  int64_t cells, cap, scan_tiles;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ inline WsLayout ws_layout(int64_t n, int32_t frames, int32_t nx, int32_t ny) {
  WsLayout L;
  L.cells = (int64_t)frames * nx * ny;
  L.cap = n < L.cells ? n : L.cells;
  L.scan_tiles = (L.cells + kScanTileCells - 1) / kScanTileCells;
  size_t o = 0;
  L.hdr = o;         o = align_up(o + sizeof(int32_t) * kHdrInts, 256);
  L.scan_state = o;  o = align_up(o + sizeof(uint64_t) * (size_t)(L.scan_tiles + 1), 256);
  L.cell = o;        o = align_up(o + sizeof(int32_t) * (size_t)L.cells, 256);
  L.clear_bytes = o;
  L.key = o;         o = align_up(o + sizeof(int32_t) * (size_t)(n + 1), 256);
  L.within = o;      o = align_up(o + sizeof(int32_t) * (size_t)(n + 1), 256);
  L.seg_off = o;     o = align_up(o + sizeof(int32_t) * (size_t)(L.cap + 2), 256);
  L.sorted_idx = o;  o = align_up(o + sizeof(int32_t) * (size_t)(n + 1), 256);
  L.big_list = o;    o = align_up(o + sizeof(int32_t) * (size_t)(n / kSmallSeg + 2), 256);
  L.tile_first = o;  o = align_up(o + sizeof(int32_t) * (size_t)(n / kWin + 4), 256);
  L.long_list = o;   o = align_up(o + sizeof(int32_t) * (size_t)(n / kLongSeg + 2), 256);
  L.total = o;
  return L;
}

struct WsView {
  int32_t* hdr;
  unsigned long long* scan_state;
  int32_t* cell;
  int32_t* key;
  int32_t* within;
  int32_t* seg_off;
  int32_t* sorted_idx;
  int32_t* big_list;
  int32_t* tile_first;
  int32_t* long_list;
};

__host__ inline WsView ws_view(void* base, const WsLayout& L) {
  char* p = static_cast<char*>(base);
  WsView v;
  v.hdr = reinterpret_cast<int32_t*>(p + L.hdr);
  v.scan_state = reinterpret_cast<unsigned long long*>(p + L.scan_state);
  v.cell = reinterpret_cast<int32_t*>(p + L.cell);
  v.key = reinterpret_cast<int32_t*>(p + L.key);
  v.within = reinterpret_cast<int32_t*>(p + L.within);
  v.seg_off = reinterpret_cast<int32_t*>(p + L.seg_off);
  v.sorted_idx = reinterpret_cast<int32_t*>(p + L.sorted_idx);
  v.big_list = reinterpret_cast<int32_t*>(p + L.big_list);
  v.tile_first = reinterpret_cast<int32_t*>(p + L.tile_first);
  v.long_list = reinterpret_cast<int32_t*>(p + L.long_list);
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

__device__ __forceinline__ float ld_stream(const float* p) { return __ldg(p); }

}  // namespace pcp
