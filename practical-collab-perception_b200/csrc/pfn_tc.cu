// pfn_tc.cu - the PFN on 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 for fp32-grade accuracy.
//
// Reference: dynamic_pillar_vfe.py:110-129 + PFNLayerV2.forward :35-46 (two layers, NUM_FILTERS [64, 64]).
//
// Work unit: a GROUP = the pillars whose first sorted point lies in a window of kWin sorted positions
// (tile_first[] from the voxelize stage).  A persistent CTA (two per SM) walks its groups; inside a group
// the pillars are packed greedily into SUB-TILES of <= 128 rows (one row = one point), so a pillar is never
// split and everything a pillar needs stays on chip:
//   P1  gather rows, per-pillar mean (sequential fp32 sum in ascending row order, exact), features,
//       TF32 hi/lo split -> A0 panels in shared memory
//   M0  D0[128x32]  = A0 . W0^T                  (tcgen05.mma kind::tf32, 3 MMAs per K step: lo.hi, hi.lo, hi.hi)
//   P2  TMEM -> regs: BN(eval)+ReLU -> x0, split -> A1 panels 0..7
//   P3  per-pillar max of x0 (exact: lexicographic max over the (hi, lo) pairs), broadcast into A1 panels 8..15
//       of every row of the pillar  == torch.cat([x, x_max[unq_inv]])
//   M1  D1[128x64]  = A1 . W1^T  (K = 64)
//   P4  TMEM -> regs: BN(eval)+ReLU -> y staged in shared memory (aliases A1)
//   P5  per-pillar max of y -> pillar_features, 256-byte coalesced rows
// Pillars with more than 128 points are left to the chunk-streaming SIMT kernel (pfn.cu), launched on the
// `long_list` the scan produced.
#include "common.cuh"
#include "umma.cuh"
#include "pfn_tc.cuh"

namespace pcp {

using namespace umma;

constexpr int kRows = 128;        // MMA M = rows per sub-tile
constexpr int kTcThreads = 256;
constexpr int kTmemCols = 128;    // D0: columns [0, 32), D1: columns [32, 96)
constexpr int kMaxK0 = 24;        // layer-0 K (c_in rounded up to 8)

struct TcSmem {
  // ---- operands (16-byte aligned panels, see umma.cuh) ----
  alignas(128) float w0h[kMaxK0 * 32];
  alignas(128) float w0l[kMaxK0 * 32];
  alignas(128) float w1h[64 * 64];
  alignas(128) float w1l[64 * 64];
  alignas(128) float a1h[16 * kRows * 4];   // panels 0..7: x0, 8..15: pillar max; A0 aliases panels 8..13; y aliases all
  alignas(128) float a1l[16 * kRows * 4];
  // ---- per sub-tile metadata ----
  float xyz[3][kRows];
  float mean[3][kRows];
  int off[kRows + 2];       // sorted position of the first row of each pillar of the group (+ end)
  int lp[kRows];            // group-local pillar of each row
  alignas(16) float a0[32], b0[32], a1[64], b1[64];
  alignas(8) uint64_t bar[2];
  uint32_t tmem_base;
  int sub_begin, sub_end;   // current sub-tile's pillar range (group-local)
};


__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
// lexicographic max over (hi, lo) pairs == split of the max of the exact values (both maps are monotone)
__device__ __forceinline__ void lexmax(float& h, float& l, float h2, float l2) {
  const bool take = (h2 > h) || (h2 == h && l2 > l);
  h = take ? h2 : h;
  l = take ? l2 : l;
}

__global__ void __launch_bounds__(kTcThreads, 2)
pfn_tc_kernel(const TcArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = A.hdr[PCP_COUNT_PILLARS];
  const int Nk = A.hdr[PCP_COUNT_KEPT];
  const int n_groups = (Nk + kWin - 1) / kWin;
  const int k0 = A.k0;

  // ---- one-time setup: weights -> smem, barriers, TMEM ----
  {
    const float4* src = reinterpret_cast<const float4*>(A.params_tc);
    const int n0 = k0 * 32 / 4, n1 = 64 * 64 / 4;
    for (int i = tid; i < n0; i += kTcThreads) {
      reinterpret_cast<float4*>(S.w0h)[i] = __ldg(src + i);
      reinterpret_cast<float4*>(S.w0l)[i] = __ldg(src + n0 + i);
    }
    for (int i = tid; i < n1; i += kTcThreads) {
      reinterpret_cast<float4*>(S.w1h)[i] = __ldg(src + 2 * n0 + i);
      reinterpret_cast<float4*>(S.w1l)[i] = __ldg(src + 2 * n0 + n1 + i);
    }
    if (tid < 32) { S.a0[tid] = A.params_simt[A.a0_off + tid]; S.b0[tid] = A.params_simt[A.a0_off + 32 + tid]; }
    if (tid < 64) { S.a1[tid] = A.params_simt[A.a1_off + tid]; S.b1[tid] = A.params_simt[A.a1_off + 64 + tid]; }
    if (tid == 0) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&S.tmem_base, kTmemCols);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  const uint32_t tmem = S.tmem_base;
  const uint32_t tmem_d0 = tmem, tmem_d1 = tmem + 32;
  const uint32_t sa1h = smem_u32(S.a1h), sa1l = smem_u32(S.a1l);
  const uint32_t sa0h = sa1h + 8 * kRows * 16, sa0l = sa1l + 8 * kRows * 16;   // A0 aliases A1 panels 8..
  float* const a0h = S.a1h + 8 * kRows * 4;
  float* const a0l = S.a1l + 8 * kRows * 4;
  const uint32_t idesc32 = idesc_tf32_m128(32), idesc64 = idesc_tf32_m128(64);
  uint32_t phase0 = 0, phase1 = 0;
  const int row = (warp & 3) * 32 + lane;      // TMEM lane == sub-tile row owned in the epilogues
  const int half = warp >> 2;                  // which half of the accumulator columns this warp reads

  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int pa = A.tile_first[grp], pb = A.tile_first[grp + 1];
    const int npil = pb - pa;
    if (npil <= 0) continue;
    __syncthreads();
    for (int i = tid; i <= npil; i += kTcThreads) S.off[i] = A.seg_off[pa + i];
    if (tid == 0) { S.sub_begin = 0; S.sub_end = 0; }
    __syncthreads();

    while (true) {
      // ---- greedy sub-tile: pillars [la, lb) with at most 128 rows; pillars longer than 128 rows are skipped ----
      if (tid == 0) {
        int la = S.sub_end;
        while (la < npil && S.off[la + 1] - S.off[la] > kRows) ++la;     // long pillar: SIMT kernel's job
        int lb = la;
        while (lb < npil && S.off[lb + 1] - S.off[la] <= kRows) ++lb;
        S.sub_begin = la; S.sub_end = lb;
      }
      __syncthreads();
      const int la = S.sub_begin, lb = S.sub_end;
      if (la >= npil) break;
      const int r0 = S.off[la];
      const int nrows = S.off[lb] - r0;

      // ================= P1a: gather =================
      if (tid < kRows) {
        float x = 0.f, y = 0.f, z = 0.f;
        int lp = -1;
        if (tid < nrows) {
          const int pos = r0 + tid;
          int lo = la, hi = lb;                      // last pillar with off <= pos
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (S.off[mid] <= pos) lo = mid; else hi = mid;
          }
          lp = lo;
          const float* rp = A.points + (int64_t)__ldg(A.sorted_idx + pos) * A.stride;
          x = __ldg(rp + 1); y = __ldg(rp + 2); z = __ldg(rp + 3);
          // raw features staged (unsplit) in the A0 hi panels
          for (int f = 0; f < A.n_raw; ++f) a0h[(f >> 2) * (kRows * 4) + tid * 4 + (f & 3)] = __ldg(rp + A.raw_col0 + f);
        }
        S.lp[tid] = lp;
        S.xyz[0][tid] = x; S.xyz[1][tid] = y; S.xyz[2][tid] = z;
      }
      __syncthreads();
      // ================= P1b: per-pillar mean (sequential, ascending row order) =================
      if (tid < lb - la) {
        const int p = la + tid;
        const int qs = S.off[p] - r0, qe = S.off[p + 1] - r0;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int q = qs; q < qe; ++q) {
          sx = __fadd_rn(sx, S.xyz[0][q]); sy = __fadd_rn(sy, S.xyz[1][q]); sz = __fadd_rn(sz, S.xyz[2][q]);
        }
        const float cnt = (float)max(qe - qs, 1);
        const float mx = __fdiv_rn(sx, cnt), my = __fdiv_rn(sy, cnt), mz = __fdiv_rn(sz, cnt);
        S.mean[0][p] = mx; S.mean[1][p] = my; S.mean[2][p] = mz;
        if (A.mean_out) {
          float* m = A.mean_out + (int64_t)(pa + p) * 3;
          m[0] = mx; m[1] = my; m[2] = mz;
        }
      }
      __syncthreads();
      // ================= P1c: features + TF32 split -> A0 =================
      // the A0 hi panels double as an fp32 staging row: raw features were written by P1a, the derived
      // features are appended here, then every panel is split in place (hi stays, lo goes to A0 lo)
      if (tid < kRows) {
        if (tid < nrows) {
          const int lp = S.lp[tid];
          const float x = S.xyz[0][tid], y = S.xyz[1][tid], z = S.xyz[2][tid];
          float e[8];
          e[0] = __fsub_rn(x, S.mean[0][lp]);                                  // f_cluster (:111)
          e[1] = __fsub_rn(y, S.mean[1][lp]);
          e[2] = __fsub_rn(z, S.mean[2][lp]);
          const float cx = quantise(x, A.g.range_min_x, A.g.voxel_x);
          const float cy = quantise(y, A.g.range_min_y, A.g.voxel_y);
          e[3] = __fsub_rn(x, __fadd_rn(__fmul_rn(cx, A.g.voxel_x), A.g.x_offset));   // f_center (:114-116)
          e[4] = __fsub_rn(y, __fadd_rn(__fmul_rn(cy, A.g.voxel_y), A.g.y_offset));
          e[5] = __fsub_rn(z, A.g.z_offset);
          e[6] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));  // :124
          e[7] = 0.f;
          const int n_derived = A.with_distance ? 7 : 6;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int f = A.n_raw + j;
            if (f < k0) a0h[(f >> 2) * (kRows * 4) + tid * 4 + (f & 3)] = (j < n_derived) ? e[j] : 0.f;
          }
          for (int f = A.n_raw + 8; f < k0; ++f) a0h[(f >> 2) * (kRows * 4) + tid * 4 + (f & 3)] = 0.f;
        }
        for (int kc = 0; kc * 4 < k0; ++kc) {
          float* ph = a0h + kc * (kRows * 4) + tid * 4;
          float4 v = (tid < nrows) ? ld4(ph) : make_float4(0.f, 0.f, 0.f, 0.f);
          float h[4], l[4];
          split_tf32(v.x, h[0], l[0]); split_tf32(v.y, h[1], l[1]);
          split_tf32(v.z, h[2], l[2]); split_tf32(v.w, h[3], l[3]);
          st4(ph, h[0], h[1], h[2], h[3]);
          st4(a0l + kc * (kRows * 4) + tid * 4, l[0], l[1], l[2], l[3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncthreads();
      // ================= M0 =================
      if (tid == 0) {
        tc_fence_after_sync();
        mma_3xtf32(tmem_d0, sa0h, sa0l, kRows, smem_u32(S.w0h), smem_u32(S.w0l), 32, k0 / 8, idesc32, false);
        mma_commit(&S.bar[0]);
      }
      mbar_wait(&S.bar[0], phase0);
      phase0 ^= 1;
      tc_fence_after_sync();
      // ================= P2: layer-0 epilogue -> A1 panels 0..7 =================
      if ((warp & 3) * 32 < nrows) {
        float v[16];
        tmem_ld16(tmem_d0 + ((uint32_t)((warp & 3) * 32) << 16) + half * 16, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float h[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = half * 16 + j * 4 + i;
            const float x = fmaxf(fmaf(v[j * 4 + i], S.a0[c], S.b0[c]), 0.f);
            split_tf32(x, h[i], l[i]);
          }
          const int kc = half * 4 + j;
          st4(S.a1h + kc * (kRows * 4) + row * 4, h[0], h[1], h[2], h[3]);
          st4(S.a1l + kc * (kRows * 4) + row * 4, l[0], l[1], l[2], l[3]);
        }
      }
      tc_fence_before_sync();
      __syncthreads();
      // ================= P3: per-pillar max of x0 -> A1 panels 8..15 of every row of the pillar =================
      {
        const int np = lb - la;
        for (int item = tid; item < np * 8; item += kTcThreads) {
          const int p = la + item % np, kc = item / np;
          const int qs = S.off[p] - r0, qe = S.off[p + 1] - r0;
          const float* ph = S.a1h + kc * (kRows * 4);
          const float* pl = S.a1l + kc * (kRows * 4);
          float4 h = ld4(ph + qs * 4), l = ld4(pl + qs * 4);
          for (int q = qs + 1; q < qe; ++q) {
            const float4 h2 = ld4(ph + q * 4), l2 = ld4(pl + q * 4);
            lexmax(h.x, l.x, h2.x, l2.x); lexmax(h.y, l.y, h2.y, l2.y);
            lexmax(h.z, l.z, h2.z, l2.z); lexmax(h.w, l.w, h2.w, l2.w);
          }
          float* qh = S.a1h + (8 + kc) * (kRows * 4);
          float* ql = S.a1l + (8 + kc) * (kRows * 4);
          for (int q = qs; q < qe; ++q) {
            *reinterpret_cast<float4*>(qh + q * 4) = h;
            *reinterpret_cast<float4*>(ql + q * 4) = l;
          }
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      // ================= M1 =================
      if (tid == 0) {
        tc_fence_after_sync();
        mma_3xtf32(tmem_d1, sa1h, sa1l, kRows, smem_u32(S.w1h), smem_u32(S.w1l), 64, 8, idesc64, false);
        mma_commit(&S.bar[1]);
      }
      mbar_wait(&S.bar[1], phase1);
      phase1 ^= 1;
      tc_fence_after_sync();
      // ================= P4: layer-1 epilogue -> y (aliases A1 hi; the MMAs that read it have completed) =================
      if ((warp & 3) * 32 < nrows) {
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          float v[16];
          tmem_ld16(tmem_d1 + ((uint32_t)((warp & 3) * 32) << 16) + half * 32 + part * 16, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = half * 32 + part * 16 + j * 4 + i;
              o[i] = fmaxf(fmaf(v[j * 4 + i], S.a1[c], S.b1[c]), 0.f);
            }
            const int f4 = half * 8 + part * 4 + j;                         // float4 slot 0..15 of the row
            st4(S.a1h + row * 64 + ((f4 ^ (row & 15)) << 2), o[0], o[1], o[2], o[3]);
          }
        }
      }
      tc_fence_before_sync();
      __syncthreads();
      // ================= P5: per-pillar max of y -> pillar_features =================
      {
        const int np = lb - la;
        for (int item = tid; item < np * 16; item += kTcThreads) {
          const int p = la + (item >> 4), f4 = item & 15;
          const int qs = S.off[p] - r0, qe = S.off[p + 1] - r0;
          float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int q = qs; q < qe; ++q) {
            const float4 t = ld4(S.a1h + q * 64 + ((f4 ^ (q & 15)) << 2));
            m.x = fmaxf(m.x, t.x); m.y = fmaxf(m.y, t.y); m.z = fmaxf(m.z, t.z); m.w = fmaxf(m.w, t.w);
          }
          *reinterpret_cast<float4*>(A.out + (int64_t)(pa + p) * 64 + f4 * 4) = m;
        }
      }
      __syncthreads();
      if (lb >= npil) break;
    }
  }
  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// TC parameter panels: w0h | w0l ([k0/4][32][4]) | w1h | w1l ([16][64][4]), round-to-nearest TF32 split
// ------------------------------------------------------------------------------------------------
__global__ void pack_tc_params_kernel(int c_in, int k0, const float* __restrict__ w0, const float* __restrict__ w1,
                                      float* __restrict__ out) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int n0 = k0 * 32, n1 = 64 * 64;
  for (int i = tid; i < n0; i += nth) {
    const int kc = i / (32 * 4), n = (i / 4) % 32, k = kc * 4 + (i & 3);
    const float w = (k < c_in) ? w0[n * c_in + k] : 0.f;
    float h, l;
    split_tf32_rn(w, h, l);
    out[i] = h; out[n0 + i] = l;
  }
  for (int i = tid; i < n1; i += nth) {
    const int kc = i / (64 * 4), n = (i / 4) % 64, k = kc * 4 + (i & 3);
    float h, l;
    split_tf32_rn(w1[n * 64 + k], h, l);
    out[2 * n0 + i] = h; out[2 * n0 + n1 + i] = l;
  }
}

// ------------------------------------------------------------------------------------------------
// self test: C[128 x N] = A[128 x K] . B[N x K]^T through the exact operand layout / descriptor / TMEM path
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads)
umma_selftest_kernel(const float* __restrict__ Ag, const float* __restrict__ Bg, int K, int N, float* __restrict__ Cg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ah = reinterpret_cast<float*>(smem_raw);
  float* al = ah + 64 * kRows;
  float* bh = al + 64 * kRows;
  float* bl = bh + 64 * 64;
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kRows * K; i += kTcThreads) {
    const int r = i / K, k = i % K;
    float h, l;
    split_tf32(Ag[i], h, l);
    ah[(k >> 2) * (kRows * 4) + r * 4 + (k & 3)] = h;
    al[(k >> 2) * (kRows * 4) + r * 4 + (k & 3)] = l;
  }
  for (int i = tid; i < N * K; i += kTcThreads) {
    const int n = i / K, k = i % K;
    float h, l;
    split_tf32_rn(Bg[i], h, l);
    bh[(k >> 2) * (N * 4) + n * 4 + (k & 3)] = h;
    bl[(k >> 2) * (N * 4) + n * 4 + (k & 3)] = l;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    mma_3xtf32(tmem, smem_u32(ah), smem_u32(al), kRows, smem_u32(bh), smem_u32(bl), (uint32_t)N, K / 8,
               idesc_tf32_m128((uint32_t)N), false);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  for (int c0 = half * 16; c0 < N; c0 += 32) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) Cg[row * N + c0 + i] = v[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

}  // namespace pcp

using namespace pcp;

extern "C" int pcp_selftest_umma(const float* a, const float* b, int32_t k, int32_t n, float* c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(a && b && c, PCP_E_INVALID, "pcp_selftest_umma: null argument");
  PCP_REQUIRE(k > 0 && k <= 64 && k % 8 == 0 && (n == 32 || n == 64), PCP_E_INVALID, "pcp_selftest_umma: bad k/n");
  const size_t smem = sizeof(float) * (2 * 64 * kRows + 2 * 64 * 64);
  PCP_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, kTcThreads, smem, stream>>>(a, b, k, n, c);
  PCP_LAUNCH_CHECK("umma_selftest_kernel");
  return 0;
}

// launched from pfn.cu
namespace pcp {
int launch_pfn_tc(const TcArgs& a, int64_t n_points, cudaStream_t stream) {
  const size_t smem = sizeof(TcSmem);
  PCP_CUDA(cudaFuncSetAttribute(pfn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t groups = (n_points + kWin - 1) / kWin;
  const unsigned blocks = (unsigned)(groups < 296 ? (groups > 0 ? groups : 1) : 296);
  pfn_tc_kernel<<<blocks, kTcThreads, smem, stream>>>(a);
  PCP_LAUNCH_CHECK("pfn_tc_kernel");
  return 0;
}
int launch_pack_tc(int c_in, int k0, const float* w0, const float* w1, float* out, cudaStream_t stream) {
  pack_tc_params_kernel<<<8, 256, 0, stream>>>(c_in, k0, w0, w1, out);
  PCP_LAUNCH_CHECK("pack_tc_params_kernel");
  return 0;
}
}  // namespace pcp
