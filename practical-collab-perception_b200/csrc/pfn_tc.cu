// pfn_tc.cu - the PFN on 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 for fp32-grade accuracy.
//
// Reference: dynamic_pillar_vfe.py:110-129 + PFNLayerV2.forward :35-46 (NUM_FILTERS [64, 64] or [64]).
//
// Layout idea: ONE PILLAR PER TENSOR-MEMORY LANE.  A group is 128 pillars of one length class (work lists
// built by the voxelize scan); slot j of the group is the j-th point of each of its pillars (pillars shorter
// than the class repeat their last point).  Every slot is one M = 128 MMA tile whose accumulator row p
// belongs to pillar p, so the per-pillar max - the reference's scatter_max - is an ELEMENTWISE max between
// successive accumulators inside the thread that owns lane p: no cross-lane traffic, no shared-memory
// transpose, no atomics.  Per slot:
//   A0   the slot's rows (gathered by cp.async), features (f_cluster needs the pillar mean, f_center the pillar's
//        cell - both per-pillar values prepared by pillar_prep_kernel), TF32 hi/lo split -> A0 in tensor memory
//   M0   D0[128 x 32] = A0 . W0^T                      (tcgen05.mma kind::tf32, 3 MMAs per K step)
//   E0   TMEM -> regs: BN(eval)+ReLU -> x0; running max0; hi/lo split -> written BACK TO TENSOR MEMORY as the
//        A operand of layer 1 (tcgen05.st): the N' x 32 activation never touches shared memory or HBM
//   M1   D1[128 x 64] = x0 . W1[:, :32]^T              (A from TMEM, B from smem)
//   E1   TMEM -> regs: running max m1 (raw accumulators: BN+ReLU are applied once per pillar, see below)
// and once per group
//   H    D [128 x 64] = max0 . W1[:, 32:]^T            the x_max half of torch.cat([x, x_max[unq_inv]]) is the same
//        for every point of a pillar, so it is evaluated once per pillar and added after the max:
//        max_i(P_i + h) == max_i(P_i) + h exactly in fp32 because rounding is monotone
//   OUT  relu(|a1| * (m1 + h) + b1): BN+ReLU after the max is exact because the sign of the BN scale is folded
//        into the weight rows, which makes the map non-decreasing
// Pillars above 32 points are cut into 32-row segments that run through the same slots; their partial maxima
// meet in a per-pillar ordered-int accumulator (atomicMax) and pfn_finish_long_kernel (pfn.cu) applies H / OUT.
#include <type_traits>

#include "internal.cuh"
#include "umma.cuh"
#include "pfn_tc.cuh"

namespace pcp {

using namespace umma;

STAGE_TABLE(g_stage_pfn);           // 0 pfn_slot_kernel
constexpr int kTcThreads = 256;     // threads of the self-test kernels
// ---- roles of the PFN kernel (one persistent CTA per SM, 25 warps) ----
//   warps  0.. 7  E0: layer-0 epilogue      (pillar p = tid & 127, column half h = (tid >> 7) & 1)
//   warps  8..15  E1: last-layer epilogue   (same mapping) + the output rows
//   warps 16..23  producers: two sets of 128 threads (one thread per pillar); set s builds the slots of parity s
//   warp  24      MMA issue
constexpr int kE0Threads = 256;
constexpr int kE1Threads = 256;
constexpr int kProdSets = 2;
constexpr int kProdThreads = kGroup * kProdSets;
// Warp numbering of the roles.  The issue arbiter of an SM sub-partition prefers the eligible warp with the HIGHEST warp id
// (B300_MICROARCH.md, "Multi-warp arbiter"), so the role on the critical path - the layer-0 epilogue, which every other
// role ends up waiting for - gets the highest ids below the MMA warp, and the producers, which run several slots ahead, the
// lowest.  Every role is 8 warps starting at a multiple of 4: warp & 3 is its tensor-memory lane quarter in any order.
#ifndef PCP_PFN_ORDER
#define PCP_PFN_ORDER 1
#endif
#if PCP_PFN_ORDER == 0       // round-1 order: E0, E1, producers
constexpr int kE0Warp0 = 0, kE1Warp0 = 8, kProdWarp0 = 16;
#elif PCP_PFN_ORDER == 1     // producers, E1, E0
constexpr int kProdWarp0 = 0, kE1Warp0 = 8, kE0Warp0 = 16;
#else                        // E1, producers, E0
constexpr int kE1Warp0 = 0, kProdWarp0 = 8, kE0Warp0 = 16;
#endif
constexpr int kMmaWarp = 24;
constexpr int kPfnThreads = (kMmaWarp + 1) * 32;                  // 800
constexpr int kTmemCols = 512;
// tensor-memory column map.  The FRONT half of the pipeline (features -> layer 0 -> E0) is NF buffers deep, the back
// half (x0 -> layer 1 -> E1) two: with 32-column operands / accumulators (two-layer PFN, k0 = 16) NF = 4, else 2.
constexpr uint32_t kColD0 = 0;      // layer-0 accumulators   [b * (128 / NF), +N0)
constexpr uint32_t kColD1 = 128;    // layer-1 / hoist accumulators [128 + b * 64, +64)
constexpr uint32_t kColA0 = 256;    // layer-0 A operand (features): hi at 256 + b * (128 / NF), lo 64 / NF columns further
constexpr uint32_t kColA1 = 384;    // layer-1 A operand (x0 / max0): hi at 384 + b * 64, lo 32 columns further
constexpr int kMaxNF = 4;
constexpr int kRowRing = 16;        // ring of per-group output-row tables (producers run a few groups ahead of the output)
constexpr int kGTab = 256;          // per-CTA table of (slots, is-segment) of its first groups (later ones are recomputed)
constexpr int kOutRowBytes = 144;   // E1 output staging: 128 bytes per thread + 16 bytes of padding (conflict-free STS.128)
// mbarrier indices
constexpr int kBarA0 = 0;           // [NF] features staged in TMEM         (128 arrivals: one producer set)
constexpr int kBarD0 = 4;           // [NF] layer-0 accumulator ready       (tcgen05.commit)
constexpr int kBarA1 = 8;           // [2] x0 / max0 staged in TMEM         (256 arrivals: E0)
constexpr int kBarD1 = 10;          // [2] layer-1 / hoist accumulator ready (tcgen05.commit)
constexpr int kBarF1 = 12;          // [2] layer-1 accumulator consumed     (256 arrivals: E1)
constexpr int kNumBars = 14;

// ---- optional per-role event trace of CTA 0 (debug build only: make dbg; tools/pfn_timing.py) ----
#ifdef PCP_PFN_TIMING
constexpr int kTraceCap = 8192;
constexpr int kTraceRoles = 5;      // mma, producer set 0, E0, E1, producer set 1
__device__ long long g_trace[kTraceRoles][kTraceCap][2];
__device__ int g_trace_n[kTraceRoles];
#define TRACE_DECL(cond) const bool trace_on_ = (blockIdx.x == 0) && (cond); int trace_cnt_ = 0;
#define TRACE(role, id)                                                                          \
  do {                                                                                           \
    if (trace_on_ && trace_cnt_ < kTraceCap) {                                                   \
      g_trace[role][trace_cnt_][0] = (id); g_trace[role][trace_cnt_][1] = clock64(); ++trace_cnt_; \
    }                                                                                            \
  } while (0)
#define TRACE_END(role) do { if (trace_on_) g_trace_n[role] = trace_cnt_; } while (0)
#else
#define TRACE_DECL(cond)
#define TRACE(role, id)
#define TRACE_END(role)
#endif

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_global_f4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// kCfg: 0 = any layout (4-byte copies, run-time feature map)
//       1 = car / early-fusion rows: c_raw 5, absolute xyz, no distance, row stride % 4 == 0, 16-byte aligned
//       2 = ego (lately fusion) rows: c_raw 11, absolute xyz, no distance, even row stride, 8-byte aligned
//   nreg  = floats of a row staged per slot, depth = rows in flight per producer thread (cp.async ring of depth + 1 stages)
template <int kCfg> struct RowCfg { static constexpr int n_raw = 0, k0 = 0, nreg = 24, depth = 1; };
template <> struct RowCfg<1> { static constexpr int n_raw = 5, k0 = 16, nreg = 8, depth = 3; };
template <> struct RowCfg<2> { static constexpr int n_raw = 11, k0 = 24, nreg = 12, depth = 2; };
// per producer set: ring of per-group work-list entries (cursor A runs up to 6 * depth + 4 groups ahead of cursor D)
__host__ __device__ constexpr int ent_ring(int depth) { return 6 * depth + 4 <= 16 ? 16 : 32; }

struct SmemPlan {   // float offsets into dynamic shared memory
  int w0h, w0l, w1ah, w1al, w1bh, w1bl, w1sh, w1sl, panels_end, prm_a0, prm_b0, prm_a1, prm_b1, ent, idx, mean, row, rows, ostage, gtab, curs, ints, total_bytes;
  int ent_set, idx_set, mean_set, row_set;   // floats per producer set
};
__host__ __device__ inline SmemPlan smem_plan(int k0, int layers, int nreg, int depth) {
  SmemPlan S{};
  const int n0 = layers == 2 ? kHidden : kCout;
  int o = 0;
  S.w0h = o; o += k0 * n0;
  S.w0l = o; o += k0 * n0;
  if (layers == 2) {
    S.w1ah = o; o += kHidden * kCout;
    S.w1al = o; o += kHidden * kCout;
    S.w1bh = o; o += kHidden * kCout;
    S.w1bl = o; o += kHidden * kCout;
    S.w1sh = o; o += kHidden * kCout;
    S.w1sl = o; o += kHidden * kCout;
  }
  S.panels_end = o;
  S.prm_a0 = o; o += n0;
  S.prm_b0 = o; o += n0;
  if (layers == 2) {
    S.prm_a1 = o; o += kCout;
    S.prm_b1 = o; o += kCout;
  }
  S.ent_set = ent_ring(depth) * kGroup * 2;     // work-list entries, [group ring][pillar] (8 bytes each)
  S.idx_set = (depth + 1) * kGroup;             // row numbers, [slot ring][pillar]
  S.mean_set = (depth + 1) * kGroup * 4;        // per-pillar mean + cell, [slot ring][pillar] (16 bytes each)
  S.row_set = (depth + 1) * nreg * kGroup;      // staged rows, [slot ring][16/8/4-byte chunk][pillar]
  S.ent = o; o += kProdSets * S.ent_set;
  S.idx = o; o += kProdSets * S.idx_set;
  S.mean = o; o += kProdSets * S.mean_set;
  S.row = o; o += kProdSets * S.row_set;
  S.rows = o; o += kRowRing * kGroup;             // output row (pillar rank) / long-pillar index of each lane
  S.ostage = o; o += (kE1Threads * kOutRowBytes) / 4;   // E1 output staging: one padded 128-byte row per thread
  S.gtab = o; o += kGTab;                         // (slots | is_seg << 8) of this CTA's first kGTab groups
  S.curs = o; o += (kProdThreads / 32) * 8;       // per producer warp: ring of packed cursor states (group, slot, is-segment)
  S.ints = o; o += 96;                            // 14 mbarriers | tmem base | group prefix | list counts | list offsets
  S.total_bytes = o * 4;
  return S;
}

// Persistent, warp-specialised kernel.  The roles talk only through mbarriers:
//   producers (2 x 4 warps, thread = pillar lane): gather each slot's row through a cp.async pipeline, build the feature
//       vector and write its TF32 hi / lo parts to tensor memory (A0); set s owns the slots of parity s == buffer s;
//   MMA warp: waits for "operand staged", issues the tcgen05.mma groups (layer 0 of the NEXT slot is queued in front
//       of layer 1 of the current one, so the tensor pipe has work while the epilogues run), commits to "accumulator ready";
//   E0 (8 warps): BN + ReLU of the layer-0 accumulator, running max0, x0 back to TMEM as layer 1's A operand, and once
//       per group max0 as the operand of the hoist;
//   E1 (8 warps): running max of the layer-1 accumulators, the hoist result, BN + ReLU and the output rows (each thread
//       owns 32 consecutive channels of its pillar: 128 contiguous bytes of pillar_features).
template <int kLayers, int kCfg>
__global__ void __launch_bounds__(kPfnThreads, 1)
pfn_slot_kernel(const TcArgs A) {
  extern __shared__ __align__(128) float smem[];
  constexpr int N0 = (kLayers == 2) ? kHidden : kCout;
  constexpr int NREG = RowCfg<kCfg>::nreg;
  constexpr int NF = (kLayers == 2 && kCfg == 1) ? 4 : 2;        // depth of the front half (A0 / D0 buffers)
  constexpr uint32_t kFS = 128 / NF;                             // columns per A0 / D0 buffer
  constexpr uint32_t kA0Lo = 64 / NF;                            // offset of the lo part inside an A0 buffer
  // the fixed row layouts have c_in = 11 / 17: a spare K column exists, pcp_pack_pfn_params folded layer 0's BN
  constexpr bool kFold0 = (kLayers == 2) && (kCfg != 0);
  static_assert(!kFold0 || RowCfg<kCfg>::k0 > RowCfg<kCfg>::n_raw + 6, "no spare K column for the folded bias");
  const int k0 = kCfg ? RowCfg<kCfg>::k0 : A.k0;
  const int n_raw = kCfg ? RowCfg<kCfg>::n_raw : A.n_raw;
  const int raw_col0 = kCfg ? 1 : A.raw_col0;
  const bool with_dist = kCfg ? false : (A.with_distance != 0);
  const SmemPlan SP = smem_plan(k0, kLayers, NREG, RowCfg<kCfg>::depth);
  const int tid = threadIdx.x, warp = tid >> 5;
  STAGE_BEGIN(g_stage_pfn, 0);
  int* const s_rows = reinterpret_cast<int*>(smem + SP.rows);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + SP.ints);
  uint32_t* const s_tmem = reinterpret_cast<uint32_t*>(smem + SP.ints + 28);
  int* const s_pre = reinterpret_cast<int*>(smem + SP.ints + 32);      // [kNumLists + 1] group prefix, processing order
  int* const s_cnt = reinterpret_cast<int*>(smem + SP.ints + 44);      // [kNumLists] entries per list
  long long* const s_loff = reinterpret_cast<long long*>(smem + SP.ints + 56);   // [kNumLists] list offsets
  int* const s_gtab = reinterpret_cast<int*>(smem + SP.gtab);

  // ---- one-time setup: parameters -> smem, barriers, TMEM, work prefix ----
  {
    // the packed parameters do not depend on the voxelize kernels: this copy runs while pillar_prep is still draining
    // (programmatic dependent launch, internal.cuh); everything below reads what voxelize wrote
    const float4* src = reinterpret_cast<const float4*>(A.params);
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < SP.panels_end / 4; i += kPfnThreads) dst[i] = __ldg(src + i);   // same order in both layouts
    pdl_wait();
    const ParamLayout PL = param_layout(A.c_in, kLayers);
    for (int i = tid; i < N0; i += kPfnThreads) {
      smem[SP.prm_a0 + i] = A.params[PL.a0 + i];
      smem[SP.prm_b0 + i] = A.params[PL.b0 + i];
    }
    if (kLayers == 2)
      for (int i = tid; i < kCout; i += kPfnThreads) {
        smem[SP.prm_a1 + i] = A.params[PL.a1 + i];
        smem[SP.prm_b1 + i] = A.params[PL.b1 + i];
      }
    if (tid == 0) {
      for (int b = 0; b < kMaxNF; ++b) {
        mbar_init(&bars[kBarA0 + b], kGroup);
        mbar_init(&bars[kBarD0 + b], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars[kBarA1 + b], kE0Threads);
        mbar_init(&bars[kBarD1 + b], 1);
        mbar_init(&bars[kBarF1 + b], kE1Threads);
      }
      fence_mbar_init();
      int acc = 0;
      for (int q = 0; q < kNumLists; ++q) {                     // processing order: segments, then classes 9 .. 0
        const int list = kNumLists - 1 - q;
        const int cnt = A.hdr[kHdrListCount + list];
        s_pre[q] = acc;
        s_cnt[list] = cnt;
        s_loff[list] = A.lo.off[list];
        acc += (cnt + kGroup - 1) / kGroup;
      }
      s_pre[kNumLists] = acc;
    }
    if (warp == kMmaWarp) tmem_alloc(s_tmem, kTmemCols);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  const uint32_t tmem = *s_tmem;
  const int total = s_pre[kNumLists];
  const int G = gridDim.x;
  auto list_of = [&](int w, int& q) {
    q = 0;
#pragma unroll
    for (int t = 1; t < kNumLists; ++t) q += (w >= s_pre[t]) ? 1 : 0;
    return kNumLists - 1 - q;
  };
  // Work items are handed out in rounds of G (CTA b takes item b + gi * G of the size-descending sequence).  Every full odd
  // round runs backwards over the CTAs ("snake"), so that the CTA that received the larger item at a class boundary of one
  // round receives the smaller one in the next: the slot count of the busiest CTA drops from 7.8 % to 4.1 % above the mean
  // on the bench batch.  `w` stays the loop variable of every role; item_of() gives the work item behind it.
  auto item_of = [&](int w, int gi) {
    const int r0 = gi * G;
    return ((gi & 1) && r0 + G <= total) ? r0 + (G - 1 - (w - r0)) : w;
  };
  auto slots_calc = [&](int w, int gi) {         // slots | is_seg << 8 of the item CTA position w takes in round gi
    int q;
    const int list = list_of(item_of(w, gi), q);
    return list == kSegList ? (kSegRows | 256) : class_slots(list);
  };
  // (slots, is-segment) of this CTA's first groups, looked up by every role at every group boundary
  for (int gi = tid; gi < kGTab; gi += kPfnThreads) {
    const long long w = (long long)blockIdx.x + (long long)gi * G;
    s_gtab[gi] = (w < total) ? slots_calc((int)w, gi) : 1;
  }
  __syncthreads();
  // gi = index of the group among this CTA's groups (w = blockIdx.x + gi * G)
  auto slots_of = [&](int w, int gi, bool& is_seg) {
    const int v = gi < kGTab ? s_gtab[gi] : slots_calc(w, gi);
    is_seg = (v & 256) != 0;
    return v & 255;
  };

  if (warp == kMmaWarp) {
    // =====================================================================================================
    // MMA warp
    // =====================================================================================================
    // The issue warp hands registers above 56 back to the CTA pool.  Nobody claims them; the instruction is here because
    // ptxas schedules the kernel measurably better with it (161 vs 168 us).  ptxas does NOT confine the code that follows
    // to the lowered budget by itself: tests/test_abi.py checks in the SASS that this role stays below 56 registers.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const uint32_t sw0h = smem_u32(smem + SP.w0h), sw0l = smem_u32(smem + SP.w0l);
    const uint32_t sw1ah = smem_u32(smem + SP.w1ah), sw1al = smem_u32(smem + SP.w1al);
    const uint32_t sw1bh = smem_u32(smem + SP.w1bh), sw1bl = smem_u32(smem + SP.w1bl);
    const uint32_t sw1sh = smem_u32(smem + SP.w1sh), sw1sl = smem_u32(smem + SP.w1sl);
    const uint32_t idesc0 = idesc_tf32_m128(N0), idesc1 = idesc_tf32_m128(kCout);
    uint32_t c0 = 0, c1 = 0;      // layer-0 ops / layer-1-type ops issued so far
    TRACE_DECL((tid & 31) == 0)
    auto issue_m0 = [&]() {
      const uint32_t b = c0 % NF;
      TRACE(0, 20);
      mbar_wait(&bars[kBarA0 + b], (c0 / NF) & 1);
      if (kLayers == 1 && c0 >= NF) mbar_wait(&bars[kBarA1 + b], ((c0 - NF) / NF) & 1);   // D0[b] consumed by E0
      TRACE(0, 21);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        mma_3xtf32_ts(tmem + kColD0 + b * kFS, tmem + kColA0 + b * kFS, tmem + kColA0 + b * kFS + kA0Lo, sw0h, sw0l, N0,
                      k0 / 8, idesc0, false);
        mma_commit(&bars[kBarD0 + b]);
      }
      __syncwarp();
      ++c0;
    };
    auto issue_m1 = [&](uint32_t wh, uint32_t wl) {
      const uint32_t b = c1 & 1;
      TRACE(0, 22);
      mbar_wait(&bars[kBarA1 + b], (c1 >> 1) & 1);
      if (c1 >= 2) mbar_wait(&bars[kBarF1 + b], ((c1 - 2) >> 1) & 1);                    // D1[b] consumed by E1
      TRACE(0, 23);
      tc_fence_after_sync();
      if (elect_one_sync()) {
        mma_3xtf32_ts(tmem + kColD1 + b * 64, tmem + kColA1 + b * 64, tmem + kColA1 + b * 64 + 32, wh, wl, kCout, kHidden / 8,
                      idesc1, false);
        mma_commit(&bars[kBarD1 + b]);
      }
      __syncwarp();
      ++c1;
    };
    // Layer 0 runs up to NF - 1 slots ahead of layer 1.  D0[b] is overwritten by the layer-0 MMA NF slots later; that
    // MMA is issued after the layer-1 MMA of the slot it replaces, which waited for A1, i.e. for E0 having read D0[b].
    uint32_t nslots = 0;          // slots of this CTA
    {
      int gi = 0;
      for (int w = blockIdx.x; w < total; w += G, ++gi) { bool sg; nslots += slots_of(w, gi, sg); }
    }
    for (int i = 0; i < NF - 1 && c0 < nslots; ++i) issue_m0();
    int gi = 0;
    for (int w = blockIdx.x; w < total; w += G, ++gi) {
      bool is_seg;
      const int slots = slots_of(w, gi, is_seg);
      // one-point pillars: x_max == x, so x . (W1[:, :32] + W1[:, 32:])^T is the whole last layer - no hoist
      const bool single = (slots == 1) && !is_seg;
      for (int j = 0; j < slots; ++j) {
        if (c0 < nslots) issue_m0();                           // layer 0 of a LATER slot goes in front of this slot's layer 1
        if (kLayers == 2) { if (single) issue_m1(sw1sh, sw1sl); else issue_m1(sw1ah, sw1al); }
      }
      if (kLayers == 2 && !is_seg && !single) issue_m1(sw1bh, sw1bl);     // hoist: max0 . W1[:, 32:]^T once per pillar
      TRACE(0, 24);
    }
    TRACE_END(0);
  } else if (warp >= kProdWarp0 && warp < kProdWarp0 + kProdThreads / 32) {
    // =====================================================================================================
    // producers
    //
    // A producer thread owns TMEM lane p, i.e. pillar p of every group of this CTA, and builds every second slot
    // (set s: the slots of parity s, which live in operand buffer s).  Everything it needs arrives through a software
    // pipeline of cp.async copies that it issues itself and that only it reads back (no cross-thread
    // synchronisation); one iteration = one of its slots, the stages are DEPTH iterations apart:
    //   cursor A  work-list entry of a group                    -> s_ent   (2 * DEPTH + 2 groups ahead of cursor B)
    //   cursor B  row number of its slot i + 2 * DEPTH           -> s_idx   (needs its group's entry)
    //   cursor C  point row + pillar mean of its slot i + DEPTH  -> s_row, s_mean (needs the row number)
    //   cursor D  slot i: features -> TF32 hi / lo -> tensor memory (A operand of layer 0)
    // One commit group per iteration and a single cp.async.wait_group<DEPTH - 1> make everything issued DEPTH or
    // more iterations ago visible, which is exactly what cursors B, C and D read.
    // =====================================================================================================
    constexpr int DEPTH = RowCfg<kCfg>::depth;
    constexpr int STAGES = DEPTH + 1;
    constexpr int LEAD = 2 * DEPTH + 2;                    // cursor A's lead over cursor B, in groups
    constexpr int kEntRing = ent_ring(DEPTH);
    static_assert(4 * DEPTH + LEAD + 2 <= kEntRing, "entry ring too small");
    const int set = (tid - kProdWarp0 * 32) >> 7;          // 0 / 1
    const int p = tid & (kGroup - 1);                      // pillar of the group == TMEM lane (kProdWarp0 * 32 % 128 == 0)
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    unsigned long long* const s_ent = reinterpret_cast<unsigned long long*>(smem + SP.ent + set * SP.ent_set);
    int* const s_idx = reinterpret_cast<int*>(smem + SP.idx + set * SP.idx_set);
    float4* const s_mean = reinterpret_cast<float4*>(smem + SP.mean + set * SP.mean_set);
    float* const s_row = smem + SP.row + set * SP.row_set;
    constexpr unsigned long long kNoEntry = ~0ull;
    const int n_cols = kCfg ? NREG : A.c_raw;              // generic layout: columns 1 .. c_raw are staged

    struct Cursor { int w, j, slots, gi; bool seg; };      // work item, slot inside it, its slot count, group counter
    auto cur_load = [&](Cursor& c) {
      c.seg = false;
      c.slots = (c.w < total) ? slots_of(c.w, c.gi, c.seg) : 1;
    };
    auto cur_next = [&](Cursor& c) -> bool {              // advance one slot; true when a new group starts
      if (++c.j < c.slots) return false;
      c.j = 0; c.w += G; ++c.gi;
      cur_load(c);
      return true;
    };
    auto cur_init = [&](Cursor& c) {                       // slot `set` of the CTA's slot sequence
      c.w = blockIdx.x; c.j = 0; c.gi = 0;
      cur_load(c);
      if (set) cur_next(c);
    };
    auto cur_next2 = [&](Cursor& c) -> bool {             // advance to this set's next slot
      const bool a = cur_next(c);
      const bool b = cur_next(c);
      return a || b;
    };
    auto issue_entry = [&](int w, int gi) {               // cursor A
      unsigned long long* dst = s_ent + (gi & (kEntRing - 1)) * kGroup + p;
      bool ok = false;
      if (w < total) {
        int q;
        const int item = item_of(w, gi);
        const int list = list_of(item, q);
        const int e = (item - s_pre[q]) * kGroup + p;
        if (e < s_cnt[list]) { cp_async8(dst, A.lists + s_loff[list] + e); ok = true; }
      }
      if (!ok) *dst = kNoEntry;
    };
    auto entry_of = [&](int gi) { return s_ent[(gi & (kEntRing - 1)) * kGroup + p]; };

    int ga_w = blockIdx.x, ga_gi = 0;                      // cursor A
    Cursor cb;
    cur_init(cb);
    unsigned long long eb = kNoEntry;                       // cached entry of cursor B's group
    // An entry requested in the iteration after cursor B entered group Y - LEAD (or Y - LEAD + 1) is committed at
    // least DEPTH iterations before B can enter group Y: B advances at most two groups per iteration.
    auto step_a = [&](int b_gi) {
      while (ga_gi <= b_gi + LEAD) { issue_entry(ga_w, ga_gi); ga_w += G; ++ga_gi; }
    };
    // Cursor B is the only one that walks the slot sequence; it leaves its state (group, slot, is-segment) in a small
    // per-warp ring that cursors C and D read DEPTH and 2 * DEPTH iterations later (every lane writes the same word).
    static_assert(2 * DEPTH + 1 <= 8, "cursor ring too small");
    uint32_t* const s_curs = reinterpret_cast<uint32_t*>(smem + SP.curs) + (warp - kProdWarp0) * 8;
    auto step_b = [&](int it) {                            // row number of cursor B's slot -> s_idx[ring]
      s_curs[it & 7] = ((uint32_t)cb.gi << 6) | ((uint32_t)cb.j << 1) | (cb.seg ? 1u : 0u);
      if (eb != kNoEntry) {
        int r, off, len;
        unpack_entry(eb, r, off, len);
        cp_async4(s_idx + (it % STAGES) * kGroup + p, A.sorted_idx + off + min(cb.j, len - 1));
      }
      if (cur_next2(cb)) eb = entry_of(cb.gi);
    };
    auto step_c = [&](int it) {                            // row and pillar mean of cursor C's slot -> s_row / s_mean[ring]
      const uint32_t u = s_curs[it & 7];
      const unsigned long long ec = entry_of((int)(u >> 6));
      if (ec != kNoEntry) {
        const int st = it % STAGES;
        const int idx = s_idx[st * kGroup + p];
        const float* row = A.points + (int64_t)idx * A.stride;
        float* dst = s_row + st * (NREG * kGroup);
        if (kCfg == 1) {
          cp_async16(dst + p * 4, row);
          cp_async16(dst + 4 * kGroup + p * 4, row + 4);
        } else if (kCfg == 2) {
#pragma unroll
          for (int c = 0; c < 6; ++c) cp_async8(dst + c * 2 * kGroup + p * 2, row + 2 * c);
        } else {
          for (int c = 0; c < n_cols; ++c) cp_async4(dst + c * kGroup + p, row + 1 + c);
        }
        const int r = (int)(ec & 0x1fffffffull);
        cp_async16(s_mean + st * kGroup + p, ((u & 1u) ? A.long_mean : A.mean) + r);
      }
    };
    int ib = 0, ic = 0;                                    // iteration numbers of cursors B and C
    // entries of the first 4 * DEPTH + LEAD + 2 groups (everything cursor B can reach during the prologue plus its lead)
    while (ga_gi < 4 * DEPTH + LEAD + 2) { issue_entry(ga_w, ga_gi); ga_w += G; ++ga_gi; }
    cp_async_commit();
    cp_async_wait<0>();
    eb = entry_of(cb.gi);
    for (int t = 0; t < DEPTH; ++t) step_b(ib++);          // row numbers of its slots 0 .. DEPTH - 1 - one round trip
    cp_async_commit();
    cp_async_wait<0>();
    // rows of its slots 0 .. DEPTH - 1 and row numbers of slots DEPTH .. 2 * DEPTH - 1, one commit group per slot:
    // DEPTH commit groups pending exactly as in the steady state (at most DEPTH + 1 row numbers are live = their ring)
    for (int t = 0; t < DEPTH; ++t) { step_b(ib++); step_c(ic++); cp_async_commit(); }

    uint32_t c0 = 0;                                       // this set's slots built so far
    int id = 0;                                            // iteration number of cursor D
    const int n_feat = n_raw + (with_dist ? 7 : 6);
    const bool fold0 = pfn_fold0(n_feat, kLayers);         // column n_feat carries 1.0 for the folded layer-0 bias
    TRACE_DECL(p == 0)
    while (true) {
      const uint32_t ud = s_curs[id & 7];                  // written by cursor B 2 * DEPTH iterations (or the prologue) ago
      const int d_gi = (int)(ud >> 6), d_j = (int)((ud >> 1) & 31u);
      const bool d_seg = (ud & 1u) != 0;
      if ((long long)blockIdx.x + (long long)d_gi * G >= total) break;
      // ---- pipeline upkeep: entries / one row number / one row request per iteration ----
      TRACE(set ? 4 : 1, 10);
      cp_async_wait<DEPTH - 1>();                          // everything issued DEPTH or more iterations ago has landed
      step_a(cb.gi);
      step_b(ib++);
      step_c(ic++);
      cp_async_commit();
      TRACE(set ? 4 : 1, 14);
      const unsigned long long ed = entry_of(d_gi);
      const bool valid = ed != kNoEntry;
      const float4 m = s_mean[(id % STAGES) * kGroup + p];
      if (d_j == 0) {
        // first slot of a group (built by exactly one of the two sets): publish the group's output rows
        const int r = valid ? (int)(ed & 0x1fffffffull) : -1;
        s_rows[(d_gi % kRowRing) * kGroup + p] = r;      // pillar rank, or long-pillar index of a segment
        if (A.mean_out && valid && !d_seg) {
          float* mo = A.mean_out + (int64_t)r * 3;
          mo[0] = m.x; mo[1] = m.y; mo[2] = m.z;
        }
      }
      // ---- this slot's row: shared memory -> registers ----
      float rw[NREG];
      const float* src = s_row + (id % STAGES) * (NREG * kGroup);
      if (kCfg == 1) {
        const float4 v0 = ld4(src + p * 4), v1 = ld4(src + 4 * kGroup + p * 4);
        rw[0] = v0.x; rw[1] = v0.y; rw[2] = v0.z; rw[3] = v0.w; rw[4] = v1.x; rw[5] = v1.y; rw[6] = v1.z; rw[7] = v1.w;
      } else if (kCfg == 2) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const float2 v = *reinterpret_cast<const float2*>(src + c * 2 * kGroup + p * 2);
          rw[2 * c] = v.x; rw[2 * c + 1] = v.y;
        }
      }
      const float* gsrc = src + p;                         // generic layout: column c at gsrc[(c - 1) * kGroup]
      // ---- features of this slot's row (dynamic_pillar_vfe.py:111-126) ----
      float x = 0.f, y = 0.f, z = 0.f;
      if (valid) {
        if (kCfg) { x = rw[1]; y = rw[2]; z = rw[3]; }
        else { x = gsrc[0]; y = gsrc[kGroup]; z = gsrc[2 * kGroup]; }
      }
      float ed_[7];
      ed_[0] = __fsub_rn(x, m.x);                                                // f_cluster (:111)
      ed_[1] = __fsub_rn(y, m.y);
      ed_[2] = __fsub_rn(z, m.z);
      // the pillar's cell (same for all its points) was packed next to the mean: no per-point division (:98, :114-116)
      const unsigned cell = __float_as_uint(m.w);
      const float cx = (float)(cell & 0xffffu), cy = (float)(cell >> 16);
      ed_[3] = __fsub_rn(x, __fadd_rn(__fmul_rn(cx, A.g.voxel_x), A.g.x_offset));   // f_center (:114-116)
      ed_[4] = __fsub_rn(y, __fadd_rn(__fmul_rn(cy, A.g.voxel_y), A.g.y_offset));
      ed_[5] = __fsub_rn(z, A.g.z_offset);
      ed_[6] = with_dist ? __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))) : 0.f;  // :124
      // slot t of the CTA lives in operand buffer t % NF, which is free once the layer-0 MMA of slot t - NF has completed
      const uint32_t t_slot = 2 * c0 + (uint32_t)set;
      const uint32_t b = t_slot % NF;
      TRACE(set ? 4 : 1, 11);
      if (t_slot >= NF) { mbar_wait(&bars[kBarD0 + b], (t_slot / NF - 1) & 1); tc_fence_after_sync(); }
      TRACE(set ? 4 : 1, 12);
      const uint32_t dh = tmem + kColA0 + b * kFS + lane_base, dl = dh + kA0Lo;
#pragma unroll
      for (int cc8 = 0; cc8 < kMaxCin; cc8 += 8) {
        if (cc8 < k0) {
          float hi[8], lo[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int f = cc8 + t;
            float val = (fold0 && f == n_feat) ? 1.f : 0.f;
            if (valid) {
              if (f < n_raw) {
                if (kCfg) val = rw[(1 + f) < NREG ? (1 + f) : (NREG - 1)];
                else val = gsrc[(raw_col0 - 1 + f) * kGroup];
              } else if (f < n_feat) {
                const int d = f - n_raw;
                val = d == 0 ? ed_[0] : d == 1 ? ed_[1] : d == 2 ? ed_[2] : d == 3 ? ed_[3] : d == 4 ? ed_[4] : d == 5 ? ed_[5] : ed_[6];
              }
            }
            split_tf32(val, hi[t], lo[t]);
          }
          tmem_st8(dh + cc8, hi);
          tmem_st8(dl + cc8, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(&bars[kBarA0 + b]);
      ++c0; ++id;
    }
    cp_async_wait<0>();
    TRACE_END(set ? 4 : 1);
  } else if (warp >= kE0Warp0 && warp < kE0Warp0 + kE0Threads / 32) {
    // =====================================================================================================
    // E0: layer-0 epilogue (for a single-layer PFN: the whole epilogue)
    // =====================================================================================================
    const int p = tid & (kGroup - 1);
    const int h = (tid - kE0Warp0 * 32) >> 7;                // column half
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t c0 = 0, c1 = 0;   // layer-0 accumulators consumed / layer-1-type operands staged
    TRACE_DECL(tid == kE0Warp0 * 32)
    int gi = 0;
    if (kLayers == 2) {
      float max0[16];          // layer-0 running max, this thread's 16 channels
      const float* pa = smem + SP.prm_a0 + 16 * h;
      const float* pb = smem + SP.prm_b0 + 16 * h;
      // Software pipeline: the accumulators of the NEXT slot are requested (tcgen05.ld) as soon as the registers of the
      // current one are consumed - half by half - so that the load latency overlaps the math, the tcgen05.st drain and
      // the hand-over of the current slot.
      uint32_t nslots = 0;
      {
        int g2 = 0;
        for (int w = blockIdx.x; w < total; w += G, ++g2) { bool sg; nslots += slots_of(w, g2, sg); }
      }
      uint32_t rr[16];
      auto d0_addr = [&](uint32_t t) { return tmem + kColD0 + (t % NF) * kFS + lane_base + 16 * h; };
      if (nslots > 0) {
        mbar_wait(&bars[kBarD0 + 0], 0);
        tc_fence_after_sync();
        tmem_ld16_nowait(d0_addr(0), rr);
      }
      for (int w = blockIdx.x; w < total; w += G, ++gi) {
        bool is_seg;
        const int slots = slots_of(w, gi, is_seg);
#pragma unroll
        for (int i = 0; i < 16; ++i) max0[i] = 0.f;
        for (int j = 0; j < slots; ++j) {
          const uint32_t b1 = c1 & 1;
          const bool more = (c0 + 1) < nslots;
          TRACE(2, 30);
          tmem_ld_wait();                                                       // rr = accumulators of slot c0
          TRACE(2, 301);
          if (c1 >= 2) mbar_wait(&bars[kBarD1 + b1], ((c1 - 2) >> 1) & 1);     // A1[b1] read by the MMA two ops ago
          if (more) mbar_wait(&bars[kBarD0 + (c0 + 1) % NF], ((c0 + 1) / NF) & 1);
          TRACE(2, 31);
          tc_fence_after_sync();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float hi[8], lo[8];
            float a8[8], b8[8];
            if (!kFold0) {
              const float4 al0 = ld4(pa + 8 * half), al1 = ld4(pa + 8 * half + 4);
              const float4 be0 = ld4(pb + 8 * half), be1 = ld4(pb + 8 * half + 4);
              a8[0] = al0.x; a8[1] = al0.y; a8[2] = al0.z; a8[3] = al0.w; a8[4] = al1.x; a8[5] = al1.y; a8[6] = al1.z; a8[7] = al1.w;
              b8[0] = be0.x; b8[1] = be0.y; b8[2] = be0.z; b8[3] = be0.w; b8[4] = be1.x; b8[5] = be1.y; b8[6] = be1.z; b8[7] = be1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // kFold0: BN already sits in the operand (scaled rows + bias column), the epilogue is a bare ReLU
              const float xv = kFold0 ? fmaxf(__uint_as_float(rr[8 * half + i]), 0.f)
                                      : fmaxf(fmaf(__uint_as_float(rr[8 * half + i]), a8[i], b8[i]), 0.f);
              max0[8 * half + i] = fmaxf(max0[8 * half + i], xv);
              split_tf32(xv, hi[i], lo[i]);
            }
            tmem_st8(tmem + kColA1 + b1 * 64 + lane_base + 16 * h + 8 * half, hi);
            tmem_st8(tmem + kColA1 + b1 * 64 + 32 + lane_base + 16 * h + 8 * half, lo);
            if (more) {                                                         // these 8 registers are free again
              if (half == 0) tmem_ld8_nowait(d0_addr(c0 + 1), rr[0], rr[1], rr[2], rr[3], rr[4], rr[5], rr[6], rr[7]);
              else tmem_ld8_nowait(d0_addr(c0 + 1) + 8, rr[8], rr[9], rr[10], rr[11], rr[12], rr[13], rr[14], rr[15]);
            }
            TRACE(2, 302 + half);
          }
          tmem_st_wait();
          TRACE(2, 304);
          tc_fence_before_sync();
          mbar_arrive(&bars[kBarA1 + b1]);
          TRACE(2, 32);
          ++c0; ++c1;
        }
        if (slots == 1 && !is_seg) {
          // one-point pillars: no hoist (the MMA warp used the summed weights)
        } else if (!is_seg) {
          // hoist: max0 -> A1; the MMA warp multiplies by W1[:, 32:]^T
          const uint32_t b1 = c1 & 1;
          if (c1 >= 2) { mbar_wait(&bars[kBarD1 + b1], ((c1 - 2) >> 1) & 1); tc_fence_after_sync(); }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split_tf32(max0[8 * half + i], hi[i], lo[i]);
            tmem_st8(tmem + kColA1 + b1 * 64 + lane_base + 16 * h + 8 * half, hi);
            tmem_st8(tmem + kColA1 + b1 * 64 + 32 + lane_base + 16 * h + 8 * half, lo);
          }
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&bars[kBarA1 + b1]);
          ++c1;
        } else {
          // long pillar segment: partial layer-0 maxima -> the pillar's accumulator
          const int li = s_rows[(gi % kRowRing) * kGroup + p];
          if (li >= 0) {
            unsigned* acc = A.long_acc + (int64_t)li * 96 + 16 * h;
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicMax(acc + i, ord_enc(max0[i]));
          }
        }
      }
    } else {
      // single layer: D0 (64 columns) is the last-layer accumulator; this thread owns 32 of them
      float m1[32];
      for (int w = blockIdx.x; w < total; w += G, ++gi) {
        bool is_seg;
        const int slots = slots_of(w, gi, is_seg);
#pragma unroll
        for (int i = 0; i < 32; ++i) m1[i] = -INFINITY;
        for (int j = 0; j < slots; ++j) {
          const uint32_t b = c0 % NF;
          mbar_wait(&bars[kBarD0 + b], (c0 / NF) & 1);
          tc_fence_after_sync();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t rr[16];
            tmem_ld16_nowait(tmem + kColD0 + b * kFS + lane_base + 32 * h + 16 * half, rr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) m1[16 * half + i] = fmaxf(m1[16 * half + i], __uint_as_float(rr[i]));
          }
          tc_fence_before_sync();
          mbar_arrive(&bars[kBarA1 + b]);             // "D0[b] consumed"
          ++c0;
        }
        const int r = s_rows[(gi % kRowRing) * kGroup + p];
        if (r >= 0) {
          if (is_seg) {
            unsigned* acc = A.long_acc + (int64_t)r * 96 + 32 + 32 * h;
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicMax(acc + i, ord_enc(m1[i]));
          } else {
            const float* pa = smem + SP.prm_a0 + 32 * h;
            const float* pb = smem + SP.prm_b0 + 32 * h;
            float* dst = A.out + (int64_t)r * kCout + 32 * h;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 al = ld4(pa + i), be = ld4(pb + i);
              st_global_f4(dst + i, fmaxf(fmaf(m1[i + 0], al.x, be.x), 0.f), fmaxf(fmaf(m1[i + 1], al.y, be.y), 0.f),
                           fmaxf(fmaf(m1[i + 2], al.z, be.z), 0.f), fmaxf(fmaf(m1[i + 3], al.w, be.w), 0.f));
            }
          }
        }
      }
    }
    TRACE_END(2);
  } else if (kLayers == 2) {
    // =====================================================================================================
    // E1: last-layer epilogue + output rows (two-layer PFN)
    // =====================================================================================================
    const int p = tid & (kGroup - 1);
    const int et = tid - kE1Warp0 * 32;                      // thread of the role
    const int h = (et >> 7) & 1;                             // column half: channels 32 h .. 32 h + 31
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t k = 0;            // layer-1-type accumulators consumed
    float m1[32];              // running max of the raw last-layer accumulators
    TRACE_DECL(tid == kE1Warp0 * 32)
    int gi = 0;
    // output staging: this thread's 128 bytes (padded rows); the warp reads the tile back transposed so that its stores are whole lines
    float* const my_stage = smem + SP.ostage + et * (kOutRowBytes / 4);
    for (int w = blockIdx.x; w < total; w += G, ++gi) {
      bool is_seg;
      const int slots = slots_of(w, gi, is_seg);
      // one-point pillars: the single accumulator (summed weights) is already max + hoist; it goes through the output path
      const bool single = (slots == 1) && !is_seg;
#pragma unroll
      for (int i = 0; i < 32; ++i) m1[i] = single ? 0.f : -INFINITY;
      for (int j = 0; j < (single ? 0 : slots); ++j) {
        const uint32_t b = k & 1;
        TRACE(3, 33);
        mbar_wait(&bars[kBarD1 + b], (k >> 1) & 1);
        TRACE(3, 34);
        tc_fence_after_sync();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t rr[16];
          tmem_ld16_nowait(tmem + kColD1 + b * 64 + lane_base + 32 * h + 16 * half, rr);
          tmem_ld_wait();
          if (half == 1) { tc_fence_before_sync(); mbar_arrive(&bars[kBarF1 + b]); }
#pragma unroll
          for (int i = 0; i < 16; ++i) m1[16 * half + i] = fmaxf(m1[16 * half + i], __uint_as_float(rr[i]));
        }
        TRACE(3, 35);
        ++k;
      }
      const int r = s_rows[(gi % kRowRing) * kGroup + p];
      if (!is_seg) {
        // the x_max half (hoist) + BN + ReLU + this thread's 128 bytes of the output row
        const uint32_t b = k & 1;
        TRACE(3, 36);
        mbar_wait(&bars[kBarD1 + b], (k >> 1) & 1);
        TRACE(3, 37);
        tc_fence_after_sync();
        const float* pa = smem + SP.prm_a1 + 32 * h;
        const float* pb = smem + SP.prm_b1 + 32 * h;
        __syncwarp();                                        // the previous group's rows have left the staging tile
        TRACE(3, 39);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t rr[16];
          tmem_ld16_nowait(tmem + kColD1 + b * 64 + lane_base + 32 * h + 16 * half, rr);
          tmem_ld_wait();
          if (half == 1) { tc_fence_before_sync(); mbar_arrive(&bars[kBarF1 + b]); }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const int c = 16 * half + i;
            const float4 al = ld4(pa + c), be = ld4(pb + c);
            float4 o;
            o.x = fmaxf(fmaf(__fadd_rn(m1[c + 0], __uint_as_float(rr[i + 0])), al.x, be.x), 0.f);
            o.y = fmaxf(fmaf(__fadd_rn(m1[c + 1], __uint_as_float(rr[i + 1])), al.y, be.y), 0.f);
            o.z = fmaxf(fmaf(__fadd_rn(m1[c + 2], __uint_as_float(rr[i + 2])), al.z, be.z), 0.f);
            o.w = fmaxf(fmaf(__fadd_rn(m1[c + 3], __uint_as_float(rr[i + 3])), al.w, be.w), 0.f);
            *reinterpret_cast<float4*>(my_stage + c) = o;
          }
        }
        __syncwarp();
        TRACE(3, 40);
        // transposed read-back: every store instruction of the warp writes 4 whole 128-byte lines
        {
          const int lane = tid & 31;
          const float* wstage = smem + SP.ostage + (et & ~31) * (kOutRowBytes / 4);
          const int* rows = s_rows + (gi % kRowRing) * kGroup + (p & ~31);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + (lane >> 3), chunk = lane & 7;
            const int rr_ = rows[row];
            const float4 v = ld4(wstage + row * (kOutRowBytes / 4) + chunk * 4);
            if (rr_ >= 0) *reinterpret_cast<float4*>(A.out + (int64_t)rr_ * kCout + 32 * h + chunk * 4) = v;
          }
        }
        TRACE(3, 38);
        ++k;
      } else if (r >= 0) {
        // long pillar segment: partial maxima -> the pillar's accumulator (r is its long-pillar index)
        unsigned* acc = A.long_acc + (int64_t)r * 96 + 32 + 32 * h;
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicMax(acc + i, ord_enc(m1[i]));
      }
    }
    TRACE_END(3);
  }
  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
  STAGE_END(g_stage_pfn, 0);
}

// ------------------------------------------------------------------------------------------------
// self test: C[128 x N] = A[128 x K] . B[N x K]^T through the exact operand layouts / descriptors / TMEM paths
// mode 0: A and B from shared memory (layer 0);  mode 1: A written to tensor memory with tcgen05.st (layer 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads)
umma_selftest_kernel(const float* __restrict__ Ag, const float* __restrict__ Bg, int K, int N, int mode,
                     float* __restrict__ Cg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ah = reinterpret_cast<float*>(smem_raw);
  float* al = ah + 64 * kGroup;
  float* bh = al + 64 * kGroup;
  float* bl = bh + 64 * 64;
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kGroup * K; i += kTcThreads) {
    const int r = i / K, k = i % K;
    float h, l;
    split_tf32(Ag[i], h, l);
    ah[(k >> 2) * (kGroup * 4) + r * 4 + (k & 3)] = h;
    al[(k >> 2) * (kGroup * 4) + r * 4 + (k & 3)] = l;
  }
  for (int i = tid; i < N * K; i += kTcThreads) {
    const int n = i / K, k = i % K;
    float h, l;
    split_tf32_rn(Bg[i], h, l);
    bh[(k >> 2) * (N * 4) + n * 4 + (k & 3)] = h;
    bl[(k >> 2) * (N * 4) + n * 4 + (k & 3)] = l;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  if (mode == 1) {
    // A (K <= 32) -> tensor memory: hi at columns [64, 64 + K), lo at [96, 96 + K); each half writes 16 columns
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = half * 16 + i;
      const float v = (k < K) ? Ag[row * K + k] : 0.f;
      split_tf32(v, hi[i], lo[i]);
    }
    tmem_st16(tmem + 64 + lane_base + 16 * half, hi);
    tmem_st16(tmem + 96 + lane_base + 16 * half, lo);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
  }
  if (warp == 0) {
    if (elect_one_sync()) {
      tc_fence_after_sync();
      if (mode == 0)
        mma_3xtf32(tmem, smem_u32(ah), smem_u32(al), kGroup, smem_u32(bh), smem_u32(bl), (uint32_t)N, K / 8,
                   idesc_tf32_m128((uint32_t)N), false);
      else
        mma_3xtf32_ts(tmem, tmem + 64, tmem + 96, smem_u32(bh), smem_u32(bl), (uint32_t)N, K / 8,
                      idesc_tf32_m128((uint32_t)N), false);
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  for (int c0 = half * 16; c0 < N; c0 += 32) {
    float v[16];
    tmem_ld16(tmem + lane_base + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) Cg[row * N + c0 + i] = v[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------
// micro-benchmark (diagnostic): cycles of `reps` back-to-back 3xTF32 groups of `ksteps` K steps (3 MMAs each),
// issue -> commit -> mbarrier wait, measured with clock64 by the issuing thread.  Operands are whatever is in
// shared / tensor memory (timing only).  out[0] = cycles, out[1] = cycles of an empty commit + wait.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
umma_cycles_kernel(int mode, int n, int ksteps, int reps, long long* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ah = reinterpret_cast<float*>(smem_raw);
  float* al = ah + 64 * kGroup;
  float* bh = al + 64 * kGroup;
  float* bl = bh + 64 * 64;
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 64 * kGroup + 2 * 64 * 64; i += 128) ah[i] = 0.f;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (warp == 0 && elect_one_sync()) {
    uint32_t phase = 0;
    const uint32_t idesc = idesc_tf32_m128((uint32_t)n);
    // warm-up
    mma_3xtf32(tmem, smem_u32(ah), smem_u32(al), kGroup, smem_u32(bh), smem_u32(bl), (uint32_t)n, 1, idesc, false);
    mma_commit(&bar);
    mbar_wait(&bar, phase); phase ^= 1;
    long long t0 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, phase); phase ^= 1;
    long long t1 = clock64();
    out[1] = t1 - t0;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode == 0)
        mma_3xtf32(tmem, smem_u32(ah), smem_u32(al), kGroup, smem_u32(bh), smem_u32(bl), (uint32_t)n, ksteps, idesc, false);
      else
        mma_3xtf32_ts(tmem, tmem + 64, tmem + 96, smem_u32(bh), smem_u32(bl), (uint32_t)n, ksteps, idesc, false);
    }
    long long t_issue = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, phase); phase ^= 1;
    t1 = clock64();
    out[0] = t1 - t0;
    out[2] = t_issue - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace pcp

using namespace pcp;

static int selftest(const float* a, const float* b, int32_t k, int32_t n, int mode, float* c, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(a && b && c, PCP_E_INVALID, "pcp_selftest_umma: null argument");
  PCP_REQUIRE(k > 0 && k <= (mode ? 32 : 64) && k % 8 == 0 && (n == 32 || n == 64), PCP_E_INVALID,
              "pcp_selftest_umma: bad k/n");
  const size_t smem = sizeof(float) * (2 * 64 * kGroup + 2 * 64 * 64);
  PCP_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, kTcThreads, smem, stream>>>(a, b, k, n, mode, c);
  PCP_LAUNCH_CHECK("umma_selftest_kernel");
  return 0;
}

extern "C" int pcp_selftest_umma(const float* a, const float* b, int32_t k, int32_t n, float* c, void* stream_) {
  return selftest(a, b, k, n, 0, c, stream_);
}
extern "C" int pcp_selftest_umma_ts(const float* a, const float* b, int32_t k, int32_t n, float* c, void* stream_) {
  return selftest(a, b, k, n, 1, c, stream_);
}

extern "C" int pcp_selftest_umma_cycles(int32_t mode, int32_t n, int32_t ksteps, int32_t reps, long long* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PCP_REQUIRE(out && (n == 32 || n == 64) && ksteps >= 1 && ksteps <= (mode ? 4 : 8) && reps >= 0, PCP_E_INVALID,
              "pcp_selftest_umma_cycles: bad argument");
  const size_t smem = sizeof(float) * (2 * 64 * kGroup + 2 * 64 * 64);
  PCP_CUDA(cudaFuncSetAttribute(umma_cycles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_cycles_kernel<<<1, 128, smem, stream>>>(mode, n, ksteps, reps, out);
  PCP_LAUNCH_CHECK("umma_cycles_kernel");
  return 0;
}

STAGE_EXPORT(pcp_debug_stage_pfn, pcp::g_stage_pfn)

#ifdef PCP_PFN_TIMING
// debug: copies the event trace of CTA 0 (3 roles x kTraceCap x (id, clock)) and the 3 event counts to host memory
extern "C" int pcp_debug_read_timing(long long* trace_host, int* counts_host) {
  PCP_CUDA(cudaDeviceSynchronize());
  PCP_CUDA(cudaMemcpyFromSymbol(trace_host, g_trace, sizeof(long long) * kTraceRoles * kTraceCap * 2));
  PCP_CUDA(cudaMemcpyFromSymbol(counts_host, g_trace_n, sizeof(int) * kTraceRoles));
  return 0;
}
#endif

// launched from pfn.cu
namespace pcp {

template <int kLayers, int kCfg>
static int launch_cfg(const TcArgs& a, int64_t n_points, cudaStream_t stream) {
  const SmemPlan SP = smem_plan(kCfg ? RowCfg<kCfg>::k0 : a.k0, kLayers, RowCfg<kCfg>::nreg, RowCfg<kCfg>::depth);
  PCP_CUDA(cudaFuncSetAttribute(pfn_slot_kernel<kLayers, kCfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP.total_bytes));
  // upper bound of the group count: every list may end in a partial group
  const int64_t groups = n_points / kGroup + kNumLists;
  const int64_t sms = sm_count();            // persistent: one CTA per SM
  const unsigned blocks = (unsigned)(groups < sms ? groups : sms);
  PCP_CUDA(launch_pdl(kPdlPfn, pfn_slot_kernel<kLayers, kCfg>, dim3(blocks), dim3(kPfnThreads), (size_t)SP.total_bytes, stream, a));
  return 0;
}

int launch_pfn_tc(const TcArgs& a, int64_t n_points, cudaStream_t stream) {
  const bool abs_nodist = (a.raw_col0 == 1) && !a.with_distance;
  const uintptr_t base = reinterpret_cast<uintptr_t>(a.points);
  if (a.num_layers == 2) {
    if (abs_nodist && a.n_raw == 5 && a.stride % 4 == 0 && a.stride >= 8 && (base & 15) == 0)
      return launch_cfg<2, 1>(a, n_points, stream);
    if (abs_nodist && a.n_raw == 11 && a.stride % 2 == 0 && a.stride >= 12 && (base & 7) == 0)
      return launch_cfg<2, 2>(a, n_points, stream);
    return launch_cfg<2, 0>(a, n_points, stream);
  }
  return launch_cfg<1, 0>(a, n_points, stream);
}

}  // namespace pcp
