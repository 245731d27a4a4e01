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
//   A0   gather the slot's rows, features (f_cluster / f_center need the pillar mean, computed per group in
//        ascending row order), TF32 hi/lo split -> A0 panels in shared memory
//   M0   D0[128 x 32] = A0 . W0^T                      (tcgen05.mma kind::tf32, SS, 3 MMAs per K step)
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

#include "common.cuh"
#include "umma.cuh"
#include "pfn_tc.cuh"

namespace pcp {

using namespace umma;

constexpr int kTcThreads = 256;     // threads of the self-test kernels
// ---- roles of the PFN kernel (one persistent CTA per SM) ----
constexpr int kEpiThreads = 512;    // epilogue threads: (pillar p = tid & 127, column quarter q = tid >> 7)
constexpr int kProdThreads = 128;   // producers: one thread per pillar (TMEM lane) of the group
constexpr int kProdWarp0 = kEpiThreads / 32;                   // 16
constexpr int kMmaWarp = kProdWarp0 + kProdThreads / 32;      // 20
constexpr int kPfnThreads = (kMmaWarp + 1) * 32;              // 672
constexpr int kTmemCols = 512;
// tensor-memory column map (everything double buffered: operand / accumulator b of op c is c & 1)
constexpr uint32_t kColD0 = 0;      // layer-0 accumulators   [b * 64, +32)  (+64 for a single-layer PFN)
constexpr uint32_t kColD1 = 128;    // layer-1 / hoist accumulators [128 + b * 64, +64)
constexpr uint32_t kColA0 = 256;    // layer-0 A operand (features): hi at 256 + b * 64, lo 32 columns further (k0 <= 24)
constexpr uint32_t kColA1 = 384;    // layer-1 A operand (x0 / max0): hi at 384 + b * 64, lo 32 columns further
constexpr int kOutLd = 68;          // padded row of the output staging tile (conflict-free 16-byte accesses)
constexpr int kRowRing = 8;         // ring of per-group output-row tables (producer runs a few groups ahead of the output)
constexpr int kEntRing = 32;        // ring of per-group work-list entries (the entry cursor runs up to 3 * depth groups ahead)
// mbarrier indices
constexpr int kBarA0 = 0;           // [2] features staged in TMEM        (128 producer arrivals)
constexpr int kBarA1 = 2;           // [2] x0 / max0 staged in TMEM       (512 epilogue arrivals)
constexpr int kBarD0 = 4;           // [2] layer-0 accumulator ready      (tcgen05.commit)
constexpr int kBarD1 = 6;           // [2] layer-1 / hoist accumulator ready

// ---- optional per-role event trace of CTA 0 (debug build only: make dbg; tools/pfn_timing.py) ----
#ifdef PCP_PFN_TIMING
constexpr int kTraceCap = 8192;
__device__ long long g_trace[3][kTraceCap][2];
__device__ int g_trace_n[3];
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

// kCfg: 0 = any layout (scalar loads, run-time feature map)
//       1 = car / early-fusion rows: c_raw 5, absolute xyz, no distance, row stride % 4 == 0, 16-byte aligned
//       2 = ego (lately fusion) rows: c_raw 11, absolute xyz, no distance, even row stride, 8-byte aligned
//   nreg  = floats of a row staged per slot, depth = rows in flight per producer thread (cp.async ring of depth + 1 stages)
template <int kCfg> struct RowCfg { static constexpr int n_raw = 0, k0 = 0, nreg = 24, depth = 3; };
template <> struct RowCfg<1> { static constexpr int n_raw = 5, k0 = 16, nreg = 8, depth = 7; };
template <> struct RowCfg<2> { static constexpr int n_raw = 11, k0 = 24, nreg = 12, depth = 7; };

struct SmemPlan {   // float offsets into dynamic shared memory
  int w0h, w0l, w1ah, w1al, w1bh, w1bl, panels_end, prm_a0, prm_b0, prm_a1, prm_b1, ent, idx, row, out, rows, ints, total_bytes;
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
  }
  S.panels_end = o;
  S.prm_a0 = o; o += n0;
  S.prm_b0 = o; o += n0;
  if (layers == 2) {
    S.prm_a1 = o; o += kCout;
    S.prm_b1 = o; o += kCout;
  }
  S.ent = o; o += kEntRing * kGroup * 2;          // work-list entries, [group ring][pillar] (8 bytes each)
  S.idx = o; o += (depth + 1) * kGroup;           // row numbers, [slot ring][pillar]
  S.row = o; o += (depth + 1) * nreg * kGroup;    // staged rows, [slot ring][16/8/4-byte chunk][pillar]
  S.out = o; o += kGroup * kOutLd;                // output staging tile (coalesced pillar_features rows)
  S.rows = o; o += kRowRing * kGroup;             // output row (pillar rank) / long-pillar index of each lane
  S.ints = o; o += 80;                            // 8 mbarriers | tmem base | group prefix | list counts | list offsets
  S.total_bytes = o * 4;
  return S;
}

// Persistent, warp-specialised kernel.  Three roles talk only through mbarriers:
//   producers (4 warps, thread = pillar lane): prefetch the work-list entries and row numbers (cp.async), compute the
//       pillar mean, gather each slot's row, build the feature vector and write its TF32 hi / lo parts to TMEM (A0);
//   MMA warp: waits for "operand staged", issues the tcgen05.mma groups (layer 0 of the NEXT slot is queued in front
//       of layer 1 of the current one, so the tensor pipe has work while the epilogue runs), commits to "accumulator ready";
//   epilogue (16 warps, thread = pillar lane x column quarter): layer-0 epilogue (BN + ReLU, running max0, x0 back to
//       TMEM as layer 1's A operand), layer-1 epilogue (running max), the per-pillar hoist and the output rows.
// The issuing thread of a tcgen05.mma is back-pressured by the tensor pipe (measured 30-60 cycles per MMA), which is
// why it owns a warp; shared-memory traffic is limited to the weight panels the MMAs read and the output staging tile.
template <int kLayers, int kCfg>
__global__ void __launch_bounds__(kPfnThreads, 1)
pfn_slot_kernel(const TcArgs A) {
  extern __shared__ __align__(128) float smem[];
  constexpr int N0 = (kLayers == 2) ? kHidden : kCout;
  constexpr int NREG = RowCfg<kCfg>::nreg;
  const int k0 = kCfg ? RowCfg<kCfg>::k0 : A.k0;
  const int n_raw = kCfg ? RowCfg<kCfg>::n_raw : A.n_raw;
  const int raw_col0 = kCfg ? 1 : A.raw_col0;
  const bool with_dist = kCfg ? false : (A.with_distance != 0);
  const SmemPlan SP = smem_plan(k0, kLayers, NREG, RowCfg<kCfg>::depth);
  const int tid = threadIdx.x, warp = tid >> 5;
  int* const s_idx = reinterpret_cast<int*>(smem + SP.idx);
  float* const s_out = smem + SP.out;
  int* const s_rows = reinterpret_cast<int*>(smem + SP.rows);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + SP.ints);
  uint32_t* const s_tmem = reinterpret_cast<uint32_t*>(smem + SP.ints + 16);
  int* const s_pre = reinterpret_cast<int*>(smem + SP.ints + 20);      // [kNumLists + 1] group prefix, processing order
  int* const s_cnt = reinterpret_cast<int*>(smem + SP.ints + 36);      // [kNumLists] entries per list
  long long* const s_loff = reinterpret_cast<long long*>(smem + SP.ints + 48);   // [kNumLists] list offsets

  // ---- one-time setup: parameters -> smem, barriers, TMEM, work prefix ----
  {
    const float4* src = reinterpret_cast<const float4*>(A.params);
    float4* dst = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < SP.panels_end / 4; i += kPfnThreads) dst[i] = __ldg(src + i);   // same order in both layouts
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
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars[kBarA0 + b], kProdThreads);
        mbar_init(&bars[kBarA1 + b], kEpiThreads);
        mbar_init(&bars[kBarD0 + b], 1);
        mbar_init(&bars[kBarD1 + b], 1);
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
  auto slots_of = [&](int w, bool& is_seg) {
    int q;
    const int list = list_of(w, q);
    is_seg = (list == kSegList);
    return is_seg ? kSegRows : class_slots(list);
  };

  if (warp == kMmaWarp) {
    // =====================================================================================================
    // MMA warp
    // =====================================================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    const uint32_t sw0h = smem_u32(smem + SP.w0h), sw0l = smem_u32(smem + SP.w0l);
    const uint32_t sw1ah = smem_u32(smem + SP.w1ah), sw1al = smem_u32(smem + SP.w1al);
    const uint32_t sw1bh = smem_u32(smem + SP.w1bh), sw1bl = smem_u32(smem + SP.w1bl);
    const uint32_t idesc0 = idesc_tf32_m128(N0), idesc1 = idesc_tf32_m128(kCout);
    uint32_t c0 = 0, c1 = 0;      // layer-0 ops / layer-1-type ops issued so far
    TRACE_DECL((tid & 31) == 0)
    auto issue_m0 = [&]() {
      const uint32_t b = c0 & 1;
      TRACE(0, 20);
      mbar_wait(&bars[kBarA0 + b], (c0 >> 1) & 1);
      TRACE(0, 21);
      if (kLayers == 1 && c0 >= 2) mbar_wait(&bars[kBarA1 + b], ((c0 - 2) >> 1) & 1);   // D0[b] consumed by the epilogue
      tc_fence_after_sync();
      if (elect_one_sync()) {
        mma_3xtf32_ts(tmem + kColD0 + b * 64, tmem + kColA0 + b * 64, tmem + kColA0 + b * 64 + 32, sw0h, sw0l, N0, k0 / 8,
                      idesc0, false);
        mma_commit(&bars[kBarD0 + b]);
      }
      __syncwarp();
      ++c0;
    };
    auto issue_m1 = [&](uint32_t wh, uint32_t wl) {
      const uint32_t b = c1 & 1;
      TRACE(0, 22);
      mbar_wait(&bars[kBarA1 + b], (c1 >> 1) & 1);
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
    if ((int)blockIdx.x < total) issue_m0();
    for (int w = blockIdx.x; w < total; w += G) {
      bool is_seg;
      const int slots = slots_of(w, is_seg);
      const bool more_groups = (w + G) < total;
      for (int j = 0; j < slots; ++j) {
        if (j + 1 < slots || more_groups) issue_m0();          // the NEXT slot's layer 0 goes in front of this slot's layer 1
        if (kLayers == 2) issue_m1(sw1ah, sw1al);
      }
      if (kLayers == 2 && !is_seg) issue_m1(sw1bh, sw1bl);     // hoist: max0 . W1[:, 32:]^T once per pillar
      TRACE(0, 24);
    }
    TRACE_END(0);
  } else if (warp >= kProdWarp0) {
    // =====================================================================================================
    // producers (register budget raised with what the other roles gave back to the CTA pool)
    //
    // A producer thread owns TMEM lane p, i.e. pillar p of every group of this CTA, and walks the CTA's slots in
    // order.  Everything it needs arrives through a software pipeline of cp.async copies that it issues itself and
    // that only it reads back (no cross-thread synchronisation), DEPTH slots apart per stage:
    //   cursor A  work-list entry of a group             -> s_ent   (DEPTH + 1 groups ahead of cursor B)
    //   cursor B  row number of slot t + 2 * DEPTH       -> s_idx   (needs its group's entry)
    //   cursor C  point row of slot t + DEPTH            -> s_row   (needs its row number)
    //   cursor D  slot t: features -> TF32 hi / lo -> tensor memory (A operand of layer 0)
    // One commit group per slot and a single cp.async.wait_group<DEPTH - 1> make everything issued DEPTH or more
    // iterations ago visible, which is exactly what cursors B, C and D read.  DEPTH rows (32 bytes each) are in
    // flight per thread, 128 threads per SM: enough outstanding gathers to cover HBM latency.
    // =====================================================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    constexpr int DEPTH = RowCfg<kCfg>::depth;
    constexpr int STAGES = DEPTH + 1;
    const int p = tid - kEpiThreads;                       // pillar of the group == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    unsigned long long* const s_ent = reinterpret_cast<unsigned long long*>(smem + SP.ent);
    float* const s_row = smem + SP.row;
    constexpr unsigned long long kNoEntry = ~0ull;
    const int n_cols = kCfg ? NREG : A.c_raw;              // generic layout: columns 1 .. c_raw are staged

    struct Cursor { int w, j, slots, gi; };               // work item, slot inside it, its slot count, group counter
    auto cur_init = [&](Cursor& c) {
      c.w = blockIdx.x; c.j = 0; c.gi = 0;
      bool sg;
      c.slots = (c.w < total) ? slots_of(c.w, sg) : 1;
    };
    auto cur_next = [&](Cursor& c) -> bool {              // advance one slot; true when a new group starts
      if (++c.j < c.slots) return false;
      c.j = 0; c.w += G; ++c.gi;
      bool sg;
      c.slots = (c.w < total) ? slots_of(c.w, sg) : 1;
      return true;
    };
    auto issue_entry = [&](int w, int gi) {               // cursor A
      unsigned long long* dst = s_ent + (gi & (kEntRing - 1)) * kGroup + p;
      bool ok = false;
      if (w < total) {
        int q;
        const int list = list_of(w, q);
        const int e = (w - s_pre[q]) * kGroup + p;
        if (e < s_cnt[list]) { cp_async8(dst, A.lists + s_loff[list] + e); ok = true; }
      }
      if (!ok) *dst = kNoEntry;
    };
    auto entry_of = [&](int gi) { return s_ent[(gi & (kEntRing - 1)) * kGroup + p]; };

    // ---- prologue: fill the pipeline (entries, then row numbers, then rows), each stage after the previous landed ----
    int ga_w = blockIdx.x, ga_gi = 0;                      // cursor A
    Cursor cb, cc, cd;
    cur_init(cb); cur_init(cc); cur_init(cd);
    unsigned long long eb = kNoEntry, ec = kNoEntry, ed = kNoEntry;   // cached entries of the cursors' groups
    // Cursor A stays DEPTH + 1 groups ahead of cursor B's group: an entry requested in the iteration after B entered
    // group Y - DEPTH - 1 is committed DEPTH iterations before B can enter group Y (every group has at least one slot).
    auto step_a = [&](int b_gi) {
      while (ga_gi <= b_gi + DEPTH + 1) { issue_entry(ga_w, ga_gi); ga_w += G; ++ga_gi; }
    };
    auto step_b = [&](int slot) {                          // row number of cursor B's slot -> s_idx[slot ring]
      if (eb != kNoEntry) {
        int r, off, len;
        unpack_entry(eb, r, off, len);
        cp_async4(s_idx + (slot % STAGES) * kGroup + p, A.sorted_idx + off + min(cb.j, len - 1));
      }
      if (cur_next(cb)) eb = entry_of(cb.gi);
    };
    auto step_c = [&](int slot) {                          // row of cursor C's slot -> s_row[slot ring]
      if (ec != kNoEntry) {
        const int idx = s_idx[(slot % STAGES) * kGroup + p];
        const float* row = A.points + (int64_t)idx * A.stride;
        float* dst = s_row + (slot % STAGES) * (NREG * kGroup);
        if (kCfg == 1) {
          cp_async16(dst + p * 4, row);
          cp_async16(dst + 4 * kGroup + p * 4, row + 4);
        } else if (kCfg == 2) {
#pragma unroll
          for (int c = 0; c < 6; ++c) cp_async8(dst + c * 2 * kGroup + p * 2, row + 2 * c);
        } else {
          for (int c = 0; c < n_cols; ++c) cp_async4(dst + c * kGroup + p, row + 1 + c);
        }
      }
      if (cur_next(cc)) ec = entry_of(cc.gi);
    };
    int sb = 0, sc = 0;                                    // absolute slot numbers of cursors B and C
    // entries of the first 2 * DEPTH + 2 groups (everything cursor B can reach during the prologue) - one round trip
    while (ga_gi < 2 * DEPTH + 2) { issue_entry(ga_w, ga_gi); ga_w += G; ++ga_gi; }
    cp_async_commit();
    cp_async_wait<0>();
    eb = entry_of(0); ec = eb; ed = eb;
    for (int t = 0; t < DEPTH; ++t) step_b(sb++);          // row numbers of slots 0 .. DEPTH - 1 - one round trip
    cp_async_commit();
    cp_async_wait<0>();
    // rows of slots 0 .. DEPTH - 1 and row numbers of slots DEPTH .. 2 * DEPTH - 1, one commit group per slot: cursor D
    // is at slot 0, C at DEPTH, B at 2 * DEPTH, with DEPTH commit groups pending exactly as in the steady state
    // (at most DEPTH + 1 row numbers are live, which is the size of their ring)
    for (int t = 0; t < DEPTH; ++t) { step_b(sb++); step_c(sc++); cp_async_commit(); }

    auto mean_of = [&](unsigned long long e, bool is_seg) -> float4 {
      if (e == kNoEntry) return make_float4(0.f, 0.f, 0.f, 0.f);
      const int r = (int)(e & 0x1fffffffull);
      return __ldg((is_seg ? A.long_mean : A.mean) + r);
    };
    auto seg_of = [&](int w) { bool sg = false; if (w < total) slots_of(w, sg); return sg; };
    // pillar means: this group's, the next group's and the one after (register prefetch, two groups ahead)
    float4 m0 = mean_of(ed, seg_of(cd.w));
    float4 m1 = mean_of(entry_of(1), seg_of(cd.w + G));
    float4 m2 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t c0 = 0;
    int sd = 0;                                            // absolute slot number of cursor D
    bool new_group = true;
    const int n_feat = n_raw + (with_dist ? 7 : 6);
    TRACE_DECL(p == 0)
    TRACE(1, 9);
    while (cd.w < total) {
      // ---- pipeline upkeep: one entry / row number / row request per slot ----
      TRACE(1, new_group ? 13 : 10);
      cp_async_wait<DEPTH - 1>();
      TRACE(1, 14);                          // everything issued DEPTH or more slots ago has landed
      step_a(cb.gi);
      step_b(sb++);
      step_c(sc++);
      cp_async_commit();
      const bool valid = ed != kNoEntry;
      if (new_group) {
        // entries of the next two groups were requested at least 2 * DEPTH slots ago
        const bool is_seg = seg_of(cd.w);
        m2 = mean_of(entry_of(cd.gi + 2), seg_of(cd.w + 2 * G));
        const int r = valid ? (int)(ed & 0x1fffffffull) : -1;
        s_rows[(cd.gi % kRowRing) * kGroup + p] = r;      // pillar rank, or long-pillar index of a segment
        if (A.mean_out && valid && !is_seg) {
          float* m = A.mean_out + (int64_t)r * 3;
          m[0] = m0.x; m[1] = m0.y; m[2] = m0.z;
        }
      }
      // ---- this slot's row: shared memory -> registers ----
      float rw[NREG];
      {
        const float* src = s_row + (sd % STAGES) * (NREG * kGroup);
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
      }
      const float* gsrc = s_row + (sd % STAGES) * (NREG * kGroup) + p;   // generic layout: column c at gsrc[(c - 1) * kGroup]
      // ---- features of this slot's row (dynamic_pillar_vfe.py:111-126) ----
      float x = 0.f, y = 0.f, z = 0.f;
      if (valid) {
        if (kCfg) { x = rw[1]; y = rw[2]; z = rw[3]; }
        else { x = gsrc[0]; y = gsrc[kGroup]; z = gsrc[2 * kGroup]; }
      }
      float ed_[7];
      ed_[0] = __fsub_rn(x, m0.x);                                               // f_cluster (:111)
      ed_[1] = __fsub_rn(y, m0.y);
      ed_[2] = __fsub_rn(z, m0.z);
      const float cx = quantise(x, A.g.range_min_x, A.g.voxel_x);
      const float cy = quantise(y, A.g.range_min_y, A.g.voxel_y);
      ed_[3] = __fsub_rn(x, __fadd_rn(__fmul_rn(cx, A.g.voxel_x), A.g.x_offset));   // f_center (:114-116)
      ed_[4] = __fsub_rn(y, __fadd_rn(__fmul_rn(cy, A.g.voxel_y), A.g.y_offset));
      ed_[5] = __fsub_rn(z, A.g.z_offset);
      ed_[6] = with_dist ? __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))) : 0.f;  // :124
      // A0[c0 & 1] is free once the layer-0 MMA that read it two ops ago has completed
      const uint32_t b = c0 & 1;
      TRACE(1, 11);
      if (c0 >= 2) { mbar_wait(&bars[kBarD0 + b], ((c0 - 2) >> 1) & 1); tc_fence_after_sync(); }
      TRACE(1, 12);
      const uint32_t dh = tmem + kColA0 + b * 64 + lane_base, dl = dh + 32;
#pragma unroll
      for (int cc8 = 0; cc8 < kMaxCin; cc8 += 8) {
        if (cc8 < k0) {
          float hi[8], lo[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int f = cc8 + t;
            float val = 0.f;
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
      ++c0; ++sd;
      new_group = cur_next(cd);
      if (new_group) {
        ed = entry_of(cd.gi);
        m0 = m1; m1 = m2;
      }
    }
    cp_async_wait<0>();
    TRACE_END(1);
  } else {
    // =====================================================================================================
    // epilogue
    // =====================================================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int p = tid & (kGroup - 1);
    const int q = tid >> 7;                                  // column quarter
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t c0 = 0, c1 = 0;   // layer-0 accumulators consumed / layer-1-type ops staged
    float max0[8];             // layer-0 running max, this thread's 8 channels (two layers only)
    float m1[16];              // last-layer running max of the raw accumulators, this thread's 16 channels
    bool pend_fin = false;     // a finished group whose hoist result / output rows are still outstanding
    uint32_t pend_k = 0;       // its hoist op
    int pend_gi = 0;
    uint32_t e1_k = 0;         // op whose accumulator the next E1 reads

    TRACE_DECL(tid == 0)
    auto e0 = [&]() {          // BN + ReLU, running max0, x0 -> TMEM as the A operand of layer 1
      const uint32_t b = c0 & 1;
      TRACE(2, 30);
      mbar_wait(&bars[kBarD0 + b], (c0 >> 1) & 1);
      TRACE(2, 31);
      tc_fence_after_sync();
      uint32_t rr[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7])
                   : "r"(tmem + kColD0 + b * 64 + lane_base + 8 * q)
                   : "memory");
      tmem_ld_wait();
      float hi[8], lo[8];
      const float4 al0 = ld4(smem + SP.prm_a0 + 8 * q), al1 = ld4(smem + SP.prm_a0 + 8 * q + 4);
      const float4 be0 = ld4(smem + SP.prm_b0 + 8 * q), be1 = ld4(smem + SP.prm_b0 + 8 * q + 4);
      const float a8[8] = {al0.x, al0.y, al0.z, al0.w, al1.x, al1.y, al1.z, al1.w};
      const float b8[8] = {be0.x, be0.y, be0.z, be0.w, be1.x, be1.y, be1.z, be1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xv = fmaxf(fmaf(__uint_as_float(rr[i]), a8[i], b8[i]), 0.f);
        max0[i] = fmaxf(max0[i], xv);
        split_tf32(xv, hi[i], lo[i]);
      }
      const uint32_t b1 = c1 & 1;
      tmem_st8(tmem + kColA1 + b1 * 64 + lane_base + 8 * q, hi);
      tmem_st8(tmem + kColA1 + b1 * 64 + 32 + lane_base + 8 * q, lo);
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(&bars[kBarA1 + b1]);
      TRACE(2, 32);
      ++c0; ++c1;
    };
    auto ld_acc16 = [&](uint32_t k, uint32_t (&rr)[16]) {       // this thread's 16 columns of layer-1-type op k
      const uint32_t b = k & 1;
      TRACE(2, 33);
      mbar_wait(&bars[kBarD1 + b], (k >> 1) & 1);
      TRACE(2, 34);
      tc_fence_after_sync();
      tmem_ld16_nowait(tmem + kColD1 + b * 64 + lane_base + 16 * q, rr);
      tmem_ld_wait();
      tc_fence_before_sync();
    };
    auto e1 = [&]() {          // running max of the raw last-layer accumulators
      uint32_t rr[16];
      ld_acc16(e1_k, rr);
#pragma unroll
      for (int i = 0; i < 16; ++i) m1[i] = fmaxf(m1[i], __uint_as_float(rr[i]));
    };
    // OUT: BN(eval) + ReLU once per pillar; the tile is staged in shared memory so that every pillar_features row
    // leaves as 256 contiguous bytes
    auto write_out = [&](int gi) {
      const float* pa = smem + (kLayers == 2 ? SP.prm_a1 : SP.prm_a0) + 16 * q;
      const float* pb = smem + (kLayers == 2 ? SP.prm_b1 : SP.prm_b0) + 16 * q;
      float* dst = s_out + p * kOutLd + 16 * q;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 al = ld4(pa + i), be = ld4(pb + i);
        float4 o;
        o.x = fmaxf(fmaf(m1[i + 0], al.x, be.x), 0.f);
        o.y = fmaxf(fmaf(m1[i + 1], al.y, be.y), 0.f);
        o.z = fmaxf(fmaf(m1[i + 2], al.z, be.z), 0.f);
        o.w = fmaxf(fmaf(m1[i + 3], al.w, be.w), 0.f);
        *reinterpret_cast<float4*>(dst + i) = o;
      }
      named_bar_sync(1, kEpiThreads);
      const int* rows = s_rows + (gi % kRowRing) * kGroup;
#pragma unroll
      for (int t = 0; t < (kGroup * kCout / 4) / kEpiThreads; ++t) {
        const int item = t * kEpiThreads + tid;
        const int row = item >> 4, c4 = item & 15;
        const int r = rows[row];
        if (r >= 0) *reinterpret_cast<float4*>(A.out + (int64_t)r * kCout + c4 * 4) = ld4(s_out + row * kOutLd + c4 * 4);
      }
      named_bar_sync(2, kEpiThreads);
    };
    auto finish_group = [&]() {   // hoist result + output of the pending group
      if (kLayers == 2) {
        uint32_t rr[16];
        ld_acc16(pend_k, rr);
#pragma unroll
        for (int i = 0; i < 16; ++i) m1[i] = __fadd_rn(m1[i], __uint_as_float(rr[i]));
      }
      TRACE(2, 37);
      write_out(pend_gi);
      TRACE(2, 38);
      pend_fin = false;
    };

    int gi = 0;
    for (int w = blockIdx.x; w < total; w += G, ++gi) {
      bool is_seg;
      const int slots = slots_of(w, is_seg);
      if (kLayers == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) max0[i] = 0.f;
        for (int j = 0; j < slots; ++j) {
          e0();                                       // stages op c1 - 1
          if (j == 0) {
            if (pend_fin) finish_group();             // previous group: its hoist ran behind our first layer-0 epilogue
#pragma unroll
            for (int i = 0; i < 16; ++i) m1[i] = -INFINITY;
          } else {
            e1();                                     // previous slot
          }
          e1_k = c1 - 1;
        }
        if (!is_seg) {
          // hoist: max0 -> A1; the MMA warp multiplies by W1[:, 32:]^T
          float hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) split_tf32(max0[i], hi[i], lo[i]);
          const uint32_t b1 = c1 & 1;
          tmem_st8(tmem + kColA1 + b1 * 64 + lane_base + 8 * q, hi);
          tmem_st8(tmem + kColA1 + b1 * 64 + 32 + lane_base + 8 * q, lo);
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&bars[kBarA1 + b1]);
          pend_k = c1;
          ++c1;
        }
        e1();                                         // last slot
        if (is_seg) {
          // long pillar segment: partial maxima -> the pillar's accumulator
          const int li = s_rows[(gi % kRowRing) * kGroup + p];
          if (li >= 0) {
            unsigned* acc = A.long_acc + (int64_t)li * 96;
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicMax(acc + 8 * q + i, ord_enc(max0[i]));
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicMax(acc + 32 + 16 * q + i, ord_enc(m1[i]));
          }
        } else {
          pend_fin = true; pend_gi = gi;
        }
      } else {
        // single layer: D0 (64 columns) is the last-layer accumulator
#pragma unroll
        for (int i = 0; i < 16; ++i) m1[i] = -INFINITY;
        for (int j = 0; j < slots; ++j) {
          const uint32_t b = c0 & 1;
          mbar_wait(&bars[kBarD0 + b], (c0 >> 1) & 1);
          tc_fence_after_sync();
          uint32_t rr[16];
          tmem_ld16_nowait(tmem + kColD0 + b * 64 + lane_base + 16 * q, rr);
          tmem_ld_wait();
          tc_fence_before_sync();
          mbar_arrive(&bars[kBarA1 + b]);             // "D0[b] consumed"
          ++c0;
#pragma unroll
          for (int i = 0; i < 16; ++i) m1[i] = fmaxf(m1[i], __uint_as_float(rr[i]));
        }
        if (is_seg) {
          const int li = s_rows[(gi % kRowRing) * kGroup + p];
          if (li >= 0) {
            unsigned* acc = A.long_acc + (int64_t)li * 96;
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicMax(acc + 32 + 16 * q + i, ord_enc(m1[i]));
          }
        } else {
          write_out(gi);
        }
      }
    }
    if (pend_fin) finish_group();
    TRACE_END(2);
  }
  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
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

#ifdef PCP_PFN_TIMING
// debug: copies the event trace of CTA 0 (3 roles x kTraceCap x (id, clock)) and the 3 event counts to host memory
extern "C" int pcp_debug_read_timing(long long* trace_host, int* counts_host) {
  PCP_CUDA(cudaDeviceSynchronize());
  PCP_CUDA(cudaMemcpyFromSymbol(trace_host, g_trace, sizeof(long long) * 3 * kTraceCap * 2));
  PCP_CUDA(cudaMemcpyFromSymbol(counts_host, g_trace_n, sizeof(int) * 3));
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
  const unsigned blocks = (unsigned)(groups < 148 ? groups : 148);
  pfn_slot_kernel<kLayers, kCfg><<<blocks, kPfnThreads, SP.total_bytes, stream>>>(a);
  PCP_LAUNCH_CHECK("pfn_slot_kernel");
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
