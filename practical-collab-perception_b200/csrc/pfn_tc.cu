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
constexpr int kWorkers = 256;       // epilogue / feature threads of the PFN kernel: (pillar p, column half h)
constexpr int kPfnThreads = kWorkers + 32;   // + one warp that only issues tensor-core instructions
constexpr int kMmaWarp = kWorkers / 32;
constexpr int kTmemCols = 256;
// tensor-memory column map of one CTA
constexpr uint32_t kColD0 = 0;      // layer-0 accumulator (32 columns; 64 for a single-layer PFN)
constexpr uint32_t kColD1 = 64;     // layer-1 / hoist accumulator (64 columns)
constexpr uint32_t kColA0h = 128;   // layer-0 A operand (features), TF32 hi part, k0 <= 24 columns
constexpr uint32_t kColA0l = 160;   //                               lo part
constexpr uint32_t kColA1h = 192;   // layer-1 A operand (x0, later max0), hi part, 32 columns
constexpr uint32_t kColA1l = 224;   //                                     lo part
constexpr int kOutLd = 68;          // padded row of the output staging tile (conflict-free 16-byte accesses)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// kCfg: 0 = any layout (scalar loads, run-time feature map)
//       1 = car / early-fusion rows: c_raw 5, absolute xyz, no distance, row stride % 4 == 0, 16-byte aligned
//       2 = ego (lately fusion) rows: c_raw 11, absolute xyz, no distance, even row stride, 8-byte aligned
template <int kCfg> struct RowCfg { static constexpr int n_raw = 0, k0 = 0, nreg = 1; };
template <> struct RowCfg<1> { static constexpr int n_raw = 5, k0 = 16, nreg = 8; };
template <> struct RowCfg<2> { static constexpr int n_raw = 11, k0 = 24, nreg = 12; };

struct SmemPlan {   // float offsets into dynamic shared memory
  int w0h, w0l, w1ah, w1al, w1bh, w1bl, panels_end, prm_a0, prm_b0, prm_a1, prm_b1, idx, out, rows, ints, total_bytes;
};
__host__ __device__ inline SmemPlan smem_plan(int k0, int layers) {
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
  S.idx = o; o += 2 * kSegRows * kGroup;     // row numbers of the current and the next group, [buffer][slot][pillar]
  S.out = o; o += kGroup * kOutLd;           // output staging tile (coalesced pillar_features rows)
  S.rows = o; o += 2 * kGroup;               // output row (pillar rank) / long-pillar index of each lane, [group parity][pillar]
  S.ints = o; o += 64;                       // 4 mbarriers | tmem base | group prefix | list counts | list offsets
  S.total_bytes = o * 4;
  return S;
}

#ifdef PCP_PFN_TIMING
__device__ long long g_pfn_timing[8192];
}  // namespace pcp
extern "C" int pcp_debug_read_timing(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, pcp::g_pfn_timing, sizeof(long long) * 8192);
}
namespace pcp {
#define PCP_T(id)                                                                             \
  do {                                                                                        \
    if (blockIdx.x == 7 && (tid == 0 || tid == 200) && dbg_n < 1000) {                        \
      g_pfn_timing[(tid ? 4096 : 0) + dbg_n * 4 + 0] = (id);                                  \
      g_pfn_timing[(tid ? 4096 : 0) + dbg_n * 4 + 1] = clock64();                             \
      dbg_n++;                                                                                \
    }                                                                                         \
  } while (0)
#else
#define PCP_T(id) do {} while (0)
#endif

struct Work {        // one thread's pillar (or long-pillar segment) of a group
  int r, off, len, li;
  bool valid;
};

// Roles (no __syncthreads inside the main loop; everything is ordered by mbarriers):
//   workers, threads 0..255 = (pillar p = tid & 127, column half h = tid >> 7)
//     h == 0 also gathers the pillar's rows, computes its mean and writes the A0 operand (features) to TMEM
//     both halves run the epilogues E0 / E1 on their half of the accumulator columns and keep the running maxima
//   MMA warp, threads 256..287: waits for "operand staged" barriers, issues the tcgen05.mma groups, commits them to
//     the "accumulator ready" barriers.  The tensor pipe back-pressures its issuer, so issuing from a worker would
//     stall the whole CTA for the duration of every MMA group (measured: 500-1000 cycles per slot).
template <int kLayers, int kCfg>
__global__ void __launch_bounds__(kPfnThreads, 2)
pfn_slot_kernel(const TcArgs A) {
  extern __shared__ __align__(128) float smem[];
  constexpr int N0 = (kLayers == 2) ? kHidden : kCout;
  constexpr int NREG = RowCfg<kCfg>::nreg;
  const int k0 = kCfg ? RowCfg<kCfg>::k0 : A.k0;
  const int n_raw = kCfg ? RowCfg<kCfg>::n_raw : A.n_raw;
  const int raw_col0 = kCfg ? 1 : A.raw_col0;
  const bool with_dist = kCfg ? false : (A.with_distance != 0);
  const SmemPlan SP = smem_plan(k0, kLayers);
  const int tid = threadIdx.x, warp = tid >> 5;
#ifdef PCP_PFN_TIMING
  int dbg_n = 0;
#endif
  const int p = tid & (kGroup - 1);      // pillar of the group == TMEM lane
  const int h = (tid >> 7) & 1;          // which half of the accumulator columns
  int* const s_idx = reinterpret_cast<int*>(smem + SP.idx);
  float* const s_out = smem + SP.out;
  int* const s_rows = reinterpret_cast<int*>(smem + SP.rows);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + SP.ints);   // 0: a0 staged, 1: a1 staged, 2: d0 ready, 3: d1 ready
  uint32_t* const s_tmem = reinterpret_cast<uint32_t*>(smem + SP.ints + 8);
  int* const s_pre = reinterpret_cast<int*>(smem + SP.ints + 12);      // [kNumLists + 1] group prefix, processing order
  int* const s_cnt = reinterpret_cast<int*>(smem + SP.ints + 28);      // [kNumLists] entries per list
  long long* const s_loff = reinterpret_cast<long long*>(smem + SP.ints + 40);   // [kNumLists] list offsets

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
      mbar_init(&bars[0], kGroup);      // A0 staged: one arrival per pillar (the h == 0 workers)
      mbar_init(&bars[1], kWorkers);    // A1 staged: every worker
      mbar_init(&bars[2], 1);           // D0 ready: tcgen05.commit
      mbar_init(&bars[3], 1);           // D1 ready: tcgen05.commit
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
  const uint32_t t_d0 = tmem + kColD0, t_d1 = tmem + kColD1;
  const uint32_t t_a0h = tmem + kColA0h, t_a0l = tmem + kColA0l, t_a1h = tmem + kColA1h, t_a1l = tmem + kColA1l;
  const int total = s_pre[kNumLists];
  const int G = gridDim.x;
  auto list_of = [&](int w, int& q) {
    q = 0;
#pragma unroll
    for (int t = 1; t < kNumLists; ++t) q += (w >= s_pre[t]) ? 1 : 0;
    return kNumLists - 1 - q;
  };

  if (warp == kMmaWarp) {
    // =====================================================================================================
    // MMA warp
    // =====================================================================================================
    const uint32_t sw0h = smem_u32(smem + SP.w0h), sw0l = smem_u32(smem + SP.w0l);
    const uint32_t sw1ah = smem_u32(smem + SP.w1ah), sw1al = smem_u32(smem + SP.w1al);
    const uint32_t sw1bh = smem_u32(smem + SP.w1bh), sw1bl = smem_u32(smem + SP.w1bl);
    const uint32_t idesc0 = idesc_tf32_m128(N0), idesc1 = idesc_tf32_m128(kCout);
    uint32_t pa0 = 0, pa1 = 0;
    for (int w = blockIdx.x; w < total; w += G) {
      int q;
      const int list = list_of(w, q);
      const bool is_seg = (list == kSegList);
      const int slots = is_seg ? kSegRows : class_slots(list);
      for (int j = 0; j < slots; ++j) {
        mbar_wait(&bars[0], pa0); pa0 ^= 1;
        tc_fence_after_sync();
        if (elect_one_sync()) {
          mma_3xtf32_ts(t_d0, t_a0h, t_a0l, sw0h, sw0l, N0, k0 / 8, idesc0, false);
          mma_commit(&bars[2]);
        }
        __syncwarp();
        if (kLayers == 2) {
          mbar_wait(&bars[1], pa1); pa1 ^= 1;
          tc_fence_after_sync();
          if (elect_one_sync()) {
            mma_3xtf32_ts(t_d1, t_a1h, t_a1l, sw1ah, sw1al, kCout, kHidden / 8, idesc1, false);
            mma_commit(&bars[3]);
          }
          __syncwarp();
        }
      }
      if (kLayers == 2 && !is_seg) {
        // hoist: max0 . W1[:, 32:]^T once per pillar
        mbar_wait(&bars[1], pa1); pa1 ^= 1;
        tc_fence_after_sync();
        if (elect_one_sync()) {
          mma_3xtf32_ts(t_d1, t_a1h, t_a1l, sw1bh, sw1bl, kCout, kHidden / 8, idesc1, false);
          mma_commit(&bars[3]);
        }
        __syncwarp();
      }
    }
  } else {
    // =====================================================================================================
    // workers
    // =====================================================================================================
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t pd0 = 0, pd1 = 0;
    // raw 16-byte descriptor of this thread's pillar in group w (loads only; decoded later so they stay in flight)
    auto fetch = [&](int w, int4& raw) {
      int q;
      const int list = list_of(w, q);
      const int e = (w - s_pre[q]) * kGroup + p;
      raw = make_int4(0, 0, 0, -1);
      if (e < s_cnt[list]) {
        if (list == kSegList) {
          raw = __ldg(A.seg_table + e);
          raw.w = 1;
        } else {
          const unsigned long long v = __ldg(A.lists + s_loff[list] + e);
          raw.x = (int)(v & 0xffffffffull); raw.y = (int)(v >> 32); raw.w = 0;
        }
      }
    };
    auto decode = [&](const int4& raw, Work& W) {
      W.valid = raw.w >= 0;
      W.r = -1; W.off = 0; W.len = 0; W.li = -1;
      if (raw.w == 1) { W.off = raw.x; W.len = raw.y; W.li = raw.z; }
      else if (raw.w == 0) unpack_entry(((unsigned long long)(unsigned)raw.y << 32) | (unsigned)raw.x, W.r, W.off, W.len);
    };
    // row numbers of every slot of a group -> s_idx[b] (cp.async: no registers, lands while the previous group computes)
    auto issue_idx = [&](const Work& W, int slots, int b) {
      if (W.valid) {
        int* dst = s_idx + b * (kSegRows * kGroup) + p;
        for (int j = 0; j < slots; ++j) cp_async4(dst + j * kGroup, A.sorted_idx + W.off + min(j, W.len - 1));
      }
    };
    auto slots_of = [&](int w) {
      int q;
      const int list = list_of(w, q);
      return list == kSegList ? kSegRows : class_slots(list);
    };

    Work cur, nxt;
    cur.valid = false; cur.r = -1; cur.off = 0; cur.len = 0; cur.li = -1;
    nxt = cur;
    if (h == 0) {
      int4 raw;
      if ((int)blockIdx.x < total) { fetch(blockIdx.x, raw); decode(raw, cur); issue_idx(cur, slots_of(blockIdx.x), 0); }
      cp_async_commit();
      if ((int)blockIdx.x + G < total) { fetch(blockIdx.x + G, raw); decode(raw, nxt); }
    }
    int buf = 0;
    bool pending = false;      // a finished group whose hoist MMA is in flight / whose output is not yet written
    int pend_par = 0;
    float max0[16];            // layer-0 running max, this thread's 16 channels (two layers only)
    float m1[32];              // last-layer running max of the raw accumulators, this thread's 32 channels

    // ======== OUT of a finished group: h = max0 . W1[:, 32:]^T landed in D1; BN(eval) + ReLU once per pillar; the tile is
    // staged in shared memory so that every pillar_features row leaves as 256 contiguous bytes =========================
    auto finish_pending = [&]() {
      if (kLayers == 2) {
        mbar_wait(&bars[3], pd1); pd1 ^= 1;
        tc_fence_after_sync();
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          uint32_t rr[16];
          tmem_ld16_nowait(t_d1 + lane_base + 32 * h + 16 * part, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) m1[part * 16 + i] = __fadd_rn(m1[part * 16 + i], __uint_as_float(rr[i]));
        }
        tc_fence_before_sync();
      }
      {
        const float* pa = smem + (kLayers == 2 ? SP.prm_a1 : SP.prm_a0) + 32 * h;
        const float* pb = smem + (kLayers == 2 ? SP.prm_b1 : SP.prm_b0) + 32 * h;
        float* dst = s_out + p * kOutLd + 32 * h;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 al = ld4(pa + i), be = ld4(pb + i);
          float4 o;
          o.x = fmaxf(fmaf(m1[i + 0], al.x, be.x), 0.f);
          o.y = fmaxf(fmaf(m1[i + 1], al.y, be.y), 0.f);
          o.z = fmaxf(fmaf(m1[i + 2], al.z, be.z), 0.f);
          o.w = fmaxf(fmaf(m1[i + 3], al.w, be.w), 0.f);
          *reinterpret_cast<float4*>(dst + i) = o;
        }
      }
      named_bar_sync(1, kWorkers);
      const int* rows = s_rows + pend_par * kGroup;
#pragma unroll
      for (int t = 0; t < (kGroup * kCout / 4) / kWorkers; ++t) {
        const int item = t * kWorkers + tid;
        const int row = item >> 4, c4 = item & 15;
        const int r = rows[row];
        if (r >= 0) *reinterpret_cast<float4*>(A.out + (int64_t)r * kCout + c4 * 4) = ld4(s_out + row * kOutLd + c4 * 4);
      }
      pending = false;
    };

    int par = 0;
    for (int w = blockIdx.x; w < total; w += G, par ^= 1) {
      int q;
      const int list = list_of(w, q);
      const bool is_seg = (list == kSegList);
      const int slots = is_seg ? kSegRows : class_slots(list);
      PCP_T(100 + slots);
      const int* my_idx = s_idx + buf * (kSegRows * kGroup) + p;
      int4 nn_raw = make_int4(0, 0, 0, -1);
      const bool have_nn = (w + 2 * G) < total;
      float mean_x = 0.f, mean_y = 0.f, mean_z = 0.f;
      const bool valid = cur.valid;
      const int len = cur.len;
      if (h == 0) {
        // ---- prefetch: row numbers of the next group, descriptor of the one after ----
        if (w + G < total) issue_idx(nxt, slots_of(w + G), buf ^ 1);
        cp_async_commit();
        if (have_nn) fetch(w + 2 * G, nn_raw);
        cp_async_wait<1>();                               // this thread's copies of the current group have landed
        s_rows[par * kGroup + p] = is_seg ? cur.li : cur.r;
        // ---- pillar mean: scatter_mean = sum in ascending row order / count (dynamic_pillar_vfe.py:110) ----
        if (valid) {
          if (is_seg) {
            const float4 m = __ldg(A.long_mean + cur.li);
            mean_x = m.x; mean_y = m.y; mean_z = m.z;
          } else {
            float sx = 0.f, sy = 0.f, sz = 0.f;
            for (int j0 = 0; j0 < len; j0 += 8) {
              float vx[8], vy[8], vz[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                if (j0 + t < len) {
                  const float* row = A.points + (int64_t)my_idx[(j0 + t) * kGroup] * A.stride;
                  if (kCfg == 1) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(row));
                    vx[t] = v.y; vy[t] = v.z; vz[t] = v.w;
                  } else {
                    vx[t] = __ldg(row + 1); vy[t] = __ldg(row + 2); vz[t] = __ldg(row + 3);
                  }
                }
              }
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                if (j0 + t < len) { sx = __fadd_rn(sx, vx[t]); sy = __fadd_rn(sy, vy[t]); sz = __fadd_rn(sz, vz[t]); }
              }
            }
            const float cnt = (float)len;
            mean_x = __fdiv_rn(sx, cnt); mean_y = __fdiv_rn(sy, cnt); mean_z = __fdiv_rn(sz, cnt);
            if (A.mean_out) {
              float* m = A.mean_out + (int64_t)cur.r * 3;
              m[0] = mean_x; m[1] = mean_y; m[2] = mean_z;
            }
          }
        }
      }
      PCP_T(103);
      // the previous group's hoist MMA ran while the loads above were in flight
      if (pending) finish_pending();
      PCP_T(105);

#pragma unroll
      for (int i = 0; i < 16; ++i) max0[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) m1[i] = -INFINITY;

      // ---- row fetch (registers; issued one slot ahead) and A0 = TF32 hi / lo features -> tensor memory (h == 0) ----
      float rw[NREG];
      const float* rowp = A.points;
      auto load_row = [&](int j) {
        if (!valid) return;
        rowp = A.points + (int64_t)my_idx[j * kGroup] * A.stride;
        if (kCfg == 1) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(rowp));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(rowp) + 1);
          rw[0] = v0.x; rw[1] = v0.y; rw[2] = v0.z; rw[3] = v0.w; rw[4] = v1.x; rw[5] = v1.y; rw[6] = v1.z; rw[7] = v1.w;
        } else if (kCfg == 2) {
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(rowp) + c);
            rw[2 * c] = v.x; rw[2 * c + 1] = v.y;
          }
        }
      };
      auto build_a0 = [&]() {
        float x = 0.f, y = 0.f, z = 0.f;
        if (valid) {
          if (kCfg) { x = rw[1]; y = rw[2]; z = rw[3]; }
          else { x = __ldg(rowp + 1); y = __ldg(rowp + 2); z = __ldg(rowp + 3); }
        }
        float ed[7];
        ed[0] = __fsub_rn(x, mean_x);                                              // f_cluster (:111)
        ed[1] = __fsub_rn(y, mean_y);
        ed[2] = __fsub_rn(z, mean_z);
        const float cx = quantise(x, A.g.range_min_x, A.g.voxel_x);
        const float cy = quantise(y, A.g.range_min_y, A.g.voxel_y);
        ed[3] = __fsub_rn(x, __fadd_rn(__fmul_rn(cx, A.g.voxel_x), A.g.x_offset));   // f_center (:114-116)
        ed[4] = __fsub_rn(y, __fadd_rn(__fmul_rn(cy, A.g.voxel_y), A.g.y_offset));
        ed[5] = __fsub_rn(z, A.g.z_offset);
        ed[6] = with_dist ? __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))) : 0.f;  // :124
        const int n_feat = n_raw + (with_dist ? 7 : 6);
        const uint32_t dh = t_a0h + lane_base, dl = t_a0l + lane_base;
#pragma unroll
        for (int c0 = 0; c0 < kMaxCin; c0 += 8) {
          if (c0 < k0) {
            float hi[8], lo[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int f = c0 + t;
              float val = 0.f;
              if (valid) {
                if (f < n_raw) {
                  if (kCfg) val = rw[(1 + f) < NREG ? (1 + f) : (NREG - 1)];
                  else val = __ldg(rowp + raw_col0 + f);
                } else if (f < n_feat) {
                  const int d = f - n_raw;
                  val = d == 0 ? ed[0] : d == 1 ? ed[1] : d == 2 ? ed[2] : d == 3 ? ed[3] : d == 4 ? ed[4] : d == 5 ? ed[5] : ed[6];
                }
              }
              split_tf32(val, hi[t], lo[t]);
            }
            tmem_st8(dh + c0, hi);
            tmem_st8(dl + c0, lo);
          }
        }
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(&bars[0]);
      };

      if (h == 0) {
        load_row(0);
        build_a0();
        if (slots > 1) load_row(1);
      }
      PCP_T(108);

      for (int j = 0; j < slots; ++j) {
        PCP_T(1);
        mbar_wait(&bars[2], pd0); pd0 ^= 1;
        tc_fence_after_sync();
        PCP_T(2);
        if (kLayers == 2) {
          // ================= E0: BN+ReLU, running max0, x0 -> TMEM as the A operand of layer 1 =================
          {
            uint32_t rr[16];
            tmem_ld16_nowait(t_d0 + lane_base + 16 * h, rr);
            tmem_ld_wait();
            float hi[16], lo[16];
            const float* pa = smem + SP.prm_a0 + 16 * h;
            const float* pb = smem + SP.prm_b0 + 16 * h;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 al = ld4(pa + i), be = ld4(pb + i);
              const float a4[4] = {al.x, al.y, al.z, al.w}, b4[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float xv = fmaxf(fmaf(__uint_as_float(rr[i + t]), a4[t], b4[t]), 0.f);
                max0[i + t] = fmaxf(max0[i + t], xv);
                split_tf32(xv, hi[i + t], lo[i + t]);
              }
            }
            tmem_st16(t_a1h + lane_base + 16 * h, hi);
            tmem_st16(t_a1l + lane_base + 16 * h, lo);
            tmem_st_wait();
          }
          tc_fence_before_sync();
          mbar_arrive(&bars[1]);
          PCP_T(3);
          // ================= next slot's A0 goes in behind M1 =================
          if (h == 0 && j + 1 < slots) {
            build_a0();
            if (j + 2 < slots) load_row(j + 2);
          }
          PCP_T(6);
          mbar_wait(&bars[3], pd1); pd1 ^= 1;
          tc_fence_after_sync();
          PCP_T(8);
        }
        // ================= E1: running max of the raw last-layer accumulators =================
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          uint32_t rr[16];
          tmem_ld16_nowait((kLayers == 2 ? t_d1 : t_d0) + lane_base + 32 * h + 16 * part, rr);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) m1[part * 16 + i] = fmaxf(m1[part * 16 + i], __uint_as_float(rr[i]));
        }
        tc_fence_before_sync();
        PCP_T(9);
        if (kLayers == 1) {
          // single layer: D0 is free again only now; the A1 barrier doubles as "D0 consumed" for the MMA warp
          if (j + 1 < slots) {
            named_bar_sync(1, kWorkers);
            if (h == 0) { build_a0(); if (j + 2 < slots) load_row(j + 2); }
          }
        }
      }

      if (is_seg) {
        // ---- long pillar segment: partial maxima -> the pillar's accumulator ----
        const int li = s_rows[par * kGroup + p];
        if (li >= 0) {
          unsigned* acc = A.long_acc + (int64_t)li * 96;
          if (kLayers == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicMax(acc + 16 * h + i, ord_enc(max0[i]));
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicMax(acc + 32 + 32 * h + i, ord_enc(m1[i]));
        }
        if (kLayers == 1) named_bar_sync(1, kWorkers);   // D0 consumed before the next group's first MMA
      } else {
        if (kLayers == 2) {
          // ============ H: max0 -> A1; the MMA warp multiplies by W1[:, 32:]^T; collected by finish_pending() ============
          float hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) split_tf32(max0[i], hi[i], lo[i]);
          tmem_st16(t_a1h + lane_base + 16 * h, hi);
          tmem_st16(t_a1l + lane_base + 16 * h, lo);
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&bars[1]);
        }
        pending = true; pend_par = par;
        if (kLayers == 1) { finish_pending(); named_bar_sync(1, kWorkers); }
      }
      PCP_T(110);
      // ---- rotate the prefetch pipeline ----
      cur = nxt;
      if (h == 0 && have_nn) decode(nn_raw, nxt);
      buf ^= 1;
    }
    if (pending) finish_pending();
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

// launched from pfn.cu
namespace pcp {

template <int kLayers, int kCfg>
static int launch_cfg(const TcArgs& a, int64_t n_points, cudaStream_t stream) {
  const SmemPlan SP = smem_plan(kCfg ? RowCfg<kCfg>::k0 : a.k0, kLayers);
  PCP_CUDA(cudaFuncSetAttribute(pfn_slot_kernel<kLayers, kCfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, SP.total_bytes));
  // upper bound of the group count: every list may end in a partial group
  const int64_t groups = n_points / kGroup + kNumLists;
  const unsigned blocks = (unsigned)(groups < 148 * 2 ? groups : 148 * 2);
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
