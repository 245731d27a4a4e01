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
constexpr int kIdxBufs = 3;         // row-number buffers: group g (in use), g + 1 (mean prefetch), g + 2 (cp.async in flight)
constexpr int kRowRing = 8;         // ring of per-group output-row tables (producer runs a few groups ahead of the output)
// mbarrier indices
constexpr int kBarA0 = 0;           // [2] features staged in TMEM        (128 producer arrivals)
constexpr int kBarA1 = 2;           // [2] x0 / max0 staged in TMEM       (512 epilogue arrivals)
constexpr int kBarD0 = 4;           // [2] layer-0 accumulator ready      (tcgen05.commit)
constexpr int kBarD1 = 6;           // [2] layer-1 / hoist accumulator ready

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
  S.idx = o; o += kIdxBufs * kSegRows * kGroup;   // row numbers, [buffer][slot][pillar]
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
  const SmemPlan SP = smem_plan(k0, kLayers);
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
    auto issue_m0 = [&]() {
      const uint32_t b = c0 & 1;
      mbar_wait(&bars[kBarA0 + b], (c0 >> 1) & 1);
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
      mbar_wait(&bars[kBarA1 + b], (c1 >> 1) & 1);
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
    }
  } else if (warp >= kProdWarp0) {
    // =====================================================================================================
    // producers (register budget raised with what the other roles gave back to the CTA pool)
    // =====================================================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int p = tid - kEpiThreads;                       // pillar of the group == TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    struct Ent { int r, off, len, li; bool valid; };
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
    auto decode = [&](const int4& raw, Ent& E) {
      E.valid = raw.w >= 0;
      E.r = -1; E.off = 0; E.len = 0; E.li = -1;
      if (raw.w == 1) { E.off = raw.x; E.len = raw.y; E.li = raw.z; }
      else if (raw.w == 0) unpack_entry(((unsigned long long)(unsigned)raw.y << 32) | (unsigned)raw.x, E.r, E.off, E.len);
    };
    auto issue_idx = [&](const Ent& E, int w, int b) {
      if (E.valid) {
        bool sg;
        const int slots = slots_of(w, sg);
        int* dst = s_idx + b * (kSegRows * kGroup) + p;
        for (int j = 0; j < slots; ++j) cp_async4(dst + j * kGroup, A.sorted_idx + E.off + min(j, E.len - 1));
      }
    };
    // xyz of up to 8 rows of a pillar (loads only: summed later, in ascending row order)
    auto load_xyz8 = [&](const Ent& E, const int* idx, int j0, float (&vx)[8], float (&vy)[8], float (&vz)[8]) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        if (E.valid && j0 + t < E.len) {
          const float* row = A.points + (int64_t)idx[(j0 + t) * kGroup] * A.stride;
          if (kCfg == 1) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row));
            vx[t] = v.y; vy[t] = v.z; vz[t] = v.w;
          } else {
            vx[t] = __ldg(row + 1); vy[t] = __ldg(row + 2); vz[t] = __ldg(row + 3);
          }
        }
      }
    };
    // scatter_mean = sum in ascending row order / count (dynamic_pillar_vfe.py:110); first 8 rows already in registers
    auto finish_mean = [&](const Ent& E, const int* idx, bool is_seg, const float (&vx)[8], const float (&vy)[8],
                           const float (&vz)[8], float& mx, float& my, float& mz) {
      mx = my = mz = 0.f;
      if (!E.valid) return;
      if (is_seg) {
        const float4 m = __ldg(A.long_mean + E.li);
        mx = m.x; my = m.y; mz = m.z;
        return;
      }
      float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (t < E.len) { sx = __fadd_rn(sx, vx[t]); sy = __fadd_rn(sy, vy[t]); sz = __fadd_rn(sz, vz[t]); }
      for (int j0 = 8; j0 < E.len; j0 += 8) {
        float wx[8], wy[8], wz[8];
        load_xyz8(E, idx, j0, wx, wy, wz);
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (j0 + t < E.len) { sx = __fadd_rn(sx, wx[t]); sy = __fadd_rn(sy, wy[t]); sz = __fadd_rn(sz, wz[t]); }
      }
      const float cnt = (float)E.len;
      mx = __fdiv_rn(sx, cnt); my = __fdiv_rn(sy, cnt); mz = __fdiv_rn(sz, cnt);
      if (A.mean_out) {
        float* m = A.mean_out + (int64_t)E.r * 3;
        m[0] = mx; m[1] = my; m[2] = mz;
      }
    };

    // ---- prime the prefetch pipeline: entries of groups 0..2, row numbers of groups 0..1, mean of group 0 ----
    Ent cur, nxt, nn;
    cur.valid = nxt.valid = nn.valid = false;
    cur.r = nxt.r = nn.r = -1; cur.off = nxt.off = nn.off = 0; cur.len = nxt.len = nn.len = 0; cur.li = nxt.li = nn.li = -1;
    const int w0 = blockIdx.x;
    {
      int4 raw;
      if (w0 < total) { fetch(w0, raw); decode(raw, cur); issue_idx(cur, w0, 0); }
      cp_async_commit();
      if (w0 + G < total) { fetch(w0 + G, raw); decode(raw, nxt); issue_idx(nxt, w0 + G, 1); }
      cp_async_commit();
      if (w0 + 2 * G < total) { fetch(w0 + 2 * G, raw); decode(raw, nn); }
    }
    float mean_x = 0.f, mean_y = 0.f, mean_z = 0.f;
    if (w0 < total) {
      cp_async_wait<1>();
      bool sg;
      slots_of(w0, sg);
      float vx[8], vy[8], vz[8];
      load_xyz8(cur, s_idx + p, 0, vx, vy, vz);
      finish_mean(cur, s_idx + p, sg, vx, vy, vz, mean_x, mean_y, mean_z);
    }
    uint32_t c0 = 0;
    int ib = 0;                 // row-number buffer of the current group
    int gi = 0;                 // group counter (ring index of s_rows)
    float rw[NREG];
    const float* rowp = A.points;
    for (int w = w0; w < total; w += G, ++gi) {
      bool is_seg, nseg = false;
      const int slots = slots_of(w, is_seg);
      const bool have_next = (w + G) < total;
      if (have_next) slots_of(w + G, nseg);
      const int* my_idx = s_idx + ib * (kSegRows * kGroup) + p;
      const int ib1 = (ib + 1) % kIdxBufs, ib2 = (ib + 2) % kIdxBufs;
      const int* nx_idx = s_idx + ib1 * (kSegRows * kGroup) + p;
      const bool valid = cur.valid;
      s_rows[(gi % kRowRing) * kGroup + p] = is_seg ? cur.li : cur.r;
      // ---- prefetch for the following groups (all in flight while this group's slots are built) ----
      int4 raw3 = make_int4(0, 0, 0, -1);
      if (w + 2 * G < total) issue_idx(nn, w + 2 * G, ib2);       // row numbers of group g + 2
      cp_async_commit();
      if (w + 3 * G < total) fetch(w + 3 * G, raw3);              // entry of group g + 3
      cp_async_wait<1>();                                         // row numbers of group g + 1 have landed
      float vx[8], vy[8], vz[8];
      if (have_next && !nseg) load_xyz8(nxt, nx_idx, 0, vx, vy, vz);   // rows of group g + 1 for its mean

      auto load_row = [&](const int* idx, int j, bool ok) {
        if (!ok) return;
        rowp = A.points + (int64_t)idx[j * kGroup] * A.stride;
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
      if (w == w0) load_row(my_idx, 0, valid);                    // later groups: fetched at the end of the previous group

      for (int j = 0; j < slots; ++j) {
        // ---- features of this slot's row (dynamic_pillar_vfe.py:111-126) ----
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
        // A0[c0 & 1] is free once the layer-0 MMA that read it two ops ago has completed
        const uint32_t b = c0 & 1;
        if (c0 >= 2) { mbar_wait(&bars[kBarD0 + b], ((c0 - 2) >> 1) & 1); tc_fence_after_sync(); }
        const uint32_t dh = tmem + kColA0 + b * 64 + lane_base, dl = dh + 32;
#pragma unroll
        for (int cc = 0; cc < kMaxCin; cc += 8) {
          if (cc < k0) {
            float hi[8], lo[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int f = cc + t;
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
            tmem_st8(dh + cc, hi);
            tmem_st8(dl + cc, lo);
          }
        }
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(&bars[kBarA0 + b]);
        ++c0;
        // ---- next row: this group's next slot, or slot 0 of the next group ----
        if (j + 1 < slots) load_row(my_idx, j + 1, valid);
        else if (have_next) load_row(nx_idx, 0, nxt.valid);
      }
      // ---- rotate: mean of the next group from the rows fetched above ----
      if (have_next) finish_mean(nxt, nx_idx, nseg, vx, vy, vz, mean_x, mean_y, mean_z);
      cur = nxt; nxt = nn;
      decode(raw3, nn);
      ib = ib1;
    }
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

    auto e0 = [&]() {          // BN + ReLU, running max0, x0 -> TMEM as the A operand of layer 1
      const uint32_t b = c0 & 1;
      mbar_wait(&bars[kBarD0 + b], (c0 >> 1) & 1);
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
      ++c0; ++c1;
    };
    auto ld_acc16 = [&](uint32_t k, uint32_t (&rr)[16]) {       // this thread's 16 columns of layer-1-type op k
      const uint32_t b = k & 1;
      mbar_wait(&bars[kBarD1 + b], (k >> 1) & 1);
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
      write_out(pend_gi);
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
