// TQC loss, group form: device code shared by tqc.cu (stand-alone kernel) and sample.cu (fused pass kernel).  See tqc.cu for the
// algorithm and the reference lines it replaces.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace fdql {

struct TqcArgs {
  int64_t M;
  int32_t n_atoms, n_z, n_drop;  // predicted atoms, pooled target atoms, target atoms cut from the top
  const float* next_z;
  const float* q_pred;
  const float* next_log_pi;
  const float* reward;
  const float* mask;
  const float* mc_return;
  const float* grad_scale;
  float alpha, gamma;
  float* loss;
  float* grad_q;
  float* td_target;
  double* stats;
  const float* alpha_dev;  // when set, overrides alpha
  int32_t grp_red_alias;   // group kernel: the per-lane partial sums live in the dead rows of the q_pred staging (see GrpCfg)
  // group kernel, optional: {next unclaimed group, blocks done} in device memory, both zero at launch.  When set, the warps of ALL blocks
  // claim their groups from this one counter (the SMs then finish together whatever their memory latencies were; the claims still form
  // one dense moving front) and the last block to finish zeroes the pair for the next launch.
  int* work_ctr;
};

__device__ __forceinline__ float2 lds2_at(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}

// =================================================================================================
// Group kernel (n_atoms, n_z <= 128): a warp owns G = 32/LPT transitions per round and works in two layouts.
//   phase A, LPT lanes per transition, E = 16 values per lane: the pooled target atoms come out of a shared-memory staging
//     copy (filled one round ahead by a 1-D bulk copy on a per-warp mbarrier), are sorted by a bitonic network that is in registers except for
//     log2(LPT) partner exchanges per merge phase, turned into the soft target, centred, prefix-summed (serial walk per lane +
//     a log2(LPT)-step scan) and written as G search tables.
//   phase B, the whole warp per transition: lane l handles predicted atoms l, l+32, ...; all 32 lanes search the SAME table,
//     whose layout makes every search level bank-conflict free, and read {-P1, P2} at the three split points.
// Against the warp-per-transition kernel above this removes most of the cross-lane sort traffic (6 instead of 15 exchange
// stages of 128 values), all scans/reductions/bounds checks that were paid per transition by 32 lanes, and the global-load
// address arithmetic (rows arrive by bulk copies).  One 16-warp block per SM; its warps pull groups from a shared counter.
//   table layout: sorted index i lives at phys(i) = (i % R) * 32 + i / R, R = NT / 32 rows, table pitch NT + 1 entries so
//   that the G tables of a warp start one bank apart (phase A stores are conflict free, too).
// =================================================================================================
constexpr int kGrpE = 16;

template <int NT>
struct GrpCfg {
  static constexpr int E = kGrpE;
  static constexpr int LPT = NT / E;    // lanes per transition in phase A
  static constexpr int G = 32 / LPT;    // transitions per warp and round
  static constexpr int R = NT / 32;     // table rows = predicted atoms per lane in phase B
  static constexpr int TP = NT + 1;     // table pitch in entries
  static constexpr int kZY = G * TP;    // floats: staged rows of next_z, later the G sorted tables (two buffers)
  static constexpr int kSc = 8;         // per-transition scalars
  static constexpr int kRedPitch = 36;  // rows of 32 per-lane partial sums, 16-byte aligned, consecutive rows 4 banks apart
  static constexpr int kRed = 3 * G * kRedPitch;  // [G][3] rows: {loss, sum q, sum q^2} of transition t at row 3 t + quantity
  static constexpr int kIn = 8 * G;     // staged per-transition inputs {reward, mask, log_pi, mc_return, grad_scale}[G], two buffers
  static constexpr int kBar = 8;        // three mbarriers (next_z buffer 0 / 1, q_pred) + padding
  // | zy[0] | zy[1] | QT float2[G*TP] | q_pred rows | scalars | in[0] | in[1] | mbarriers | (red) |
  // The per-lane partial sums of transition t are parked when its phase B ends; by then rows 0..t of the q_pred staging are dead, so
  // with at least 3 * kRedPitch predicted atoms per row `red` lives there (rows 3 t .. 3 t + 2 end before row t + 1 begins) and the
  // block holds a fifth warp per scheduler in the shared memory that frees; narrower rows keep a separate `red`.
  static constexpr int kWarpFloatsAlias = 2 * kZY + 2 * kZY + G * NT + kSc * G + 2 * kIn + kBar;
  static constexpr int kWarpFloats = kWarpFloatsAlias + kRed;
  static constexpr int kWarps = NT >= 128 ? 16 : NT >= 64 ? 12 : 8;       // one block per SM (shared memory fills it)
  static constexpr int kWarpsAlias = NT >= 128 ? 20 : NT >= 64 ? 12 : 8;
  static constexpr int kMaxWarps = kWarpsAlias;
  static_assert(kZY % 4 == 0 && kWarpFloats % 4 == 0 && kWarpFloatsAlias % 4 == 0, "16-byte alignment of the staging buffers");
};

// (LS = lane stride of a transition's lanes: sub-lane sl of transition grp is lane sl * LS + grp)
template <int E, int NT, int K, int LS>
__device__ __forceinline__ void grp_flip(float (&e)[E], int sl) {
  if constexpr (K <= E) {
#pragma unroll
    for (int s = 0; s < E; ++s)
      if ((s & (K >> 1)) == 0) {
        const int p = s ^ (K - 1);
        const float x = e[s], y = e[p];
        e[s] = fminf(x, y);
        e[p] = fmaxf(x, y);
      }
  } else {
    constexpr int LM = K / E - 1;
    const bool lower = (sl & (K / 2 / E)) == 0;
    float o[E];
#pragma unroll
    for (int s = 0; s < E; ++s) o[s] = __shfl_xor_sync(kFull, e[E - 1 - s], LM * LS);
#pragma unroll
    for (int s = 0; s < E; ++s) e[s] = lower ? fminf(e[s], o[s]) : fmaxf(e[s], o[s]);
  }
}
template <int E, int J, int LS>
__device__ __forceinline__ void grp_half(float (&e)[E], int sl) {
  if constexpr (J >= E) {
    constexpr int LM = J / E;
    const bool lower = (sl & LM) == 0;
#pragma unroll
    for (int s = 0; s < E; ++s) {
      const float o = __shfl_xor_sync(kFull, e[s], LM * LS);
      e[s] = lower ? fminf(e[s], o) : fmaxf(e[s], o);
    }
  } else {
#pragma unroll
    for (int s = 0; s < E; ++s)
      if ((s & J) == 0) {
        const float x = e[s], y = e[s | J];
        e[s] = fminf(x, y);
        e[s | J] = fmaxf(x, y);
      }
  }
  if constexpr (J > 1) grp_half<E, J / 2, LS>(e, sl);
}
// 60-comparator, 10-layer sorting network for 16 inputs (the bitonic network needs 80 for the same job); checked on all 2^16 0/1
// inputs (tests/test_cabi_and_host.py::test_sort16_network_sorts_every_01_input reads this table)
#define FDQL_SORT16_NETWORK(CE)                                                                             \
  CE(0, 13) CE(1, 12) CE(2, 15) CE(3, 14) CE(4, 8) CE(5, 6) CE(7, 11) CE(9, 10)                           \
  CE(0, 5) CE(1, 7) CE(2, 9) CE(3, 4) CE(6, 13) CE(8, 14) CE(10, 15) CE(11, 12)                           \
  CE(0, 1) CE(2, 3) CE(4, 5) CE(6, 8) CE(7, 9) CE(10, 11) CE(12, 13) CE(14, 15)                           \
  CE(0, 2) CE(1, 3) CE(4, 10) CE(5, 11) CE(6, 7) CE(8, 9) CE(12, 14) CE(13, 15)                           \
  CE(1, 2) CE(3, 12) CE(4, 6) CE(5, 7) CE(8, 10) CE(9, 11) CE(13, 14)                                     \
  CE(1, 4) CE(2, 6) CE(5, 8) CE(7, 10) CE(9, 13) CE(11, 14)                                               \
  CE(2, 4) CE(3, 6) CE(9, 12) CE(11, 13)                                                                  \
  CE(3, 5) CE(6, 8) CE(7, 9) CE(10, 12)                                                                   \
  CE(3, 4) CE(5, 6) CE(7, 8) CE(9, 10) CE(11, 12)                                                         \
  CE(6, 7) CE(8, 9)
__device__ __forceinline__ void grp_sort16(float (&e)[16]) {
#define FDQL_CE(a, b)                  \
  {                                    \
    const float x = e[a], y = e[b];    \
    e[a] = fminf(x, y);                \
    e[b] = fmaxf(x, y);                \
  }
  FDQL_SORT16_NETWORK(FDQL_CE)
#undef FDQL_CE
}

template <int E, int NT, int K, int LS>
__device__ __forceinline__ void grp_sort_from(float (&e)[E], int sl) {
  grp_flip<E, NT, K, LS>(e, sl);
  if constexpr (K >= 4) grp_half<E, K / 4, LS>(e, sl);
  if constexpr (K < NT) grp_sort_from<E, NT, K * 2, LS>(e, sl);
}

// one level of the branch-free search in the pitch-32 layout: lo is a multiple of 2*STEP, probe logical lo + STEP - 1.
// The conditional advance is issued as a predicated IMAD (addr = one * imm + addr with an opaque register holding 1): the
// sort and the compares already load the ALU pipe, IMAD goes down the FMA pipe.
template <int R, int STEP, bool LE>
__device__ __forceinline__ void grp_search_steps(uint32_t& addr, float x, uint32_t one) {
  constexpr int kProbe = STEP >= R ? (R - 1) * 32 + STEP / R - 1 : (STEP - 1) * 32;
  constexpr int kAdvance = STEP >= R ? STEP / R : STEP * 32;
  if constexpr (LE)
    asm volatile("{\n .reg .pred p;\n .reg .f32 v;\n ld.shared.f32 v, [%0+%3];\n setp.le.f32 p, v, %1;\n @p mad.lo.u32 %0, %2, %4, %0;\n}"
                 : "+r"(addr)
                 : "f"(x), "r"(one), "n"(4 * kProbe), "n"(4 * kAdvance));
  else
    asm volatile("{\n .reg .pred p;\n .reg .f32 v;\n ld.shared.f32 v, [%0+%3];\n setp.lt.f32 p, v, %1;\n @p mad.lo.u32 %0, %2, %4, %0;\n}"
                 : "+r"(addr)
                 : "f"(x), "r"(one), "n"(4 * kProbe), "n"(4 * kAdvance));
  if constexpr (STEP > 1) grp_search_steps<R, STEP / 2, LE>(addr, x, one);
}
// the first two levels of a search with their three pivots in registers (every search of a transition probes the same three
// table entries there): compare, select the level-2 pivot, compare; both advances are predicated IMADs like in grp_search_steps
template <int ADV1, int ADV2, bool LE>
__device__ __forceinline__ void grp_search_top2(uint32_t& addr, float x, float piv1, float piv2lo, float piv2hi, uint32_t one) {
  if constexpr (LE)
    asm volatile("{\n .reg .pred p, q;\n .reg .f32 v;\n setp.le.f32 p, %2, %1;\n selp.f32 v, %4, %3, p;\n @p mad.lo.u32 %0, %5, %6, %0;\n"
                 " setp.le.f32 q, v, %1;\n @q mad.lo.u32 %0, %5, %7, %0;\n}"
                 : "+r"(addr)
                 : "f"(x), "f"(piv1), "f"(piv2lo), "f"(piv2hi), "r"(one), "n"(ADV1), "n"(ADV2));
  else
    asm volatile("{\n .reg .pred p, q;\n .reg .f32 v;\n setp.lt.f32 p, %2, %1;\n selp.f32 v, %4, %3, p;\n @p mad.lo.u32 %0, %5, %6, %0;\n"
                 " setp.lt.f32 q, v, %1;\n @q mad.lo.u32 %0, %5, %7, %0;\n}"
                 : "+r"(addr)
                 : "f"(x), "f"(piv1), "f"(piv2lo), "f"(piv2hi), "r"(one), "n"(ADV1), "n"(ADV2));
}
// The same two functions on byte offsets RELATIVE to the table (`base` = the table's shared-memory address, warp-uniform: the probe
// address base + offset + constant goes into the load's [register + uniform register + immediate] form, so the offset costs no
// instruction; the search then starts without a copy of the base and ends without subtracting it again).
template <int R, int STEP, bool LE>
__device__ __forceinline__ void grp_search_steps_rel(uint32_t& off, uint32_t base, float x, uint32_t one) {
  constexpr int kProbe = STEP >= R ? (R - 1) * 32 + STEP / R - 1 : (STEP - 1) * 32;
  constexpr int kAdvance = STEP >= R ? STEP / R : STEP * 32;
  if constexpr (LE)
    asm volatile("{\n .reg .pred p;\n .reg .f32 v;\n .reg .u32 a;\n add.u32 a, %0, %5;\n ld.shared.f32 v, [a+%3];\n setp.le.f32 p, v, %1;\n"
                 " @p mad.lo.u32 %0, %2, %4, %0;\n}"
                 : "+r"(off)
                 : "f"(x), "r"(one), "n"(4 * kProbe), "n"(4 * kAdvance), "r"(base));
  else
    asm volatile("{\n .reg .pred p;\n .reg .f32 v;\n .reg .u32 a;\n add.u32 a, %0, %5;\n ld.shared.f32 v, [a+%3];\n setp.lt.f32 p, v, %1;\n"
                 " @p mad.lo.u32 %0, %2, %4, %0;\n}"
                 : "+r"(off)
                 : "f"(x), "r"(one), "n"(4 * kProbe), "n"(4 * kAdvance), "r"(base));
  if constexpr (STEP > 1) grp_search_steps_rel<R, STEP / 2, LE>(off, base, x, one);
}
template <int ADV1, int ADV2, bool LE>
__device__ __forceinline__ uint32_t grp_search_top2_rel(float x, float piv1, float piv2lo, float piv2hi, uint32_t one) {
  uint32_t off;
  if constexpr (LE)
    asm volatile("{\n .reg .pred p, q;\n .reg .f32 v;\n setp.le.f32 p, %2, %1;\n selp.f32 v, %4, %3, p;\n selp.u32 %0, %6, 0, p;\n"
                 " setp.le.f32 q, v, %1;\n @q mad.lo.u32 %0, %5, %7, %0;\n}"
                 : "=&r"(off)
                 : "f"(x), "f"(piv1), "f"(piv2lo), "f"(piv2hi), "r"(one), "n"(ADV1), "n"(ADV2));
  else
    asm volatile("{\n .reg .pred p, q;\n .reg .f32 v;\n setp.lt.f32 p, %2, %1;\n selp.f32 v, %4, %3, p;\n selp.u32 %0, %6, 0, p;\n"
                 " setp.lt.f32 q, v, %1;\n @q mad.lo.u32 %0, %5, %7, %0;\n}"
                 : "=&r"(off)
                 : "f"(x), "f"(piv1), "f"(piv2lo), "f"(piv2hi), "r"(one), "n"(ADV1), "n"(ADV2));
  return off;
}
// byte offset 4*phys -> sorted index: phys = row * 32 + col, i = col * R + row
// (o = 128 row + 4 col  ->  i = o R / 4 - (32 R - 1) row, written as multiply-high / multiply-add so that it runs down the FMA
// pipe: the ALU pipe is the busier one in this kernel)
template <int R>
__device__ __forceinline__ float grp_count(uint32_t o) {
  uint32_t row, i;
  asm("mul.hi.u32 %0, %1, 33554432;" : "=r"(row) : "r"(o));  // o >> 7
  if constexpr (R == 4) {
    asm("mad.lo.u32 %0, %1, -127, %2;" : "=r"(i) : "r"(row), "r"(o));
  } else {
    uint32_t u;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(u) : "r"(o), "n"(R << 30));  // o R / 4
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(i) : "r"(row), "n"(-(32 * R - 1)), "r"(u));
  }
  return (float)(int)i;
}

// address of the {-P1, P2} entry that belongs to byte offset o of the Y table (8-byte entries), again on the FMA pipe
__device__ __forceinline__ uint32_t grp_qaddr(uint32_t base, uint32_t o) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(r) : "r"(o), "r"(base));
  return r;
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
// ---- 1-D bulk copies (TMA engine) signalled on a per-warp mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n"
      "MBAR_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra MBAR_DONE;\n bra MBAR_WAIT;\n"
      "MBAR_DONE:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ bool grp_elect_one() {  // one lane of the converged warp
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
// rows [m0, m0 + rows) of a [M, width] matrix -> shared memory: one bulk copy issued by lane 0 and signalled on `bar` when the
// run is 16-byte aligned and whole (returns true: the reader waits on the barrier), plain loads otherwise (returns false)
__device__ __forceinline__ bool grp_stage_rows(float* dst, const float* __restrict__ src, int64_t m0, int width, int rows, int full_rows,
                                               bool aligned, int lane, uint32_t bar) {
  const float* g1 = src + m0 * width;
  const int nfl = rows * width;
#ifdef FDQL_TQC_STAGE_LDGSTS
  if (aligned && rows == full_rows && (nfl & 3) == 0) {  // 16-byte asynchronous copies by every lane; the reader waits on the group
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    for (int i = lane; i < (nfl >> 2); i += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * (uint32_t)i), "l"(g1 + 4 * i) : "memory");
    return true;
  }
#endif
  if (aligned && rows == full_rows) {
    if (grp_elect_one()) {
      mbar_expect_tx(bar, 4u * (uint32_t)nfl);
      bulk_g2s((uint32_t)__cvta_generic_to_shared(dst), g1, 4u * (uint32_t)nfl, bar);
    }
    return true;
  }
  for (int i = lane; i < nfl; i += 32) dst[i] = ld_stream1(g1 + i);
  return false;
}

constexpr int kGrpLb = 1, kGrpStats = 2, kGrpFull = 4;  // kernel flavours: lower bound (mc_return given), summaries (stats given),
                                                          // every atom slot but the last one holds 32 real atoms (n_atoms > 32 * (R - 1))

#ifndef FDQL_TQC_BOUND_THREADS
#define FDQL_TQC_BOUND_THREADS 768  // register budget of the group kernel: 65536 / 768 -> 80 per thread (it needs 72)
#endif
// The kernel body as a device function, so that the same code serves the stand-alone kernel (tqc.cu) and the loss role of the fused
// pass kernel (sample.cu: loss warps of batch k + gather warps of batch k+1 in one block).  `wib` / `n_warps`: this warp's index among
// the role's warps and their number; `blk` / `n_blk`: the block's rank among the blocks that share the work; `bar_id`: 0 = the role is
// the whole block (__syncthreads), otherwise the named barrier the role's `n_warps * 32` threads meet on.
__device__ __forceinline__ void role_barrier(int bar_id, int n_threads) {
  if (bar_id == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(n_threads) : "memory");
}
template <int NT, int FLAGS>
__device__ __forceinline__ void tqc_group_body(const TqcArgs& a, float* grp_smem, const int wib, const int n_warps, const int blk,
                                               const int n_blk, const int bar_id) {
  using C = GrpCfg<NT>;
  constexpr int E = C::E, LPT = C::LPT, G = C::G, R = C::R, TP = C::TP;
  constexpr bool LB = (FLAGS & kGrpLb) != 0, STATS = (FLAGS & kGrpStats) != 0, FULL = (FLAGS & kGrpFull) != 0;
  constexpr int NQ = STATS ? 3 : 1;  // per-transition sums reduced over the warp: loss, sum q, sum q^2
  __shared__ double sm_stats[3];
  // (the caller takes the warp index through a shuffle: the compiler then knows that it, the group index and the staging addresses
  // derived from it are warp-uniform, and the bulk copies take their operands from uniform registers without a per-lane
  // uniformisation loop)
  const int lane = lane_id(), tid = wib * 32 + lane;  // thread index within the role
  const bool red_alias = a.grp_red_alias != 0;
  float* W = grp_smem + wib * (red_alias ? C::kWarpFloatsAlias : C::kWarpFloats);
  float2* QT = reinterpret_cast<float2*>(W + 2 * C::kZY);
  float* qs = W + 4 * C::kZY;
  float* sc = qs + G * NT;
  float* inb = sc + C::kSc * G;  // [2][5][G] staged per-transition inputs
  float* red = red_alias ? qs : W + C::kWarpFloatsAlias;  // [G][3][kRedPitch]: per-lane partial sums, one row per (transition, quantity)
  const uint32_t aW = (uint32_t)__cvta_generic_to_shared(W), aQ = aW + 8 * C::kZY;
  const uint32_t aBar = aW + 4 * (uint32_t)(C::kWarpFloatsAlias - C::kBar);  // + 0 / 8: next_z buffers, + 16: q_pred
  if (lane == 0) {
    mbar_init(aBar, 1);
    mbar_init(aBar + 8, 1);
    mbar_init(aBar + 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async_smem();
  __syncwarp();
  const int n = a.n_atoms, nz = a.n_z, K = nz - a.n_drop;
  const uint32_t one = (uint32_t)min(n, 1);  // 1, opaque to the compiler (see grp_search_steps)
  const float Kf = (float)K;
  const float inv_n = 1.f / (float)n;
  const float inv_nk = 1.f / ((float)n * (float)K);
  const float inv_nm1 = n > 1 ? 1.f / (float)(n - 1) : 0.f;
  const float half_over_n = (float)(0.5 / (double)n);
  // a warp that finishes a round takes the block's next unclaimed group, so the warps of an SM end within one round of each
  // other whatever order the scheduler favours them in
  __shared__ int sm_next;
  if (STATS && tid < 3) sm_stats[tid] = 0.0;
  if (tid == 0) sm_next = n_warps;
  role_barrier(bar_id, n_warps * 32);
  double st_sum = 0.0, st_var = 0.0;
  int viol = 0;
  // phase B constants of this lane: :98, tau over the pooled atoms: fl32(fl32(j / n) + fl32(1/2/n)) for j = lane + 32 s
  float taus[R], validf[R];
#pragma unroll
  for (int s = 0; s < R; ++s) {
    taus[s] = __fadd_rn(__fdiv_rn((float)(lane + 32 * s), (float)n), half_over_n);
    validf[s] = lane + 32 * s < n ? 1.f : 0.f;
  }
  // phase A constants
  // the lanes of a transition are G apart (lane = sl * G + grp): a half-warp then holds every transition with half of its
  // sub-lanes, which makes the 64-bit stores of the {-P1, P2} tables conflict free (bank pair = grp + 4 sl + const mod 16)
  const int grp = lane % G, sl = lane / G;
  const int kk = K - sl * E;                  // kept slots of this lane: s < kk
  const int c_src = ((K / 2) / E) * G + grp;  // lane holding a kept target near the median in slot 0
  const uint32_t physK4 = 4u * (uint32_t)((K % R) * 32 + K / R);
  const bool zal = (reinterpret_cast<uintptr_t>(a.next_z) & 15) == 0, qal = (reinterpret_cast<uintptr_t>(a.q_pred) & 15) == 0;
  // per-transition inputs travel one round ahead: lane l < 5*G copies array l / G of transition l % G (4-byte cp.async)
  // (NT = 32 has G = 16: lanes take up to three slots)
  const float alpha = a.alpha_dev ? __ldg(a.alpha_dev) : a.alpha;
  // (the five pointers sit side by side in TqcArgs in this order: one indexed constant load instead of a select chain)
  static_assert(offsetof(TqcArgs, grad_scale) - offsetof(TqcArgs, next_log_pi) == 4 * sizeof(const float*), "input pointer block");
  const float* const* in_ptrs = &a.next_log_pi;  // {next_log_pi, reward, mask, mc_return, grad_scale}
  auto stage_inputs = [&](int b, int64_t m0) {
#pragma unroll
    for (int slot = lane; slot < 5 * G; slot += 32) {
      const int which = slot / G, t = slot % G;
      const float* src = in_ptrs[which];
      if (src != nullptr && m0 + t < a.M) cp_async4(aW + 4 * (uint32_t)((inb - W) + b * C::kIn + which * G + t), src + m0 + t);
    }
  };

  // the c-th claim of block b is group c * gridDim + b: the groups in flight over the whole GPU form one dense moving front
  // (a contiguous run per block made 148 x 3 separate DRAM streams and was 10 % slower at 800K transitions)
  const int64_t n_groups = (a.M + G - 1) / G;
  const int64_t g_begin = blk, g_step = n_blk;
  int* const wctr = a.work_ctr;
  auto claim_group = [&]() -> int64_t {
    int claimed = 0;
    if (wctr != nullptr) {
      if (lane == 0) claimed = atomicAdd(wctr, 1);
      return (int64_t)__shfl_sync(kFull, claimed, 0);
    }
    if (lane == 0) claimed = atomicAdd(&sm_next, 1);
    return g_begin + __shfl_sync(kFull, claimed, 0) * g_step;
  };
  int64_t gi = wctr != nullptr ? claim_group() : g_begin + wib * g_step;
  int buf = 0;
  uint32_t it = 0;  // round counter: the barrier of next_z buffer b completes once per use (parity (it >> 1) & 1), q_pred's every round
  bool z_bulk = false;
  if (gi < n_groups) {
    z_bulk = grp_stage_rows(W, a.next_z, gi * G, nz, (int)min((int64_t)G, a.M - gi * G), G, zal, lane, aBar);
    stage_inputs(0, gi * G);
  }
  cp_async_commit();
  int64_t gnext = 0;
  for (; gi < n_groups; gi = gnext, buf ^= 1, ++it) {
    const int64_t m0 = gi * G;
    const int rows = (int)min((int64_t)G, a.M - m0);
    float* Zb = W + buf * C::kZY;
    const uint32_t aZb = aW + 4 * buf * C::kZY;
    const bool q_bulk = grp_stage_rows(qs, a.q_pred, m0, n, rows, G, qal, lane, aBar + 16);
#ifdef FDQL_TQC_STAGE_LDGSTS
    cp_async_commit();   // (group of this round's q_pred rows)
    cp_async_wait<1>();  // everything older has landed: this round's per-transition inputs and next_z rows
#else
    cp_async_wait<0>();  // this round's per-transition inputs have landed
    if (z_bulk) mbar_wait(aBar + 8 * buf, (it >> 1) & 1);  // ... and its next_z rows
#endif
    __syncwarp();

    // ================= phase A: LPT lanes per transition =================
    {
      const bool live = grp < rows;
      const float* zst = Zb + grp * nz;
      float e[E];
      if constexpr (NT == 128) {
        // lane (grp, sl) reads element 32 (s / 4) + r_k of its row, k = s % 4, r_k = (8 ((grp + k) % 4) - grp nz + sl) mod 32: at every
        // step the four transitions of the warp read four different octants of the 32 banks, whatever nz is (rows of 125 floats
        // start 29 banks apart: with the plain split below every load is a 3-way bank conflict)
        const int r0 = (8 * grp - grp * nz + sl) & 31;
        if (nz >= 96) {  // only the last four slots can fall off the row
#pragma unroll
          for (int s = 0; s < E; ++s) {
            const int j = 32 * (s / 4) + ((r0 + 8 * (s % 4)) & 31);
            e[s] = (s < 12 || j < nz) ? zst[j] : CUDART_INF_F;
          }
        } else {
#pragma unroll
          for (int s = 0; s < E; ++s) {
            const int j = 32 * (s / 4) + ((r0 + 8 * (s % 4)) & 31);
            e[s] = j < nz ? zst[j] : CUDART_INF_F;
          }
        }
      } else if (nz >= NT - LPT) {  // only the last slot can fall off the row
#pragma unroll
        for (int s = 0; s < E - 1; ++s) e[s] = zst[s * LPT + sl];  // any split of the row over the lanes will do: it is sorted next
        e[E - 1] = (E - 1) * LPT + sl < nz ? zst[(E - 1) * LPT + sl] : CUDART_INF_F;
      } else {
#pragma unroll
        for (int s = 0; s < E; ++s) {
          const int j = s * LPT + sl;
          e[s] = j < nz ? zst[j] : CUDART_INF_F;
        }
      }
      if (!live) {  // rows past the end of the batch: harmless finite values, nothing is written for them
#pragma unroll
        for (int s = 0; s < E; ++s) e[s] = 0.f;
      }
      const float* in = inb + buf * C::kIn + grp;
      const float rew = (a.reward && live) ? in[1 * G] : 0.f, msk = (a.mask && live) ? in[2 * G] : 1.f;
      const float ent = (a.next_log_pi && live) ? __fmul_rn(alpha, -in[0 * G]) : 0.f;
      const float Gv = (LB && live) ? in[3 * G] : 0.f;
      const float gs = (a.grad_scale && live) ? in[4 * G] : 1.f;
      const float mg = __fmul_rn(msk, a.gamma);

      static_assert(E == 16, "the in-lane sorter is a 16-input network");
      grp_sort16(e);                    // each lane's 16 values ascending
      grp_sort_from<E, NT, 2 * E, G>(e, sl);  // bitonic merges across the lanes; sorted position of (sl, s) is i = sl * E + s

      // centre: a kept target near the median.  Fetched here, before any per-slot predicate is live: a shuffle inside the loop
      // has an out-of-line non-converged path, and ptxas would pack and unpack every live predicate around it.
      float c0 = __shfl_sync(kFull, e[0], c_src);
      int kkl = kk;
      asm volatile("" : "+r"(kkl));  // keeps the `s < kk` compares below the shuffle

      // soft target (:50-58); raw mode (quantile_huber_loss_f): targets as given.  Then cut the top n_drop (+inf), centre, this lane's
      // sums of y and y^2.  Entries at sorted index >= K hold +inf and make every later prefix non-finite; no search ever lands past K.
      float l1 = 0.f, l2 = 0.f;
#ifdef FDQL_TQC_NO_FOLD  // (A/B builds)
      if (true) {
#else
      if (a.td_target != nullptr) {  // the targets themselves are wanted: the reference's operator order, bit for bit
#endif
        if (a.reward) {
          if (a.next_log_pi) {
#pragma unroll
            for (int s = 0; s < E; ++s) e[s] = __fadd_rn(rew, __fmul_rn(mg, __fadd_rn(e[s], ent)));
            c0 = __fadd_rn(rew, __fmul_rn(mg, __fadd_rn(c0, ent)));
          } else {
#pragma unroll
            for (int s = 0; s < E; ++s) e[s] = __fadd_rn(rew, __fmul_rn(mg, e[s]));
            c0 = __fadd_rn(rew, __fmul_rn(mg, c0));
          }
        }
        if (a.td_target != nullptr && live) {
#pragma unroll
          for (int s = 0; s < E; ++s)
            if (s < kkl) a.td_target[(m0 + grp) * K + sl * E + s] = e[s];
        }
#pragma unroll
        for (int s = 0; s < E; ++s) {
          const float y = s < kkl ? e[s] - c0 : CUDART_INF_F;
          e[s] = y;
          l1 += y;
          l2 = fmaf(y, y, l2);
        }
      } else {
        // only the loss is wanted: the centred target directly, y_k - y_c = mask gamma (z_k - z_c) -- the reward and the entropy
        // term cancel, two operations per atom instead of four, and no cancellation between large offsets
        const float zc = c0, scale = a.reward ? mg : 1.f;
        if (a.reward) c0 = a.next_log_pi ? __fadd_rn(rew, __fmul_rn(mg, __fadd_rn(c0, ent))) : __fadd_rn(rew, __fmul_rn(mg, c0));
#pragma unroll
        for (int s = 0; s < E; ++s) {
          const float y = s < kkl ? (e[s] - zc) * scale : CUDART_INF_F;
          e[s] = y;
          l1 += y;
          l2 = fmaf(y, y, l2);
        }
      }
      float x1 = l1, x2 = l2;  // inclusive scan over the LPT lanes of the transition
#pragma unroll
      for (int d = 1; d < LPT; d <<= 1) {
        const float o1 = __shfl_up_sync(kFull, x1, d * G), o2 = __shfl_up_sync(kFull, x2, d * G);
        if (sl >= d) {
          x1 += o1;
          x2 += o2;
        }
      }
      float p1 = __shfl_up_sync(kFull, x1, G), p2 = __shfl_up_sync(kFull, x2, G);  // exclusive
      if (sl == 0) p1 = p2 = 0.f;
      p1 = -p1;
      __syncwarp();  // every lane has read its staged row: the buffer becomes the tables
      float* Yt = Zb + grp * TP + sl * (E / R);
      float2* Qt = QT + grp * TP + sl * (E / R);
#pragma unroll
      for (int s = 0; s < E; ++s) {
        const int ph = (s % R) * 32 + s / R;
        Yt[ph] = e[s];
        Qt[ph] = make_float2(p1, p2);  // {-P1_i, P2_i}
        p1 -= e[s];
        p2 = fmaf(e[s], e[s], p2);
      }
      if (sl == 0) {
        sc[grp * C::kSc + 0] = c0;
        sc[grp * C::kSc + 1] = Gv - c0;
        sc[grp * C::kSc + 2] = gs;
      }
    }
    fence_proxy_async_smem();  // the tables were written through the generic proxy, the bulk copy below overwrites the older ones
    __syncwarp();
    {  // next round's next_z rows into the other buffer (its tables are dead)
      gnext = claim_group();
      z_bulk = false;
      if (gnext < n_groups) {
        z_bulk = grp_stage_rows(W + (buf ^ 1) * C::kZY, a.next_z, gnext * G, nz, (int)min((int64_t)G, a.M - gnext * G), G, zal, lane,
                                aBar + 8 * (buf ^ 1));
        stage_inputs(buf ^ 1, gnext * G);
      }
      cp_async_commit();
    }
#ifdef FDQL_TQC_STAGE_LDGSTS
    cp_async_wait<1>();  // all but the group committed just above: this round's q_pred rows have landed
#else
    if (q_bulk) mbar_wait(aBar + 16, it & 1);  // this round's q_pred rows have landed
#endif
    __syncwarp();

    // ================= phase B: the warp per transition =================
    const bool has_grad = a.grad_q != nullptr;
    float* __restrict__ grow_ = a.grad_q + m0 * n + lane;  // advanced by one row per transition; only dereferenced under has_grad
    for (int t = 0; t < rows; ++t, grow_ += n) {
      const float c0 = sc[t * C::kSc + 0], gs = sc[t * C::kSc + 2];
      const float Gc = LB ? sc[t * C::kSc + 1] : 0.f;
      const uint32_t aYt = aZb + 4 * t * TP, aQt = aQ + 8 * t * TP;
      const float nT1 = lds_f32(aQt + 2 * physK4);  // -(sum of the kept centred targets)
      // level-1 and level-2 pivots (logical indices NT/2-1, NT/4-1, 3NT/4-1) and the byte advances of these two levels
      constexpr int kAdv1 = 4 * (NT / 2 >= R ? NT / 2 / R : NT / 2 * 32), kAdv2 = 4 * (NT / 4 >= R ? NT / 4 / R : NT / 4 * 32);
      constexpr int kPrb1 = 4 * (NT / 2 >= R ? (R - 1) * 32 + NT / 2 / R - 1 : (NT / 2 - 1) * 32);
      constexpr int kPrb2 = 4 * (NT / 4 >= R ? (R - 1) * 32 + NT / 4 / R - 1 : (NT / 4 - 1) * 32);
      float piv1 = 0.f, piv2lo = 0.f, piv2hi = 0.f;
      if constexpr (NT >= 128) {
        piv1 = lds_f32(aYt + kPrb1);
        piv2lo = lds_f32(aYt + kPrb2);
        piv2hi = lds_f32(aYt + kAdv1 + kPrb2);
      }
      const float gscale = inv_nk * gs, glb = -inv_n * gs;
      const float* qrow = qs + t * n + lane;
      float acc = 0.f, lbacc = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const bool ok = (FULL && s < R - 1) ? true : validf[s] != 0.f;
        const float qc = ok ? qrow[32 * s] - c0 : 0.f;
        uint32_t oa = aYt, ob = aYt, oc = aYt;
#ifndef FDQL_TQC_ABS_SEARCH
        if constexpr (NT >= 128) {  // (offsets relative to the table: see grp_search_steps_rel)
          const float xa = qc - 1.f, xc = qc + 1.f;
          oa = grp_search_top2_rel<kAdv1, kAdv2, false>(xa, piv1, piv2lo, piv2hi, one);
          ob = grp_search_top2_rel<kAdv1, kAdv2, false>(qc, piv1, piv2lo, piv2hi, one);
          oc = grp_search_top2_rel<kAdv1, kAdv2, true>(xc, piv1, piv2lo, piv2hi, one);
          grp_search_steps_rel<R, NT / 8, false>(oa, aYt, xa, one);  // a = #(y < q-1)
          grp_search_steps_rel<R, NT / 8, false>(ob, aYt, qc, one);  // b = #(y < q)
          grp_search_steps_rel<R, NT / 8, true>(oc, aYt, xc, one);   // c = #(y <= q+1)
          oa += aYt;
          ob += aYt;
          oc += aYt;  // (cancels against the subtraction below at compile time)
        } else
#endif
        if constexpr (NT >= 128) {  // levels 1-2 from the three pivots in registers, levels 3.. from the table
          const float xa = qc - 1.f, xc = qc + 1.f;
          grp_search_top2<kAdv1, kAdv2, false>(oa, xa, piv1, piv2lo, piv2hi, one);
          grp_search_top2<kAdv1, kAdv2, false>(ob, qc, piv1, piv2lo, piv2hi, one);
          grp_search_top2<kAdv1, kAdv2, true>(oc, xc, piv1, piv2lo, piv2hi, one);
          grp_search_steps<R, NT / 8, false>(oa, xa, one);  // a = #(y < q-1)
          grp_search_steps<R, NT / 8, false>(ob, qc, one);  // b = #(y < q)
          grp_search_steps<R, NT / 8, true>(oc, xc, one);   // c = #(y <= q+1)
        } else {
          grp_search_steps<R, NT / 2, false>(oa, qc - 1.f, one);
          grp_search_steps<R, NT / 2, false>(ob, qc, one);
          grp_search_steps<R, NT / 2, true>(oc, qc + 1.f, one);
        }
        oa -= aYt;
        ob -= aYt;
        oc -= aYt;
        const float2 Qa = lds2_at(grp_qaddr(aQt, oa)), Qb = lds2_at(grp_qaddr(aQt, ob)), Qc = lds2_at(grp_qaddr(aQt, oc));
        const float ia = grp_count<R>(oa), ib = grp_count<R>(ob), ic = grp_count<R>(oc);
        // L_i(q) = sum_{k<i} (q - y_k),  F_i(q) = sum_{k<i} (y_k - q)^2   (table: x = -P1_i, y = P2_i)
        const float La = fmaf(ia, qc, Qa.x), Lb = fmaf(ib, qc, Qb.x), Lc = fmaf(ic, qc, Qc.x);
        const float Fa = fmaf(qc, La + Qa.x, Qa.y), Fb = fmaf(qc, Lb + Qb.x, Qb.y), Fc = fmaf(qc, Lc + Qc.x, Qc.y);
        const float LK = fmaf(Kf, qc, nT1);
        const float kc = Kf - ic;
        const float neg = fmaf(0.5f, Fb - Fa, fmaf(-0.5f, ia, La));       // delta < 0 : weight 1 - tau
        const float pos = fmaf(0.5f, Fc - Fb, fmaf(-0.5f, kc, Lc - LK));  // delta >= 0: weight tau
        const float lj = fmaf(taus[s], pos - neg, neg);  // (1 - tau) neg + tau pos, without a register for 1 - tau
        const float gneg = ia + (Lb - La);
        const float gpos = (Lc - Lb) - kc;
        const float gj = fmaf(taus[s], gpos - gneg, gneg);
        float gl = 0.f;
        if constexpr (LB) {  // :76-79 lower bound relu(mc_return - q)
          const float lbj = fmaxf(Gc - qc, 0.f);
          const bool on = lbj > 0.f;
          gl = on ? glb : 0.f;
          lbacc = (FULL && s < R - 1) ? lbacc + lbj : fmaf(lbj, validf[s], lbacc);
          if constexpr (STATS) viol += (on && ok) ? 1 : 0;
        }
        acc = (FULL && s < R - 1) ? acc + lj : fmaf(lj, validf[s], acc);
        if (has_grad && ok) st_stream1(grow_ + 32 * s, fmaf(gj, gscale, gl));
        if constexpr (STATS) {
          s1 += qc;  // padded slots hold 0
          s2 = fmaf(qc, qc, s2);
        }
      }
      __syncwarp();  // every lane has read row t of the q_pred staging (aliased layout: `red` overwrites its head)
      red[(3 * t + 0) * C::kRedPitch + lane] = fmaf(acc, inv_nk, lbacc * inv_n);
      if constexpr (STATS) {
        red[(3 * t + 1) * C::kRedPitch + lane] = s1;
        red[(3 * t + 2) * C::kRedPitch + lane] = s2;
      }
    }
    __syncwarp();
    // per-transition sums over the lanes: the 32 / G lanes of transition t each add G columns of its rows (128-bit loads), then a
    // log2(32 / G)-step butterfly; lane t * (32 / G) ends up with the loss, sum q and sum q^2 of transition t
    {
      constexpr int LB_ = 32 / G;  // lanes per transition
      const int t = lane / LB_, part = lane % LB_;
      float sums[NQ];
#pragma unroll
      for (int qn = 0; qn < NQ; ++qn) {
        const float4* rrow = reinterpret_cast<const float4*>(red + (3 * t + qn) * C::kRedPitch + part * G);
        float sacc = 0.f;
#pragma unroll
        for (int k = 0; k < G / 4; ++k) {
          const float4 v = rrow[k];
          sacc += (v.x + v.y) + (v.z + v.w);
        }
#pragma unroll
        for (int d = LB_ / 2; d >= 1; d >>= 1) sacc += __shfl_xor_sync(kFull, sacc, d);
        sums[qn] = sacc;
      }
      if (part == 0 && t < rows) {
        if (a.loss) a.loss[m0 + t] = sums[0];
        if constexpr (STATS) {  // :66-67,80-82 q_pred mean, mean row variance (unbiased) from the centred moments
          const double S1 = (double)sums[1], S2 = (double)sums[2];
          st_sum += S1 + (double)n * (double)sc[t * C::kSc + 0];
          st_var += (S2 - S1 * S1 * (double)inv_n) * (double)inv_nm1;
        }
      }
    }
    __syncwarp();  // the q_pred staging, the scalars and `red` are rewritten in the next round
  }
  cp_async_wait<0>();
  if (wctr != nullptr) {  // every warp of this block has made its last claim: the last block re-arms the counter for the next launch
    role_barrier(bar_id, n_warps * 32);
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(wctr + 1, 1) == n_blk - 1) {
        wctr[0] = 0;
        wctr[1] = 0;
        __threadfence();
      }
    }
  }
  if constexpr (STATS) {
    const int vsum = __reduce_add_sync(kFull, viol);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      st_sum += shfl_xor_f64(st_sum, d);
      st_var += shfl_xor_f64(st_var, d);
    }
    if (lane == 0) {
      atomicAdd(&sm_stats[0], st_sum);
      atomicAdd(&sm_stats[1], st_var);
      atomicAdd(&sm_stats[2], (double)vsum);
    }
    role_barrier(bar_id, n_warps * 32);
    if (tid < 3) atomicAdd(a.stats + tid, sm_stats[tid]);
    if (tid == 3 && blk == 0) atomicAdd(a.stats + 3, (double)a.M);
  }
}

// (tqc.cu) picks the kernel for the shape and launches it
int launch_tqc(const TqcArgs& a, cudaStream_t st);

}  // namespace fdql
