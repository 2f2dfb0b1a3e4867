// TQC target + quantile-Huber loss, forward and backward in one pass.
//   DistributionalSoftActorCritic.q_loss   franQ/Agent/components/distributional_soft_actor_critic.py:50-58,66-67,70,76-82
//   quantile_huber_loss_f                  franQ/Agent/components/distributional_soft_actor_critic.py:90-103
//   SoftActorCritic.q_loss                 franQ/Agent/components/soft_actor_critic.py:63-99,134
//
// The reference materialises the [n_atoms, K] pairwise tensor (14,375 pair evaluations per transition at 5x25 atoms).
// Here the pooled target atoms are sorted in registers by a bitonic network, the top n_drop are cut, and because the K kept
// targets y are sorted, the sum over k for one predicted atom q splits at a = #(y < q-1), b = #(y < q), c = #(y <= q+1) into
// pieces that are linear in prefix sums of y and y^2.  y and q are shifted by a per-transition centre
// first so the fp32 prefix sums do not cancel.
// Two kernels: tqc_loss_group_kernel (up to 128 atoms: sub-warp sort, warp-wide search; the one bench.py times) and
// tqc_loss_kernel (one warp per transition, kept for 129..256 atoms and as a cross-check behind fdql_debug_tqc_warp_kernel).
#include <math_constants.h>

#include "common.cuh"
#include "tqc_group.cuh"

namespace fdql {


// ---- warp-level bitonic sort of 32*VPL values; sorted position of (lane, slot) is i = lane*VPL + slot ----------------
// "Flip" formulation: phase K first compares i with i ^ (K-1), then i with i ^ J for J = K/4 ... 1, and every
// compare-exchange is ascending (the lower index keeps the minimum).  So an in-lane stage needs no predicate at all and a
// cross-lane stage needs one (am I the lower lane of the pair).
template <int VPL, int K>
__device__ __forceinline__ void bitonic_flip(float (&e)[VPL], int lane) {
  if constexpr (K <= VPL) {
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
      if ((s & (K >> 1)) == 0) {
        const int p = s ^ (K - 1);
        const float x = e[s], y = e[p];
        e[s] = fminf(x, y);
        e[p] = fmaxf(x, y);
      }
    }
  } else {
    constexpr int LM = K / VPL - 1;  // partner lane = lane ^ LM, partner slot = VPL-1-s
    const bool lower = (lane & (K / 2 / VPL)) == 0;
    float o[VPL];
#pragma unroll
    for (int s = 0; s < VPL; ++s) o[s] = __shfl_xor_sync(kFull, e[VPL - 1 - s], LM);
#pragma unroll
    for (int s = 0; s < VPL; ++s) e[s] = lower ? fminf(e[s], o[s]) : fmaxf(e[s], o[s]);
  }
}
template <int VPL, int J>
__device__ __forceinline__ void bitonic_half(float (&e)[VPL], int lane) {
  if constexpr (J >= VPL) {
    constexpr int LM = J / VPL;
    const bool lower = (lane & LM) == 0;
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
      const float o = __shfl_xor_sync(kFull, e[s], LM);
      e[s] = lower ? fminf(e[s], o) : fmaxf(e[s], o);
    }
  } else {
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
      if ((s & J) == 0) {
        const float x = e[s], y = e[s | J];
        e[s] = fminf(x, y);
        e[s | J] = fmaxf(x, y);
      }
    }
  }
  if constexpr (J > 1) bitonic_half<VPL, J / 2>(e, lane);
}
template <int VPL, int K>
__device__ __forceinline__ void bitonic_sort_from(float (&e)[VPL], int lane) {
  bitonic_flip<VPL, K>(e, lane);
  if constexpr (K >= 4) bitonic_half<VPL, K / 4>(e, lane);
  if constexpr (K < 32 * VPL) bitonic_sort_from<VPL, K * 2>(e, lane);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  return v;
}

// count of sorted[0..N) strictly below x (LE=false) or <= x (LE=true); entries past the kept range hold +inf
// `base` is the shared-window byte address of sorted[0]; the result is the byte offset 4*count, so that one step is
// LDS [addr + imm], FSETP, predicated IADD and the prefix-sum tables are indexed by adding the same offset.
// Bank-conflict-free layout: sorted index i = lane*VPL + slot lives at phys(i) = slot*33 + lane (VPL rows of 32 columns with
// a pitch of 33).  Stores of one slot by 32 lanes are consecutive, and the probes of one search level -- logical indices
// that differ by multiples of 2*STEP -- fall into distinct banks down to the last level.  Index 32*VPL (the "count = all"
// entry of the prefix tables) lands in the pad column of row 0.
template <int VPL>
__host__ __device__ constexpr int phys(int i) { return (i % VPL) * 33 + i / VPL; }
template <int VPL, int STEP, bool LE>
__device__ __forceinline__ void search_steps(uint32_t& addr, float x) {
  // lo is a multiple of 2*STEP; probe logical lo + STEP - 1, then advance lo by STEP
  constexpr int kProbe = STEP >= VPL ? (VPL - 1) * 33 + STEP / VPL - 1 : (STEP - 1) * 33;
  constexpr int kAdvance = STEP >= VPL ? STEP / VPL : STEP * 33;
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(4 * kProbe));
  const bool take = LE ? (v <= x) : (v < x);
  if (take) addr += 4 * kAdvance;
  if constexpr (STEP > 1) search_steps<VPL, STEP / 2, LE>(addr, x);
}
// returns 4 * phys(count)
template <int VPL, bool LE>
__device__ __forceinline__ uint32_t count_below_bytes(uint32_t base, float x) {
  uint32_t addr = base;
  search_steps<VPL, 16 * VPL, LE>(addr, x);
  return addr - base;
}
template <int VPL>
__device__ __forceinline__ float count_from_offset(uint32_t o) {
  const uint32_t slot = (o * 993u) >> 17;
  return (float)(int)((VPL * o >> 2) - (33u * VPL - 1u) * slot);
}

constexpr int kTqcWarps = 8;

template <int VPL>
__global__ void __launch_bounds__(kTqcWarps * 32, VPL <= 4 ? 3 : 2) tqc_loss_kernel(const __grid_constant__ TqcArgs a) {
  constexpr int N = 32 * VPL;
  // per warp: sorted centred targets Y[N], exclusive prefix sums P1[N+1], P2[N+1]  (+pad to dodge bank aliasing)
  constexpr int kTab = 33 * VPL + (VPL & 1);  // phys() table of N+1 entries, even length
  constexpr int kStride = 3 * kTab;           // Y | Q = float2 {prefix sum of y, prefix sum of y^2}
  __shared__ __align__(16) float sm[kTqcWarps * kStride];
  __shared__ double sm_stats[3];
  const int lane = lane_id(), wib = threadIdx.x >> 5;
  float* Y = sm + wib * kStride;
  float2* Qt = reinterpret_cast<float2*>(Y + kTab);
  const uint32_t aY = (uint32_t)__cvta_generic_to_shared(Y), aQ = aY + 4 * kTab;
  const int n = a.n_atoms, nz = a.n_z, K = nz - a.n_drop;
  const float inv_n = 1.f / (float)n;
  const float inv_nk = 1.f / ((float)n * (float)K);
  const float half_over_n = (float)(0.5 / (double)n);
  if (a.stats && threadIdx.x < 3) sm_stats[threadIdx.x] = 0.0;
  if (a.stats) __syncthreads();
  double st_sum = 0.0, st_var = 0.0, st_viol = 0.0;
  float taus[VPL];  // :98, tau over the pooled atoms: fl32(fl32(j / n) + fl32(1/2/n)), for this lane's atoms j = lane + 32 s
#pragma unroll
  for (int s = 0; s < VPL; ++s) taus[s] = __fadd_rn(__fdiv_rn((float)(lane + 32 * s), (float)n), half_over_n);
  const float inv_nm1 = n > 1 ? 1.f / (float)(n - 1) : 0.f;

  const int64_t nwarps = (int64_t)gridDim.x * kTqcWarps;
  for (int64_t m = (int64_t)blockIdx.x * kTqcWarps + wib; m < a.M; m += nwarps) {
    // ---- load: target atoms (coalesced), predicted atoms, per-transition scalars ---------------
    const float* __restrict__ zrow = a.next_z + m * nz;
    const float* __restrict__ qrow = a.q_pred + m * n;
    float e[VPL], q[VPL];
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
      const int j = lane + 32 * s;
      e[s] = j < nz ? ld_stream1(zrow + j) : CUDART_INF_F;
      q[s] = j < n ? ld_stream1(qrow + j) : 0.f;
    }
    const float rew = a.reward ? __ldg(a.reward + m) : 0.f, msk = a.mask ? __ldg(a.mask + m) : 1.f;
    const float alpha = a.alpha_dev ? __ldg(a.alpha_dev) : a.alpha;
    const float ent = a.next_log_pi ? __fmul_rn(alpha, -__ldg(a.next_log_pi + m)) : 0.f;
    const float G = a.mc_return ? __ldg(a.mc_return + m) : 0.f;
    const float gs = a.grad_scale ? __ldg(a.grad_scale + m) : 1.f;
    const float mg = __fmul_rn(msk, a.gamma);

    // ---- sort ascending, cut the top n_drop, soft target (:50-58) -------------------------------
    bitonic_sort_from<VPL, 2>(e, lane);
    float y[VPL];
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
      float z = e[s];
      if (a.next_log_pi) z = __fadd_rn(z, ent);
      y[s] = a.reward ? __fadd_rn(rew, __fmul_rn(mg, z)) : z;  // raw mode (quantile_huber_loss_f): targets as given
    }
    if (a.td_target) {
#pragma unroll
      for (int s = 0; s < VPL; ++s) {
        const int i = lane * VPL + s;
        if (i < K) a.td_target[m * K + i] = y[s];
      }
    }
    // centre: a kept target near the median
    const float c0 = __shfl_sync(kFull, y[0], (K / 2) / VPL);
    float l1 = 0.f, l2 = 0.f;  // this lane's sums of centred y, y^2 over its kept slots
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
      const int i = lane * VPL + s;
      y[s] = i < K ? y[s] - c0 : CUDART_INF_F;
      if (i < K) {
        l1 += y[s];
        l2 = fmaf(y[s], y[s], l2);
      }
    }
    float x1 = l1, x2 = l2;  // inclusive scan over lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float o1 = __shfl_up_sync(kFull, x1, d), o2 = __shfl_up_sync(kFull, x2, d);
      if (lane >= d) {
        x1 += o1;
        x2 += o2;
      }
    }
    float p1 = __shfl_up_sync(kFull, x1, 1), p2 = __shfl_up_sync(kFull, x2, 1);  // exclusive
    if (lane == 0) p1 = p2 = 0.f;
    __syncwarp();
    {
      float pa[VPL], pb[VPL];
#pragma unroll
      for (int s = 0; s < VPL; ++s) {
        const int i = lane * VPL + s;
        pa[s] = p1;
        pb[s] = p2;
        if (i < K) {
          p1 += y[s];
          p2 = fmaf(y[s], y[s], p2);
        }
      }
#pragma unroll
      for (int s = 0; s < VPL; ++s) {
        Y[s * 33 + lane] = y[s];
        Qt[s * 33 + lane] = make_float2(pa[s], pb[s]);
      }
      if (lane == 31) Qt[phys<VPL>(N)] = make_float2(p1, p2);
    }
    __syncwarp();
    const float T1 = Qt[phys<VPL>(K)].x;

    // ---- per predicted atom: loss and gradient from the three split points ----------------------
    float acc = 0.f;  // this lane's share of the per-transition loss
    float qsum = 0.f;
    int viol = 0;
    float gout[VPL];
#pragma unroll
    for (int s = 0; s < VPL; ++s) {
      const int j = lane + 32 * s;
      const float qc = q[s] - c0;
      const uint32_t oa = count_below_bytes<VPL, false>(aY, qc - 1.f);
      const uint32_t ob = count_below_bytes<VPL, false>(aY, qc);
      const uint32_t oc = count_below_bytes<VPL, true>(aY, qc + 1.f);
      const float2 Qa = lds2_at(aQ + 2 * oa), Qb = lds2_at(aQ + 2 * ob), Qc = lds2_at(aQ + 2 * oc);
      const float P1a = Qa.x, P1b = Qb.x, P1c = Qc.x, P2a = Qa.y, P2b = Qb.y, P2c = Qc.y;
      // byte offset o = 4*phys, phys = slot*33 + lane  ->  count = lane*VPL + slot = VPL*phys - (33*VPL-1)*slot with
      // slot = phys/33 = (o*993) >> 17 (exact for phys < 2^13)
      const float fa = count_from_offset<VPL>(oa), fb = count_from_offset<VPL>(ob), fc = count_from_offset<VPL>(oc);
      const float na = fa, nab = fb - fa, nbc = fc - fb, nc = (float)K - fc;
      const float tau = taus[s];
      const float d1ab = P1b - P1a, d1bc = P1c - P1b;
      // sum over a<=k<b of (y-q)^2 = dP2 - 2q dP1 + n q^2, same for b<=k<c
      const float sqab = fmaf(qc, fmaf(qc, nab, -2.f * d1ab), P2b - P2a);
      const float sqbc = fmaf(qc, fmaf(qc, nbc, -2.f * d1bc), P2c - P2b);
      const float neg = fmaf(na, qc - 0.5f, -P1a) + 0.5f * sqab;               // delta < 0 : weight 1 - tau
      const float pos = 0.5f * sqbc + ((T1 - P1c) - nc * (qc + 0.5f));         // delta >= 0: weight tau
      const float lj = fmaf(1.f - tau, neg, tau * pos);
      // d/dq of the same sum
      const float gneg = na + fmaf(nab, qc, -d1ab);
      const float gpos = fmaf(nbc, qc, -d1bc) - nc;
      float gj = fmaf(1.f - tau, gneg, tau * gpos) * inv_nk;
      float lbj = 0.f;
      if (a.mc_return) {  // :76-79 lower bound relu(mc_return - q)
        lbj = fmaxf(G - q[s], 0.f);
        const bool on = lbj > 0.f;
        gj -= on ? inv_n : 0.f;
        viol += (on && j < n) ? 1 : 0;
      }
      const float contrib = fmaf(lj, inv_nk, lbj * inv_n);
      acc += j < n ? contrib : 0.f;
      qsum += q[s];  // padded slots hold 0
      gout[s] = gj * gs;
    }
    if (a.grad_q) {
      float* __restrict__ grow_ = a.grad_q + m * n;
#pragma unroll
      for (int s = 0; s < VPL; ++s) {
        const int j = lane + 32 * s;
        if (j < n) st_stream1(grow_ + j, gout[s]);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0 && a.loss) a.loss[m] = acc;
    if (a.stats) {  // :66-67,80-82 q_pred mean, mean row variance (unbiased), constraint violations
      const float mean = warp_sum(qsum) * inv_n;
      float dv = 0.f;
#pragma unroll
      for (int s = 0; s < VPL; ++s) {
        const int j = lane + 32 * s;
        const float d = q[s] - mean;
        if (j < n) dv = fmaf(d, d, dv);
      }
      dv = warp_sum(dv);
      const int vsum = __reduce_add_sync(kFull, viol);
      if (lane == 0) {
        st_sum += (double)mean * n;
        st_var += (double)(dv * inv_nm1);
        st_viol += (double)vsum;
      }
    }
    __syncwarp();
  }
  if (a.stats) {
    if (lane == 0) {
      atomicAdd(&sm_stats[0], st_sum);
      atomicAdd(&sm_stats[1], st_var);
      atomicAdd(&sm_stats[2], st_viol);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(a.stats + threadIdx.x, sm_stats[threadIdx.x]);
    if (threadIdx.x == 3 && blockIdx.x == 0) atomicAdd(a.stats + 3, (double)a.M);
  }
}

// the stand-alone kernel: the whole block is the loss role
template <int NT, int FLAGS>
__global__ void __launch_bounds__(FDQL_TQC_BOUND_THREADS, 1) tqc_loss_group_kernel(const __grid_constant__ TqcArgs a) {
  extern __shared__ __align__(16) float grp_smem_dyn[];
  const int wib = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
  tqc_group_body<NT, FLAGS>(a, grp_smem_dyn, wib, (int)(blockDim.x >> 5), (int)blockIdx.x, (int)gridDim.x, 0);
}

// ---- non-distributional variant: min over atoms, smooth-L1, lower bound replaces the TD term where active ----
struct SacArgs {
  int64_t M;
  int32_t n_atoms;
  const float* target_z;
  const float* q_pred;
  const float* next_log_pi;
  const float* reward;
  const float* mask;
  const float* mc_return;
  const float* grad_scale;
  float alpha, gamma;
  float* loss;
  float* grad_q;
  double* stats;
  const float* alpha_dev;  // when set, overrides alpha
};

__global__ void __launch_bounds__(256) sac_min_target_kernel(const __grid_constant__ SacArgs a) {
  __shared__ double sm_stats[3];
  const int lane = lane_id(), wib = threadIdx.x >> 5;
  const int n = a.n_atoms;
  const float inv_n = 1.f / (float)n;
  if (a.stats && threadIdx.x < 3) sm_stats[threadIdx.x] = 0.0;
  if (a.stats) __syncthreads();
  double st_sum = 0.0, st_var = 0.0, st_viol = 0.0;
  const float alpha = a.alpha_dev ? __ldg(a.alpha_dev) : a.alpha;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t m = (int64_t)blockIdx.x * 8 + wib; m < a.M; m += nwarps) {
    const float* __restrict__ zrow = a.target_z + m * n;
    const float* __restrict__ qrow = a.q_pred + m * n;
    const float ent = a.next_log_pi ? __fmul_rn(alpha, -__ldg(a.next_log_pi + m)) : 0.f;
    float zmin = CUDART_INF_F;
    for (int j = lane; j < n; j += 32) {
      float z = ld_stream1(zrow + j);
      if (a.next_log_pi) z = __fadd_rn(z, ent);
      zmin = fminf(zmin, z);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) zmin = fminf(zmin, __shfl_xor_sync(kFull, zmin, d));
    const float td = __fadd_rn(__ldg(a.reward + m), __fmul_rn(__fmul_rn(__ldg(a.mask + m), a.gamma), zmin));
    const float G = a.mc_return ? __ldg(a.mc_return + m) : 0.f;
    const float gs = a.grad_scale ? __ldg(a.grad_scale + m) : 1.f;
    float acc = 0.f, qsum = 0.f, qsq = 0.f;
    int viol = 0;
    for (int j = lane; j < n; j += 32) {
      const float q = ld_stream1(qrow + j);
      const float d = q - td, ad = fabsf(d);
      float l = ad < 1.f ? 0.5f * d * d : ad - 0.5f;  // F.smooth_l1_loss, beta = 1
      float gq = fminf(fmaxf(d, -1.f), 1.f);
      if (a.mc_return) {  // soft_actor_critic.py:93-97: q_loss = q_loss * (lb == 0) + lb
        const float lb = fmaxf(G - q, 0.f);
        if (lb != 0.f) {
          l = lb;
          gq = -1.f;
          ++viol;
        }
      }
      acc += l;
      qsum += q;
      qsq = fmaf(q, q, qsq);
      if (a.grad_q) st_stream1(a.grad_q + m * n + j, gq * inv_n * gs);
    }
    acc = warp_sum(acc);
    if (lane == 0 && a.loss) a.loss[m] = acc * inv_n;
    if (a.stats) {
      const float mean = warp_sum(qsum) * inv_n;
      float dv = 0.f;
      for (int j = lane; j < n; j += 32) {
        const float d = __ldg(qrow + j) - mean;
        dv = fmaf(d, d, dv);
      }
      dv = warp_sum(dv);
      const int vsum = __reduce_add_sync(kFull, viol);
      if (lane == 0) {
        st_sum += (double)mean * n;
        st_var += n > 1 ? (double)dv / (double)(n - 1) : 0.0;
        st_viol += (double)vsum;
      }
    }
  }
  if (a.stats) {
    if (lane == 0) {
      atomicAdd(&sm_stats[0], st_sum);
      atomicAdd(&sm_stats[1], st_var);
      atomicAdd(&sm_stats[2], st_viol);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(a.stats + threadIdx.x, sm_stats[threadIdx.x]);
    if (threadIdx.x == 3 && blockIdx.x == 0) atomicAdd(a.stats + 3, (double)a.M);
  }
}

// the same for narrow heads (n_atoms <= 16, e.g. 5 critics x 1 value): one THREAD per transition -- a warp per transition would
// leave most lanes idle and the launch latency-bound; neighbouring threads read neighbouring rows, so the loads still coalesce
__global__ void __launch_bounds__(256) sac_min_target_thread_kernel(const __grid_constant__ SacArgs a) {
  __shared__ double sm_stats[3];
  const int n = a.n_atoms;
  const float inv_n = 1.f / (float)n;
  if (a.stats && threadIdx.x < 3) sm_stats[threadIdx.x] = 0.0;
  if (a.stats) __syncthreads();
  const float alpha = a.alpha_dev ? __ldg(a.alpha_dev) : a.alpha;
  double st_sum = 0.0, st_var = 0.0, st_viol = 0.0;
  for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < a.M; m += (int64_t)gridDim.x * blockDim.x) {
    const float* __restrict__ zrow = a.target_z + m * n;
    const float* __restrict__ qrow = a.q_pred + m * n;
    const float ent = a.next_log_pi ? __fmul_rn(alpha, -__ldg(a.next_log_pi + m)) : 0.f;
    float zmin = CUDART_INF_F;
    for (int j = 0; j < n; ++j) {
      float z = ld_stream1(zrow + j);
      if (a.next_log_pi) z = __fadd_rn(z, ent);
      zmin = fminf(zmin, z);
    }
    const float td = __fadd_rn(__ldg(a.reward + m), __fmul_rn(__fmul_rn(__ldg(a.mask + m), a.gamma), zmin));
    const float G = a.mc_return ? __ldg(a.mc_return + m) : 0.f;
    const float gs = a.grad_scale ? __ldg(a.grad_scale + m) : 1.f;
    float acc = 0.f, qsum = 0.f;
    int viol = 0;
    for (int j = 0; j < n; ++j) {
      const float q = ld_stream1(qrow + j);
      const float d = q - td, ad = fabsf(d);
      float l = ad < 1.f ? 0.5f * d * d : ad - 0.5f;  // F.smooth_l1_loss, beta = 1
      float gq = fminf(fmaxf(d, -1.f), 1.f);
      if (a.mc_return) {  // soft_actor_critic.py:93-97: q_loss = q_loss * (lb == 0) + lb
        const float lb = fmaxf(G - q, 0.f);
        if (lb != 0.f) {
          l = lb;
          gq = -1.f;
          ++viol;
        }
      }
      acc += l;
      qsum += q;
      if (a.grad_q) st_stream1(a.grad_q + m * n + j, gq * inv_n * gs);
    }
    if (a.loss) a.loss[m] = acc * inv_n;
    if (a.stats) {
      const float mean = qsum * inv_n;
      float dv = 0.f;
      for (int j = 0; j < n; ++j) {
        const float d = __ldg(qrow + j) - mean;
        dv = fmaf(d, d, dv);
      }
      st_sum += (double)mean * n;
      st_var += n > 1 ? (double)dv / (double)(n - 1) : 0.0;
      st_viol += (double)viol;
    }
  }
  if (a.stats) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      st_sum += shfl_xor_f64(st_sum, d);
      st_var += shfl_xor_f64(st_var, d);
      st_viol += shfl_xor_f64(st_viol, d);
    }
    if (lane_id() == 0) {
      atomicAdd(&sm_stats[0], st_sum);
      atomicAdd(&sm_stats[1], st_var);
      atomicAdd(&sm_stats[2], st_viol);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(a.stats + threadIdx.x, sm_stats[threadIdx.x]);
    if (threadIdx.x == 3 && blockIdx.x == 0) atomicAdd(a.stats + 3, (double)a.M);
  }
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int g_tqc_warp_kernel = 0;  // test hook: 1 = always the warp-per-transition kernel

int g_tqc_grp_warps = 0;  // test / tuning hook: warps per block of the group kernel (0 = automatic)
int g_tqc_coresident = 0; // fdql_set_coresident: leave room on every SM for one lean gather block of another stream

template <int NT, int FLAGS>
static int launch_tqc_group_f(const TqcArgs& a0, cudaStream_t st) {
  using C = GrpCfg<NT>;
  TqcArgs a = a0;
  const int64_t n_groups = (a.M + C::G - 1) / C::G;
  // the per-lane sums of transition t (three rows) fit into the dead q_pred rows 0..t when a row is at least as long as they are
  a.grp_red_alias = a.n_atoms >= 3 * C::kRedPitch ? 1 : 0;
  const int auto_warps = (a.grp_red_alias && !g_tqc_coresident) ? C::kWarpsAlias : C::kWarps;
  const int full_warps = g_tqc_grp_warps > 0 && g_tqc_grp_warps < auto_warps ? g_tqc_grp_warps : auto_warps;
  const int warp_floats = a.grp_red_alias ? C::kWarpFloatsAlias : C::kWarpFloats;
  // large batches: one block per SM (its warps share a work counter, see the kernel); batches of at most two rounds per warp: the
  // same kernel in 4-warp blocks, which leave room on an SM for kernels of other streams (the learner's prefetch pattern)
  const bool small = n_groups <= (int64_t)2 * full_warps * num_sms();
  const int kGrpWarps = small ? 4 : full_warps;
  const size_t smem = (size_t)kGrpWarps * warp_floats * sizeof(float);
  static int per_sm_cached[2][2] = {{0, 0}, {0, 0}};
  static int warps_cached = 0;
  if (warps_cached != full_warps) {
    per_sm_cached[0][0] = per_sm_cached[0][1] = per_sm_cached[1][0] = per_sm_cached[1][1] = 0;
    warps_cached = full_warps;
    constexpr size_t kNeed = sizeof(float) * (C::kWarps * C::kWarpFloats > C::kWarpsAlias * C::kWarpFloatsAlias
                                                  ? C::kWarps * C::kWarpFloats
                                                  : C::kWarpsAlias * C::kWarpFloatsAlias);
    FDQL_CUDA(cudaFuncSetAttribute(tqc_loss_group_kernel<NT, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNeed));
    // the whole unified L1 / shared memory as shared memory: the SM's split is fixed while a block is resident, and a block of a
    // co-resident gather (FDQL_OPT_CORESIDENT) only fits next to this one under the largest carve-out
    FDQL_CUDA(cudaFuncSetAttribute(tqc_loss_group_kernel<NT, FLAGS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
  }
  int& per_sm = per_sm_cached[small ? 1 : 0][a.grp_red_alias];
  if (per_sm == 0) {
    FDQL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tqc_loss_group_kernel<NT, FLAGS>, kGrpWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
  }
  // block b takes groups b, b + blocks, b + 2 blocks, ... and hands them to its warps one by one; small batches spread over all SMs
  int64_t blocks = small ? (n_groups + kGrpWarps - 1) / kGrpWarps : n_groups;
  const int64_t resident = (int64_t)num_sms() * per_sm;
  if (blocks > resident) blocks = resident;
  tqc_loss_group_kernel<NT, FLAGS><<<(unsigned)blocks, kGrpWarps * 32, smem, st>>>(a);
  FDQL_CUDA(cudaGetLastError());
  return FDQL_OK;
}
template <int NT>
static int launch_tqc_group(const TqcArgs& a, cudaStream_t st) {
  const int flags = (a.mc_return ? kGrpLb : 0) | (a.stats ? kGrpStats : 0) | (a.n_atoms > 32 * (NT / 32 - 1) ? kGrpFull : 0);
  switch (flags) {
    case 0: return launch_tqc_group_f<NT, 0>(a, st);
    case 1: return launch_tqc_group_f<NT, 1>(a, st);
    case 2: return launch_tqc_group_f<NT, 2>(a, st);
    case 3: return launch_tqc_group_f<NT, 3>(a, st);
    case 4: return launch_tqc_group_f<NT, 4>(a, st);
    case 5: return launch_tqc_group_f<NT, 5>(a, st);
    case 6: return launch_tqc_group_f<NT, 6>(a, st);
    default: return launch_tqc_group_f<NT, 7>(a, st);
  }
}

int launch_tqc(const TqcArgs& a, cudaStream_t st) {
  // the sort network holds n_z - n_drop < capacity kept targets plus +inf padding
  int need = a.n_atoms > a.n_z ? a.n_atoms : a.n_z;
  if (a.n_z - a.n_drop + 1 > need) need = a.n_z - a.n_drop + 1;
  if (need <= 128 && !g_tqc_warp_kernel) {
    if (need <= 32) return launch_tqc_group<32>(a, st);
    if (need <= 64) return launch_tqc_group<64>(a, st);
    return launch_tqc_group<128>(a, st);
  }
  int64_t blocks = (a.M + kTqcWarps - 1) / kTqcWarps;
  const int64_t max_blocks = (int64_t)num_sms() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  if (need <= 32) tqc_loss_kernel<1><<<(unsigned)blocks, kTqcWarps * 32, 0, st>>>(a);
  else if (need <= 64) tqc_loss_kernel<2><<<(unsigned)blocks, kTqcWarps * 32, 0, st>>>(a);
  else if (need <= 128) tqc_loss_kernel<4><<<(unsigned)blocks, kTqcWarps * 32, 0, st>>>(a);
  else tqc_loss_kernel<8><<<(unsigned)blocks, kTqcWarps * 32, 0, st>>>(a);
  FDQL_CUDA(cudaGetLastError());
  return FDQL_OK;
}

}  // namespace fdql

using namespace fdql;

extern "C" {

int fdql_set_coresident(int on) {
  const int old = g_tqc_coresident;
  g_tqc_coresident = on != 0;
  return old;
}

int fdql_debug_tqc_warp_kernel(int on) {
  const int old = g_tqc_warp_kernel | (g_tqc_grp_warps << 8);
  g_tqc_warp_kernel = on & 1;
  g_tqc_grp_warps = (on >> 8) & 0xff;  // bits 8..15: warps per block of the group kernel (0 = automatic)
  return old;
}

int fdql_tqc_loss(int64_t M, int32_t n_atoms, int32_t n_drop, const float* next_z, const float* q_pred,
                  const float* next_log_pi, const float* reward, const float* mask, const float* mc_return,
                  const float* grad_scale, float alpha, float gamma, float* loss, float* grad_q, float* td_target,
                  double* stats, void* stream) {
  FDQL_REQUIRE(M >= 0, "negative M");
  FDQL_REQUIRE(n_atoms >= 2 && n_atoms <= 256, "n_atoms must be in [2, 256], got %d", n_atoms);
  // quirk Q8: int(p*CQ)==0 makes the reference slice [:-0], an empty target -> refuse
  FDQL_REQUIRE(n_drop >= 1 && n_drop < n_atoms, "n_drop must be in [1, n_atoms) (reference: int(top_quantiles_to_drop*CQ)); got %d",
               n_drop);
  if (M == 0) return FDQL_OK;
  FDQL_REQUIRE(next_z && q_pred && reward && mask, "null input");
  TqcArgs a{M, n_atoms, n_atoms, n_drop, next_z, q_pred, next_log_pi, reward, mask, mc_return, grad_scale, alpha, gamma, loss, grad_q,
            td_target, stats, nullptr, 0};
  return launch_tqc(a, (cudaStream_t)stream);
}

int fdql_tqc_loss_dev_alpha(int64_t M, int32_t n_atoms, int32_t n_drop, const float* next_z, const float* q_pred,
                            const float* next_log_pi, const float* reward, const float* mask, const float* mc_return,
                            const float* grad_scale, const float* alpha_dev, float gamma, float* loss, float* grad_q,
                            float* td_target, double* stats, void* stream) {
  FDQL_REQUIRE(M >= 0, "negative M");
  FDQL_REQUIRE(n_atoms >= 2 && n_atoms <= 256, "n_atoms must be in [2, 256], got %d", n_atoms);
  FDQL_REQUIRE(n_drop >= 1 && n_drop < n_atoms, "n_drop must be in [1, n_atoms) (reference: int(top_quantiles_to_drop*CQ)); got %d",
               n_drop);
  if (M == 0) return FDQL_OK;
  FDQL_REQUIRE(next_z && q_pred && reward && mask && alpha_dev, "null input");
  TqcArgs a{M, n_atoms, n_atoms, n_drop, next_z, q_pred, next_log_pi, reward, mask, mc_return, grad_scale, 0.f, gamma, loss, grad_q,
            td_target, stats, alpha_dev, 0};
  return launch_tqc(a, (cudaStream_t)stream);
}

int fdql_quantile_huber(int64_t M, int32_t n_quantiles, int32_t n_samples, const float* quantiles, const float* samples,
                        const float* grad_scale, float* loss, float* grad_q, void* stream) {
  FDQL_REQUIRE(M >= 0, "negative M");
  FDQL_REQUIRE(n_quantiles >= 1 && n_quantiles <= 256 && n_samples >= 1 && n_samples <= 255,
               "need 1 <= n_quantiles <= 256 and 1 <= n_samples <= 255, got %d, %d", n_quantiles, n_samples);
  if (M == 0) return FDQL_OK;
  FDQL_REQUIRE(quantiles && samples, "null input");
  TqcArgs a{M, n_quantiles, n_samples, 0, samples, quantiles, nullptr, nullptr, nullptr, nullptr, grad_scale, 1.f, 1.f,
            loss, grad_q, nullptr, nullptr, nullptr, 0};
  return launch_tqc(a, (cudaStream_t)stream);
}

static int launch_sac(const SacArgs& a, cudaStream_t st) {
  if (a.n_atoms <= 16 && !g_tqc_warp_kernel) {
    int64_t tb = (a.M + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (tb > cap) tb = cap;
    sac_min_target_thread_kernel<<<(unsigned)tb, 256, 0, st>>>(a);
    FDQL_CUDA(cudaGetLastError());
    return FDQL_OK;
  }
  int64_t blocks = (a.M + 7) / 8;
  const int64_t max_blocks = (int64_t)num_sms() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  sac_min_target_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  FDQL_CUDA(cudaGetLastError());
  return FDQL_OK;
}

int fdql_sac_min_target_loss(int64_t M, int32_t n_atoms, const float* target_z, const float* q_pred, const float* next_log_pi,
                             const float* reward, const float* mask, const float* mc_return, const float* grad_scale,
                             float alpha, float gamma, float* loss, float* grad_q, double* stats, void* stream) {
  FDQL_REQUIRE(M >= 0 && n_atoms >= 1, "bad sizes");
  if (M == 0) return FDQL_OK;
  FDQL_REQUIRE(target_z && q_pred && reward && mask, "null input");
  SacArgs a{M, n_atoms, target_z, q_pred, next_log_pi, reward, mask, mc_return, grad_scale, alpha, gamma, loss, grad_q, stats, nullptr};
  return launch_sac(a, (cudaStream_t)stream);
}

int fdql_sac_min_target_loss_dev_alpha(int64_t M, int32_t n_atoms, const float* target_z, const float* q_pred,
                                       const float* next_log_pi, const float* reward, const float* mask, const float* mc_return,
                                       const float* grad_scale, const float* alpha_dev, float gamma, float* loss, float* grad_q,
                                       double* stats, void* stream) {
  FDQL_REQUIRE(M >= 0 && n_atoms >= 1, "bad sizes");
  if (M == 0) return FDQL_OK;
  FDQL_REQUIRE(target_z && q_pred && reward && mask && alpha_dev, "null input");
  SacArgs a{M, n_atoms, target_z, q_pred, next_log_pi, reward, mask, mc_return, grad_scale, 0.f, gamma, loss, grad_q, stats, alpha_dev};
  return launch_sac(a, (cudaStream_t)stream);
}

}  // extern "C"
