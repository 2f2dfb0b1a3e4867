// Warp-cooperative evaluation of the reward functor R(achieved_goal, goal) -> (reward, done) over 32 consecutive
// rows of one episode (device form of the Python callable franQ's HER wrapper calls per row,
// franQ/Replay/wrappers/her.py:62,67; functors: franQ/Env/bitflip.py:143-152,
// franQ/Env/classic_control_goal/classic_goal.py:88-93,306-311, franQ/Env/eleurent_parking.py:42-55).
//
// Lane layout: a goal row of `vecs` float4 is covered by LPR lanes (LPR = power of two >= vecs), so one warp-wide
// 128-bit load touches 32/LPR rows; LPR such passes cover 32 rows.  Per-row results are then transposed to
// "lane l <-> row jbase+l" with one ballot (flag functors) or one shuffle (norm functor) per pass.
#pragma once
#include "common.cuh"

namespace fdql {

template <int LPR>
struct GoalRegs {
  float4 g;  // this lane's slice (vec index lane % LPR) of the broadcast goal g*
};

template <int LPR>
__device__ __forceinline__ float4 load_goal_slice(const ArenaDev& A, int64_t goal_row) {
  const WideSlab& ag = A.wide[A.wide_ag];
  const int v = lane_id() % LPR;
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  return (v < ag.vecs) ? ldg4(ag.base + goal_row * (int64_t)ag.stride + 4 * v) : z;
}

__device__ __forceinline__ bool flag_partial(int op, int v, const float4& a, const float4& g) {
  switch (op) {
    case FDQL_REWARD_BITFLIP:
      return (a.x == g.x) & (a.y == g.y) & (a.z == g.z) & (a.w == g.w);
    case FDQL_REWARD_ALL_GEQ:
      return (a.x >= g.x) & (a.y >= g.y) & (a.z >= g.z) & (a.w >= g.w);
    default:  // FDQL_REWARD_FIRST_GEQ
      return v == 0 ? (a.x >= g.x) : true;
  }
}

// Results for this lane's row j = jbase + lane (meaningless when j > jmax; callers mask).
// PER_ROW_GOAL: compare each row against its own desired_goal (the goal-agnostic term, her.py:65-68) instead of g*.
template <int LPR, bool PER_ROW_GOAL>
__device__ __forceinline__ void eval_chunk(const ArenaDev& A, const RewardSpec& rs, int64_t ep_first, int32_t jbase,
                                           int32_t jmax, const float4& gstar, float& reward, bool& done) {
  constexpr int RPP = 32 / LPR;              // rows per pass
  constexpr int PB = LPR < 4 ? LPR : 4;      // passes whose loads are issued back to back
  constexpr unsigned GM = LPR == 32 ? 0xffffffffu : ((1u << LPR) - 1u);
  const int lane = lane_id();
  const int v = lane % LPR, rloc = lane / LPR;
  const WideSlab& ag = A.wide[A.wide_ag];
  const bool vok = v < ag.vecs;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  reward = 0.f;
  done = false;
#pragma unroll 1
  for (int p0 = 0; p0 < LPR; p0 += PB) {
    float4 a[PB], g[PB];
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int j = jbase + (p0 + q) * RPP + rloc;
      const bool ok = vok && j <= jmax;
      const int64_t row = ring_row(ep_first, j, A.capacity);
      a[q] = ok ? ldg4(ag.base + row * (int64_t)ag.stride + 4 * v) : zero;
      if (PER_ROW_GOAL) {
        const WideSlab& dg = A.wide[A.wide_dg];
        g[q] = ok ? ldg4(dg.base + row * (int64_t)dg.stride + 4 * v) : zero;
      } else {
        g[q] = ok ? gstar : zero;
      }
    }
#pragma unroll
    for (int q = 0; q < PB; ++q) {
      const int p = p0 + q;
      const bool mine = (lane / RPP) == p;
      const int src = (lane % RPP) * LPR;
      if (rs.op == FDQL_REWARD_WEIGHTED_PNORM) {
        double acc = 0.0;
        if (vok) {
          const float* w = rs.params + 2 + 4 * v;
          const int n = min(4, ag.width - 4 * v);
          // the reference evaluates the functor on float64 goals (eleurent_parking.py:53): difference, weighting and sum in fp64
          const double da[4] = {(double)a[q].x - (double)g[q].x, (double)a[q].y - (double)g[q].y, (double)a[q].z - (double)g[q].z,
                                (double)a[q].w - (double)g[q].w};
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < n) acc += fabs(da[c]) * (double)w[c];
        }
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) acc += shfl_xor_f64(acc, m);
        const double t = shfl_idx_f64(acc, src);
        if (mine) {
          const double r = -pow(t, (double)rs.params[0]);
          reward = (float)r;
          done = r > -(double)rs.params[1];
        }
      } else {
        const bool okl = flag_partial(rs.op, v, a[q], g[q]);
        const unsigned bal = __ballot_sync(kFull, okl);
        if (mine) {
          const bool m = ((bal >> src) & GM) == GM;
          reward = m ? 0.f : -1.f;
          done = m;
        }
      }
    }
  }
}

// ---- 64-bit hash of achieved_goal rows (lane <-> row jbase+lane on return), same lane layout as eval_chunk -------------
__device__ __forceinline__ uint64_t mix64(uint64_t h) {
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  return h;
}
__device__ __forceinline__ uint32_t canon_bits(float x) { return x == 0.f ? 0u : __float_as_uint(x); }

template <int LPR>
__device__ __forceinline__ void hash_chunk(const ArenaDev& A, int64_t ep_first, int32_t jbase, int32_t jmax, uint64_t& hash,
                                           bool& has_nan) {
  constexpr int RPP = 32 / LPR;
  const int lane = lane_id();
  const int v = lane % LPR, rloc = lane / LPR;
  const WideSlab& ag = A.wide[A.wide_ag];
  const bool vok = v < ag.vecs;
  hash = 0;
  has_nan = false;
#pragma unroll 1
  for (int p = 0; p < LPR; ++p) {
    const int j = jbase + p * RPP + rloc;
    const bool ok = vok && j <= jmax;
    const int64_t row = ring_row(ep_first, j, A.capacity);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) a = ldg4(ag.base + row * (int64_t)ag.stride + 4 * v);
    const int m = min(4, ag.width - 4 * v);  // floats of this slice that belong to the key
    const float xs[4] = {a.x, a.y, a.z, a.w};
    uint64_t h = 0;
    bool nan = false;
    if (ok) {
      h = 0x9E3779B97F4A7C15ull * (uint64_t)(v + 1);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < m) {
          h = mix64(h ^ ((uint64_t)canon_bits(xs[c]) + ((uint64_t)(c + 1) << 32)));
          nan |= (xs[c] != xs[c]);
        }
    }
    uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
#pragma unroll
    for (int mk = LPR / 2; mk >= 1; mk >>= 1) {
      lo ^= __shfl_xor_sync(kFull, lo, mk);
      hi ^= __shfl_xor_sync(kFull, hi, mk);
    }
    const unsigned nb = __ballot_sync(kFull, nan);
    const int src = (lane % RPP) * LPR;
    const uint32_t tlo = __shfl_sync(kFull, lo, src), thi = __shfl_sync(kFull, hi, src);
    if ((lane / RPP) == p) {
      constexpr unsigned GM = LPR == 32 ? 0xffffffffu : ((1u << LPR) - 1u);
      hash = ((uint64_t)thi << 32) | tlo;
      has_nan = ((nb >> src) & GM) != 0u;
    }
  }
}

// full-vector equality of achieved_goal[row] and achieved_goal[goal_row], by one lane (verification of a hash match)
__device__ __forceinline__ bool rows_equal(const ArenaDev& A, int64_t row, int64_t goal_row) {
  const WideSlab& ag = A.wide[A.wide_ag];
  const float* a = ag.base + row * (int64_t)ag.stride;
  const float* g = ag.base + goal_row * (int64_t)ag.stride;
  bool eq = true;
  for (int c = 0; c < ag.width; c += 4) {
    const float4 x = ldg4(a + c), y = ldg4(g + c);
    const int m = ag.width - c;
    eq &= (x.x == y.x) & (m < 2 || x.y == y.y) & (m < 3 || x.z == y.z) & (m < 4 || x.w == y.w);
  }
  return eq;
}

// lanes-per-row for a goal slab of `vecs` float4 (power of two, <= 32); vecs > 32 is rejected on the host
inline int lanes_per_row(int vecs) {
  int l = 1;
  while (l < vecs) l <<= 1;
  return l;
}

}  // namespace fdql
