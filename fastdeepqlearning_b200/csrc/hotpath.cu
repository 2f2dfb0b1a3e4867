// Host-buffer form of one whole pass of the hot path (include/fdql.h: fdql_hotpath_step_host): what a caller that
// owns no device memory binds.  Index/goal streams and the critics' outputs come from (pinned) host memory, the
// gathered + relabelled batch stays in HBM for the device-side MLPs, loss and dloss/dq_pred go back to host memory.
// The batch is cut into slices; one internal stream carries every H2D copy back to back, a second one the kernels and
// a third one the D2H copies, chained per slice by events, so that the H2D copy of slice i+1, the kernels of slice i and
// the D2H copy of slice i-1 overlap and the upstream copy engine never idles; the caller's stream is joined on both ends.
#include <stdlib.h>

#include "common.cuh"

using namespace fdql;

extern "C" {

int fdql_hotpath_step_host(fdql_arena* a, int64_t n, int32_t T, int64_t len, const int64_t* starts_host,
                           const uint8_t* flags_host, const int64_t* goal_rows_host, int32_t reward_op,
                           const float* reward_params_host, int32_t n_params, double gamma, uint32_t opts, float* const* out,
                           int32_t n_atoms, int32_t n_drop, const float* next_z_host, const float* q_pred_host,
                           const float* next_log_pi_host, float alpha, float* loss_host, float* grad_q_host, void* stream) {
  FDQL_REQUIRE(a != nullptr && starts_host && out && next_z_host && q_pred_host && loss_host, "null argument");
  FDQL_REQUIRE(T >= 2 && n >= 1, "need T >= 2 and at least one window");
  FDQL_REQUIRE(len >= T && len <= a->dev.capacity, "need T <= len <= capacity");
  FDQL_REQUIRE((flags_host == nullptr) == (goal_rows_host == nullptr), "flags and goal_rows come together");
  const ArenaDev& D = a->dev;
  FDQL_REQUIRE(D.col_reward >= 0 && D.col_task_done >= 0 && D.col_ep_step >= 0, "arena lacks reward/task_done/episode_step keys");
  float* o_reward = out[D.scal_key[D.col_reward]];
  float* o_ret = D.col_mc_return >= 0 ? out[D.scal_key[D.col_mc_return]] : nullptr;
  FDQL_REQUIRE(o_reward != nullptr, "the reward output is needed by the target");
  cudaStream_t user = (cudaStream_t)stream;
  const int64_t M = (int64_t)(T - 1) * n;

  // ---- device staging, grown only when a larger call arrives --------------------------------------
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  size_t off = 0;
  const size_t o_starts = off; off += al(n * 8);
  const size_t o_goal = off;   off += al(n * 8);
  const size_t o_flags = off;  off += al(n);
  const size_t o_z = off;      off += al((size_t)M * n_atoms * 4);
  const size_t o_q = off;      off += al((size_t)M * n_atoms * 4);
  const size_t o_lp = off;     off += al((size_t)M * 4);
  const size_t o_mask = off;   off += al((size_t)T * n * 4);
  const size_t o_contig = off; off += al((size_t)M * 4);
  const size_t o_weight = off; off += al((size_t)M * 4);
  const size_t o_loss = off;   off += al((size_t)M * 4);
  const size_t o_grad = off;   off += al((size_t)M * n_atoms * 4);
  if (!a->step_sync_ready) {
    for (int i = 0; i < 3; ++i) FDQL_CUDA(cudaStreamCreateWithFlags(&a->step_streams[i], cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) FDQL_CUDA(cudaEventCreateWithFlags(&a->step_events[i], cudaEventDisableTiming));
    a->step_sync_ready = 1;
  }
  if (off > a->step_bytes) {
    FDQL_CUDA(cudaDeviceSynchronize());
    if (a->step_dev) cudaFree(a->step_dev);
    a->step_dev = nullptr;
    a->step_bytes = 0;
    FDQL_CUDA(cudaMalloc(&a->step_dev, off));
    a->step_bytes = off;
  }
  char* base = static_cast<char*>(a->step_dev);
  int64_t* d_starts = reinterpret_cast<int64_t*>(base + o_starts);
  int64_t* d_goal = reinterpret_cast<int64_t*>(base + o_goal);
  uint8_t* d_flags = reinterpret_cast<uint8_t*>(base + o_flags);
  float* d_z = reinterpret_cast<float*>(base + o_z);
  float* d_q = reinterpret_cast<float*>(base + o_q);
  float* d_lp = reinterpret_cast<float*>(base + o_lp);
  float* d_mask = reinterpret_cast<float*>(base + o_mask);
  float* d_contig = reinterpret_cast<float*>(base + o_contig);
  float* d_weight = reinterpret_cast<float*>(base + o_weight);
  float* d_loss = reinterpret_cast<float*>(base + o_loss);
  float* d_grad = reinterpret_cast<float*>(base + o_grad);
  const bool relabel = flags_host != nullptr;

  // ---- slices --------------------------------------------------------------------------------------
  int64_t slice = n;
  if (n >= 16384) slice = 8192;
  else if (n >= 4096) slice = (n + 1) / 2;
  const int64_t n_slices = (n + slice - 1) / slice;
  if (2 * n_slices > a->n_slice_events) {
    cudaEvent_t* ev = static_cast<cudaEvent_t*>(realloc(a->slice_events, sizeof(cudaEvent_t) * 2 * n_slices));
    FDQL_REQUIRE(ev != nullptr, "out of host memory");
    a->slice_events = ev;
    for (int64_t i = a->n_slice_events; i < 2 * n_slices; ++i) {
      FDQL_CUDA(cudaEventCreateWithFlags(&a->slice_events[i], cudaEventDisableTiming));
      a->n_slice_events = (int)i + 1;
    }
  }
  cudaStream_t s_in = a->step_streams[0], s_run = a->step_streams[1], s_out = a->step_streams[2];
  FDQL_CUDA(cudaEventRecord(a->step_events[0], user));
  for (int i = 0; i < 3; ++i) FDQL_CUDA(cudaStreamWaitEvent(a->step_streams[i], a->step_events[0], 0));
  // the small per-window streams travel whole and first (a handful of copies instead of four per slice): the gathers can then
  // run ahead of the critic outputs, which make up 98% of the upstream bytes
  FDQL_CUDA(cudaMemcpyAsync(d_starts, starts_host, (size_t)n * 8, cudaMemcpyHostToDevice, s_in));
  if (relabel) {
    FDQL_CUDA(cudaMemcpyAsync(d_flags, flags_host, (size_t)n, cudaMemcpyHostToDevice, s_in));
    FDQL_CUDA(cudaMemcpyAsync(d_goal, goal_rows_host, (size_t)n * 8, cudaMemcpyHostToDevice, s_in));
  }
  if (next_log_pi_host) FDQL_CUDA(cudaMemcpyAsync(d_lp, next_log_pi_host, (size_t)M * 4, cudaMemcpyHostToDevice, s_in));
  for (int64_t c = 0; c < n_slices; ++c) {
    const int64_t b0 = c * slice, b1 = (b0 + slice < n) ? b0 + slice : n, nb = b1 - b0;
    cudaEvent_t landed = a->slice_events[2 * c], done = a->slice_events[2 * c + 1];
    // ---- upstream: critic outputs of this slice ----
    for (int t = 0; t + 1 < T; ++t) {
      const int64_t m0 = (int64_t)t * n + b0;
      FDQL_CUDA(cudaMemcpyAsync(d_z + m0 * n_atoms, next_z_host + m0 * n_atoms, (size_t)nb * n_atoms * 4, cudaMemcpyHostToDevice, s_in));
      FDQL_CUDA(cudaMemcpyAsync(d_q + m0 * n_atoms, q_pred_host + m0 * n_atoms, (size_t)nb * n_atoms * 4, cudaMemcpyHostToDevice, s_in));
    }
    FDQL_CUDA(cudaEventRecord(landed, s_in));
    // ---- kernels ----
    FDQL_CUDA(cudaStreamWaitEvent(s_run, landed, 0));
    int rc = launch_gather(a, n, b0, b1, T, len, d_starts, relabel ? d_flags : nullptr, relabel ? d_goal : nullptr, reward_op,
                           reward_params_host, n_params, gamma, opts | FDQL_OPT_EMIT_LEARNER_AUX, (int32_t)n, out, d_mask,
                           d_contig, d_weight, s_run);
    if (rc) return rc;
    for (int t = 0; t + 1 < T; ++t) {
      const int64_t m0 = (int64_t)t * n + b0;
      // the target reads reward / mask / mc_return of the NEXT row (t+1), quirk Q10
      rc = fdql_tqc_loss(nb, n_atoms, n_drop, d_z + m0 * n_atoms, d_q + m0 * n_atoms, next_log_pi_host ? d_lp + m0 : nullptr,
                         o_reward + m0 + n, d_mask + m0 + n, o_ret ? o_ret + m0 + n : nullptr, d_weight + m0, alpha, (float)gamma,
                         d_loss + m0, grad_q_host ? d_grad + m0 * n_atoms : nullptr, nullptr, nullptr, s_run);
      if (rc) return rc;
    }
    FDQL_CUDA(cudaEventRecord(done, s_run));
    // ---- downstream: loss and dloss/dq of this slice ----
    FDQL_CUDA(cudaStreamWaitEvent(s_out, done, 0));
    for (int t = 0; t + 1 < T; ++t) {
      const int64_t m0 = (int64_t)t * n + b0;
      FDQL_CUDA(cudaMemcpyAsync(loss_host + m0, d_loss + m0, (size_t)nb * 4, cudaMemcpyDeviceToHost, s_out));
      if (grad_q_host)
        FDQL_CUDA(cudaMemcpyAsync(grad_q_host + m0 * n_atoms, d_grad + m0 * n_atoms, (size_t)nb * n_atoms * 4, cudaMemcpyDeviceToHost, s_out));
    }
  }
  // every kernel precedes the last event of s_run, every upstream copy precedes a kernel: joining s_run and s_out is enough
  for (int i = 1; i < 3; ++i) {
    FDQL_CUDA(cudaEventRecord(a->step_events[1 + i], a->step_streams[i]));
    FDQL_CUDA(cudaStreamWaitEvent(user, a->step_events[1 + i], 0));
  }
  return FDQL_OK;
}

}  // extern "C"
