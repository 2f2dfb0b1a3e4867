// Probe (not product code): rate of per-thread 1-D bulk copies (cp.async.bulk, TMA engine) of small rows gathered at random from
// HBM into shared memory, followed by one contiguous bulk store per (key, t, tile).  Shapes of the headline gather: rows of
// 256 / 32 / 64 / 64 bytes from four slabs of 1e7 rows, T = 2 consecutive rows per window, outputs [T][n][w].
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe profiles/tma_gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);                \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int NK = 4;
struct Args {
  const char* slab[NK];
  char* out[NK];
  int rb[NK];  // row bytes
  const int* starts;
  int64_t n;
  int T;
};

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
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n"
      "W_:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_;\n bra W_;\n"
      "D_:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// W windows per tile, one thread per window, STAGES tiles in flight per block
template <int W, int STAGES>
__global__ void __launch_bounds__(W) tma_gather(const __grid_constant__ Args a) {
  extern __shared__ __align__(128) char sm[];
  __shared__ __align__(8) unsigned long long bars[STAGES];
  int off[NK + 1];
  off[0] = 0;
  for (int k = 0; k < NK; ++k) off[k + 1] = off[k] + a.T * W * a.rb[k];
  const int stage_bytes = off[NK];
  int row_bytes = 0;
  for (int k = 0; k < NK; ++k) row_bytes += a.rb[k];
  const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * s, W);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t n_tiles = a.n / W;
  auto issue = [&](int64_t tile, int s) {
    const int64_t b = tile * W + threadIdx.x;
    const int64_t row = a.starts[b];
    const uint32_t base = sm0 + s * stage_bytes;
    mbar_expect_tx(bar0 + 8 * s, (uint32_t)(a.T * row_bytes));
#pragma unroll
    for (int k = 0; k < NK; ++k)
      for (int t = 0; t < a.T; ++t)
        bulk_g2s(base + off[k] + (t * W + threadIdx.x) * a.rb[k], a.slab[k] + (row + t) * a.rb[k], a.rb[k], bar0 + 8 * s);
  };
  int64_t tile = blockIdx.x;
  int it = 0;
  for (int s = 0; s < STAGES - 1; ++s) {
    if (tile + (int64_t)s * gridDim.x < n_tiles) issue(tile + (int64_t)s * gridDim.x, s);
  }
  for (; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = it % STAGES;
    // the stage that the next issue overwrites must have been read out by its stores
    const int64_t nxt = tile + (int64_t)(STAGES - 1) * gridDim.x;
    if (STAGES > 1) {
      if (threadIdx.x < NK * 2) bulk_wait_read<0>();
      __syncthreads();
      if (nxt < n_tiles) issue(nxt, (it + STAGES - 1) % STAGES);
    } else {
      issue(tile, 0);
    }
    mbar_wait(bar0 + 8 * s, (it / STAGES) & 1);
    if (threadIdx.x < NK * a.T) {
      const int k = threadIdx.x / a.T, t = threadIdx.x % a.T;
      bulk_s2g(a.out[k] + ((int64_t)t * a.n + tile * W) * a.rb[k], sm0 + s * stage_bytes + off[k] + t * W * a.rb[k], W * a.rb[k]);
      bulk_commit();
      if (STAGES == 1) bulk_wait_read<0>();
    }
    if (STAGES == 1) __syncthreads();
  }
  if (threadIdx.x < NK * 2) bulk_wait_read<0>();
}

// reference point: the same bytes moved by plain 128-bit loads / stores, one warp per window (the shape of r1's phase 2)
__global__ void __launch_bounds__(256) ldg_gather(const __grid_constant__ Args a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int k = 0, v = lane;
  for (; k < NK; ++k) {
    if (v < a.rb[k] / 16) break;
    v -= a.rb[k] / 16;
  }
  if (k == NK) return;
  for (int64_t b = warp; b < a.n; b += nwarps) {
    const int64_t row = a.starts[b];
    float4 x[2];
    for (int t = 0; t < 2; ++t) x[t] = __ldg(reinterpret_cast<const float4*>(a.slab[k] + (row + t) * a.rb[k]) + v);
    for (int t = 0; t < 2; ++t) __stcs(reinterpret_cast<float4*>(a.out[k] + ((int64_t)t * a.n + b) * a.rb[k]) + v, x[t]);
  }
}

template <int W, int STAGES>
float run(const Args& a, int blocks_per_sm, int iters, const int* starts3[3]) {
  int stage = 0;
  for (int k = 0; k < NK; ++k) stage += a.T * W * a.rb[k];
  const size_t smem = (size_t)stage * STAGES;
  CK(cudaFuncSetAttribute(tma_gather<W, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tma_gather<W, STAGES>, W, smem));
  if (blocks_per_sm > occ) blocks_per_sm = occ;
  if (blocks_per_sm < 1) return -1.f;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  Args b = a;
  for (int i = 0; i < 3; ++i) {
    b.starts = starts3[i % 3];
    tma_gather<W, STAGES><<<148 * blocks_per_sm, W, smem>>>(b);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) {
    b.starts = starts3[i % 3];
    tma_gather<W, STAGES><<<148 * blocks_per_sm, W, smem>>>(b);
  }
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  int rb = 0;
  for (int k = 0; k < NK; ++k) rb += a.rb[k];
  const double bytes = 2.0 * a.n * a.T * rb;
  printf("tma W=%3d stages=%d blocks/SM=%d (occ %d, smem %zu): %.4f ms  %.0f GB/s\n", W, STAGES, blocks_per_sm, occ, smem, ms, bytes / ms * 1e-6);
  return ms;
}

int main() {
  const int64_t cap = 10000000, n = 262144;
  Args a;
  const int rb[NK] = {256, 32, 64, 64};
  for (int k = 0; k < NK; ++k) {
    a.rb[k] = rb[k];
    char* p;
    CK(cudaMalloc(&p, (size_t)cap * rb[k]));
    CK(cudaMemset(p, k + 1, (size_t)cap * rb[k]));
    a.slab[k] = p;
    CK(cudaMalloc(&p, (size_t)2 * n * rb[k]));
    a.out[k] = p;
  }
  a.n = n;
  a.T = 2;
  const int* starts3[3];
  int* h = (int*)malloc(n * sizeof(int));
  uint64_t s = 88172645463325252ull;
  for (int i = 0; i < 3; ++i) {
    for (int64_t j = 0; j < n; ++j) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      h[j] = (int)(s % (uint64_t)(cap - 2));
    }
    int* d;
    CK(cudaMalloc(&d, n * sizeof(int)));
    CK(cudaMemcpy(d, h, n * sizeof(int), cudaMemcpyHostToDevice));
    starts3[i] = d;
  }
  // verify one variant against the plain gather
  {
    Args b = a;
    b.starts = starts3[0];
    ldg_gather<<<148 * 8, 256>>>(b);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 20; ++i) {
      b.starts = starts3[i % 3];
      ldg_gather<<<148 * 8, 256>>>(b);
    }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= 20;
    printf("ldg warp-per-window: %.4f ms  %.0f GB/s\n", ms, 2.0 * n * 2 * 416 / ms * 1e-6);
  }
  run<64, 1>(a, 4, 20, starts3);
  run<64, 1>(a, 3, 20, starts3);
  run<64, 1>(a, 2, 20, starts3);
  run<64, 2>(a, 2, 20, starts3);
  run<64, 2>(a, 1, 20, starts3);
  run<128, 1>(a, 2, 20, starts3);
  run<128, 1>(a, 1, 20, starts3);
  run<128, 2>(a, 1, 20, starts3);
  run<32, 1>(a, 8, 20, starts3);
  run<32, 2>(a, 4, 20, starts3);
  run<32, 2>(a, 2, 20, starts3);
  run<32, 2>(a, 1, 20, starts3);
  run<32, 4>(a, 2, 20, starts3);
  // only the wide rows (obs) to see the per-copy cost: rb = {256, 256, 256, 256} and {32, 32, 32, 32}
  for (int v = 0; v < 2; ++v) {
    Args b = a;
    for (int k = 0; k < NK; ++k) b.rb[k] = v == 0 ? 64 : 32;
    printf("all rows %d bytes:\n", b.rb[0]);
    run<64, 2>(b, 4, 20, starts3);
    run<128, 2>(b, 4, 20, starts3);
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
