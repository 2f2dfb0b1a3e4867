// probe: how does the block scheduler place a small co-resident kernel (2 blocks of 4 warps / 27 KB per SM wanted) next to a one-block-per-SM
// kernel (16 warps, 172 KB) when both are launched from two streams in the bench's lockstep pattern?  Prints the histogram of
// small blocks per SM and the start-time spread of the big kernel's blocks.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(768, 1) big(unsigned long long* rec, int slot, unsigned long long ns) {
  extern __shared__ float sm[];
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) { rec[(slot * 148 + blockIdx.x) * 3 + 0] = t0; rec[(slot * 148 + blockIdx.x) * 3 + 1] = smid(); }
  while (gtime() - t0 < ns) { sm[threadIdx.x] += 1.f; }
  if (threadIdx.x == 0) rec[(slot * 148 + blockIdx.x) * 3 + 2] = gtime();
}
__global__ void __launch_bounds__(128) small(unsigned long long* rec, int slot, unsigned long long ns) {
  extern __shared__ float sm[];
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) { rec[(slot * 296 + blockIdx.x) * 3 + 0] = t0; rec[(slot * 296 + blockIdx.x) * 3 + 1] = smid(); }
  while (gtime() - t0 < ns) { sm[threadIdx.x] += 1.f; }
  if (threadIdx.x == 0) rec[(slot * 296 + blockIdx.x) * 3 + 2] = gtime();
}
int main() {
  const int K = 12;
  unsigned long long *rb, *rs;
  cudaMalloc(&rb, K * 148 * 3 * 8); cudaMalloc(&rs, K * 296 * 3 * 8);
  const int sb = 171520, ss = 26624;
  cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, sb);
  cudaFuncSetAttribute(big, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(small, cudaFuncAttributeMaxDynamicSharedMemorySize, ss);
  cudaFuncSetAttribute(small, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaStream_t A, B; cudaStreamCreate(&A); cudaStreamCreate(&B);
  std::vector<cudaEvent_t> dg(K + 1), dt(K + 1);
  for (auto& e : dg) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  for (auto& e : dt) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  // lockstep: loss(k) on A waits gather(k); gather(k+2) on B waits loss(k)
  small<<<296, 128, ss, B>>>(rs, 0, 150000); cudaEventRecord(dg[0], B);
  for (int k = 0; k < K; ++k) {
    if (k + 1 < K) {
      if (k >= 1) cudaStreamWaitEvent(B, dt[k - 1], 0);
      small<<<296, 128, ss, B>>>(rs, k + 1 < K ? k + 1 : 0, 150000); cudaEventRecord(dg[k + 1], B);
    }
    cudaStreamWaitEvent(A, dg[k], 0);
    big<<<148, 512, sb, A>>>(rb, k, 240000); cudaEventRecord(dt[k], A);
  }
  cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  std::vector<unsigned long long> hb(K * 148 * 3), hs(K * 296 * 3);
  cudaMemcpy(hb.data(), rb, hb.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hs.data(), rs, hs.size() * 8, cudaMemcpyDeviceToHost);
  unsigned long long t00 = hs[0];
  for (int k = 0; k < K; ++k) {
    int cnt[256] = {0}; unsigned long long smin = ~0ull, smax = 0, emax = 0;
    for (int b = 0; b < 296; ++b) { cnt[hs[(k * 296 + b) * 3 + 1]]++; smin = std::min(smin, hs[(k * 296 + b) * 3]); smax = std::max(smax, hs[(k * 296 + b) * 3]); emax = std::max(emax, hs[(k * 296 + b) * 3 + 2]); }
    int hist[9] = {0}; for (int s = 0; s < 256; ++s) if (cnt[s]) hist[std::min(cnt[s], 8)]++;
    unsigned long long bmin = ~0ull, bmax = 0, bemax = 0;
    for (int b = 0; b < 148; ++b) { bmin = std::min(bmin, hb[(k * 148 + b) * 3]); bmax = std::max(bmax, hb[(k * 148 + b) * 3]); bemax = std::max(bemax, hb[(k * 148 + b) * 3 + 2]); }
    printf("pass %2d: small blocks/SM histogram 1:%d 2:%d 3:%d 4:%d 5+:%d | small start %.1f..%.1f end %.1f us | big start %.1f..%.1f end %.1f us\n", k,
           hist[1], hist[2], hist[3], hist[4], hist[5] + hist[6] + hist[7] + hist[8], (smin - t00) / 1e3, (smax - t00) / 1e3, (emax - t00) / 1e3,
           (bmin - t00) / 1e3, (bmax - t00) / 1e3, (bemax - t00) / 1e3);
  }
  return 0;
}
