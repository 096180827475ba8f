// How fast are remote shared-memory reductions inside a cluster?  Cluster of CL CTAs x 512 threads; every thread issues
// ITERS red.shared::cluster.add.u32 of a 16-bit-lane increment to pseudo-random words of a 72 KB array in a
// pseudo-random CTA of its cluster (3/4 remote at CL = 4) -- the access pattern of a one-scan event histogram.
// Prints reductions per clock per SM; compare with plain atomicAdd on the CTA's own shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/dsmem_red_probe scripts/dsmem_red_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int WORDS = 18240, ITERS = 256, THREADS = 512;
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
template <int MODE>
__global__ void __launch_bounds__(THREADS) probe(int cl, unsigned long long* cycles, uint32_t* sink) {
  extern __shared__ uint32_t cnt[];
  for (int i = threadIdx.x; i < WORDS; i += THREADS) cnt[i] = 0;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(cnt);
  uint32_t s = hash(blockIdx.x * THREADS + threadIdx.x);
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    s = hash(s + i);
    const uint32_t w = s % WORDS, inc = 1u << ((s >> 20) & 16), tgt = (s >> 24) % (uint32_t)cl;
    if (MODE == 0) {
      atomicAdd(cnt + w, inc);
    } else {
      uint32_t ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + 4 * w), "r"(MODE == 1 ? tgt : rank));
      asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(ra), "r"(inc) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  uint32_t acc = 0;
  for (int i = threadIdx.x; i < WORDS; i += THREADS) acc += cnt[i];
  if (acc == 0xdeadbeef) sink[0] = acc;
}
template <int MODE>
void run(int cl, int per_sm, const char* name) {
  const int grid = 148 / cl * cl * per_sm;
  unsigned long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, grid * 8); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = WORDS * 4;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<MODE>, cl, cyc, sink);
    if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); return; }
    cudaDeviceSynchronize();
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); cudaLaunchKernelEx(&cfg, probe<MODE>, cl, cyc, sink); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  unsigned long long h[148 * 8]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
  printf("%-34s cluster %d, %d CTAs/SM: %.0f cycles for %d reds per thread -> %.2f reds / clock / SM (kernel %.3f ms, %s)\n", name, cl,
         per_sm, mean, ITERS, (double)ITERS * THREADS * per_sm / mean, ms, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  for (int per_sm = 1; per_sm <= 3; per_sm += 2) {
    run<0>(1, per_sm, "local atomicAdd");
    run<2>(1, per_sm, "red.shared::cluster to own CTA");
    run<1>(2, per_sm, "red.shared::cluster, 1/2 remote");
    run<1>(4, per_sm, "red.shared::cluster, 3/4 remote");
  }
  return 0;
}
