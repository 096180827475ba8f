// Hardware probe (not product code): cycles per tcgen05.mma (M=128, K=16, bf16, A and B from shared
// memory, SWIZZLE_64B rows) as a function of N and of the number of independent accumulators the
// issuer rotates over.  Explains why small-N accumulation chains are latency bound and what the
// shared-memory operand read costs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/bin/umma_rate_probe scripts/umma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ uint64_t sw64_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

constexpr int NCASE = 6 * 5;  // N in {16,32,48,64,128,256} x chains in {1,2,3,4,8}

__global__ void __launch_bounds__(128) probe(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;              // 1024 rows x 64 B = 64 KB (zeros)
  uint8_t* sB = smem + 64 * 1024;  // 256 rows x 64 B = 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 80 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 80 * 1024 + 64);
  for (int i = threadIdx.x; i < 80 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const int Ns[6] = {16, 32, 48, 64, 128, 256};
    const int Cs[5] = {1, 2, 3, 4, 8};
    uint32_t phase = 0;
    for (int ci = 0; ci < NCASE; ++ci) {
      const int N = Ns[ci / 5], chains = Cs[ci % 5];
      if (N * chains > 512) { out[ci] = -1; continue; }
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // descriptors precomputed: 8 rotating A windows (rows shifted by 21, as in the sampler), 2 K halves
      uint64_t ad[8], bd[2];
      uint32_t dd[8];
      for (int r = 0; r < 8; ++r) {
        ad[r] = sw64_desc(smem_u32(sA) + (uint32_t)(r * 21) * 64 + (r & 1) * 32);
        dd[r] = tmem + (uint32_t)((r % chains) * N);
      }
      bd[0] = sw64_desc(smem_u32(sB)), bd[1] = sw64_desc(smem_u32(sB) + 32);
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd[u]), "l"(ad[u]), "l"(bd[u & 1]), "r"(idesc), "r"(1u));
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
      mbar_wait(bar, phase);
      phase ^= 1;
      out[ci] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, NCASE * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int reps = 512;
  for (int pass = 0; pass < 2; ++pass) {
    probe<<<1, 128, 96 * 1024>>>(d, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long h[NCASE];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const int Ns[6] = {16, 32, 48, 64, 128, 256};
  const int Cs[5] = {1, 2, 3, 4, 8};
  printf("cycles per tcgen05.mma (M=128, K=16, SS, SW64 rows), %d back-to-back issues\n", reps);
  for (int i = 0; i < NCASE; ++i)
    printf("N=%3d accumulators=%d : %7.1f cycles/MMA\n", Ns[i / 5], Cs[i % 5], h[i] < 0 ? -1.0 : (double)h[i] / reps);
  return 0;
}
