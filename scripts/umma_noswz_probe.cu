// Hardware probe (not product code): tcgen05.mma with K-major NO-SWIZZLE operand descriptors whose
// 8-row groups are contiguous (SBO = 128 B => row r of a 16-byte K chunk sits at r*16: dense slabs)
// and whose two K chunks are LBO apart.  Checks (1) that the layout is read as assumed for any row
// shift of the start address, (2) A = fp16 with B = bf16 in one kind::f16 instruction, (3) cycles
// per MMA versus N for this dense layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/bin/umma_noswz_probe scripts/umma_noswz_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cstdlib>

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
// no-swizzle K-major descriptor: LBO = byte distance between the two 16 B K chunks, SBO = distance
// between 8-row groups
__device__ __forceinline__ uint64_t ns_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}

constexpr int AROWS = 1024;            // rows per A chunk slab
constexpr int A_LBO = AROWS * 16;      // 16 KB between the two K chunks
constexpr int B_LBO = 256 * 16;        // 4 KB
constexpr int NCORR = 12 * 2;          // shifts x (A bf16 | A fp16)
constexpr int NRATE = 7;

__host__ __device__ inline int logical(int r, int k) { return ((r * 7 + k * 3) % 61) - 30; }

__global__ void __launch_bounds__(128) probe(float* out, long long* cyc, int ncorr, int b_fp16) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                    // 2 chunks x 16 KB
  uint8_t* sB = smem + 32 * 1024;        // 2 chunks x 4 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 40 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 40 * 1024 + 64);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  uint32_t phase = 0;

  for (int ci = 0; ci < ncorr; ++ci) {
    const int shift = ci % 12, a_fp16 = ci / 12;
    for (int i = threadIdx.x; i < AROWS * 16; i += 128) {
      const int r = i / 16, k = i % 16;
      const float v = (float)logical(r, k);
      uint8_t* p = sA + (k / 8) * A_LBO + r * 16 + (k % 8) * 2;
      if (a_fp16) *reinterpret_cast<__half*>(p) = __float2half(v);
      else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(v);
    }
    for (int i = threadIdx.x; i < 256 * 16; i += 128) {
      const int n = i / 16, k = i % 16;
      uint8_t* p = sB + (k / 8) * B_LBO + n * 16 + (k % 8) * 2;
      if (b_fp16) *reinterpret_cast<__half*>(p) = __float2half(n == k ? 1.0f : 0.0f);
      else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16(n == k ? 1.0f : 0.0f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      // a_format bits 7-9 (0 = f16, 1 = bf16), b_format bits 10-12
      const uint32_t idesc = (1u << 4) | ((a_fp16 ? 0u : 1u) << 7) | ((b_fp16 ? 0u : 1u) << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      mma(tmem, ns_desc(smem_u32(sA) + shift * 16, A_LBO, 128), ns_desc(smem_u32(sB), B_LBO, 128), idesc, 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[((size_t)ci * 128 + threadIdx.x) * 16 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }

  // ---- rate: 512 MMAs, 8 rotating A windows, dense no-swizzle layout ----
  if (threadIdx.x == 0) {
    const int Ns[NRATE] = {16, 32, 48, 64, 96, 128, 256};
    for (int ci = 0; ci < NRATE; ++ci) {
      const int N = Ns[ci];
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint64_t ad[8];
      for (int r = 0; r < 8; ++r) ad[r] = ns_desc(smem_u32(sA) + (uint32_t)(r * 21) * 16, A_LBO, 128);
      const uint64_t bd = ns_desc(smem_u32(sB), B_LBO, 128);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const long long t0 = clock64();
      for (int r = 0; r < 512; r += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)((u & 1) * N)), "l"(ad[u]), "l"(bd), "r"(idesc), "r"(1u));
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
      mbar_wait(bar, phase);
      phase ^= 1;
      cyc[ci] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;  // 0: bf16 x bf16 only; 1: + A fp16 x B bf16; 2: A fp16 x B fp16
  const int ncorr = mode == 0 ? 12 : 24, b_fp16 = mode == 2;
  float* d;
  long long* c;
  const size_t n = (size_t)NCORR * 128 * 16;
  cudaMalloc(&d, n * 4);
  cudaMalloc(&c, NRATE * 8);
  cudaMemset(d, 0, n * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  for (int pass = 0; pass < 2; ++pass) {
    probe<<<1, 128, 48 * 1024>>>(d, c, ncorr, b_fp16);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  std::vector<float> h(n);
  long long hc[NRATE];
  cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
  for (int ci = 0; ci < ncorr; ++ci) {
    const int shift = ci % 12, a_fp16 = ci / 12;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int j = 0; j < 16; ++j)
        if (h[((size_t)ci * 128 + m) * 16 + j] != (float)logical(m + shift, j)) ++bad;
    printf("no-swizzle dense slabs, A=%s, row shift %2d : %s (%d bad)\n", a_fp16 ? "fp16" : "bf16", shift,
           bad ? "MISMATCH" : "ok", bad);
  }
  const int Ns[NRATE] = {16, 32, 48, 64, 96, 128, 256};
  for (int i = 0; i < NRATE; ++i) printf("dense no-swizzle A: N=%3d : %6.1f cycles/MMA\n", Ns[i], (double)hc[i] / 512);
  return 0;
}
