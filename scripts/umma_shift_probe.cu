// Hardware probe (not product code): does a tcgen05 K-major swizzled A operand accept a start address
// that is shifted by whole rows which are NOT a multiple of the swizzle repeat (8 rows), when the data
// was written with the absolute-address XOR swizzle?  The tensor-core sampler's Toeplitz window trick
// depends on the answer.  Prints, per (swizzle mode, row shift, base_offset policy), whether
// D[m][n] == A_logical[m + shift][16*kstep + n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/bin/umma_shift_probe scripts/umma_shift_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <vector>

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

constexpr int ROWS = 160;     // logical rows available (128 + max shift + slack)
constexpr int NCFG = 2 * 12 * 2 * 2;  // mode x shift x policy x kstep

struct Cfg { int mode, shift, policy, kstep; };

__host__ __device__ inline Cfg cfg_of(int i) {
  Cfg c;
  c.kstep = i & 1; i >>= 1;
  c.policy = i & 1; i >>= 1;
  c.shift = i % 12; i /= 12;
  c.mode = i;  // 0 = SW128 (128 B rows), 1 = SW64 (64 B rows)
  return c;
}

__host__ __device__ inline int logical(int r, int k) { return ((r * 7 + k * 3) % 61) - 30; }

__global__ void __launch_bounds__(128) probe(float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // ROWS x 128 B max = 20 KB
  uint8_t* sB = smem + 24 * 1024;           // 16 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 28 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 28 * 1024 + 64);
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  uint32_t phase = 0;

  for (int ci = 0; ci < NCFG; ++ci) {
    const Cfg c = cfg_of(ci);
    const int RB = c.mode == 0 ? 128 : 64;         // row bytes
    const int xmask = c.mode == 0 ? 7 : 3;         // XOR of address bits [7..] into bits [4..]
    // fill A and B with the absolute-address swizzle
    for (int i = threadIdx.x; i < ROWS * (RB / 2); i += 128) {
      const int r = i / (RB / 2), k = i % (RB / 2);
      uint32_t a = smem_u32(sA) + r * RB + k * 2;
      a ^= ((a >> 7) & xmask) << 4;
      *reinterpret_cast<__nv_bfloat16*>(sA + (a - smem_u32(sA))) = __float2bfloat16((float)logical(r, k));
    }
    for (int i = threadIdx.x; i < 16 * (RB / 2); i += 128) {
      const int n = i / (RB / 2), k = i % (RB / 2);
      uint32_t a = smem_u32(sB) + n * RB + k * 2;
      a ^= ((a >> 7) & xmask) << 4;
      *reinterpret_cast<__nv_bfloat16*>(sB + (a - smem_u32(sB))) = __float2bfloat16(k == c.kstep * 16 + n ? 1.0f : 0.0f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint64_t layout = c.mode == 0 ? 2 : 4;
      const uint32_t a_start = smem_u32(sA) + c.shift * RB + c.kstep * 32;
      const uint32_t b_start = smem_u32(sB) + c.kstep * 32;
      uint64_t ad = 0, bd = 0;
      ad |= (uint64_t)((a_start & 0x3FFFFu) >> 4);
      ad |= (uint64_t)((8u * RB) >> 4) << 32;
      ad |= (uint64_t)1 << 46;
      if (c.policy == 1) ad |= (uint64_t)((a_start >> 7) & 7) << 49;
      ad |= layout << 61;
      bd |= (uint64_t)((b_start & 0x3FFFFu) >> 4);
      bd |= (uint64_t)((8u * RB) >> 4) << 32;
      bd |= (uint64_t)1 << 46;
      bd |= layout << 61;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u)
          : "memory");
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
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
  }
}

int main() {
  float* d;
  const size_t n = (size_t)NCFG * 128 * 16;
  cudaMalloc(&d, n * 4);
  cudaMemset(d, 0, n * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  probe<<<1, 128, 32 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> h(n);
  cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost);
  for (int ci = 0; ci < NCFG; ++ci) {
    const Cfg c = cfg_of(ci);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int j = 0; j < 16; ++j)
        if (h[((size_t)ci * 128 + m) * 16 + j] != (float)logical(m + c.shift, c.kstep * 16 + j)) ++bad;
    printf("mode=%s shift=%2d base_offset=%s kstep=%d : %s (%d bad)\n", c.mode == 0 ? "SW128" : "SW64 ", c.shift,
           c.policy ? "phase" : "0    ", c.kstep, bad ? "MISMATCH" : "ok", bad);
  }
  return 0;
}
