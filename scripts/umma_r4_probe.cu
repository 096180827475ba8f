// Hardware probe (not product code) for the round-2 sampler kernel (sampler_tc2.cu):
//  (1) does a K-major SWIZZLE_32B A operand (32 B rows = one K-step) accept a start address shifted by
//      an arbitrary number of rows, like SW64 / SW128 do (umma_shift_probe.cu)?
//  (2) can one B array of 5 blocks x 32 rows (SWIZZLE_64B) be used through windows of 1..4 blocks
//      (N = 32..128, start shifted by whole blocks) writing to a shifted TMEM column range?
//  (3) cycles per MMA for A = SW32 / SW64 rows and N = 32..256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/bin/umma_r4_probe scripts/umma_r4_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
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
// K-major descriptor: layout 6 = SWIZZLE_32B (32 B rows, 8-row groups 256 B apart), 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, uint32_t sbo, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t idesc_f16(int N) {  // fp16 x fp16 -> fp32, M = 128
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int AROWS = 192;
__host__ __device__ inline int a_val(int r, int k) { return ((r * 7 + k * 3) % 13) - 6; }
// B[n][k] (5 blocks x 32 rows, 32 K elements = two K-steps): selects A column (n % 16) of K-step h with weight
__host__ __device__ inline int b_val(int n, int k) { return (k % 16) == (n % 16) ? 1 + (n / 16) + 20 * (k / 16) : 0; }

constexpr int NSHIFT = 34;
constexpr int NWIN = 8;  // the eight (start block, blocks, column offset) windows of the R = 4 scheme
__constant__ int c_win[NWIN][3] = {{4, 1, 0}, {3, 2, 0}, {2, 3, 0}, {1, 4, 0}, {0, 4, 0}, {0, 3, 32}, {0, 2, 64}, {0, 1, 96}};
static const int h_win[NWIN][3] = {{4, 1, 0}, {3, 2, 0}, {2, 3, 0}, {1, 4, 0}, {0, 4, 0}, {0, 3, 32}, {0, 2, 64}, {0, 1, 96}};

__global__ void __launch_bounds__(128) probe(float* out, long long* cyc, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                  // AROWS x 32 B (SW32) -- timing part uses up to 64 KB here
  uint8_t* sB = smem + 64 * 1024;      // 160 rows x 64 B (SW64) = 10 KB (timing: 256 rows = 16 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 80 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 80 * 1024 + 64);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // A: SW32 absolute-address swizzle (16 B chunk bit 4 ^= address bit 7)
  for (int i = threadIdx.x; i < AROWS * 16; i += 128) {
    const int r = i / 16, k = i % 16;
    uint32_t a = smem_u32(sA) + r * 32 + k * 2;
    a ^= ((a >> 7) & 1) << 4;
    *reinterpret_cast<__half*>(sA + (a - smem_u32(sA))) = __float2half((float)a_val(r, k));
  }
  for (int i = threadIdx.x; i < 160 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    uint32_t a = smem_u32(sB) + n * 64 + k * 2;
    a ^= ((a >> 7) & 3) << 4;
    *reinterpret_cast<__half*>(sB + (a - smem_u32(sB))) = __float2half((float)b_val(n, k));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  uint32_t phase = 0;

  // ---- (1) + (2): every shift with window (shift % NWIN), K-step h = shift & 1 ----
  for (int s = 0; s < NSHIFT; ++s) {
    const int wi = s % NWIN, h = s & 1;
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // zero the 128 columns first with a full-window MMA against... simply: first MMA acc = 0 over N = 128
      // using window 3 (all four blocks), then the tested window accumulates on top.
      const uint64_t ad = kdesc(smem_u32(sA) + s * 32, 256, 6);
      const uint64_t b_full = kdesc(smem_u32(sB) + 1 * 32 * 64, 512, 4);
      mma(tmem, ad, b_full, idesc_f16(128), 0u);
      const uint64_t bw = kdesc(smem_u32(sB) + c_win[wi][0] * 32 * 64 + h * 32, 512, 4);
      mma(tmem + c_win[wi][2], ad, bw, idesc_f16(32 * c_win[wi][1]), 1u);
      commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 128; c += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) out[((size_t)s * 128 + threadIdx.x) * 128 + c + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }

  // ---- (3) rates ----
  if (threadIdx.x == 0) {
    const int Ns[5] = {32, 64, 96, 128, 256};
    for (int mode = 0; mode < 2; ++mode)       // 0: A = SW32 rows, 1: A = SW64 rows
      for (int ni = 0; ni < 5; ++ni)
        for (int chains = 1; chains <= 2; ++chains) {
          const int N = Ns[ni];
          const uint32_t idesc = idesc_f16(N);
          uint64_t ad[8], bd[2];
          uint32_t dd[8];
          for (int r = 0; r < 8; ++r) {
            ad[r] = mode == 0 ? kdesc(smem_u32(sA) + (uint32_t)(r * 21) * 32, 256, 6)
                              : kdesc(smem_u32(sA) + (uint32_t)(r * 21) * 64 + (r & 1) * 32, 512, 4);
            dd[r] = tmem + (uint32_t)((r % chains) * N);
          }
          bd[0] = kdesc(smem_u32(sB), 512, 4), bd[1] = kdesc(smem_u32(sB) + 32, 512, 4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const long long t0 = clock64();
          for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) mma(dd[u], ad[u], bd[u & 1], idesc, 1u);
          }
          commit(bar);
          mbar_wait(bar, phase);
          phase ^= 1;
          cyc[(mode * 5 + ni) * 2 + chains - 1] = clock64() - t0;
        }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  float* d;
  long long* dc;
  const size_t n = (size_t)NSHIFT * 128 * 128;
  cudaMalloc(&d, n * 4);
  cudaMalloc(&dc, 20 * 8);
  cudaMemset(d, 0, n * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int reps = 512;
  for (int pass = 0; pass < 2; ++pass) {
    probe<<<1, 128, 96 * 1024>>>(d, dc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  std::vector<float> h(n);
  long long hc[20];
  cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
  int total_bad = 0;
  for (int s = 0; s < NSHIFT; ++s) {
    const int wi = s % NWIN, hh = s & 1;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int col = 0; col < 128; ++col) {
        // full window (blocks 1..4, K-step 0) + tested window (K-step hh)
        double ref = 0;
        for (int k = 0; k < 16; ++k) ref += (double)a_val(m + s, k) * b_val(32 + col, k);
        const int c0 = h_win[wi][2], nb = h_win[wi][1], b0 = h_win[wi][0];
        if (col >= c0 && col < c0 + 32 * nb)
          for (int k = 0; k < 16; ++k) ref += (double)a_val(m + s, k) * b_val(32 * b0 + (col - c0), 16 * hh + k);
        if (h[((size_t)s * 128 + m) * 128 + col] != (float)ref) ++bad;
      }
    printf("A=SW32 shift=%2d window(start block %d, %d blocks, col %3d) kstep=%d : %s (%d bad)\n", s, h_win[wi][0],
           h_win[wi][1], h_win[wi][2], hh, bad ? "MISMATCH" : "ok", bad);
    total_bad += bad;
  }
  const int Ns[5] = {32, 64, 96, 128, 256};
  for (int mode = 0; mode < 2; ++mode)
    for (int ni = 0; ni < 5; ++ni)
      for (int c = 0; c < 2; ++c)
        printf("rate A=%s N=%3d accumulators=%d : %6.1f cycles/MMA\n", mode ? "SW64" : "SW32", Ns[ni], c + 1,
               (double)hc[(mode * 5 + ni) * 2 + c] / reps);
  printf("total bad %d\n", total_bad);
  return total_bad ? 2 : 0;
}
