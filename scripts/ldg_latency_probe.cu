// Hardware probe (not product code): how long does a warp wait for a batch of 8 independent 512 B row loads
// (the sampler producer's access pattern: 4 image rows x 2 polarity planes of one strip) out of a 150 MB array,
// as a function of the warps per SM doing it?  Prints cycles per batch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/bin/ldg_latency_probe scripts/ldg_latency_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__global__ void probe(const uint4* __restrict__ src, long long* out, int iters, int W4, int HW4, int planes, int rows_per_cta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint32_t accx = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // unit: 4 consecutive rows, 2 planes; strips of 28 quads: lanes 0..27 useful
    const int unit = it * nw + warp;
    const int row0 = (blockIdx.x * rows_per_cta + unit * 4) % (240 - 4);
    const int win = (blockIdx.x * 7 + unit / 60) % planes;
    uint4 v[8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const size_t off = (size_t)(win * 2 + c) * HW4 + (size_t)(row0 + r) * W4 + lane;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[r * 2 + c].x), "=r"(v[r * 2 + c].y), "=r"(v[r * 2 + c].z), "=r"(v[r * 2 + c].w) : "l"(src + off));
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) accx += v[i].x ^ v[i].w;
  }
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * nw + warp] = (t1 - t0) + (accx == 0x12345 ? 1 : 0);
}
int main() {
  const int W4 = 76, HW4 = 240 * 76, planes = 256;  // 256 (window, micro-bin) planes x 2 polarities = 150 MB
  uint4* d; long long* o;
  cudaMalloc(&d, (size_t)planes * 2 * HW4 * 16);
  cudaMemset(d, 1, (size_t)planes * 2 * HW4 * 16);
  cudaMalloc(&o, 148 * 32 * 8);
  for (int nw = 1; nw <= 16; nw *= 2) {
    const int iters = 64;
    probe<<<148, nw * 32>>>(d, o, iters, W4, HW4, planes, 311);
    cudaDeviceSynchronize();
    probe<<<148, nw * 32>>>(d, o, iters, W4, HW4, planes, 311);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148 * 32];
    cudaMemcpy(h, o, 148 * nw * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148 * nw; ++i) s += h[i];
    printf("warps/SM %2d: %8.0f cycles per 8-load batch (dependent batches), %6.1f GB/s at 1.9 GHz\n", nw, s / (148 * nw) / iters,
           148.0 * nw * 8 * 512 / (s / (148 * nw) / iters) * 1.9);
  }
  return 0;
}
