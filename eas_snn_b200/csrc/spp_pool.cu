// SPP max-pools of the spiking CSPDarknet (SPPBottleneck, yolox/models/network_blocks.py:128-147):
// cat = [x | maxpool_k1(x) | maxpool_k2(x) | maxpool_k3(x)] with stride 1 and padding k/2, channels-last
// fp16, written straight into the channel slices of the concat buffer the next 1x1 conv reads.
// One CTA per (image, 32-channel chunk): the image chunk sits in shared memory, the three pools are
// computed separably (row maxima for the three radii, then column maxima) from one read of x.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

__global__ void __launch_bounds__(256)
spp_pool_kernel(__half* __restrict__ cat, int H, int W, int C, int ld, int chunks, int r1, int r2, int r3) {
  extern __shared__ __align__(16) uint4 sm[];   // [4][H*W][4 octets]: x, row maxima for r1, r2, r3
  const int HW = H * W, items = HW * 4;
  uint4* sx = sm;
  uint4* h1 = sm + items;
  uint4* h2 = h1 + items;
  uint4* h3 = h2 + items;
  const int n = blockIdx.x / chunks, chunk = blockIdx.x - n * chunks;
  const int c0 = chunk * 32;
  __half* img = cat + (int64_t)n * HW * ld;
  const __half2 ninf2 = __float2half2_rn(-65504.0f);
  uint4 ninf;
  reinterpret_cast<__half2*>(&ninf)[0] = reinterpret_cast<__half2*>(&ninf)[1] = ninf2;
  reinterpret_cast<__half2*>(&ninf)[2] = reinterpret_cast<__half2*>(&ninf)[3] = ninf2;
  for (int i = threadIdx.x; i < items; i += 256) {
    const int p = i >> 2, o = i & 3;
    const int c = c0 + o * 8;
    sx[i] = c < C ? *reinterpret_cast<const uint4*>(img + (int64_t)p * ld + c) : ninf;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < items; i += 256) {
    const int p = i >> 2, o = i & 3;
    const int y = p / W, x = p - y * W;
    uint4 m = sx[i];
    for (int d = 1; d <= r3; ++d) {
      if (x - d >= 0) m = hmax8(m, sx[((p - d) << 2) + o]);
      if (x + d < W) m = hmax8(m, sx[((p + d) << 2) + o]);
      if (d == r1) h1[i] = m;
      if (d == r2) h2[i] = m;
    }
    h3[i] = m;
    (void)y;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < items; i += 256) {
    const int p = i >> 2, o = i & 3;
    const int y = p / W;
    const int c = c0 + o * 8;
    if (c >= C) continue;
    uint4 m1 = h1[i], m2 = h2[i], m3 = h3[i];
    for (int d = 1; d <= r3; ++d) {
      const bool up = y - d >= 0, dn = y + d < H;
      const int iu = ((p - d * W) << 2) + o, id = ((p + d * W) << 2) + o;
      if (d <= r1) {
        if (up) m1 = hmax8(m1, h1[iu]);
        if (dn) m1 = hmax8(m1, h1[id]);
      }
      if (d <= r2) {
        if (up) m2 = hmax8(m2, h2[iu]);
        if (dn) m2 = hmax8(m2, h2[id]);
      }
      if (up) m3 = hmax8(m3, h3[iu]);
      if (dn) m3 = hmax8(m3, h3[id]);
    }
    __half* dst = img + (int64_t)p * ld + c;
    *reinterpret_cast<uint4*>(dst + C) = m1;
    *reinterpret_cast<uint4*>(dst + 2 * C) = m2;
    *reinterpret_cast<uint4*>(dst + 3 * C) = m3;
  }
}

}  // namespace

extern "C" int eas_spp_pool_fwd(void* cat, int64_t n_images, int H, int W, int C, int ld, int k1, int k2, int k3,
                                void* stream) {
  EAS_REQUIRE(n_images >= 0 && H > 0 && W > 0 && C > 0, EAS_E_SHAPE);
  EAS_REQUIRE(C % 8 == 0 && ld % 8 == 0 && ld >= 4 * C, EAS_E_SHAPE);
  EAS_REQUIRE((k1 & 1) && (k2 & 1) && (k3 & 1) && 1 <= k1 && k1 <= k2 && k2 <= k3, EAS_E_UNSUPPORTED);
  if (n_images == 0) return EAS_OK;
  EAS_REQUIRE(cat, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)cat % 16 == 0, EAS_E_ALIGN);
  const size_t smem = (size_t)4 * H * W * 4 * sizeof(uint4);
  EAS_REQUIRE(smem <= 200 * 1024, EAS_E_UNSUPPORTED);
  cudaError_t e = cudaFuncSetAttribute(spp_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int chunks = (C + 31) / 32;
  EAS_REQUIRE(n_images * chunks < (1ll << 31), EAS_E_SHAPE);
  spp_pool_kernel<<<(unsigned)(n_images * chunks), 256, smem, (cudaStream_t)stream>>>(
      (__half*)cat, H, W, C, ld, chunks, k1 / 2, k2 / 2, k3 / 2);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}
