// (f-2) Glue kernels between the spiking backbone, the ANN PAFPN / head and the detections, all HBM bound:
//   * eas_time_mean_planes   -- `out_features[f].mean(axis=0)` (yolox/models/spiking_yolo_pafpn.py:98): firing rate
//                               over the T SNN steps, written as two fp16 planes hi + lo (the fp32 value to 2^-22,
//                               the layout the tensor-core conv kernel takes for real-valued inputs) into a channel
//                               slice of a concat buffer (the torch.cat of :103 / :108);
//   * eas_upsample2x_planes  -- nn.Upsample(scale_factor=2, mode="nearest") (:36, :102, :107) into a channel slice;
//   * eas_yolox_decode       -- YOLOXHead inference tail (yolox/models/yolo_head.py:187-199, 232-250): sigmoid on
//                               objectness / class logits, flatten + concat over levels, grid / stride decode.
// Activations are channels-last fp16 [plane][image][H][W][ld]; 8 channels (16 B) per thread.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ void split8(const float (&v)[8], uint4* hi, uint4* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    h[j] = *reinterpret_cast<const uint32_t*>(&h2), l[j] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  *hi = make_uint4(h[0], h[1], h[2], h[3]);
  *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(256)
time_mean_planes_kernel(const __half* __restrict__ x, int T, int64_t n_pix, int C8, int x_ld, __half* __restrict__ out,
                        int out_ld, int64_t out_plane) {
  const int64_t total = n_pix * C8;
  const float Tf = (float)T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / C8;
    const int c = (int)(i - pix * C8) * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    for (int t = 0; t < T; ++t) {
      const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(x + ((int64_t)t * n_pix + pix) * x_ld + c));
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        acc[2 * j] += f.x, acc[2 * j + 1] += f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __fdiv_rn(acc[j], Tf);   // torch.mean = sum / T
    uint4 hi, lo;
    split8(acc, &hi, &lo);
    __half* dst = out + pix * out_ld + c;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + out_plane) = lo;
  }
}

__global__ void __launch_bounds__(256)
upsample2x_planes_kernel(const __half* __restrict__ in, int n_planes, int64_t in_plane, int64_t n_img, int H, int W,
                         int C8, int in_ld, __half* __restrict__ out, int out_ld, int64_t out_plane) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t per_plane = n_img * Ho * Wo * C8;
  const int64_t total = per_plane * n_planes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / per_plane);
    int64_t r = i - (int64_t)p * per_plane;
    const int c = (int)(r % C8) * 8;
    r /= C8;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int64_t b = r / Ho;
    const uint4 v = *reinterpret_cast<const uint4*>(in + p * in_plane + ((b * H + (ho >> 1)) * W + (wo >> 1)) * in_ld + c);
    *reinterpret_cast<uint4*>(out + p * out_plane + ((b * Ho + ho) * Wo + wo) * out_ld + c) = v;
  }
}

__global__ void __launch_bounds__(256)
yolox_decode_kernel(const float* __restrict__ preds, int64_t n_img, int H, int W, int n_ch, int ld, float stride,
                    int decode, float* __restrict__ out, int64_t a_off, int64_t A_total) {
  const int64_t total = n_img * H * W * n_ch;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % n_ch);
    int64_t r = i / n_ch;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const int64_t b = r / H;
    float v = preds[((b * H + h) * W + w) * ld + c];
    if (c >= 4) {
      v = 1.0f / (1.0f + expf(-v));                     // obj_output.sigmoid(), cls_output.sigmoid()
    } else if (decode) {
      if (c < 2) v = __fmul_rn(__fadd_rn(v, (float)(c == 0 ? w : h)), stride);   // (xy + grid) * stride
      else v = __fmul_rn(expf(v), stride);                                      // exp(wh) * stride
    }
    out[(b * A_total + a_off + (int64_t)h * W + w) * n_ch + c] = v;
  }
}

inline unsigned grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  if (g > 8 * EAS_NUM_SMS) g = 8 * EAS_NUM_SMS;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" int eas_time_mean_planes(const void* spikes, int T, int64_t n_pix, int C, int x_ld, void* out, int out_ld,
                                    int64_t out_plane_stride, void* stream) {
  EAS_REQUIRE(T >= 1 && n_pix >= 0 && C >= 8 && C % 8 == 0, EAS_E_SHAPE);
  EAS_REQUIRE(x_ld >= C && x_ld % 8 == 0 && out_ld >= C && out_ld % 8 == 0 && out_plane_stride % 8 == 0, EAS_E_SHAPE);
  if (n_pix == 0) return EAS_OK;
  EAS_REQUIRE(spikes && out, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)spikes % 16 == 0 && (uintptr_t)out % 16 == 0, EAS_E_ALIGN);
  const int64_t total = n_pix * (C / 8);
  time_mean_planes_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)spikes, T, n_pix, C / 8, x_ld, (__half*)out, out_ld, out_plane_stride);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

extern "C" int eas_upsample2x_planes(const void* in, int n_planes, int64_t in_plane_stride, int64_t n_images, int H,
                                     int W, int C, int in_ld, void* out, int out_ld, int64_t out_plane_stride,
                                     void* stream) {
  EAS_REQUIRE(n_planes >= 1 && n_images >= 0 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0, EAS_E_SHAPE);
  EAS_REQUIRE(in_ld >= C && in_ld % 8 == 0 && out_ld >= C && out_ld % 8 == 0 && in_plane_stride % 8 == 0 &&
                  out_plane_stride % 8 == 0,
              EAS_E_SHAPE);
  if (n_images == 0) return EAS_OK;
  EAS_REQUIRE(in && out, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0, EAS_E_ALIGN);
  const int64_t total = (int64_t)n_planes * n_images * 4 * H * W * (C / 8);
  upsample2x_planes_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)in, n_planes, in_plane_stride, n_images, H, W, C / 8, in_ld, (__half*)out, out_ld,
      out_plane_stride);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

extern "C" int eas_yolox_decode(const float* preds, int64_t n_images, int H, int W, int n_ch, int ld, float stride,
                                int decode, float* out, int64_t anchor_offset, int64_t n_anchors_total, void* stream) {
  EAS_REQUIRE(n_images >= 0 && H >= 1 && W >= 1 && n_ch >= 5 && ld >= n_ch, EAS_E_SHAPE);
  EAS_REQUIRE(anchor_offset >= 0 && anchor_offset + (int64_t)H * W <= n_anchors_total, EAS_E_SHAPE);
  if (n_images == 0) return EAS_OK;
  EAS_REQUIRE(preds && out, EAS_E_NULL);
  const int64_t total = n_images * H * W * n_ch;
  yolox_decode_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(preds, n_images, H, W, n_ch, ld, stride,
                                                                         decode, out, anchor_offset, n_anchors_total);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Focus stem front end (yolox/models/network_blocks.py:191-213): space-to-depth (4 phase planes concatenated on
// channels: top-left, bottom-left, top-right, bottom-right) followed by the stem's 3x3 conv.  An 8-channel 3x3
// conv is a poor fit for TMA (nine 32-byte-row boxes per tile), so this kernel writes the im2col rows of the
// space-to-depth tensor instead -- [pixel][tap (ky, kx)][phase q = dy + 2 dx][c] = 72 values, zero padded to 80 --
// as two fp16 planes hi + lo, and the stem runs as a 1x1 conv with K = 80 on the tensor cores.
// One thread per (pixel, tap slot): 8 fp32 reads (L2 resident frames), one 16-byte store per plane.
namespace {

__global__ void __launch_bounds__(256)
focus_im2col_kernel(const float* __restrict__ frames, int64_t n_img, int H, int W, __half* __restrict__ out,
                    int64_t out_plane) {
  const int H2 = H >> 1, W2 = W >> 1;
  const int64_t total = n_img * H2 * W2 * 10;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int slot = (int)(i % 10);
    int64_t r = i / 10;
    const int w2 = (int)(r % W2);
    r /= W2;
    const int h2 = (int)(r % H2);
    const int64_t img = r / H2;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.0f;
    if (slot < 9) {
      const int ky = slot / 3, kx = slot - ky * 3;
      const int hs = h2 + ky - 1, ws = w2 + kx - 1;          // pixel of the space-to-depth map (zero padding outside)
      if (hs >= 0 && hs < H2 && ws >= 0 && ws < W2) {
        const float* src = frames + (img * 2 * H + 2 * hs) * W + 2 * ws;     // channel 0, row 2 hs, col 2 ws
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float2 r0 = *reinterpret_cast<const float2*>(src + (int64_t)c * H * W);        // dy = 0: dx = 0, 1
          const float2 r1 = *reinterpret_cast<const float2*>(src + (int64_t)c * H * W + W);    // dy = 1
          v[0 * 2 + c] = r0.x, v[1 * 2 + c] = r1.x, v[2 * 2 + c] = r0.y, v[3 * 2 + c] = r1.y;  // q = dy + 2 dx
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fminf(fmaxf(v[j], -65504.0f), 65504.0f);
    uint4 hi, lo;
    split8(v, &hi, &lo);
    __half* dst = out + (i / 10) * 80 + slot * 8;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + out_plane) = lo;
  }
}

}  // namespace

extern "C" int eas_focus_im2col(const float* frames, int64_t n_images, int H, int W, void* out,
                                int64_t out_plane_stride, void* stream) {
  EAS_REQUIRE(n_images >= 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, EAS_E_SHAPE);
  EAS_REQUIRE(out_plane_stride % 8 == 0, EAS_E_SHAPE);
  if (n_images == 0) return EAS_OK;
  EAS_REQUIRE(frames && out, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)frames % 8 == 0 && (uintptr_t)out % 16 == 0, EAS_E_ALIGN);
  const int64_t total = n_images * (H / 2) * (W / 2) * 10;
  focus_im2col_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(frames, n_images, H, W, (__half*)out,
                                                                         out_plane_stride);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}
