// (f-4) RVT-preprocessed stacked histograms -> per-polarity event counts.  Replaces the 'event_sum' branch of
// RVTGEN4Dataset.generate_slices (yolox/data/datasets/rvt_gen4.py:120-122):
//   ev_repr.reshape(n, 2, -1, H, W).sum(axis=2)      uint8 [n][2*nb][H][W] -> [n][2][H][W]
// HBM bound: 2*nb bytes read and 8 bytes (two fp32 counts, exact) written per pixel; 16 pixels per thread
// with 16 B loads.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
rvt_event_sum_kernel(const uint8_t* __restrict__ repr, float* __restrict__ out, int64_t n2, int nb, int64_t HW) {
  // one (frame, polarity) plane group per blockIdx.y; 16 pixels per thread
  const int64_t np16 = (HW + 15) / 16;
  for (int64_t g = blockIdx.y; g < n2; g += gridDim.y) {
    const uint8_t* src = repr + g * nb * HW;
    float* dst = out + g * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < np16; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t p0 = i * 16;
      uint32_t acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0;
      if (p0 + 16 <= HW && (HW & 15) == 0) {
        for (int b = 0; b < nb; ++b) {
          const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(src + b * HW + p0));
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[q * 4 + j] += (w[q] >> (8 * j)) & 0xffu;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st_stream_f4(reinterpret_cast<float4*>(dst + p0 + 4 * q),
                       make_float4((float)acc[4 * q], (float)acc[4 * q + 1], (float)acc[4 * q + 2], (float)acc[4 * q + 3]));
      } else {
        for (int j = 0; j < 16 && p0 + j < HW; ++j) {
          uint32_t a = 0;
          for (int b = 0; b < nb; ++b) a += src[b * HW + p0 + j];
          dst[p0 + j] = (float)a;
        }
      }
    }
  }
}

}  // namespace

extern "C" int eas_rvt_event_sum(const uint8_t* repr, int64_t n, int nb, int H, int W, float* out, void* stream) {
  EAS_REQUIRE(n >= 0 && nb >= 1 && nb <= 255 && H > 0 && W > 0, EAS_E_SHAPE);
  if (n == 0) return EAS_OK;
  EAS_REQUIRE(repr && out, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)repr % 16 == 0 && (uintptr_t)out % 16 == 0, EAS_E_ALIGN);
  const int64_t HW = (int64_t)H * W;
  const int64_t np16 = (HW + 15) / 16;
  int64_t gx = (np16 + 255) / 256;
  if (gx > 4 * EAS_NUM_SMS) gx = 4 * EAS_NUM_SMS;
  int64_t gy = 2 * n;
  if (gy > 65535) gy = 65535;
  rvt_event_sum_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(repr, out, 2 * n, nb, HW);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

// (f-4) SpikeCountEmbedding (yolox/models/embedding.py:9-24): the micro-bin histograms of a window summed over the Tm
// micro-bins, `events.transpose(0, 1).sum(axis=0)`: hist [n][Tm][P] (f32 or i32 counts) -> out f32 [n][P], P = 2*H*W.
// HBM bound: 4*Tm bytes read + 4 written per element; 4 elements per thread, 16 B loads.
namespace {

template <bool kInt>
__global__ void __launch_bounds__(256)
hist_time_sum_kernel(const float* __restrict__ hist, float* __restrict__ out, int64_t n, int Tm, int64_t P) {
  const int64_t P4 = P >> 2;                       // P % 4 == 0 checked by the caller
  const int64_t total = n * P4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / P4, q = i - b * P4;
    const float4* src = reinterpret_cast<const float4*>(hist + b * Tm * P) + q;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < Tm; ++t) {
      float4 v = ld_stream_f4(src + (int64_t)t * P4);
      if (kInt) {
        v.x = (float)__float_as_int(v.x), v.y = (float)__float_as_int(v.y);
        v.z = (float)__float_as_int(v.z), v.w = (float)__float_as_int(v.w);
      }
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;     // t ascending, like torch.sum over dim 0
    }
    st_stream_f4(reinterpret_cast<float4*>(out + b * P) + q, acc);
  }
}

}  // namespace

extern "C" int eas_hist_time_sum(const void* hist, int in_dtype, int64_t n, int Tm, int64_t plane_elems, float* out,
                                 void* stream) {
  EAS_REQUIRE(n >= 0 && Tm >= 1 && plane_elems > 0 && plane_elems % 4 == 0, EAS_E_SHAPE);
  EAS_REQUIRE(in_dtype == EAS_F32 || in_dtype == EAS_I32, EAS_E_UNSUPPORTED);
  if (n == 0) return EAS_OK;
  EAS_REQUIRE(hist && out, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)hist % 16 == 0 && (uintptr_t)out % 16 == 0, EAS_E_ALIGN);
  const int64_t total = n * (plane_elems / 4);
  int64_t g = (total + 255) / 256;
  if (g > 8 * EAS_NUM_SMS) g = 8 * EAS_NUM_SMS;
  if (in_dtype == EAS_I32)
    hist_time_sum_kernel<true><<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const float*)hist, out, n, Tm, plane_elems);
  else
    hist_time_sum_kernel<false><<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const float*)hist, out, n, Tm, plane_elems);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

// (f-3) Letterbox + bilinear resize of micro-frames: GEN1Dataset.get_random_data with random=False
// (yolox/data/datasets/gen1.py:433-483): cv2.resize(INTER_LINEAR) of every [ih][iw] plane to [nh][nw], pasted at
// (dy, dx) into a zero canvas [oh][ow].  The tap tables (source index and float32 weight per output column / row:
// cv2's half-pixel rule, built on the host) make the kernel a pure gather: 4 output pixels per thread, 16 B stores;
// HBM bound on the write (4 B per output pixel).
namespace {

template <bool kInt>
__global__ void __launch_bounds__(256)
letterbox_bilinear_kernel(const float* __restrict__ in, int64_t n_planes, int ih, int iw, const int* __restrict__ x0t,
                          const float* __restrict__ fxt, const int* __restrict__ y0t, const float* __restrict__ fyt,
                          int nh, int nw, int dy, int dx, float* __restrict__ out, int oh, int ow) {
  const int ow4 = ow >> 2;                               // ow % 4 == 0 checked by the caller
  const int64_t total = n_planes * oh * ow4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % ow4);
    int64_t r = i / ow4;
    const int oy = (int)(r % oh);
    const int64_t pl = r / oh;
    float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
    const int y = oy - dy;
    if (y >= 0 && y < nh) {
      const int ya = y0t[y], yb = min(ya + 1, ih - 1);
      const float fy = fyt[y], gy = 1.0f - fy;
      const float* ra = in + (pl * ih + ya) * iw;
      const float* rb = in + (pl * ih + yb) * iw;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = q * 4 + j - dx;
        v[j] = 0.0f;
        if (x >= 0 && x < nw) {
          const int xa = x0t[x], xb = min(xa + 1, iw - 1);
          const float fx = fxt[x], gx = 1.0f - fx;
          float a0 = ra[xa], a1 = ra[xb], b0 = rb[xa], b1 = rb[xb];
          if (kInt) {
            a0 = (float)__float_as_int(a0), a1 = (float)__float_as_int(a1);
            b0 = (float)__float_as_int(b0), b1 = (float)__float_as_int(b1);
          }
          // horizontal pass then vertical pass, like cv2 (weights are float32 there too)
          v[j] = (a0 * gx + a1 * fx) * gy + (b0 * gx + b1 * fx) * fy;
        }
      }
      res = make_float4(v[0], v[1], v[2], v[3]);
    }
    st_stream_f4(reinterpret_cast<float4*>(out + (pl * oh + oy) * ow) + q, res);
  }
}

}  // namespace

extern "C" int eas_letterbox_bilinear(const void* in, int in_dtype, int64_t n_planes, int ih, int iw, const int32_t* x0,
                                      const float* fx, const int32_t* y0, const float* fy, int nh, int nw, int dy,
                                      int dx, float* out, int oh, int ow, void* stream) {
  EAS_REQUIRE(n_planes >= 0 && ih >= 1 && iw >= 1 && nh >= 1 && nw >= 1 && oh >= 1 && ow >= 4 && ow % 4 == 0, EAS_E_SHAPE);
  EAS_REQUIRE(dy >= 0 && dx >= 0 && dy + nh <= oh && dx + nw <= ow, EAS_E_SHAPE);
  EAS_REQUIRE(in_dtype == EAS_F32 || in_dtype == EAS_I32, EAS_E_UNSUPPORTED);
  if (n_planes == 0) return EAS_OK;
  EAS_REQUIRE(in && out && x0 && fx && y0 && fy, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)out % 16 == 0, EAS_E_ALIGN);
  const int64_t total = n_planes * oh * (ow / 4);
  int64_t g = (total + 255) / 256;
  if (g > 8 * EAS_NUM_SMS) g = 8 * EAS_NUM_SMS;
  if (in_dtype == EAS_I32)
    letterbox_bilinear_kernel<true><<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(
        (const float*)in, n_planes, ih, iw, x0, fx, y0, fy, nh, nw, dy, dx, out, oh, ow);
  else
    letterbox_bilinear_kernel<false><<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(
        (const float*)in, n_planes, ih, iw, x0, fx, y0, fy, nh, nw, dy, dx, out, oh, ow);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}


// ---------------------------------------------------------------------------------------------
// (f-4) Voxel grid with bilinear interpolation in time: to_voxel_grid_numpy, yolox/utils/event_reps.py:30-89
// (Zhu et al. 2019).  Per window: ts = n_bins * (t - t_first) / (t_last - t_first) in float64 like the reference,
// ti = int(ts), dt = ts - ti; the event adds pol * (1 - dt) to bin ti (if ti < n_bins) and pol * dt to bin ti + 1 (if
// ti + 1 < n_bins) of its pixel -- np.add.at, here fp32 reductions into the zero-filled grid.
// signed_polarity = 1: pol = +1 / -1, what the reference's comment says ("polarity should be +1 / -1").
// signed_polarity = 0: pol = +1 for every event -- what the reference COMPUTES on its own event dtype: p is a bool
// field (events_struct, yolox/utils/util.py:119-121) and `pols[pols == 0] = -1` stores bool(-1) = True (:62-63).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
voxel_grid_kernel(const int16_t* __restrict__ x, const int16_t* __restrict__ y, const int64_t* __restrict__ t,
                  const uint8_t* __restrict__ p, const int64_t* __restrict__ offsets, int64_t B, int64_t n, int H, int W,
                  int n_bins, int signed_polarity, float* __restrict__ out) {
  const int64_t HW = (int64_t)H * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t lo = 0, hi = B;   // window of event i: offsets[lo] <= i < offsets[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (offsets[mid] <= i) lo = mid;
      else hi = mid;
    }
    const int64_t s = offsets[lo], e = offsets[lo + 1];
    if (i < s || i >= e) continue;
    const int64_t t0 = t[s], t1 = t[e - 1];
    if (t1 == t0) continue;   // the reference divides by zero here (NaN bins); such a window stays all zero
    const int xi = x[i], yi = y[i];
    if ((unsigned)xi >= (unsigned)W || (unsigned)yi >= (unsigned)H) continue;
    const double ts = (double)n_bins * (double)(t[i] - t0) / (double)(t1 - t0);
    const int ti = (int)ts;
    const double dt = ts - (double)ti;
    const double pol = (signed_polarity && p[i] == 0) ? -1.0 : 1.0;
    float* cell = out + ((int64_t)lo * n_bins + ti) * HW + (int64_t)yi * W + xi;
    if (ti < n_bins) atomicAdd(cell, (float)(pol * (1.0 - dt)));
    if (ti + 1 < n_bins) atomicAdd(cell + HW, (float)(pol * dt));
  }
}
}  // namespace

extern "C" int eas_voxel_grid(const int16_t* x, const int16_t* y, const int64_t* t, const uint8_t* p,
                              const int64_t* offsets, int64_t B, int64_t n_events, int H, int W, int n_bins,
                              int signed_polarity, float* out, void* stream) {
  EAS_REQUIRE(B >= 0 && n_events >= 0 && H > 0 && W > 0 && n_bins > 0, EAS_E_SHAPE);
  if (B == 0) return EAS_OK;
  EAS_REQUIRE(offsets && out, EAS_E_NULL);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * n_bins * H * W, st);
  if (e != cudaSuccess) return (int)e;
  if (n_events == 0) return EAS_OK;
  EAS_REQUIRE(x && y && t && p, EAS_E_NULL);
  int64_t grid = eas_ceil_div(n_events, 256 * 4);
  if (grid > 8 * EAS_NUM_SMS) grid = 8 * EAS_NUM_SMS;
  voxel_grid_kernel<<<(unsigned)grid, 256, 0, st>>>(x, y, t, p, offsets, B, n_events, H, W, n_bins, signed_polarity, out);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}
