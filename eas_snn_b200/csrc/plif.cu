// (a-4) Multi-step parametric LIF: forward and surrogate-gradient backward.
// Replaces spikingjelly 0.0.0.0.14 neuron.ParametricLIFNode (step_mode 'm', backend 'torch') as
// configured at yolox/utils/utils_snn.py:44-53.
//
// Layout: x / spikes / grads are [T][N]; every thread owns VEC consecutive neurons and walks all T
// steps with the membrane potential in registers, so HBM traffic is the compulsory one:
//   forward  : read x, write s           (8 B per element-step in fp32, 4 B in bf16)
//   backward : read x, read g, write dx  (12 B per element-step in fp32); v is recomputed.
// The charge step uses separate IEEE mul and add (__fmul_rn/__fadd_rn), like eager PyTorch, so the
// potentials -- and therefore the spikes -- are bit-identical to the reference given identical x.
#include "lif.cuh"

namespace {

template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T, int V> struct Pack { T v[V]; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* p, float (&out)[V]) {
  if constexpr (V == 1) {
    out[0] = (float)p[0];
  } else {
    uint4 raw = ld_stream_u4(reinterpret_cast<const uint4*>(p));
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int j = 0; j < V; ++j) out[j] = (float)e[j];
  }
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* p, const float (&in)[V]) {
  if constexpr (V == 1) {
    p[0] = (T)in[0];
  } else {
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int j = 0; j < V; ++j) e[j] = (T)in[j];
    st_stream_u4(reinterpret_cast<uint4*>(p), raw);
  }
}

using Dyn = LifDyn;
__device__ __forceinline__ float charge(const Dyn& d, float v, float x) { return lif_charge(d, v, x); }
__device__ __forceinline__ float fire(const Dyn& d, float h) { return lif_fire(d, h); }
__device__ __forceinline__ float reset(const Dyn& d, float h, float s) { return lif_reset(d, h, s); }
__device__ __forceinline__ Dyn make_dyn(const eas_plif_cfg& c, const float* w) {
  return make_lif(*w, c.v_threshold, c.hard_reset, c.v_reset, c.decay_input);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
plif_fwd_kernel(eas_plif_cfg c, const T* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ v0, T* __restrict__ spikes, float* __restrict__ v_out) {
  const Dyn d = make_dyn(c, w);
  const int64_t nvec = c.N / V;
  for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec;
       iv += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = iv * V;
    float v[V], xc[V], xn[V], s[V];
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = v0 ? v0[i + j] : (d.hard ? d.vr : 0.0f);
    load_vec<T, V>(x + i, xc);
    for (int64_t t = 0; t < c.T; ++t) {
      if (t + 1 < c.T) load_vec<T, V>(x + (t + 1) * c.N + i, xn);  // next step in flight
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float h = charge(d, v[j], xc[j]);
        s[j] = fire(d, h);
        v[j] = reset(d, h, s[j]);
        xc[j] = xn[j];
      }
      store_vec<T, V>(spikes + t * c.N + i, s);
    }
    if (v_out) {
#pragma unroll
      for (int j = 0; j < V; ++j) v_out[i + j] = v[j];
    }
  }
}

__device__ __forceinline__ float surrogate_grad(int kind, float alpha, float xx) {
  if (kind == EAS_SG_ATAN) {
    const float z = 1.5707963267948966f * alpha * xx;
    return __fdividef(alpha * 0.5f, 1.0f + z * z);
  }
  if (kind == EAS_SG_SIGMOID) {
    const float sg = eas_sigmoid(alpha * xx);
    return alpha * sg * (1.0f - sg);
  }
  return fabsf(xx) < 0.5f / alpha ? alpha : 0.0f;  // Rectangle, yolox/models/activation.py:26-30
}

// FAST = the configuration the reference builds (utils_snn.py:44-53): soft reset, decay_input False,
// ATan surrogate, reset not detached.  The kernel is instruction bound (~40 instructions per element-step
// against 12 B), so the run-time switches of the general path are folded at compile time here.
template <typename T, int V, int TMAX, bool FAST>
__global__ void __launch_bounds__(256)
plif_bwd_kernel(eas_plif_cfg c, const T* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ v0, const T* __restrict__ g, T* __restrict__ dx,
                float* __restrict__ partial) {
  Dyn d = make_dyn(c, w);
  if (FAST) {
    d.hard = false, d.decay_in = false, d.vr_eff = 0.0f;
    c.surrogate = EAS_SG_ATAN, c.detach_reset = 0;
  }
  const int64_t nvec = c.N / V;
  const int Tn = (int)c.T;
  float dsw = 0.0f;
  for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec;
       iv += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = iv * V;
    float h[TMAX][V];
    float vinit[V], v[V], xt[V];
    // short sequences: every load of x and of the incoming gradient is issued before any arithmetic
    // (2T independent 16 B loads in flight per thread); h is then recomputed from registers
    constexpr bool kPreload = TMAX <= 4;
    float xs[kPreload ? TMAX : 1][V], gs[kPreload ? TMAX : 1][V];
    if (kPreload) {
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < Tn) load_vec<T, V>(x + (int64_t)t * c.N + i, xs[t]);
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < Tn) load_vec<T, V>(g + (int64_t)t * c.N + i, gs[t]);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) vinit[j] = v[j] = v0 ? v0[i + j] : (d.hard ? d.vr : 0.0f);
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
      if (t < Tn) {
        if (!kPreload) load_vec<T, V>(x + (int64_t)t * c.N + i, xt);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          h[t][j] = charge(d, v[j], kPreload ? xs[t][j] : xt[j]);
          v[j] = reset(d, h[t][j], fire(d, h[t][j]));
        }
      }
    }
    float dv[V];
#pragma unroll
    for (int j = 0; j < V; ++j) dv[j] = 0.0f;
#pragma unroll
    for (int t = TMAX - 1; t >= 0; --t) {
      if (t < Tn) {
        float gt[V], dxt[V];
        if (kPreload) {
#pragma unroll
          for (int j = 0; j < V; ++j) gt[j] = gs[t][j], xt[j] = xs[t][j];
        } else {
          load_vec<T, V>(g + (int64_t)t * c.N + i, gt);
          if (d.decay_in) load_vec<T, V>(x + (int64_t)t * c.N + i, xt);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float hh = h[t][j];
          const float s = fire(d, hh);
          const float sg = surrogate_grad(c.surrogate, c.alpha, hh - d.vth);
          float dv_ds, dv_dh;
          if (d.hard) {
            dv_ds = d.vr - hh;
            dv_dh = 1.0f - s;
          } else {
            dv_ds = -d.vth;
            dv_dh = 1.0f;
          }
          if (c.detach_reset) dv_ds = 0.0f;
          const float ds = gt[j] + dv[j] * dv_ds;
          const float dh = dv[j] * dv_dh + ds * sg;
          float vprev;
          if (t == 0) vprev = vinit[j];
          else vprev = reset(d, h[t > 0 ? t - 1 : 0][j], fire(d, h[t > 0 ? t - 1 : 0][j]));
          const float u = vprev - d.vr_eff;
          if (d.decay_in) {
            dxt[j] = dh * d.sw;
            dsw += dh * (xt[j] - u);
          } else {
            dxt[j] = dh;
            dsw += dh * (-u);
          }
          dv[j] = dh * d.k;
        }
        store_vec<T, V>(dx + (int64_t)t * c.N + i, dxt);
      }
    }
  }
  // block reduction of d loss / d sigmoid(w)
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dsw += __shfl_xor_sync(0xffffffffu, dsw, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dsw;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
  }
}

__global__ void plif_bwd_finish_kernel(const float* __restrict__ partial, int n, const float* __restrict__ w,
                                       float* __restrict__ grad_w) {
  // one block, fixed order => deterministic
  __shared__ float red[32];
  float acc = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (threadIdx.x == 0) {
      const float sw = eas_sigmoid(*w);
      *grad_w = r * sw * (1.0f - sw);
    }
  }
}

int check_cfg(const eas_plif_cfg* c) {
  EAS_REQUIRE(c, EAS_E_NULL);
  EAS_REQUIRE(c->T >= 1 && c->N >= 0, EAS_E_SHAPE);
  EAS_REQUIRE(c->dtype == EAS_F32 || c->dtype == EAS_BF16, EAS_E_UNSUPPORTED);
  return EAS_OK;
}

unsigned grid_for(int64_t nvec) {
  int64_t blocks = eas_ceil_div(nvec, 256);
  const int64_t cap = (int64_t)EAS_NUM_SMS * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

template <typename T>
bool can_vec(const eas_plif_cfg* c, const void* a, const void* b, const void* e) {
  constexpr int V = Vec<T>::N;
  return c->N % V == 0 && (uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0 && (uintptr_t)e % 16 == 0;
}

template <typename T>
int fwd_impl(const eas_plif_cfg* c, const void* x, const float* w, const float* v0, void* spikes, float* v_out,
             cudaStream_t st) {
  constexpr int V = Vec<T>::N;
  if (can_vec<T>(c, x, spikes, nullptr)) {
    plif_fwd_kernel<T, V><<<grid_for(c->N / V), 256, 0, st>>>(*c, (const T*)x, w, v0, (T*)spikes, v_out);
  } else {
    plif_fwd_kernel<T, 1><<<grid_for(c->N), 256, 0, st>>>(*c, (const T*)x, w, v0, (T*)spikes, v_out);
  }
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

template <typename T, int V>
int bwd_launch(const eas_plif_cfg* c, const void* x, const float* w, const float* v0, const void* g, void* dx,
               float* partial, unsigned grid, cudaStream_t st) {
  const T* xp = (const T*)x;
  const T* gp = (const T*)g;
  T* dp = (T*)dx;
  const bool fast = !c->hard_reset && !c->decay_input && c->surrogate == EAS_SG_ATAN && !c->detach_reset;
  if (c->T <= 4 && fast) plif_bwd_kernel<T, V, 4, true><<<grid, 256, 0, st>>>(*c, xp, w, v0, gp, dp, partial);
  else if (c->T <= 4) plif_bwd_kernel<T, V, 4, false><<<grid, 256, 0, st>>>(*c, xp, w, v0, gp, dp, partial);
  else if (c->T <= 8) plif_bwd_kernel<T, V, 8, false><<<grid, 256, 0, st>>>(*c, xp, w, v0, gp, dp, partial);
  else if (c->T <= 16) plif_bwd_kernel<T, V, 16, false><<<grid, 256, 0, st>>>(*c, xp, w, v0, gp, dp, partial);
  else return EAS_E_UNSUPPORTED;
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

}  // namespace

extern "C" int eas_plif_fwd(const eas_plif_cfg* cfg, const void* x, const float* w, const float* v0, void* spikes,
                            float* v_out, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (cfg->N == 0) return EAS_OK;
  EAS_REQUIRE(x && w && spikes, EAS_E_NULL);
  cudaStream_t st = (cudaStream_t)stream;
  if (cfg->dtype == EAS_F32) return fwd_impl<float>(cfg, x, w, v0, spikes, v_out, st);
  return fwd_impl<__nv_bfloat16>(cfg, x, w, v0, spikes, v_out, st);
}

extern "C" size_t eas_plif_bwd_ws_bytes(const eas_plif_cfg* cfg) {
  if (!cfg || cfg->N <= 0) return 256;
  return eas_align_up((size_t)grid_for(cfg->N) * sizeof(float), 256);
}

extern "C" int eas_plif_bwd(const eas_plif_cfg* cfg, const void* x, const float* w, const float* v0,
                            const void* grad_spikes, void* grad_x, float* grad_w, void* ws, size_t ws_bytes,
                            void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  EAS_REQUIRE(cfg->T <= 16, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(cfg->surrogate >= EAS_SG_ATAN && cfg->surrogate <= EAS_SG_RECT, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(w && grad_w && ws, EAS_E_NULL);
  EAS_REQUIRE(ws_bytes >= eas_plif_bwd_ws_bytes(cfg), EAS_E_WORKSPACE);
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)ws;
  unsigned grid = 1;
  if (cfg->N > 0) {
    EAS_REQUIRE(x && grad_spikes && grad_x, EAS_E_NULL);
    if (cfg->dtype == EAS_F32) {
      if (can_vec<float>(cfg, x, grad_spikes, grad_x)) {
        grid = grid_for(cfg->N / 4);
        rc = bwd_launch<float, 4>(cfg, x, w, v0, grad_spikes, grad_x, partial, grid, st);
      } else {
        grid = grid_for(cfg->N);
        rc = bwd_launch<float, 1>(cfg, x, w, v0, grad_spikes, grad_x, partial, grid, st);
      }
    } else {
      if (can_vec<__nv_bfloat16>(cfg, x, grad_spikes, grad_x)) {
        grid = grid_for(cfg->N / 8);
        rc = bwd_launch<__nv_bfloat16, 8>(cfg, x, w, v0, grad_spikes, grad_x, partial, grid, st);
      } else {
        grid = grid_for(cfg->N);
        rc = bwd_launch<__nv_bfloat16, 1>(cfg, x, w, v0, grad_spikes, grad_x, partial, grid, st);
      }
    }
    if (rc) return rc;
    plif_bwd_finish_kernel<<<1, 1024, 0, st>>>(partial, (int)grid, w, grad_w);
  } else {
    plif_bwd_finish_kernel<<<1, 1024, 0, st>>>(partial, 0, w, grad_w);
  }
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}
