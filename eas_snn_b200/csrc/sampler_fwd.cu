// (a-2) Adaptive event sampler, forward.  Replaces AdaptiveRSNNEmbedding.forward / update
// (yolox/models/embedding.py:132-226) with the Rectangle spike function (activation.py:17-30).
//
// One launch per sampler step t (the launch boundary is the grid-wide dependency the recurrent
// gate_conv(spike_{t-1}) needs: its two stacked KxK convolutions see a (4R+1)^2 neighbourhood of
// the previous step's spikes).  Inside a launch one CTA owns a TH x TW pixel tile of one window and
// fuses everything the reference does in ~35 ATen launches and 3 host syncs:
//   load  : micro-bin counts (fp32 or the int32 histogram straight from eas_bin_events) and the
//           previous spikes (u8), tile + 2R*depth halo, zero padded, into shared memory
//   conv 1: input_conv[0] (2->4) and gate_conv[0] (2->4) + ReLU on tile + R halo -> shared memory
//   conv 2: input_conv[2] + gate_conv[2] as ONE 8->4 convolution (the reference adds the two
//           stacks' outputs, embedding.py:175-176), 4 x 8 register tile per thread
//   update: sigmoid gate, membrane update, strict threshold, reset, running no-reset sum,
//           spike-triggered read-out into agg[seg], seg/t_last bookkeeping; on the last step the
//           residual write (RPD: write_zero) and the optional ReLU.
// Per-pixel state (vm, acc: f32; seg, t_last: u8; spikes: u8, double buffered) lives in the
// caller's workspace between launches: 12 B per state element.  The kernel is FP32-pipe bound
// (2400 FLOP per pixel-step for depth 2, k 5) not HBM bound; see DESIGN.md.
#include "common.cuh"

namespace {

constexpr int ru4(int a) { return (a + 3) / 4 * 4; }

struct StepArgs {
  const void* events;   // [B][Tm][2][H][W]
  const uint8_t* s_prev;  // [B][2][H][W]
  uint8_t* s_next;
  float* vm;
  float* acc;
  uint8_t* seg;
  uint8_t* tl;          // t_last + 1
  float* out;           // [Ts][B][2][H][W]
  float* v_seq;         // [Tm][B][2][H][W] or null
  float* gate_seq;
  eas_sampler_weights w;
  int B, H, W, Tm, Ts;
  int t;                // sampler step (0 = newest micro-bin)
  int readout, hard_reset, write_zero, use_abs;
  float vreset, thresh;
};

// acc[co][px] += sum_{ci,ky,kx} src[ci][ky][px+kx] * w[ci][ky][kx][co]
template <int CI, int CO, int K, int PX>
__device__ __forceinline__ void conv_acc(const float* __restrict__ src, int ch_stride, int row_stride,
                                         const float* __restrict__ wsm, float (&acc)[CO][PX]) {
  constexpr int NIN = PX + K - 1;
  constexpr int NV = (NIN + 3) / 4;
#pragma unroll 1
  for (int ci = 0; ci < CI; ++ci) {
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      float in[NV * 4];
      const float4* rowp = reinterpret_cast<const float4*>(src + ci * ch_stride + ky * row_stride);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 q = rowp[v];
        in[4 * v + 0] = q.x, in[4 * v + 1] = q.y, in[4 * v + 2] = q.z, in[4 * v + 3] = q.w;
      }
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float4 wq = *reinterpret_cast<const float4*>(wsm + ((ci * K + ky) * K + kx) * CO);
        const float wv[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
        for (int co = 0; co < CO; ++co)
#pragma unroll
          for (int px = 0; px < PX; ++px) acc[co][px] = fmaf(in[px + kx], wv[co], acc[co][px]);
      }
    }
  }
}

template <int K, int DEPTH, int TH, int TW>
struct Geo {
  static constexpr int R = K / 2;
  static constexpr int PX = 8;
  static constexpr int NT = TH * TW / PX;
  // second-layer input (h1 for depth 2, the raw tile for depth 1)
  static constexpr int HR = TH + 2 * R;
  static constexpr int HC = ru4(TW + 2 * R);
  static constexpr int HS = HC + 4;
  // first-layer input (depth 2 only)
  static constexpr int IR = HR + 2 * R;
  static constexpr int IC = HC + 2 * R;
  static constexpr int IS = ru4(IC) + 4;
  static constexpr int W1 = 2 * K * K * 4;   // one first-layer stack [2][K][K][4]
  static constexpr int W2 = (DEPTH == 2 ? 8 : 4) * K * K * 4;
  static constexpr int SM_H = (DEPTH == 2 ? 8 : 4) * HR * HS + 16;
  static constexpr int SM_I = DEPTH == 2 ? 4 * IR * IS + 16 : 0;
  static constexpr int SM_W = W2 + (DEPTH == 2 ? 2 * W1 : 0) + 16;
  static constexpr size_t SMEM = sizeof(float) * (size_t)(SM_H + SM_I + SM_W);
};

template <typename IN_T>
__device__ __forceinline__ float ld_in(const IN_T* p) { return (float)__ldg(p); }

template <int K, int DEPTH, int TH, int TW, typename IN_T>
__global__ void __launch_bounds__(TH* TW / 8)
sampler_step_kernel(const StepArgs a) {
  using G = Geo<K, DEPTH, TH, TW>;
  constexpr int R = G::R;
  extern __shared__ __align__(16) float smem[];
  float* sh_h = smem;                 // layer-2 input
  float* sh_i = sh_h + G::SM_H;       // layer-1 input (depth 2)
  float* sh_w2 = sh_i + G::SM_I;      // [CI2][K][K][4]
  float* sh_w1 = sh_w2 + G::W2;       // [2 stacks][2][K][K][4]
  __shared__ float sh_b[16];          // [0..3] layer-2 bias sum, [4..7] in b0, [8..11] gate b0

  const int tid = threadIdx.x;
  const int tiles_x = (a.W + TW - 1) / TW;
  const int tiles_y = (a.H + TH - 1) / TH;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x;
  bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int x0 = tx * TW, y0 = ty * TH;
  const int64_t HW = (int64_t)a.H * a.W;
  const bool first = a.t == 0, last = a.t == a.Tm - 1;
  const int tm = a.Tm - 1 - a.t;  // newest micro-bin first (embedding.py:155-156)

  // ---- weights -> shared, re-laid out as [ci][ky][kx][co] -------------------------------------
  if (DEPTH == 2) {
    for (int i = tid; i < 2 * G::W1; i += G::NT) {
      const int stack = i / G::W1;
      int r = i - stack * G::W1;
      const int co = r & 3;
      r >>= 2;
      const int kx = r % K;
      r /= K;
      const int ky = r % K;
      const int ci = r / K;
      const float* w0 = stack == 0 ? a.w.in_w0 : a.w.gate_w0;
      sh_w1[i] = w0[((co * 2 + ci) * K + ky) * K + kx];
    }
    for (int i = tid; i < G::W2; i += G::NT) {
      int r = i;
      const int co = r & 3;
      r >>= 2;
      const int kx = r % K;
      r /= K;
      const int ky = r % K;
      const int ci = r / K;  // 0..3 input stack, 4..7 gate stack
      const float* w1 = ci < 4 ? a.w.in_w1 : a.w.gate_w1;
      sh_w2[i] = w1[((co * 4 + (ci & 3)) * K + ky) * K + kx];
    }
    if (tid < 4) {
      sh_b[tid] = a.w.in_b1[tid] + a.w.gate_b1[tid];
      sh_b[4 + tid] = a.w.in_b0[tid];
      sh_b[8 + tid] = a.w.gate_b0[tid];
    }
  } else {
    for (int i = tid; i < G::W2; i += G::NT) {
      int r = i;
      const int co = r & 3;
      r >>= 2;
      const int kx = r % K;
      r /= K;
      const int ky = r % K;
      const int ci = r / K;  // 0..1 events, 2..3 spikes
      const float* w0 = ci < 2 ? a.w.in_w0 : a.w.gate_w0;
      sh_w2[i] = w0[((co * 2 + (ci & 1)) * K + ky) * K + kx];
    }
    if (tid < 4) sh_b[tid] = a.w.in_b0[tid] + a.w.gate_b0[tid];
  }

  // ---- input tile (events + previous spikes) -> shared, zero padded ---------------------------
  {
    constexpr int ROWS = DEPTH == 2 ? G::IR : G::HR;
    constexpr int COLS = DEPTH == 2 ? G::IC : G::HC;
    constexpr int STR = DEPTH == 2 ? G::IS : G::HS;
    constexpr int HALO = R * DEPTH;
    float* dst = DEPTH == 2 ? sh_i : sh_h;
    const IN_T* ev = reinterpret_cast<const IN_T*>(a.events) + ((int64_t)b * a.Tm + tm) * 2 * HW;
    const uint8_t* sp = a.s_prev + (int64_t)b * 2 * HW;
    for (int i = tid; i < 4 * ROWS * COLS; i += G::NT) {
      const int c = i / (ROWS * COLS);
      const int rem = i - c * (ROWS * COLS);
      const int r = rem / COLS, cc = rem - r * COLS;
      const int gy = y0 - HALO + r, gx = x0 - HALO + cc;
      float v = 0.0f;
      if ((unsigned)gy < (unsigned)a.H && (unsigned)gx < (unsigned)a.W) {
        const int64_t off = (int64_t)gy * a.W + gx;
        if (c < 2) v = ld_in<IN_T>(ev + c * HW + off);
        else if (!first) v = (float)sp[(c - 2) * HW + off];
      }
      dst[(c * ROWS + r) * STR + cc] = v;
    }
  }
  __syncthreads();

  // ---- layer 1 (depth 2): 2->4 per stack, bias, ReLU, zero outside the image ------------------
  if (DEPTH == 2) {
    constexpr int PX1 = 4;
    constexpr int NSTRIP = G::HR * (G::HC / PX1);
    for (int idx = tid; idx < NSTRIP; idx += G::NT) {
      const int r = idx / (G::HC / PX1);
      const int c0 = (idx - r * (G::HC / PX1)) * PX1;
      const int gy = y0 - R + r;
      const bool row_in = (unsigned)gy < (unsigned)a.H;
#pragma unroll
      for (int stack = 0; stack < 2; ++stack) {
        float acc[4][PX1];
#pragma unroll
        for (int co = 0; co < 4; ++co)
#pragma unroll
          for (int px = 0; px < PX1; ++px) acc[co][px] = 0.0f;
        if (row_in && !(stack == 1 && first))
          conv_acc<2, 4, K, PX1>(sh_i + (stack * 2 * G::IR + r) * G::IS + c0, G::IR * G::IS, G::IS,
                                 sh_w1 + stack * G::W1, acc);
#pragma unroll
        for (int co = 0; co < 4; ++co) {
          const float bias = sh_b[4 + stack * 4 + co];
          float4 o;
          float* op = reinterpret_cast<float*>(&o);
#pragma unroll
          for (int px = 0; px < PX1; ++px) {
            const int gx = x0 - R + c0 + px;
            const bool in_img = row_in && (unsigned)gx < (unsigned)a.W;
            op[px] = in_img ? fmaxf(acc[co][px] + bias, 0.0f) : 0.0f;
          }
          *reinterpret_cast<float4*>(sh_h + ((stack * 4 + co) * G::HR + r) * G::HS + c0) = o;
        }
      }
    }
    __syncthreads();
  }

  // ---- layer 2: (8|4) -> 4, one 4 x 8 register tile per thread --------------------------------
  constexpr int PX = G::PX;
  const int r = tid / (TW / PX);
  const int c0 = (tid - r * (TW / PX)) * PX;
  float acc2[4][PX];
#pragma unroll
  for (int co = 0; co < 4; ++co)
#pragma unroll
    for (int px = 0; px < PX; ++px) acc2[co][px] = 0.0f;
  conv_acc<(DEPTH == 2 ? 8 : 4), 4, K, PX>(sh_h + r * G::HS + c0, G::HR * G::HS, G::HS, sh_w2, acc2);

  // ---- membrane update + spike-triggered aggregation (embedding.py:132-139, 177-217) ----------
  const int gy = y0 + r;
  if (gy >= a.H) return;
  const int64_t BHW2 = (int64_t)a.B * 2 * HW;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float bg = sh_b[c], bc = sh_b[2 + c];
    const int64_t base = ((int64_t)b * 2 + c) * HW + (int64_t)gy * a.W + x0 + c0;
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      if (x0 + c0 + px >= a.W) break;
      const int64_t e = base + px;
      const float gate = eas_sigmoid(acc2[c][px] + bg);
      const float cur = acc2[2 + c][px] + bc;
      float vm = first ? 0.0f : a.vm[e];
      float ac = first ? 0.0f : a.acc[e];
      int seg = first ? 0 : (int)a.seg[e];
      int tl = first ? -1 : (int)a.tl[e] - 1;
      const float v = __fadd_rn(__fmul_rn(gate, vm), cur);
      const bool s = __fsub_rn(v, a.thresh) > 0.0f;
      if (a.hard_reset) vm = s ? a.vreset : v;
      else vm = s ? __fsub_rn(v, a.thresh) : v;
      ac = __fadd_rn(ac, v);
      if (a.v_seq) {
        const int64_t se = (int64_t)a.t * BHW2 + e;
        a.v_seq[se] = v;
        a.gate_seq[se] = gate;
      }
      const bool valid = s && seg < a.Ts;
      float val = 0.0f;
      if (valid) {
        if (a.readout == EAS_READOUT_SUM) val = ac;
        else if (a.readout == EAS_READOUT_LAST) val = vm;
        else val = ac / (float)(a.t - tl);
      }
      float* outp = a.out + e;  // plane k at outp + k*BHW2
      if (first) {
        for (int k = 0; k < a.Ts; ++k) outp[k * BHW2] = (valid && k == 0) ? val : 0.0f;
      } else if (valid) {
        outp[seg * BHW2] += val;
      }
      if (valid) {
        ++seg;
        tl = a.t;
      }
      if (s) ac = 0.0f;
      if (last) {
        if (!s && seg < a.Ts && !a.write_zero) {
          float tv;
          if (a.readout == EAS_READOUT_SUM) tv = ac;
          else if (a.readout == EAS_READOUT_LAST) tv = vm;
          else tv = ac / (float)(a.Tm - 1 - tl);
          outp[seg * BHW2] += tv;
        }
        if (a.use_abs)
          for (int k = 0; k < a.Ts; ++k) outp[k * BHW2] = fmaxf(outp[k * BHW2], 0.0f);
      } else {
        a.vm[e] = vm;
        a.acc[e] = ac;
        a.seg[e] = (uint8_t)seg;
        a.tl[e] = (uint8_t)(tl + 1);
        a.s_next[e] = s ? 1 : 0;
      }
    }
  }
}

template <int K, int DEPTH, typename IN_T>
int launch_steps(const eas_sampler_cfg* cfg, StepArgs a, uint8_t* s0, uint8_t* s1, cudaStream_t st) {
  constexpr int TH = 16, TW = 64;
  using G = Geo<K, DEPTH, TH, TW>;
  auto kern = sampler_step_kernel<K, DEPTH, TH, TW, IN_T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int64_t tiles = (int64_t)((cfg->W + TW - 1) / TW) * ((cfg->H + TH - 1) / TH) * cfg->B;
  EAS_REQUIRE(tiles < (1ll << 31), EAS_E_SHAPE);
  for (int t = 0; t < cfg->Tm; ++t) {
    a.t = t;
    a.s_prev = (t & 1) ? s0 : s1;  // step t reads what step t-1 wrote
    a.s_next = (t & 1) ? s1 : s0;
    kern<<<(unsigned)tiles, G::NT, G::SMEM, st>>>(a);
    EAS_LAUNCH_CHECK();
  }
  return EAS_OK;
}

template <typename IN_T>
int dispatch(const eas_sampler_cfg* c, const StepArgs& a, uint8_t* s0, uint8_t* s1, cudaStream_t st) {
  if (c->depth == 2) {
    if (c->ksize == 3) return launch_steps<3, 2, IN_T>(c, a, s0, s1, st);
    if (c->ksize == 5) return launch_steps<5, 2, IN_T>(c, a, s0, s1, st);
    if (c->ksize == 7) return launch_steps<7, 2, IN_T>(c, a, s0, s1, st);
  } else {
    if (c->ksize == 3) return launch_steps<3, 1, IN_T>(c, a, s0, s1, st);
    if (c->ksize == 5) return launch_steps<5, 1, IN_T>(c, a, s0, s1, st);
    if (c->ksize == 7) return launch_steps<7, 1, IN_T>(c, a, s0, s1, st);
  }
  return EAS_E_UNSUPPORTED;
}

int check(const eas_sampler_cfg* c) {
  EAS_REQUIRE(c, EAS_E_NULL);
  EAS_REQUIRE(c->B >= 0 && c->H > 0 && c->W > 0, EAS_E_SHAPE);
  EAS_REQUIRE(c->Tm >= 1 && c->Tm <= 254 && c->Ts >= 1 && c->Ts <= 254, EAS_E_SHAPE);
  EAS_REQUIRE(c->depth == 1 || c->depth == 2, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->ksize == 3 || c->ksize == 5 || c->ksize == 7, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->readout >= EAS_READOUT_SUM && c->readout <= EAS_READOUT_AVG, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->in_dtype == EAS_F32 || c->in_dtype == EAS_I32, EAS_E_UNSUPPORTED);
  return EAS_OK;
}

}  // namespace

extern "C" size_t eas_sampler_fwd_ws_bytes(const eas_sampler_cfg* c) {
  if (check(c) != EAS_OK) return 0;
  const size_t n = (size_t)c->B * 2 * c->H * c->W;
  // vm, acc (f32) + seg, tl, s0, s1 (u8), each segment 256 B aligned
  return 2 * eas_align_up(n * 4, 256) + 4 * eas_align_up(n, 256) + 256;
}

extern "C" int eas_sampler_fwd(const eas_sampler_cfg* c, const void* events, const eas_sampler_weights* w,
                               float* out, float* v_seq, float* gate_seq, void* ws, size_t ws_bytes,
                               void* stream) {
  int rc = check(c);
  if (rc) return rc;
  if (c->B == 0) return EAS_OK;
  EAS_REQUIRE(events && w && out && ws, EAS_E_NULL);
  EAS_REQUIRE(w->in_w0 && w->in_b0 && w->gate_w0 && w->gate_b0, EAS_E_NULL);
  if (c->depth == 2) EAS_REQUIRE(w->in_w1 && w->in_b1 && w->gate_w1 && w->gate_b1, EAS_E_NULL);
  EAS_REQUIRE((v_seq == nullptr) == (gate_seq == nullptr), EAS_E_NULL);
  EAS_REQUIRE(ws_bytes >= eas_sampler_fwd_ws_bytes(c), EAS_E_WORKSPACE);
  EAS_REQUIRE((uintptr_t)ws % 16 == 0, EAS_E_ALIGN);
  const size_t n = (size_t)c->B * 2 * c->H * c->W;
  char* p = (char*)ws;
  StepArgs a{};
  a.events = events;
  a.vm = (float*)p;
  p += eas_align_up(n * 4, 256);
  a.acc = (float*)p;
  p += eas_align_up(n * 4, 256);
  a.seg = (uint8_t*)p;
  p += eas_align_up(n, 256);
  a.tl = (uint8_t*)p;
  p += eas_align_up(n, 256);
  uint8_t* s0 = (uint8_t*)p;
  p += eas_align_up(n, 256);
  uint8_t* s1 = (uint8_t*)p;
  a.out = out;
  a.v_seq = v_seq;
  a.gate_seq = gate_seq;
  a.w = *w;
  a.B = c->B, a.H = c->H, a.W = c->W, a.Tm = c->Tm, a.Ts = c->Ts;
  a.readout = c->readout, a.hard_reset = c->hard_reset, a.write_zero = c->write_zero, a.use_abs = c->use_abs;
  a.vreset = c->vreset, a.thresh = c->thresh;
  cudaStream_t st = (cudaStream_t)stream;
  if (c->in_dtype == EAS_F32) return dispatch<float>(c, a, s0, s1, st);
  return dispatch<int32_t>(c, a, s0, s1, st);
}
