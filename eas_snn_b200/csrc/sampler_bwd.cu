// (a-2) Adaptive event sampler, backward (BPTT with the Rectangle surrogate).
// Autograd of AdaptiveRSNNEmbedding.forward (yolox/models/embedding.py:141-226), surrogate gradient
// 1[|v - thresh| < 0.5] (yolox/models/activation.py:26-30).  Gradients reach the spike through the
// reset v*(1-s), through the next step's gate_conv(s) and -- with Spike-Aware Training
// (spike_attach) -- through the read-out val*s; seg / t_last / masks are not differentiable and the
// running sum's gradient is cut where it was zeroed (embedding.py:197).
//
// Per sampler step, newest step first (reverse of the forward order), two launches:
//   point kernel : per pixel, replays the forward bookkeeping from the saved potentials v_seq[0..t]
//                  (spike, seg, t_last, running sum), combines the carried d vm / d acc, the read-out
//                  gradient and the spike gradient coming back from step t+1's gate_conv, and emits
//                  d pre-activation [B][4][H][W] (gate logits, currents) + the carries for step t-1.
//   conv kernel  : per 16x64 tile: recompute h1 = relu(conv1(.)) (never stored by the forward),
//                  d h1 = conv2^T(d pre) * [h1 > 0], d spike_{t-1} / d events = conv1^T(d h1), and the
//                  weight gradients as register-tiled correlations reduced through shared memory
//                  into the caller's gradient buffers (fp32 atomics).
// All convolutions reuse the forward's FFMA2 register-tiled routine with flipped weights.
#include "sampler_common.cuh"

namespace {

using namespace eas_sampler;

struct BwdArgs {
  const void* events;     // [B][Tm][2][H][W]
  const float* v_seq;     // [Tm][B][2][H][W]
  const float* gate_seq;
  const float* grad_out;  // [Ts][B][2][H][W]
  float* d_vm;            // carries [B][2][H][W]
  float* d_acc;
  float* d_s;             // d loss / d spike_t coming from step t+1's gate_conv
  float* dpre;            // [B][4][H][W]
  float* grad_events;     // [B][Tm][2][H][W] or null
  eas_sampler_weights w;
  eas_sampler_grads g;
  int B, H, W, Tm, Ts, t;
  int readout, hard_reset, write_zero, use_abs, spike_attach, in_is_int;
  float vreset, thresh;
  float surr_alpha, surr_half;  // Rectangle: grad * alpha inside |v - thresh| < 0.5 / alpha
};

// ---------------------------------------------------------------------------------------------
// point kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sampler_bwd_point_kernel(const BwdArgs a) {
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t n = (int64_t)a.B * 2 * HW;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int t = a.t;
  const bool is_last = t == a.Tm - 1;
  // ---- replay the forward bookkeeping up to step t ----
  float vm = 0.0f, acc = 0.0f;
  int seg = 0, tl = -1;
  float v = 0.0f, vm_prev = 0.0f, accp = 0.0f, vm_t = 0.0f;
  int seg_prev = 0, tl_prev = -1;
  bool s = false, valid = false;
  for (int tau = 0; tau <= t; ++tau) {
    v = a.v_seq[(int64_t)tau * n + e];
    s = __fsub_rn(v, a.thresh) > 0.0f;
    vm_prev = vm, seg_prev = seg, tl_prev = tl;
    vm = s ? (a.hard_reset ? a.vreset : __fsub_rn(v, a.thresh)) : v;
    accp = __fadd_rn(acc, v);
    valid = s && seg < a.Ts;
    vm_t = vm;
    if (valid) ++seg, tl = tau;
    acc = s ? 0.0f : accp;
  }
  // ---- carried gradients ----
  float d_vm = is_last ? 0.0f : a.d_vm[e];
  float d_acc = is_last ? 0.0f : a.d_acc[e];
  float d_s = is_last ? 0.0f : a.d_s[e];
  if (is_last && !s && seg < a.Ts && !a.write_zero) {  // residual write (embedding.py:203-217)
    float tv = a.readout == EAS_READOUT_SUM ? acc : vm;
    const float den = (float)(a.Tm - 1 - tl);
    if (a.readout == EAS_READOUT_AVG) tv = acc / den;
    float g = a.grad_out[(int64_t)seg * n + e];
    if (a.use_abs && !(tv > 0.0f)) g = 0.0f;
    if (a.readout == EAS_READOUT_SUM) d_acc += g;
    else if (a.readout == EAS_READOUT_LAST) d_vm += g;
    else d_acc += g / den;
  }
  float d_accp = s ? 0.0f : d_acc;  // acc = where(s, 0, acc')
  if (valid) {                      // spike-triggered read-out into agg[seg_prev] (embedding.py:181-194)
    const float den = (float)(t - tl_prev);
    float val = a.readout == EAS_READOUT_SUM ? accp : vm_t;
    if (a.readout == EAS_READOUT_AVG) val = accp / den;
    float g = a.grad_out[(int64_t)seg_prev * n + e];
    if (a.use_abs && !(val > 0.0f)) g = 0.0f;
    if (a.readout == EAS_READOUT_SUM) d_accp += g;
    else if (a.readout == EAS_READOUT_LAST) d_vm += g;
    else d_accp += g / den;
    if (a.spike_attach) d_s += g * val;  // SAT: val * s
  }
  float d_v;
  if (a.hard_reset) {  // vm = v*(1-s) + vreset*s
    d_v = s ? 0.0f : d_vm;
    d_s += d_vm * (a.vreset - v);
  } else {             // vm = v - thresh*s
    d_v = d_vm;
    d_s -= a.thresh * d_vm;
  }
  d_v += d_accp;                                                 // acc' = acc + v
  if (fabsf(__fsub_rn(v, a.thresh)) < a.surr_half) d_v += d_s * a.surr_alpha;   // Rectangle surrogate (activation.py:26-30)
  const float gate = a.gate_seq[(int64_t)t * n + e];
  // v = gate*vm_prev + cur
  const float d_gate = d_v * vm_prev;
  a.d_vm[e] = d_v * gate;
  a.d_acc[e] = d_accp;
  const int64_t bc = e / HW;          // b*2 + c
  const int64_t b = bc >> 1, c = bc & 1;
  const int64_t pix = e - bc * HW;
  a.dpre[(b * 4 + c) * HW + pix] = d_gate * gate * (1.0f - gate);  // gate logit
  a.dpre[(b * 4 + 2 + c) * HW + pix] = d_v;                        // current
}

// ---------------------------------------------------------------------------------------------
// conv kernel
// ---------------------------------------------------------------------------------------------
template <int K, int DEPTH, int TH, int TW>
struct BGeo {
  using G = Geo<K, DEPTH, TH, TW>;
  static constexpr int R = G::R;
  // number of weight-gradient "combos": one (input channel, ky) row of KX x 4 gradients each
  static constexpr int NCOMBO = DEPTH == 2 ? 12 * K : 4 * K;
  static constexpr int NGROUP = G::NT / NCOMBO;
  static constexpr int NGW = NCOMBO * K * 4;
  // buffers (floats)
  static constexpr int SM_IN = 4 * G::LR * G::LS + 16;                 // events + spikes, tile + HALO
  static constexpr int SM_DP = 4 * G::LR * G::LS + 16;                 // d pre, same geometry
  static constexpr int SM_H = DEPTH == 2 ? 8 * G::HR * G::HS + 16 : 0; // h1, tile + R
  static constexpr int SM_DH = SM_H;                                   // d h1
  static constexpr int SM_W = DEPTH == 2 ? (2 * G::W1 + 2 * 4 * K * K * 4 + 8 * K * K * 4) : 4 * K * K * 4;
  static constexpr size_t SMEM = sizeof(float) * (size_t)(SM_IN + SM_DP + SM_H + SM_DH + SM_W + NGW + 32);
};

template <int K, int DEPTH, int TH, int TW, typename IN_T>
__global__ void __launch_bounds__(TH* TW / 4, (TH * TW <= 512) ? 2 : 1) sampler_bwd_conv_kernel(const BwdArgs a) {
  using G = Geo<K, DEPTH, TH, TW>;
  using BG = BGeo<K, DEPTH, TH, TW>;
  constexpr int R = G::R;
  constexpr int PX = 4;
  constexpr int KK4 = K * K * 4;
  extern __shared__ __align__(16) float smem[];
  float* sh_in = smem;                       // [4][LR][LS]
  float* sh_dp = sh_in + BG::SM_IN;          // [4][LR][LS]
  float* sh_h = sh_dp + BG::SM_DP;           // [8][HR][HS]   (depth 2)
  float* sh_dh = sh_h + BG::SM_H;            // [8][HR][HS]
  float* sh_w = sh_dh + BG::SM_DH;
  // depth 2: w1 fwd [2][2][K][K][4] | w2t [2 halves][4][K][K][4] | w1t [8][K][K][4]
  // depth 1: w0t [4][K][K][4]
  float* sh_w1 = sh_w;
  float* sh_w2t = sh_w + (DEPTH == 2 ? 2 * G::W1 : 0);
  float* sh_w1t = sh_w2t + (DEPTH == 2 ? 2 * 4 * KK4 : 0);
  float* sh_gw = sh_w + BG::SM_W;            // [NCOMBO][K][4]
  float* sh_gb = sh_gw + BG::NGW;            // [12] bias gradients
  __shared__ float sh_b0[8];

  const int tid = threadIdx.x;
  const int tiles_x = (a.W + TW - 1) / TW, tiles_y = (a.H + TH - 1) / TH;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x;
  bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int x0 = tx * TW, y0 = ty * TH;
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t n = (int64_t)a.B * 2 * HW;
  const int t = a.t, tm = a.Tm - 1 - t;
  const bool first = t == 0;

  // ---- weights -> shared ----
  if (DEPTH == 2) {
    for (int i = tid; i < 2 * G::W1; i += G::NT) {  // forward layer 1: [stack][ci][ky][kx][co]
      const int stack = i / G::W1;
      int r = i - stack * G::W1;
      const int co = r & 3;
      r >>= 2;
      const int kx = r % K;
      r /= K;
      const int ky = r % K, ci = r / K;
      const float* w0 = stack == 0 ? a.w.in_w0 : a.w.gate_w0;
      sh_w1[i] = w0[((co * 2 + ci) * K + ky) * K + kx];
    }
    // layer 2 transposed: d h1[half*4 + o] = sum_{co,k'} dpre[co][. + k'] * w2[co][half*4+o][K-1-k']
    for (int i = tid; i < 2 * 4 * KK4; i += G::NT) {
      const int half = i / (4 * KK4);
      int r = i - half * 4 * KK4;
      const int o = r & 3;
      r >>= 2;
      const int kx = r % K;
      r /= K;
      const int ky = r % K, co = r / K;
      const float* w1 = half == 0 ? a.w.in_w1 : a.w.gate_w1;
      sh_w2t[i] = w1[((co * 4 + o) * K + (K - 1 - ky)) * K + (K - 1 - kx)];
    }
    // layer 1 transposed: out = [d ev0, d ev1, d s0, d s1]; in = d h1[8] (block structured)
    for (int i = tid; i < 8 * KK4; i += G::NT) {
      int r = i;
      const int o = r & 3;
      r >>= 2;
      const int kx = r % K;
      r /= K;
      const int ky = r % K, ci = r / K;
      float val = 0.0f;
      if (ci < 4 && o < 2) val = a.w.in_w0[((ci * 2 + o) * K + (K - 1 - ky)) * K + (K - 1 - kx)];
      if (ci >= 4 && o >= 2) val = a.w.gate_w0[(((ci - 4) * 2 + (o - 2)) * K + (K - 1 - ky)) * K + (K - 1 - kx)];
      sh_w1t[i] = val;
    }
    if (tid < 4) sh_b0[tid] = a.w.in_b0[tid], sh_b0[4 + tid] = a.w.gate_b0[tid];
  } else {
    // single layer transposed: out c = [ev0, ev1, s0, s1]; in = dpre[4]
    for (int i = tid; i < 4 * KK4; i += G::NT) {
      int r = i;
      const int o = r & 3;
      r >>= 2;
      const int kx = r % K;
      r /= K;
      const int ky = r % K, co = r / K;
      const float* w0 = o < 2 ? a.w.in_w0 : a.w.gate_w0;
      sh_w2t[i] = w0[((co * 2 + (o & 1)) * K + (K - 1 - ky)) * K + (K - 1 - kx)];
    }
  }
  for (int i = tid; i < BG::NGW + 32; i += G::NT) sh_gw[i] = 0.0f;

  // ---- inputs (events, previous spikes) and d pre, tile + HALO, zero padded ----
  {
    const IN_T* ev = reinterpret_cast<const IN_T*>(a.events) + ((int64_t)b * a.Tm + tm) * 2 * HW;
    const float* vprev = first ? nullptr : a.v_seq + (int64_t)(t - 1) * n + (int64_t)b * 2 * HW;
    const float* dp = a.dpre + (int64_t)b * 4 * HW;
    for (int i = tid; i < 4 * G::LR * G::LC; i += G::NT) {
      const int c = i / (G::LR * G::LC);
      const int rem = i - c * (G::LR * G::LC);
      const int r = rem / G::LC, cc = rem - r * G::LC;
      const int gy = y0 - G::HALO + r, gx = x0 - G::HALO + cc;
      float vin = 0.0f, vdp = 0.0f;
      if ((unsigned)gy < (unsigned)a.H && (unsigned)gx < (unsigned)a.W) {
        const int64_t off = (int64_t)gy * a.W + gx;
        if (c < 2) vin = (float)ev[c * HW + off];
        else if (!first) vin = __fsub_rn(vprev[(c - 2) * HW + off], a.thresh) > 0.0f ? 1.0f : 0.0f;
        vdp = dp[c * HW + off];
      }
      sh_in[(c * G::LR + r) * G::LS + cc] = vin;
      sh_dp[(c * G::LR + r) * G::LS + cc] = vdp;
    }
  }
  __syncthreads();

  if (DEPTH == 2) {
    // ---- recompute h1 and compute d h1 = conv2^T(d pre) * [h1 > 0], both on tile + R ----
    constexpr int SPR = G::HC / PX;
    constexpr int NITEM = 2 * G::HR * SPR;  // (half, row, strip)
    for (int idx = tid; idx < NITEM; idx += G::NT) {
      const int half = idx / (G::HR * SPR);
      const int rem = idx - half * (G::HR * SPR);
      const int r = rem / SPR;
      const int c0 = (rem - r * SPR) * PX;
      const int gy = y0 - R + r;
      const bool row_in = (unsigned)gy < (unsigned)a.H;
      unsigned long long hp[2][PX], dp2[2][PX];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int px = 0; px < PX; ++px) hp[h][px] = 0ull, dp2[h][px] = 0ull;
      if (row_in) {
        if (!(half == 1 && first))
          conv_acc<2, 4, K, PX>(sh_in + (half * 2 * G::IR + r) * G::IS + c0, G::IR * G::IS, G::IS, sh_w1 + half * G::W1, hp);
        conv_acc<4, 4, K, PX>(sh_dp + r * G::IS + c0, G::IR * G::IS, G::IS, sh_w2t + half * 4 * KK4, dp2);
      }
      float hv[4][PX], dv[4][PX];
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        unpack2(hp[0][px], hv[0][px], hv[1][px]);
        unpack2(hp[1][px], hv[2][px], hv[3][px]);
        unpack2(dp2[0][px], dv[0][px], dv[1][px]);
        unpack2(dp2[1][px], dv[2][px], dv[3][px]);
      }
#pragma unroll
      for (int co = 0; co < 4; ++co) {
        const float bias = sh_b0[half * 4 + co];
        float4 oh, od;
        float* ph = reinterpret_cast<float*>(&oh);
        float* pd = reinterpret_cast<float*>(&od);
#pragma unroll
        for (int px = 0; px < PX; ++px) {
          const int gx = x0 - R + c0 + px;
          const bool in_img = row_in && (unsigned)gx < (unsigned)a.W;
          const float h = in_img ? fmaxf(hv[co][px] + bias, 0.0f) : 0.0f;
          ph[px] = h;
          pd[px] = h > 0.0f ? dv[co][px] : 0.0f;
        }
        *reinterpret_cast<float4*>(sh_h + ((half * 4 + co) * G::HR + r) * G::HS + c0) = oh;
        *reinterpret_cast<float4*>(sh_dh + ((half * 4 + co) * G::HR + r) * G::HS + c0) = od;
      }
    }
    __syncthreads();
  }

  // ---- d inputs on the tile: [d ev0, d ev1, d s_prev0, d s_prev1] ----
  const int r = tid / (TW / PX);
  const int c0 = (tid - r * (TW / PX)) * PX;
  {
    unsigned long long dp2[2][PX];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int px = 0; px < PX; ++px) dp2[h][px] = 0ull;
    if (DEPTH == 2) conv_acc<8, 4, K, PX>(sh_dh + r * G::HS + c0, G::HR * G::HS, G::HS, sh_w1t, dp2);
    else conv_acc<4, 4, K, PX>(sh_dp + r * G::HS + c0, G::HR * G::HS, G::HS, sh_w2t, dp2);
    float d4[4][PX];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      unpack2(dp2[0][px], d4[0][px], d4[1][px]);
      unpack2(dp2[1][px], d4[2][px], d4[3][px]);
    }
    const int gy = y0 + r;
    if (gy < a.H) {
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        const int gx = x0 + c0 + px;
        if (gx < a.W) {
          const int64_t off = (int64_t)gy * a.W + gx;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (!first) a.d_s[((int64_t)b * 2 + c) * HW + off] = d4[2 + c][px];
            if (a.grad_events) a.grad_events[(((int64_t)b * a.Tm + tm) * 2 + c) * HW + off] = d4[c][px];
          }
        }
      }
    }
  }

  // ---- bias gradients: sums over the tile's own pixels ----
  {
    float bs[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) bs[j] = 0.0f;
    // d pre at the tile pixel (r, c0 + px) sits at offset HALO in the padded buffer
#pragma unroll
    for (int co = 0; co < 4; ++co)
#pragma unroll
      for (int px = 0; px < PX; ++px) bs[co] += sh_dp[(co * G::LR + r + G::HALO) * G::LS + c0 + G::HALO + px];
    if (DEPTH == 2) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
#pragma unroll
        for (int px = 0; px < PX; ++px) bs[4 + ch] += sh_dh[(ch * G::HR + r + R) * G::HS + c0 + R + px];
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      float vsum = bs[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
      if ((tid & 31) == 0 && (DEPTH == 2 || j < 4)) atomicAdd(sh_gb + j, vsum);
    }
  }

  // ---- weight gradients: thread = (combo, row group); KX x 4 accumulators each ----
  {
    const int combo = tid % BG::NCOMBO, grp = tid / BG::NCOMBO;
    if (grp < BG::NGROUP) {
      // A: 4 "output side" channels at the tile pixel; Bv: one "input side" channel shifted by (ky, kx)
      const float* A;
      const float* Bv;
      int a_cs, a_rs, b_rs;
      if (DEPTH == 2) {
        if (combo < 8 * K) {  // layer 2: d w2[co][ci][ky][kx] += d pre[co][p] * h1[ci][p + k - R]
          const int ci = combo / K, ky = combo - ci * K;
          A = sh_dp + (2 * R) * G::IS + 2 * R, a_cs = G::IR * G::IS, a_rs = G::IS;
          Bv = sh_h + (ci * G::HR + ky) * G::HS, b_rs = G::HS;
        } else {              // layer 1: d w1[stack][co][cj][ky][kx] += d h1[stack*4+co][q] * in[stack*2+cj][q + k - R]
          const int id = combo - 8 * K;
          const int stack = id / (2 * K), cj = (id / K) & 1, ky = id % K;
          A = sh_dh + (stack * 4 * G::HR + R) * G::HS + R, a_cs = G::HR * G::HS, a_rs = G::HS;
          Bv = sh_in + ((stack * 2 + cj) * G::IR + R + ky) * G::IS + R, b_rs = G::IS;
        }
      } else {                // single layer: d w0[co][c][ky][kx] += d pre[co][p] * in[c][p + k - R]
        const int c = combo / K, ky = combo - c * K;
        A = sh_dp + R * G::HS + R, a_cs = G::HR * G::HS, a_rs = G::HS;
        Bv = sh_in + (c * G::HR + ky) * G::HS, b_rs = G::HS;
      }
      float wacc[K][4];
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
#pragma unroll
        for (int co = 0; co < 4; ++co) wacc[kx][co] = 0.0f;
      for (int rr = grp; rr < TH; rr += BG::NGROUP) {
#pragma unroll 2
        for (int cc = 0; cc < TW; cc += PX) {
          float av[4][PX], bv[PX + K - 1];
#pragma unroll
          for (int co = 0; co < 4; ++co)
#pragma unroll
            for (int px = 0; px < PX; ++px) av[co][px] = A[co * a_cs + rr * a_rs + cc + px];
#pragma unroll
          for (int j = 0; j < PX + K - 1; ++j) bv[j] = Bv[rr * b_rs + cc + j];
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int co = 0; co < 4; ++co)
#pragma unroll
              for (int px = 0; px < PX; ++px) wacc[kx][co] = fmaf(av[co][px], bv[px + kx], wacc[kx][co]);
        }
      }
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
#pragma unroll
        for (int co = 0; co < 4; ++co) atomicAdd(sh_gw + (combo * K + kx) * 4 + co, wacc[kx][co]);
    }
  }
  __syncthreads();

  // ---- flush the tile's gradient sums into the caller's buffers (PyTorch layouts) ----
  for (int i = tid; i < BG::NGW; i += G::NT) {
    const int co = i & 3;
    const int kx = (i >> 2) % K;
    const int combo = (i >> 2) / K;
    float* dst;
    if (DEPTH == 2) {
      if (combo < 8 * K) {
        const int ci = combo / K, ky = combo - ci * K;
        float* base = ci < 4 ? a.g.in_w1 : a.g.gate_w1;
        dst = base + ((co * 4 + (ci & 3)) * K + ky) * K + kx;
      } else {
        const int id = combo - 8 * K;
        const int stack = id / (2 * K), cj = (id / K) & 1, ky = id % K;
        float* base = stack == 0 ? a.g.in_w0 : a.g.gate_w0;
        dst = base + ((co * 2 + cj) * K + ky) * K + kx;
      }
    } else {
      const int c = combo / K, ky = combo - c * K;
      float* base = c < 2 ? a.g.in_w0 : a.g.gate_w0;
      dst = base + ((co * 2 + (c & 1)) * K + ky) * K + kx;
    }
    const float vsum = sh_gw[i];
    if (vsum != 0.0f) atomicAdd(dst, vsum);
  }
  if (tid < 12) {
    const float vsum = sh_gb[tid];
    if (DEPTH == 2) {
      if (tid < 4) {
        atomicAdd(a.g.in_b1 + tid, vsum);
        atomicAdd(a.g.gate_b1 + tid, vsum);
      } else if (tid < 8) {
        atomicAdd(a.g.in_b0 + (tid - 4), vsum);
      } else {
        atomicAdd(a.g.gate_b0 + (tid - 8), vsum);
      }
    } else if (tid < 4) {
      atomicAdd(a.g.in_b0 + tid, vsum);
      atomicAdd(a.g.gate_b0 + tid, vsum);
    }
  }
}

template <int K, int DEPTH, typename IN_T>
int launch_bwd(const eas_sampler_cfg* c, BwdArgs a, cudaStream_t st) {
  // 16 x 32 tiles: ~98 KB of shared memory, i.e. TWO resident CTAs per SM whose load / recompute / gradient phases
  // overlap.  With 16 x 64 tiles (164 KB, one CTA of 8 warps per SM) ncu showed 12 % of the warp slots active and the
  // kernel latency bound; the narrower tile recomputes 11 % more halo and is still 8 % (B = 64) to 17 % (B = 8) faster.
  constexpr int TH = 16, TW = 32;
  using G = Geo<K, DEPTH, TH, TW>;
  using BG = BGeo<K, DEPTH, TH, TW>;
  static_assert(BG::NGROUP >= 1, "not enough threads for the weight-gradient combos");
  auto kern = sampler_bwd_conv_kernel<K, DEPTH, TH, TW, IN_T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BG::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int64_t tiles = (int64_t)((c->W + TW - 1) / TW) * ((c->H + TH - 1) / TH) * c->B;
  EAS_REQUIRE(tiles < (1ll << 31), EAS_E_SHAPE);
  const int64_t n = (int64_t)c->B * 2 * c->H * c->W;
  for (int t = c->Tm - 1; t >= 0; --t) {
    a.t = t;
    sampler_bwd_point_kernel<<<(unsigned)eas_ceil_div(n, 256), 256, 0, st>>>(a);
    EAS_LAUNCH_CHECK();
    kern<<<(unsigned)tiles, G::NT, BG::SMEM, st>>>(a);
    EAS_LAUNCH_CHECK();
  }
  return EAS_OK;
}

template <typename IN_T>
int dispatch_bwd(const eas_sampler_cfg* c, const BwdArgs& a, cudaStream_t st) {
  if (c->depth == 2) {
    if (c->ksize == 3) return launch_bwd<3, 2, IN_T>(c, a, st);
    if (c->ksize == 5) return launch_bwd<5, 2, IN_T>(c, a, st);
    if (c->ksize == 7) return launch_bwd<7, 2, IN_T>(c, a, st);
  } else {
    if (c->ksize == 3) return launch_bwd<3, 1, IN_T>(c, a, st);
    if (c->ksize == 5) return launch_bwd<5, 1, IN_T>(c, a, st);
    if (c->ksize == 7) return launch_bwd<7, 1, IN_T>(c, a, st);
  }
  return EAS_E_UNSUPPORTED;
}

size_t seg_bytes(const eas_sampler_cfg* c, int ch) {
  return eas_align_up((size_t)c->B * ch * c->H * c->W * sizeof(float), 256);
}

}  // namespace

extern "C" size_t eas_sampler_bwd_ws_bytes(const eas_sampler_cfg* c) {
  if (!c || c->B < 0 || c->H <= 0 || c->W <= 0) return 0;
  // d_vm, d_acc, d_s ([B][2][H][W]) + d pre ([B][4][H][W])
  return 3 * seg_bytes(c, 2) + seg_bytes(c, 4) + 256;
}

extern "C" int eas_sampler_bwd(const eas_sampler_cfg* c, const void* events, const eas_sampler_weights* w,
                               const float* v_seq, const float* gate_seq, const float* grad_out,
                               const eas_sampler_grads* gw, float* grad_events, void* ws, size_t ws_bytes,
                               void* stream) {
  EAS_REQUIRE(c, EAS_E_NULL);
  EAS_REQUIRE(c->B >= 0 && c->H > 0 && c->W > 0, EAS_E_SHAPE);
  EAS_REQUIRE(c->Tm >= 1 && c->Tm <= 254 && c->Ts >= 1 && c->Ts <= 254, EAS_E_SHAPE);
  EAS_REQUIRE(c->depth == 1 || c->depth == 2, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->ksize == 3 || c->ksize == 5 || c->ksize == 7, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->readout >= EAS_READOUT_SUM && c->readout <= EAS_READOUT_AVG, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->in_dtype == EAS_F32 || c->in_dtype == EAS_I32, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(events && w && v_seq && gate_seq && grad_out && gw && ws, EAS_E_NULL);
  EAS_REQUIRE(w->in_w0 && w->in_b0 && w->gate_w0 && w->gate_b0, EAS_E_NULL);
  EAS_REQUIRE(gw->in_w0 && gw->in_b0 && gw->gate_w0 && gw->gate_b0, EAS_E_NULL);
  if (c->depth == 2) {
    EAS_REQUIRE(w->in_w1 && w->in_b1 && w->gate_w1 && w->gate_b1, EAS_E_NULL);
    EAS_REQUIRE(gw->in_w1 && gw->in_b1 && gw->gate_w1 && gw->gate_b1, EAS_E_NULL);
  }
  EAS_REQUIRE(ws_bytes >= eas_sampler_bwd_ws_bytes(c), EAS_E_WORKSPACE);
  EAS_REQUIRE((uintptr_t)ws % 16 == 0, EAS_E_ALIGN);
  cudaStream_t st = (cudaStream_t)stream;
  // the gradient buffers are accumulated into with atomics: clear them first
  const size_t k2 = (size_t)c->ksize * c->ksize;
  cudaError_t e;
  if ((e = cudaMemsetAsync(gw->in_w0, 0, 4 * 2 * k2 * 4, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(gw->gate_w0, 0, 4 * 2 * k2 * 4, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(gw->in_b0, 0, 16, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(gw->gate_b0, 0, 16, st)) != cudaSuccess) return (int)e;
  if (c->depth == 2) {
    if ((e = cudaMemsetAsync(gw->in_w1, 0, 4 * 4 * k2 * 4, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemsetAsync(gw->gate_w1, 0, 4 * 4 * k2 * 4, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemsetAsync(gw->in_b1, 0, 16, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemsetAsync(gw->gate_b1, 0, 16, st)) != cudaSuccess) return (int)e;
  }
  if (c->B == 0) return EAS_OK;
  char* p = (char*)ws;
  BwdArgs a{};
  a.events = events, a.v_seq = v_seq, a.gate_seq = gate_seq, a.grad_out = grad_out;
  a.d_vm = (float*)p;
  p += seg_bytes(c, 2);
  a.d_acc = (float*)p;
  p += seg_bytes(c, 2);
  a.d_s = (float*)p;
  p += seg_bytes(c, 2);
  a.dpre = (float*)p;
  a.grad_events = grad_events;
  a.w = *w, a.g = *gw;
  a.B = c->B, a.H = c->H, a.W = c->W, a.Tm = c->Tm, a.Ts = c->Ts;
  a.readout = c->readout, a.hard_reset = c->hard_reset, a.write_zero = c->write_zero, a.use_abs = c->use_abs;
  a.spike_attach = c->spike_attach, a.vreset = c->vreset, a.thresh = c->thresh;
  a.surr_alpha = c->surr_alpha > 0.0f ? c->surr_alpha : 1.0f;
  a.surr_half = 0.5f / a.surr_alpha;
  if (c->in_dtype == EAS_F32) return dispatch_bwd<float>(c, a, st);
  return dispatch_bwd<int32_t>(c, a, st);
}
