// (a-2) Adaptive event sampler, forward.  Replaces AdaptiveRSNNEmbedding.forward / update
// (yolox/models/embedding.py:132-226) with the Rectangle spike function (activation.py:17-30).
//
// One launch per sampler step t (the launch boundary is the grid-wide dependency the recurrent
// gate_conv(spike_{t-1}) needs: its two stacked KxK convolutions see a (4R+1)^2 neighbourhood of
// the previous step's spikes).  A launch is a persistent grid (2 CTAs per SM) walking TH x TW pixel
// tiles; per tile a CTA fuses everything the reference does in ~35 ATen launches and 3 host syncs:
//   load  : micro-bin counts (fp32, or the int32 histogram straight from eas_bin_events) and the
//           previous spikes, tile + R*depth halo, zero padded, by cp.async (zfill) into shared
//           memory -- issued one tile ahead, so the copy overlaps the previous tile's conv 2
//   conv 1: input_conv[0] (2->4) and gate_conv[0] (2->4) + ReLU on tile + R halo -> shared memory
//   conv 2: input_conv[2] + gate_conv[2] as ONE 8->4 convolution (the reference adds the two
//           stacks' outputs, embedding.py:175-176), 4 x 8 register tile per thread
//   update: sigmoid gate, membrane update, strict threshold, reset, running no-reset sum,
//           spike-triggered read-out into agg[seg], seg/t_last bookkeeping; on the last step the
//           residual write (RPD: write_zero) and the optional ReLU.  The per-pixel state tile is
//           also fetched by cp.async while the convolutions run.
// Per-pixel state between launches (caller's workspace): vm, acc, spikes x2 (f32), seg|t_last (u16)
// = 18 B per state element.  The kernel is FP32-pipe bound (2400 FLOP per pixel-step for depth 2,
// k 5), not HBM bound; see DESIGN.md.
#include <type_traits>
#include "sampler_common.cuh"
#include "hist_u8.cuh"

namespace {

using namespace eas_sampler;

// Barrier over the CTAs of a cooperative launch (all resident): a monotonic counter, `epoch` counts this CTA's passes.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    __threadfence();
    atomicAdd(bar, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(bar) < epoch * nblocks) {
    }
    __threadfence();
  }
  __syncthreads();
}

// MULTI: the cooperative all-steps form (fall-back only); false compiles to exactly the one-step kernel.
template <int K, int DEPTH, int TH, int TW, typename IN_T, bool VEC, bool MULTI>
__global__ void __launch_bounds__(TH* TW / 4, 2)
sampler_step_kernel(const StepArgs a) {
  using G = Geo<K, DEPTH, TH, TW>;
  constexpr int R = G::R;
  constexpr int PX = G::PX;
  constexpr bool kInt = !std::is_same<IN_T, float>::value;
  extern __shared__ __align__(16) float smem[];
  float* sh_h = smem;                 // layer-2 input
  float* sh_i = sh_h + G::SM_H;       // layer-1 input (depth 2)
  float* sh_w2 = sh_i + G::SM_I;      // [CI2][K][K][4]
  float* sh_w1 = sh_w2 + G::W2;       // [2 stacks][2][K][K][4]
  float* sh_vm = sh_w2 + G::SM_W;     // [2][TH][TW]
  float* sh_acc = sh_vm + G::SM_ST;
  uint16_t* sh_meta = reinterpret_cast<uint16_t*>(sh_acc + G::SM_ST);
  float* sh_l = DEPTH == 2 ? sh_i : sh_h;  // the tile filled from global memory
  __shared__ float sh_b[16];          // [0..3] layer-2 bias sum, [4..7] in b0, [8..11] gate b0

  if (a.run_if != nullptr && *a.run_if == 0) return;  // fall-back launch that is not needed
  const int tid = threadIdx.x;
  unsigned int bar_epoch = 0;
  if (MULTI && a.expand_src != nullptr) {   // (cooperative launch) compact byte histogram -> the dense counts this kernel reads
    hist_u8_expand_f32(reinterpret_cast<const uint8_t*>(a.expand_src), (int64_t)a.B * a.Tm * 2 * a.H * a.W,
                       reinterpret_cast<float*>(const_cast<void*>(a.events)));
    grid_barrier(a.grid_bar, gridDim.x, bar_epoch);
  }
  const int tiles_x = (a.W + TW - 1) / TW;
  const int tiles_y = (a.H + TH - 1) / TH;
  const int ntiles = tiles_x * tiles_y * a.B;
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t BHW2 = (int64_t)a.B * 2 * HW;
  const IN_T* ev_base = reinterpret_cast<const IN_T*>(a.events);

  // ---- weights -> shared (once per CTA), re-laid out as [ci][ky][kx][co] ----------------------
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

  // One launch = one sampler step; as the predicated fall-back one cooperative launch runs all of them with a grid
  // barrier in between (the recurrent gate conv reads the spikes every CTA wrote in the step before).
  const int t_end = MULTI ? a.t + a.t_count : a.t + 1;
  for (int t = a.t; t < t_end; ++t) {
    const bool first = t == 0, last = t == a.Tm - 1;
    const int tm = a.Tm - 1 - t;  // newest micro-bin first (embedding.py:155-156)
    const float* s_prev = MULTI ? ((t & 1) ? a.s0 : a.s1) : a.s_prev;  // step t reads what step t-1 wrote
    float* s_next = MULTI ? ((t & 1) ? a.s1 : a.s0) : a.s_next;
    // ---- tile loaders ---------------------------------------------------------------------------
    // events (raw 4-byte words) and previous spikes, tile + halo, zero filled outside the image
    auto issue_in = [&](int tile) {
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int b = tile / (tiles_x * tiles_y);
      const int gx0 = tx * TW - G::HALO, gy0 = ty * TH - G::HALO;
      const IN_T* ev = ev_base + ((int64_t)b * a.Tm + tm) * 2 * HW;
      const float* sp = s_prev + (int64_t)b * 2 * HW;
      const int nch = first ? 2 : 4;  // step 0: previous spikes are all zero, nothing to load
      constexpr bool kGran4 = VEC && (G::HALO % 4 == 0) && (G::LC % 4 == 0);
      if (kGran4) {
        constexpr int GPR = G::LC / 4;
        for (int i = tid; i < nch * G::LR * GPR; i += G::NT) {
          const int c = i / (G::LR * GPR);
          const int rem = i - c * (G::LR * GPR);
          const int r = rem / GPR, g4 = (rem - r * GPR) * 4;
          const int gy = gy0 + r, gx = gx0 + g4;
          const bool ok = (unsigned)gy < (unsigned)a.H && (unsigned)gx < (unsigned)a.W;
          const int64_t off = ok ? (int64_t)gy * a.W + gx : 0;
          const void* src = c < 2 ? (const void*)(ev + c * HW + off) : (const void*)(sp + (c - 2) * HW + off);
          cp_async_16(sh_l + (c * G::LR + r) * G::LS + g4, src, ok);
        }
      } else {
        for (int i = tid; i < nch * G::LR * G::LC; i += G::NT) {
          const int c = i / (G::LR * G::LC);
          const int rem = i - c * (G::LR * G::LC);
          const int r = rem / G::LC, cc = rem - r * G::LC;
          const int gy = gy0 + r, gx = gx0 + cc;
          const bool ok = (unsigned)gy < (unsigned)a.H && (unsigned)gx < (unsigned)a.W;
          const int64_t off = ok ? (int64_t)gy * a.W + gx : 0;
          const void* src = c < 2 ? (const void*)(ev + c * HW + off) : (const void*)(sp + (c - 2) * HW + off);
          cp_async_4(sh_l + (c * G::LR + r) * G::LS + cc, src, ok);
        }
      }
    };
    // vm, acc (f32) and seg|t_last (u16) of the tile
    auto issue_state = [&](int tile) {
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int b = tile / (tiles_x * tiles_y);
      if (VEC) {
        constexpr int GPR = TW / 4;
        for (int i = tid; i < 2 * TH * GPR; i += G::NT) {
          const int c = i / (TH * GPR);
          const int rem = i - c * (TH * GPR);
          const int r = rem / GPR, g4 = (rem - r * GPR) * 4;
          const int gy = ty * TH + r, gx = tx * TW + g4;
          const bool ok = gy < a.H && gx < a.W;
          const int64_t e = ok ? ((int64_t)b * 2 + c) * HW + (int64_t)gy * a.W + gx : 0;
          const int so = (c * TH + r) * TW + g4;
          cp_async_16(sh_vm + so, a.vm + e, ok);
          cp_async_16(sh_acc + so, a.acc + e, ok);
          cp_async_8(sh_meta + so, a.meta + e, ok);
        }
      } else {
        for (int i = tid; i < 2 * TH * TW; i += G::NT) {
          const int c = i / (TH * TW);
          const int rem = i - c * (TH * TW);
          const int r = rem / TW, cc = rem - r * TW;
          const int gy = ty * TH + r, gx = tx * TW + cc;
          const bool ok = gy < a.H && gx < a.W;
          const int64_t e = ok ? ((int64_t)b * 2 + c) * HW + (int64_t)gy * a.W + gx : 0;
          sh_vm[i] = ok ? a.vm[e] : 0.0f;
          sh_acc[i] = ok ? a.acc[e] : 0.0f;
          sh_meta[i] = ok ? a.meta[e] : (uint16_t)0;
        }
      }
    };

    int tile = blockIdx.x;
    if (tile < ntiles) issue_in(tile);
    cp_async_commit();

    for (; tile < ntiles; tile += gridDim.x) {
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int b = tile / (tiles_x * tiles_y);
      const int x0 = tx * TW, y0 = ty * TH;
      if (!first) issue_state(tile);
      cp_async_commit();
      cp_async_wait<1>();  // this tile's input has landed (the state group may still be in flight)
      __syncthreads();

      if (kInt || first) {
        // int32 counts -> fp32 in place; on step 0 also clear the spike channels
        for (int i = tid; i < 4 * G::LR * G::LC; i += G::NT) {
          const int c = i / (G::LR * G::LC);
          const int rem = i - c * (G::LR * G::LC);
          const int r = rem / G::LC, cc = rem - r * G::LC;
          float* q = sh_l + (c * G::LR + r) * G::LS + cc;
          if (c < 2) {
            if (kInt) *q = (float)__float_as_int(*q);
          } else if (first) {
            *q = 0.0f;
          }
        }
        __syncthreads();
      }

      // ---- layer 1 (depth 2): 2->4 per stack, bias, ReLU, zero outside the image ----------------
      if (DEPTH == 2) {
        constexpr int PX1 = 4;
        constexpr int SPR = G::HC / PX1;            // strips per row
        constexpr int NITEM = 2 * G::HR * SPR;      // (stack, row, strip)
        for (int idx = tid; idx < NITEM; idx += G::NT) {
          const int stack = idx / (G::HR * SPR);
          const int rem = idx - stack * (G::HR * SPR);
          const int r = rem / SPR;
          const int c0 = (rem - r * SPR) * PX1;
          const int gy = y0 - R + r;
          const bool row_in = (unsigned)gy < (unsigned)a.H;
          unsigned long long accp[2][PX1];
  #pragma unroll
          for (int h = 0; h < 2; ++h)
  #pragma unroll
            for (int px = 0; px < PX1; ++px) accp[h][px] = 0ull;
          if (row_in && !(stack == 1 && first))
            conv_acc<2, 4, K, PX1>(sh_i + (stack * 2 * G::IR + r) * G::IS + c0, G::IR * G::IS, G::IS,
                                   sh_w1 + stack * G::W1, accp);
          float acc[4][PX1];
  #pragma unroll
          for (int px = 0; px < PX1; ++px) {
            unpack2(accp[0][px], acc[0][px], acc[1][px]);
            unpack2(accp[1][px], acc[2][px], acc[3][px]);
          }
          const int gxs = x0 - R + c0;
          float* hp = sh_h + (stack * 4 * G::HR + r) * G::HS + c0;
          if (row_in && gxs >= 0 && gxs + PX1 <= a.W) {  // strip fully inside the image (the common case)
  #pragma unroll
            for (int co = 0; co < 4; ++co) {
              const float bias = sh_b[4 + stack * 4 + co];
              *reinterpret_cast<float4*>(hp + co * G::HR * G::HS) =
                  make_float4(fmaxf(acc[co][0] + bias, 0.0f), fmaxf(acc[co][1] + bias, 0.0f),
                              fmaxf(acc[co][2] + bias, 0.0f), fmaxf(acc[co][3] + bias, 0.0f));
            }
          } else {
  #pragma unroll
            for (int co = 0; co < 4; ++co) {
              const float bias = sh_b[4 + stack * 4 + co];
              float4 o;
              float* op = reinterpret_cast<float*>(&o);
  #pragma unroll
              for (int px = 0; px < PX1; ++px) {
                const bool in_img = row_in && (unsigned)(gxs + px) < (unsigned)a.W;
                op[px] = in_img ? fmaxf(acc[co][px] + bias, 0.0f) : 0.0f;
              }
              *reinterpret_cast<float4*>(hp + co * G::HR * G::HS) = o;
            }
          }
        }
        __syncthreads();
        // the layer-1 input buffer is free: fetch the next tile while layer 2 runs
        if (tile + (int)gridDim.x < ntiles) issue_in(tile + gridDim.x);
        cp_async_commit();
      }

      // ---- layer 2: (8|4) -> 4, one 4 x 8 register tile per thread ------------------------------
      const int r = tid / (TW / PX);
      const int c0 = (tid - r * (TW / PX)) * PX;
      float acc2[4][PX];
      {
        unsigned long long accp[2][PX];
  #pragma unroll
        for (int h = 0; h < 2; ++h)
  #pragma unroll
          for (int px = 0; px < PX; ++px) accp[h][px] = 0ull;
        conv_acc<(DEPTH == 2 ? 8 : 4), 4, K, PX>(sh_h + r * G::HS + c0, G::HR * G::HS, G::HS, sh_w2, accp);
  #pragma unroll
        for (int px = 0; px < PX; ++px) {
          unpack2(accp[0][px], acc2[0][px], acc2[1][px]);
          unpack2(accp[1][px], acc2[2][px], acc2[3][px]);
        }
      }

      if (DEPTH == 1) {
        __syncthreads();  // everyone is done reading the raw tile
        if (tile + (int)gridDim.x < ntiles) issue_in(tile + gridDim.x);
        cp_async_commit();
      }
      cp_async_wait<1>();  // this tile's state has landed (the next tile's input may be in flight)
      __syncthreads();

      // ---- membrane update + spike-triggered aggregation (embedding.py:132-139, 177-217) --------
      // Every agg[k] element is written exactly once: by the k-th valid spike of its pixel or by the
      // residual write at the end (the reference's "+=" always lands on a still-zero element), so
      // the read-out is a plain store and ReLU (abs) can be applied at write time.
      const int gy = y0 + r;
      const int nv = min(PX, a.W - (x0 + c0));  // valid pixels of this strip (<= 0: none)
      if (gy < a.H && nv > 0) {
  #pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float bg = sh_b[c], bc = sh_b[2 + c];
          const int64_t base = ((int64_t)b * 2 + c) * HW + (int64_t)gy * a.W + x0 + c0;
          const int so = (c * TH + r) * TW + c0;
          __align__(16) float vm4[PX], ac4[PX], v4[PX], g4[PX], s4[PX], o4[PX];
          __align__(8) uint16_t m4[PX];
          if (first) {
  #pragma unroll
            for (int px = 0; px < PX; ++px) vm4[px] = 0.0f, ac4[px] = 0.0f, m4[px] = 0;
          } else {
            *reinterpret_cast<float4*>(vm4) = *reinterpret_cast<const float4*>(sh_vm + so);
            *reinterpret_cast<float4*>(ac4) = *reinterpret_cast<const float4*>(sh_acc + so);
            *reinterpret_cast<uint2*>(m4) = *reinterpret_cast<const uint2*>(sh_meta + so);
          }
          float* outp = a.out + base;  // plane k at outp + k*BHW2
  #pragma unroll
          for (int px = 0; px < PX; ++px) {
            const float gate = __fdividef(1.0f, 1.0f + __expf(-(acc2[c][px] + bg)));
            const float cur = acc2[2 + c][px] + bc;
            int seg = m4[px] & 0xff;
            int tl = (int)(m4[px] >> 8) - 1;
            const float v = __fadd_rn(__fmul_rn(gate, vm4[px]), cur);
            const bool s = __fsub_rn(v, a.thresh) > 0.0f;
            const float vm = s ? (a.hard_reset ? a.vreset : __fsub_rn(v, a.thresh)) : v;
            float ac = __fadd_rn(ac4[px], v);
            const bool valid = s && seg < a.Ts;
            float val = a.readout == EAS_READOUT_SUM ? ac : vm;
            if (a.readout == EAS_READOUT_AVG) val = ac / (float)(t - tl);
            if (a.use_abs) val = fmaxf(val, 0.0f);
            o4[px] = valid ? val : 0.0f;        // plane 0 on the first step
            if (!first && valid && px < nv) outp[(int64_t)seg * BHW2 + px] = val;
            seg += valid ? 1 : 0;
            tl = valid ? t : tl;
            ac = s ? 0.0f : ac;
            if (last && !s && seg < a.Ts && !a.write_zero && px < nv) {
              float tv = a.readout == EAS_READOUT_SUM ? ac : vm;
              if (a.readout == EAS_READOUT_AVG) tv = ac / (float)(a.Tm - 1 - tl);
              if (a.use_abs) tv = fmaxf(tv, 0.0f);
              if (first && seg == 0) o4[px] = tv;   // Tm == 1: still inside the zero-initialising store
              else outp[(int64_t)seg * BHW2 + px] = tv;
            }
            vm4[px] = vm, ac4[px] = ac, v4[px] = v, g4[px] = gate, s4[px] = s ? 1.0f : 0.0f;
            m4[px] = (uint16_t)(seg | ((tl + 1) << 8));
          }
          if (VEC) {  // nv is 4 here (W % 4 == 0)
            if (first) {
              *reinterpret_cast<float4*>(outp) = *reinterpret_cast<const float4*>(o4);
              for (int k = 1; k < a.Ts; ++k) *reinterpret_cast<float4*>(outp + k * BHW2) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (!last) {
              *reinterpret_cast<float4*>(a.vm + base) = *reinterpret_cast<const float4*>(vm4);
              *reinterpret_cast<float4*>(a.acc + base) = *reinterpret_cast<const float4*>(ac4);
              *reinterpret_cast<uint2*>(a.meta + base) = *reinterpret_cast<const uint2*>(m4);
              *reinterpret_cast<float4*>(s_next + base) = *reinterpret_cast<const float4*>(s4);
            }
            if (a.v_seq) {
              const int64_t se = (int64_t)t * BHW2 + base;
              *reinterpret_cast<float4*>(a.v_seq + se) = *reinterpret_cast<const float4*>(v4);
              *reinterpret_cast<float4*>(a.gate_seq + se) = *reinterpret_cast<const float4*>(g4);
            }
          } else {
  #pragma unroll
            for (int px = 0; px < PX; ++px) {
              if (px < nv) {
                if (first) {
                  outp[px] = o4[px];
                  for (int k = 1; k < a.Ts; ++k) outp[k * BHW2 + px] = 0.0f;
                }
                if (!last) {
                  a.vm[base + px] = vm4[px];
                  a.acc[base + px] = ac4[px];
                  a.meta[base + px] = m4[px];
                  s_next[base + px] = s4[px];
                }
                if (a.v_seq) {
                  const int64_t se = (int64_t)t * BHW2 + base + px;
                  a.v_seq[se] = v4[px];
                  a.gate_seq[se] = g4[px];
                }
              }
            }
          }
        }
      }
      __syncthreads();  // state tile and layer-2 input are free for the next tile
    }
    cp_async_wait<0>();
    if (MULTI && t + 1 < t_end) grid_barrier(a.grid_bar, gridDim.x, bar_epoch);
  }
}

template <int K, int DEPTH, typename IN_T, bool VEC>
int launch_steps(const eas_sampler_cfg* cfg, StepArgs a, float* s0, float* s1, cudaStream_t st) {
  constexpr int TH = 16, TW = 64;
  using G = Geo<K, DEPTH, TH, TW>;
  auto kern = sampler_step_kernel<K, DEPTH, TH, TW, IN_T, VEC, false>;
  auto kern_all = sampler_step_kernel<K, DEPTH, TH, TW, IN_T, VEC, true>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(kern_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
  if (e != cudaSuccess) return (int)e;
  const int64_t tiles = (int64_t)((cfg->W + TW - 1) / TW) * ((cfg->H + TH - 1) / TH) * cfg->B;
  EAS_REQUIRE(tiles < (1ll << 31), EAS_E_SHAPE);
  const int per_sm = G::SMEM + 1024 <= 113 * 1024 ? 2 : 1;
  const int64_t grid = tiles < (int64_t)EAS_NUM_SMS * per_sm ? tiles : (int64_t)EAS_NUM_SMS * per_sm;
  if (a.run_if != nullptr && a.grid_bar != nullptr) {
    // predicated fall-back: ONE cooperative launch over all steps (an idle launch costs ~3.5 us of stream time; Tm + 1
    // of them behind every tensor-core forward were 3.6 % of it)
    a.t = 0, a.t_count = cfg->Tm, a.s0 = s0, a.s1 = s1;
    void* kargs[] = {(void*)&a};
    cudaError_t ce = cudaLaunchCooperativeKernel((const void*)kern_all, dim3((unsigned)grid), dim3(G::NT), kargs, G::SMEM, st);
    if (ce == cudaSuccess) return EAS_OK;
    (void)cudaGetLastError();   // not available here: the step launches below do the same work
    if (a.expand_src != nullptr) {
      int rc = eas_hist_u8_expand_if(a.expand_src, (int64_t)cfg->B * cfg->Tm * 2 * cfg->H * cfg->W,
                                     const_cast<void*>(a.events), EAS_F32, a.run_if, st);
      if (rc) return rc;
    }
    a.t_count = 0, a.expand_src = nullptr;
  }
  for (int t = 0; t < cfg->Tm; ++t) {
    a.t = t;
    a.s_prev = (t & 1) ? s0 : s1;  // step t reads what step t-1 wrote
    a.s_next = (t & 1) ? s1 : s0;
    kern<<<(unsigned)grid, G::NT, G::SMEM, st>>>(a);
    EAS_LAUNCH_CHECK();
  }
  return EAS_OK;
}

template <typename IN_T, bool VEC>
int dispatch(const eas_sampler_cfg* c, const StepArgs& a, float* s0, float* s1, cudaStream_t st) {
  if (c->depth == 2) {
    if (c->ksize == 3) return launch_steps<3, 2, IN_T, VEC>(c, a, s0, s1, st);
    if (c->ksize == 5) return launch_steps<5, 2, IN_T, VEC>(c, a, s0, s1, st);
    if (c->ksize == 7) return launch_steps<7, 2, IN_T, VEC>(c, a, s0, s1, st);
  } else {
    if (c->ksize == 3) return launch_steps<3, 1, IN_T, VEC>(c, a, s0, s1, st);
    if (c->ksize == 5) return launch_steps<5, 1, IN_T, VEC>(c, a, s0, s1, st);
    if (c->ksize == 7) return launch_steps<7, 1, IN_T, VEC>(c, a, s0, s1, st);
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
  EAS_REQUIRE(c->in_dtype == EAS_F32 || c->in_dtype == EAS_I32 || c->in_dtype == EAS_U8, EAS_E_UNSUPPORTED);
  if (c->in_dtype == EAS_U8) EAS_REQUIRE((int64_t)c->B * c->Tm * 2 * c->H * c->W < (1ll << 32), EAS_E_SHAPE);
  EAS_REQUIRE(c->algo >= EAS_SAMPLER_AUTO && c->algo <= EAS_SAMPLER_TENSOR_SPLIT, EAS_E_UNSUPPORTED);
  return EAS_OK;
}

}  // namespace

extern "C" size_t eas_sampler_fwd_ws_bytes(const eas_sampler_cfg* c) {
  if (check(c) != EAS_OK) return 0;
  const size_t n = (size_t)c->B * 2 * c->H * c->W;
  // vm, acc, s0, s1 (f32) + meta (u16), each segment 256 B aligned
  // + the packed weight tiles of the tensor-core path
  const size_t wimg = eas_sampler_tc_wimg_bytes() > eas_sampler_tc2_wimg_bytes() ? eas_sampler_tc_wimg_bytes()
                                                                               : eas_sampler_tc2_wimg_bytes();
  // + for the compact byte histogram: room for its dense fp32 form, which only the kernels that cannot read bytes
  //   (FP32-pipe fall-back, first tensor-core kernel) ever touch
  const size_t dense = c->in_dtype == EAS_U8 ? eas_align_up(n * (size_t)c->Tm * 4, 256) : 0;
  return 4 * eas_align_up(n * 4, 256) + eas_align_up(n * 2, 256) + 256 + eas_align_up(wimg, 256) + dense;
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
  EAS_REQUIRE((uintptr_t)ws % 16 == 0 && (uintptr_t)events % 4 == 0, EAS_E_ALIGN);
  const size_t n = (size_t)c->B * 2 * c->H * c->W;
  char* p = (char*)ws;
  StepArgs a{};
  a.events = events;
  a.vm = (float*)p;
  p += eas_align_up(n * 4, 256);
  a.acc = (float*)p;
  p += eas_align_up(n * 4, 256);
  float* s0 = (float*)p;
  p += eas_align_up(n * 4, 256);
  float* s1 = (float*)p;
  p += eas_align_up(n * 4, 256);
  a.meta = (uint16_t*)p;
  p += eas_align_up(n * 2, 256);
  void* wimg = (void*)p;
  p += eas_align_up(eas_sampler_tc_wimg_bytes() > eas_sampler_tc2_wimg_bytes() ? eas_sampler_tc_wimg_bytes()
                                                                              : eas_sampler_tc2_wimg_bytes(), 256) + 256;
  float* dense = (float*)p;   // in_dtype EAS_U8 only
  a.out = out;
  a.v_seq = v_seq;
  a.gate_seq = gate_seq;
  a.w = *w;
  a.B = c->B, a.H = c->H, a.W = c->W, a.Tm = c->Tm, a.Ts = c->Ts;
  a.readout = c->readout, a.hard_reset = c->hard_reset, a.write_zero = c->write_zero, a.use_abs = c->use_abs;
  a.vreset = c->vreset, a.thresh = c->thresh;
  cudaStream_t st = (cudaStream_t)stream;
  // depth 2, k 5 (the published configuration): tensor-core kernels.
  //   TENSOR / AUTO : row-folded kernel (sampler_tc2.cu; inputs exact in one fp16 plane = event counts)
  //   TENSOR_SPLIT  : first kernel (sampler_tc.cu; hi + lo input planes = real-valued inputs such as
  //                   letterboxed frames), also what AUTO uses when only it supports the shape
  // AUTO: inputs the kernel cannot hold exactly (or magnitudes beyond the fp16 range, never seen on event
  // data) raise a device flag; the FP32-pipe launches below then recompute the whole step sequence,
  // otherwise they exit at once.
  const bool tc2_ok = eas_sampler_tc2_supported(c, events, out, v_seq, gate_seq);
  const bool tc1_ok = c->in_dtype != EAS_U8 && eas_sampler_tc_supported(c, events, out, v_seq, gate_seq);
  if (c->algo == EAS_SAMPLER_TENSOR) EAS_REQUIRE(tc2_ok, EAS_E_UNSUPPORTED);
  if (c->algo == EAS_SAMPLER_TENSOR_SPLIT) EAS_REQUIRE(tc1_ok, EAS_E_UNSUPPORTED);
  if (tc2_ok && (c->algo == EAS_SAMPLER_AUTO || c->algo == EAS_SAMPLER_TENSOR)) {
    // the compact state lives in the same workspace segments: spike bytes in s0 / s1, one meta byte per element
    rc = eas_sampler_tc2_run(c, a, (uint8_t*)a.meta, (uint8_t*)s0, (uint8_t*)s1, wimg, st);
    if (rc != EAS_OK || c->algo == EAS_SAMPLER_TENSOR) return rc;
    a.run_if = eas_sampler_tc2_flag(wimg);
    a.grid_bar = reinterpret_cast<unsigned int*>(const_cast<int*>(a.run_if)) + 1;   // zeroed by the weight pack
  } else if (tc1_ok && c->algo != EAS_SAMPLER_FP32) {
    rc = eas_sampler_tc_run(c, a, s0, s1, wimg, st);
    if (rc != EAS_OK || c->algo == EAS_SAMPLER_TENSOR_SPLIT) return rc;
    a.run_if = eas_sampler_tc_flag(wimg);
    a.grid_bar = reinterpret_cast<unsigned int*>(const_cast<int*>(a.run_if)) + 1;   // zeroed by the weight pack
  }
  // 16-byte vector path: rows must keep 16 B alignment (W % 4 == 0) and so must every base pointer
  const bool vec = (c->W % 4 == 0) && ((uintptr_t)events % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                   (!v_seq || ((uintptr_t)v_seq % 16 == 0 && (uintptr_t)gate_seq % 16 == 0));
  if (c->in_dtype == EAS_U8) {
    // the FP32-pipe kernel reads dense counts: expand the byte histogram (a no-op, like the step launches after it,
    // unless the tensor-core kernel raised its flag)
    a.events = dense;
    if (a.run_if != nullptr && a.grid_bar != nullptr) {
      a.expand_src = events;                     // expanded inside the one cooperative fall-back launch
    } else {
      rc = eas_hist_u8_expand_if(events, (int64_t)n * c->Tm, dense, EAS_F32, a.run_if, st);
      if (rc) return rc;
    }
    return vec ? dispatch<float, true>(c, a, s0, s1, st) : dispatch<float, false>(c, a, s0, s1, st);
  }
  if (c->in_dtype == EAS_F32)
    return vec ? dispatch<float, true>(c, a, s0, s1, st) : dispatch<float, false>(c, a, s0, s1, st);
  return vec ? dispatch<int32_t, true>(c, a, s0, s1, st) : dispatch<int32_t, false>(c, a, s0, s1, st);
}
