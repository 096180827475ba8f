// (a-2) Adaptive event sampler, forward, on the 5th-gen tensor cores (depth 2, kernel 5: the
// published EAS-SNN configuration).  Replaces AdaptiveRSNNEmbedding.forward / update
// (yolox/models/embedding.py:132-226) with the Rectangle spike function (activation.py:17-30);
// same per-step contract and workspace as the FP32-pipe kernel in sampler_fwd.cu.
//
// Why tensor cores: one sampler step is 2400 FLOP per pixel (two stacked 5x5 convolutions of both
// the input and the recurrent stack) against ~10 compulsory bytes, i.e. it is arithmetic bound; on
// the FP32 pipe that is ~400 us per step for 64 Gen1 frames.  The convolutions have only 2/8 input
// and 4 output channels, far too thin for an implicit GEMM over channels, so the GEMM is built over
// the x axis instead ("Toeplitz window"):
//   * a matrix row = one QUAD of 4 horizontally adjacent pixels, all channels, one bf16 plane: 64 B;
//   * for filter row ky the A operand of output quad m is the 8-pixel window made of quad rows
//     m and m+1 of image row y+ky: two K halves whose shared-memory descriptors start at
//     (f + ky*QPR + h) * 64 B, f = flattened quad index of the strip -- a plain start-address shift.
//     tcgen05 swizzles on absolute shared-memory address bits, so any row shift is legal as long as
//     the data was written with the same absolute-address XOR (verified on B200 by
//     scripts/umma_shift_probe.cu);
//   * B is the block-Toeplitz expansion of the 5 taps of that filter row: [N = 4 px x 4 channels]
//     x [K = 8 px x channels], built once per forward by sampler_tc_pack_weights;
//   * fp32 accuracy on bf16 tensor cores: weights are split into 3 bf16 planes (hi+mid+lo), the
//     real-valued operands (counts, hidden activations) likewise; spikes are exact.  Product terms
//     below fp32 rounding are skipped and the small terms are accumulated first (the TMEM
//     accumulator truncates).
// One persistent CTA per SM walks vertical strips of <= 116 pixels (QPR <= 31 quads per row incl.
// halo) top to bottom as a rolling pipeline over 128-quad tiles:
//   producers (4 warps): counts + previous spikes -> 3 bf16 planes -> X0 ring (zero padded)
//   MMA (1 thread)     : layer 1 (X0 -> D1 in TMEM), layer 2 (X1 -> D2 in TMEM)
//   epilogue 1 (4 warps): D1 + bias, ReLU, image mask, 3-plane split -> X1 ring
//   epilogue 2 (4 warps): D2 -> sigmoid gate, membrane update, threshold/reset, spike-triggered
//                         read-out, state write-back (same arithmetic as sampler_step_kernel)
// Rings hold S tiles plus a mirrored copy of slot 0 so that a 128-row operand never wraps.
#include "sampler_common.cuh"

namespace eas_sampler {
namespace {

constexpr int TILE = 128;       // quads (MMA rows) per tile
constexpr int ROWB = 64;        // bytes per quad row per plane (4 px x 8 ch bf16)
constexpr int MAX_QPR = 31;     // 4*QPR + 1 <= 125: an operand reaches at most one tile ahead
constexpr int S0 = 4;           // X0 ring slots
constexpr int S1 = 3;           // X1 ring slots
constexpr int RS0 = S0 * TILE, RS1 = S1 * TILE;
constexpr int X0_BYTES = (S0 + 1) * TILE * ROWB;          // + mirrored slot 0
constexpr int X1_PLANE = (S1 + 1) * TILE * ROWB;
constexpr int X1_BYTES = 3 * X1_PLANE;
constexpr int BT = 1024;        // one B tile: [16 rows (N)][32 bf16 (K)] SWIZZLE_64B
constexpr int NB_L1 = 5 * 7;    // per ky: in0.h0, in0.h1, in1, in2, gate0, gate1, gate2
constexpr int NB_L2 = 3 * 5 * 2;
constexpr int WB_BYTES = (NB_L1 + NB_L2) * BT;            // 66560
constexpr int WIMG_BYTES = WB_BYTES + 64;                  // + 12 bias floats
constexpr int OFF_X0 = WB_BYTES;
constexpr int OFF_X1 = OFF_X0 + X0_BYTES;
constexpr int OFF_BAR = OFF_X1 + X1_BYTES;
constexpr int NBAR = 2 * S0 + 2 * S1 + 8;
constexpr int OFF_MISC = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_MISC + 128 + 1024;          // + slack for the 1024 B alignment
constexpr int NUM_THREADS = 13 * 32;  // warps 0-3 epilogue 1, 4-7 epilogue 2, 8 MMA, 9-12 producers
constexpr uint32_t SPIN_LIMIT = 1u << 26;

// TMEM columns: D1 (input stack 16 + gate stack 16) x 2 buffers, D2 16 x 2 buffers
constexpr uint32_t TM_D1 = 0, TM_D2 = 64, TM_COLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > SPIN_LIMIT) __trap();  // watchdog: trap instead of hanging the GPU
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major SWIZZLE_64B operand descriptor: 64 B rows, 8-row groups 512 B apart, base_offset 0 (the
// swizzle phase comes from the absolute address, so the start may sit on any row).
__device__ __forceinline__ uint64_t sw64_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__host__ __device__ __forceinline__ float bf16_round_f(float x) {
#ifdef __CUDA_ARCH__
  return __bfloat162float(__float2bfloat16_rn(x));
#else
  return __bfloat162float(__float2bfloat16(x));
#endif
}
// x = hi + mid + lo (three bf16 values); returned as packed pairs for two inputs.
__device__ __forceinline__ void split3_pair(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const float ra = a - hf.x, rb = b - hf.y;
  const __nv_bfloat162 m = __floats2bfloat162_rn(ra, rb);
  const float2 mf = __bfloat1622float2(m);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra - mf.x, rb - mf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  mid = *reinterpret_cast<const uint32_t*>(&m);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- weight image -------------------------------------------------------------------------------
// B tiles in the order the MMA issuer walks them, each already in its SWIZZLE_64B shared-memory
// layout (tile base 1024 B aligned), followed by 12 bias floats.
//   layer 1, per ky (7 tiles):
//     in0.h : K = [hi|mid|lo|spk] chunks of half h (32)   x w_in0 plane 0 (spk chunk: 0)
//     in1   : K = [half 0: hi|mid][half 1: hi|mid]        x w_in0 plane 1
//     in2   : same, mid chunk zero                        x w_in0 plane 2
//     gate j: K = [half 0: lo|spk][half 1: lo|spk], lo: 0 x w_gate0 plane j
//   layer 2, per (j, ky, h): K = 4 px x 8 hidden channels of half h, w_in1 | w_gate1 plane j.
// N index n = 4 * (output pixel of the quad) + output channel.
__device__ __forceinline__ float bf16_plane(float w, int plane) {
  const float hi = bf16_round_f(w);
  if (plane == 0) return hi;
  const float mid = bf16_round_f(w - hi);
  if (plane == 1) return mid;
  return bf16_round_f(w - hi - mid);
}

__global__ void sampler_tc_pack_weights(const eas_sampler_weights w, uint8_t* img) {
  constexpr int K = 5;
  const int total = (NB_L1 + NB_L2) * 16 * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int tile = idx / 512, n = (idx >> 5) & 15, k = idx & 31;
    const int jpx = n >> 2, co = n & 3;
    float val = 0.0f;
    if (tile < NB_L1) {
      const int ky = tile / 7, kind = tile % 7;
      if (kind < 2) {  // in0, half h = kind
        const int c = k >> 3, p = (k & 7) >> 1, ci = k & 1, tap = 4 * kind + p - jpx;
        if (c < 3 && tap >= 0 && tap < K) val = bf16_plane(w.in_w0[((co * 2 + ci) * K + ky) * K + tap], 0);
      } else if (kind < 4) {  // in1 / in2: k = h*16 + c*8 + p*2 + ci, c in {hi, mid}
        const int h = k >> 4, c = (k >> 3) & 1, p = (k & 7) >> 1, ci = k & 1, tap = 4 * h + p - jpx;
        const int plane = kind - 1;
        if (tap >= 0 && tap < K && !(plane == 2 && c == 1))
          val = bf16_plane(w.in_w0[((co * 2 + ci) * K + ky) * K + tap], plane);
      } else {  // gate j: k = h*16 + c*8 + p*2 + ci, c in {lo (zero), spk}
        const int h = k >> 4, c = (k >> 3) & 1, p = (k & 7) >> 1, ci = k & 1, tap = 4 * h + p - jpx;
        if (c == 1 && tap >= 0 && tap < K) val = bf16_plane(w.gate_w0[((co * 2 + ci) * K + ky) * K + tap], kind - 4);
      }
    } else {
      const int t2 = tile - NB_L1;
      const int h = t2 & 1, ky = (t2 >> 1) % 5, j = t2 / 10;
      const int p = k >> 3, ci = k & 7, tap = 4 * h + p - jpx;
      if (tap >= 0 && tap < K) {
        const float wv = ci < 4 ? w.in_w1[((co * 4 + ci) * K + ky) * K + tap]
                                : w.gate_w1[((co * 4 + (ci - 4)) * K + ky) * K + tap];
        val = bf16_plane(wv, j);
      }
    }
    const int off = tile * BT + n * 64 + (((k >> 3) ^ ((n >> 1) & 3)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(img + off) = __float2bfloat16(val);
  }
  if (blockIdx.x == 0 && threadIdx.x < 4) {
    float* b = reinterpret_cast<float*>(img + WB_BYTES);
    b[threadIdx.x] = w.in_b1[threadIdx.x] + w.gate_b1[threadIdx.x];
    b[4 + threadIdx.x] = w.in_b0[threadIdx.x];
    b[8 + threadIdx.x] = w.gate_b0[threadIdx.x];
  }
}

// ---- the step kernel ----------------------------------------------------------------------------
struct TcGeo {
  int ns;    // strips per image row
  int TW;    // strip width in pixels (multiple of 4)
  int QPR;   // quads per strip row incl. 2 halo quads = TW/4 + 2
};

struct Seg {
  int b, xs, ya, nrows;
  int nt0, nt1, nt2;
};

// CTA's share of (image, strip, row) space, cut into segments that stay inside one strip.
struct SegIter {
  int64_t r, r_end;
  int H, ns, TW, QPR;
  __device__ SegIter(const StepArgs& a, const TcGeo& g) {
    const int64_t total = (int64_t)a.B * g.ns * a.H;
    r = total * blockIdx.x / gridDim.x;
    r_end = total * (blockIdx.x + 1) / gridDim.x;
    H = a.H, ns = g.ns, TW = g.TW, QPR = g.QPR;
  }
  __device__ bool next(Seg& s) {
    if (r >= r_end) return false;
    const int64_t unit = r / H;
    s.ya = (int)(r - unit * H);
    s.b = (int)(unit / ns);
    s.xs = (int)(unit - (int64_t)s.b * ns) * TW;
    const int64_t left = r_end - r;
    s.nrows = (int)(left < (int64_t)(H - s.ya) ? left : (int64_t)(H - s.ya));
    s.nt0 = ((s.nrows + 8) * QPR + TILE - 1) / TILE;
    s.nt1 = ((s.nrows + 4) * QPR + TILE - 1) / TILE;
    s.nt2 = (s.nrows * QPR + TILE - 1) / TILE;
    r += s.nrows;
    return true;
  }
};

template <bool kInt>
__global__ void __launch_bounds__(NUM_THREADS, 1)
sampler_tc_step_kernel(const StepArgs a, const TcGeo g, const uint8_t* __restrict__ wimg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* x0_full = bars;                 // [S0]
  uint64_t* x0_empty = x0_full + S0;        // [S0]
  uint64_t* x1_full = x0_empty + S0;        // [S1]
  uint64_t* x1_empty = x1_full + S1;        // [S1]
  uint64_t* d1_full = x1_empty + S1;        // [2]
  uint64_t* d1_empty = d1_full + 2;
  uint64_t* d2_full = d1_empty + 2;
  uint64_t* d2_empty = d2_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_MISC);
  float* sh_b = reinterpret_cast<float*>(smem + OFF_MISC + 16);  // 12 floats
  const uint32_t sX0 = smem_u32(smem + OFF_X0), sX1 = smem_u32(smem + OFF_X1), sWB = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int QPR = g.QPR;
  const bool first = a.t == 0, last = a.t == a.Tm - 1;
  const int tm = a.Tm - 1 - a.t;  // newest micro-bin first (embedding.py:155-156)
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t BHW2 = (int64_t)a.B * 2 * HW;

  // ---- prologue: barriers, TMEM, weight image -> shared -------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < S0; ++i) mbar_init(x0_full + i, 4), mbar_init(x0_empty + i, 1);
    for (int i = 0; i < S1; ++i) mbar_init(x1_full + i, 4), mbar_init(x1_empty + i, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(d1_full + i, 1), mbar_init(d1_empty + i, 4);
      mbar_init(d2_full + i, 1), mbar_init(d2_empty + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(wimg);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < WB_BYTES / 16; i += NUM_THREADS) dst[i] = src[i];
    if (threadIdx.x < 12) sh_b[threadIdx.x] = reinterpret_cast<const float*>(wimg + WB_BYTES)[threadIdx.x];
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 9) {
    // ===================== producers: counts + previous spikes -> X0 ring =====================
    const int q = (warp - 9) * 32 + lane;  // quad of the tile this thread fills
    SegIter it(a, g);
    Seg s;
    int gt = 0;  // global X0 tile counter
    while (it.next(s)) {
      const float* ev_img = reinterpret_cast<const float*>(a.events) + ((int64_t)s.b * a.Tm + tm) * 2 * HW;
      const float* sp_img = a.s_prev + (int64_t)s.b * 2 * HW;
      const int nq0 = (s.nrows + 8) * QPR;
      float4 cur[4], nxt[4];
      auto load_tile = [&](int i, float4 (&v)[4]) {
        const int f = i * TILE + q;
        const int r0 = f / QPR, m = f - r0 * QPR;
        const int y = s.ya - 4 + r0, x = s.xs - 4 + 4 * m;
        const bool ok = f < nq0 && (unsigned)y < (unsigned)a.H && (unsigned)x < (unsigned)a.W;
        const int64_t off = ok ? (int64_t)y * a.W + x : 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          v[c] = ok ? ld_stream_f4(reinterpret_cast<const float4*>(ev_img + c * HW + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[2 + c] = (ok && !first) ? ld_stream_f4(reinterpret_cast<const float4*>(sp_img + c * HW + off))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      load_tile(0, nxt);
      for (int i = 0; i < s.nt0; ++i, ++gt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) cur[c] = nxt[c];
        if (i + 1 < s.nt0) load_tile(i + 1, nxt);
        if (kInt) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            cur[c].x = (float)__float_as_int(cur[c].x), cur[c].y = (float)__float_as_int(cur[c].y);
            cur[c].z = (float)__float_as_int(cur[c].z), cur[c].w = (float)__float_as_int(cur[c].w);
          }
        }
        // chunk layout: element = px*2 + ch
        uint32_t hi[4], mid[4], lo[4], sp[4];
        split3_pair(cur[0].x, cur[1].x, hi[0], mid[0], lo[0]);
        split3_pair(cur[0].y, cur[1].y, hi[1], mid[1], lo[1]);
        split3_pair(cur[0].z, cur[1].z, hi[2], mid[2], lo[2]);
        split3_pair(cur[0].w, cur[1].w, hi[3], mid[3], lo[3]);
        sp[0] = pack2_bf16(cur[2].x, cur[3].x), sp[1] = pack2_bf16(cur[2].y, cur[3].y);
        sp[2] = pack2_bf16(cur[2].z, cur[3].z), sp[3] = pack2_bf16(cur[2].w, cur[3].w);
        const int slot = gt % S0;
        mbar_wait(x0_empty + slot, ((gt / S0) & 1) ^ 1);
        const int pos = slot * TILE + q;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          if (rep == 1 && slot != 0) break;
          const uint32_t row = sX0 + (uint32_t)(pos + rep * RS0) * ROWB;
          const uint32_t sw = (row >> 7) & 3;
          st_shared_v4(row + ((0 ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(row + ((1 ^ sw) << 4), mid[0], mid[1], mid[2], mid[3]);
          st_shared_v4(row + ((2 ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
          st_shared_v4(row + ((3 ^ sw) << 4), sp[0], sp[1], sp[2], sp[3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(x0_full + slot);
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) |
                                 ((uint32_t)(TILE >> 4) << 24);
      SegIter it(a, g);
      Seg s;
      int g0 = 0, g1 = 0, g2 = 0;  // global tile counters at the segment start
      int w0 = 0, w1 = 0;          // X0 / X1 tiles already waited for (global)
      while (it.next(s)) {
        const int base0 = (g0 % S0) * TILE, base1 = (g1 % S1) * TILE;
        for (int itr = 0; itr < s.nt1 + 2; ++itr) {
          if (itr < s.nt1) {
            // ---------- layer 1, tile i: X0 -> D1 ----------
            const int i = itr, gi = g1 + i, buf = gi & 1;
            const int need0 = g0 + min(i + 2, s.nt0);
            while (w0 < need0) mbar_wait(x0_full + (w0 % S0), (w0 / S0) & 1), ++w0;
            mbar_wait(d1_empty + buf, ((gi >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_in = tmem_base + TM_D1 + buf * 32, d_gate = d_in + 16;
            uint32_t arow[10];
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                arow[ky * 2 + h] = sX0 + (uint32_t)((base0 + i * TILE + ky * QPR + h) % RS0) * ROWB;
            uint32_t acc_in = 0, acc_g = 0;
            // small product terms first: x_hi*w_lo, (x_hi+x_mid)*w_mid, x_lo*w_hi
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
              const uint32_t bt = sWB + (uint32_t)(ky * 7) * BT;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t a0 = sw64_desc(arow[ky * 2 + h]), a1 = sw64_desc(arow[ky * 2 + h] + 32);
                tc_mma(d_in, a0, sw64_desc(bt + 3 * BT + h * 32), idesc, acc_in), acc_in = 1;   // in2
                tc_mma(d_in, a0, sw64_desc(bt + 2 * BT + h * 32), idesc, 1);                    // in1
                tc_mma(d_in, a1, sw64_desc(bt + h * BT + 32), idesc, 1);                        // in0: lo chunk
                if (!first) {
                  tc_mma(d_gate, a1, sw64_desc(bt + 6 * BT + h * 32), idesc, acc_g), acc_g = 1;  // gate lo
                  tc_mma(d_gate, a1, sw64_desc(bt + 5 * BT + h * 32), idesc, 1);                 // gate mid
                }
              }
            }
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
              const uint32_t bt = sWB + (uint32_t)(ky * 7) * BT;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t a0 = sw64_desc(arow[ky * 2 + h]), a1 = sw64_desc(arow[ky * 2 + h] + 32);
                tc_mma(d_in, a0, sw64_desc(bt + h * BT), idesc, 1);                              // in0: hi|mid
                if (!first) tc_mma(d_gate, a1, sw64_desc(bt + 4 * BT + h * 32), idesc, 1);       // gate hi
              }
            }
            tc_commit(d1_full + buf);
            tc_commit(x0_empty + ((g0 + i) % S0));
            if (i == s.nt1 - 1)
              for (int k = s.nt1; k < s.nt0; ++k) tc_commit(x0_empty + ((g0 + k) % S0));
          }
          const int j = itr - 2;
          if (j >= 0 && j < s.nt2) {
            // ---------- layer 2, tile j: X1 -> D2 ----------
            const int gj = g2 + j, buf = gj & 1;
            const int need1 = g1 + min(j + 2, s.nt1);
            while (w1 < need1) mbar_wait(x1_full + (w1 % S1), (w1 / S1) & 1), ++w1;
            mbar_wait(d2_empty + buf, ((gj >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d2 = tmem_base + TM_D2 + buf * 16;
            uint32_t arow[10];
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
              for (int h = 0; h < 2; ++h)
                arow[ky * 2 + h] = sX1 + (uint32_t)((base1 + j * TILE + ky * QPR + h) % RS1) * ROWB;
            uint32_t acc = 0;
            // (x plane, w plane) terms with xp + wp < 3, smallest first
            constexpr int XP[6] = {2, 1, 0, 1, 0, 0};
            constexpr int WP[6] = {0, 1, 2, 0, 1, 0};
#pragma unroll
            for (int term = 0; term < 6; ++term) {
#pragma unroll
              for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint32_t ar = arow[ky * 2 + h] + (uint32_t)XP[term] * X1_PLANE;
                  const uint32_t bt = sWB + (uint32_t)(NB_L1 + (WP[term] * 5 + ky) * 2 + h) * BT;
                  tc_mma(d2, sw64_desc(ar), sw64_desc(bt), idesc, acc), acc = 1;
                  tc_mma(d2, sw64_desc(ar + 32), sw64_desc(bt + 32), idesc, 1);
                }
              }
            }
            tc_commit(d2_full + buf);
            tc_commit(x1_empty + ((g1 + j) % S1));
            if (j == s.nt2 - 1)
              for (int k = s.nt2; k < s.nt1; ++k) tc_commit(x1_empty + ((g1 + k) % S1));
          }
        }
        g0 += s.nt0, g1 += s.nt1, g2 += s.nt2;
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue 1: D1 -> bias, ReLU, mask, split -> X1 ring =====================
    const int q = warp * 32 + lane;
    SegIter it(a, g);
    Seg s;
    int gt = 0;
    while (it.next(s)) {
      for (int i = 0; i < s.nt1; ++i, ++gt) {
        const int buf = gt & 1, slot = gt % S1;
        const int f = i * TILE + q;
        const int r1 = f / QPR, m = f - r1 * QPR;
        const int y1 = s.ya - 2 + r1, x1 = s.xs - 2 + 4 * m;
        const bool row_in = (unsigned)y1 < (unsigned)a.H;
        mbar_wait(d1_full + buf, (gt >> 1) & 1);
        tc_fence_after();
        uint32_t din[16], dg[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + TM_D1 + buf * 32;
        tmem_ld16(taddr, din);
        if (!first) tmem_ld16(taddr + 16, dg);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d1_empty + buf);
        uint32_t ph[4][4], pm[4][4], pl[4][4];  // [px][4 x (2 channels)]
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const bool in_img = row_in && (unsigned)(x1 + p) < (unsigned)a.W;
          float hv[8];
#pragma unroll
          for (int co = 0; co < 4; ++co) {
            const float vi = fmaxf(__uint_as_float(din[p * 4 + co]) + sh_b[4 + co], 0.0f);
            const float vg = fmaxf((first ? 0.0f : __uint_as_float(dg[p * 4 + co])) + sh_b[8 + co], 0.0f);
            hv[co] = in_img ? vi : 0.0f;
            hv[4 + co] = in_img ? vg : 0.0f;
          }
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) split3_pair(hv[2 * c2], hv[2 * c2 + 1], ph[p][c2], pm[p][c2], pl[p][c2]);
        }
        mbar_wait(x1_empty + slot, ((gt / S1) & 1) ^ 1);
        const int pos = slot * TILE + q;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          if (rep == 1 && slot != 0) break;
          const uint32_t row = sX1 + (uint32_t)(pos + rep * RS1) * ROWB;
          const uint32_t sw = (row >> 7) & 3;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const uint32_t ca = row + (((uint32_t)p ^ sw) << 4);
            st_shared_v4(ca, ph[p][0], ph[p][1], ph[p][2], ph[p][3]);
            st_shared_v4(ca + X1_PLANE, pm[p][0], pm[p][1], pm[p][2], pm[p][3]);
            st_shared_v4(ca + 2 * X1_PLANE, pl[p][0], pl[p][1], pl[p][2], pl[p][3]);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(x1_full + slot);
      }
    }
  } else {
    // ===================== epilogue 2: D2 -> membrane update + spike-triggered read-out =====================
    // Same arithmetic, statement for statement, as sampler_step_kernel (embedding.py:132-139, 177-217).
    const int wq = warp - 4;
    const int q = wq * 32 + lane;
    SegIter it(a, g);
    Seg s;
    int gt = 0;
    while (it.next(s)) {
      const int nq2 = s.nrows * QPR;
      for (int j = 0; j < s.nt2; ++j, ++gt) {
        const int buf = gt & 1;
        const int f = j * TILE + q;
        const int r2 = f / QPR, m = f - r2 * QPR;
        const int gy = s.ya + r2, gx = s.xs + 4 * m;
        const bool valid = f < nq2 && m < QPR - 2 && gx < a.W;
        const int64_t base0 = valid ? ((int64_t)s.b * 2) * HW + (int64_t)gy * a.W + gx : 0;
        // state of the quad (both channels), fetched while the MMAs run
        float4 vmq[2], acq[2];
        uint2 mtq[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          vmq[c] = acq[c] = make_float4(0.f, 0.f, 0.f, 0.f);
          mtq[c] = make_uint2(0u, 0u);
          if (valid && !first) {
            vmq[c] = ld_stream_f4(reinterpret_cast<const float4*>(a.vm + base0 + c * HW));
            acq[c] = ld_stream_f4(reinterpret_cast<const float4*>(a.acc + base0 + c * HW));
            mtq[c] = ld_stream_u2(reinterpret_cast<const uint2*>(a.meta + base0 + c * HW));
          }
        }
        mbar_wait(d2_full + buf, (gt >> 1) & 1);
        tc_fence_after();
        uint32_t d[16];
        tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + TM_D2 + buf * 16, d);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty + buf);
        if (!valid) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float bg = sh_b[c], bc = sh_b[2 + c];
          const int64_t base = base0 + c * HW;
          __align__(16) float vm4[4], ac4[4], v4[4], g4[4], s4[4], o4[4];
          __align__(8) uint16_t m4[4];
          *reinterpret_cast<float4*>(vm4) = vmq[c];
          *reinterpret_cast<float4*>(ac4) = acq[c];
          *reinterpret_cast<uint2*>(m4) = mtq[c];
          float* outp = a.out + base;  // plane k at outp + k*BHW2
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const float gate = __fdividef(1.0f, 1.0f + __expf(-(__uint_as_float(d[px * 4 + c]) + bg)));
            const float cur = __uint_as_float(d[px * 4 + 2 + c]) + bc;
            int seg = m4[px] & 0xff;
            int tl = (int)(m4[px] >> 8) - 1;
            const float v = __fadd_rn(__fmul_rn(gate, vm4[px]), cur);
            const bool sp = __fsub_rn(v, a.thresh) > 0.0f;
            const float vm = sp ? (a.hard_reset ? a.vreset : __fsub_rn(v, a.thresh)) : v;
            float ac = __fadd_rn(ac4[px], v);
            const bool vld = sp && seg < a.Ts;
            float val = a.readout == EAS_READOUT_SUM ? ac : vm;
            if (a.readout == EAS_READOUT_AVG) val = ac / (float)(a.t - tl);
            if (a.use_abs) val = fmaxf(val, 0.0f);
            o4[px] = vld ? val : 0.0f;  // plane 0 on the first step
            if (!first && vld) outp[(int64_t)seg * BHW2 + px] = val;
            seg += vld ? 1 : 0;
            tl = vld ? a.t : tl;
            ac = sp ? 0.0f : ac;
            if (last && !sp && seg < a.Ts && !a.write_zero) {
              float tv = a.readout == EAS_READOUT_SUM ? ac : vm;
              if (a.readout == EAS_READOUT_AVG) tv = ac / (float)(a.Tm - 1 - tl);
              if (a.use_abs) tv = fmaxf(tv, 0.0f);
              if (first && seg == 0) o4[px] = tv;  // Tm == 1: still inside the zero-initialising store
              else outp[(int64_t)seg * BHW2 + px] = tv;
            }
            vm4[px] = vm, ac4[px] = ac, v4[px] = v, g4[px] = gate, s4[px] = sp ? 1.0f : 0.0f;
            m4[px] = (uint16_t)(seg | ((tl + 1) << 8));
          }
          if (first) {
            *reinterpret_cast<float4*>(outp) = *reinterpret_cast<const float4*>(o4);
            for (int k = 1; k < a.Ts; ++k) *reinterpret_cast<float4*>(outp + k * BHW2) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (!last) {
            *reinterpret_cast<float4*>(a.vm + base) = *reinterpret_cast<const float4*>(vm4);
            *reinterpret_cast<float4*>(a.acc + base) = *reinterpret_cast<const float4*>(ac4);
            *reinterpret_cast<uint2*>(a.meta + base) = *reinterpret_cast<const uint2*>(m4);
            *reinterpret_cast<float4*>(a.s_next + base) = *reinterpret_cast<const float4*>(s4);
          }
          if (a.v_seq) {
            const int64_t se = (int64_t)a.t * BHW2 + base;
            *reinterpret_cast<float4*>(a.v_seq + se) = *reinterpret_cast<const float4*>(v4);
            *reinterpret_cast<float4*>(a.gate_seq + se) = *reinterpret_cast<const float4*>(g4);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS) : "memory");
  }
}

TcGeo pick_geo(int W) {
  TcGeo g;
  const int max_tw = 4 * (MAX_QPR - 2);
  g.ns = (W + max_tw - 1) / max_tw;
  g.TW = ((W + g.ns - 1) / g.ns + 3) / 4 * 4;
  g.QPR = g.TW / 4 + 2;
  return g;
}

}  // namespace

size_t eas_sampler_tc_wimg_bytes() { return (size_t)WIMG_BYTES; }

bool eas_sampler_tc_supported(const eas_sampler_cfg* c, const void* events, const float* out, const float* v_seq,
                              const float* gate_seq) {
  if (c->depth != 2 || c->ksize != 5) return false;
  if (c->W % 4 != 0) return false;  // quads are loaded / stored as 16 B vectors
  if (((uintptr_t)events | (uintptr_t)out | (uintptr_t)v_seq | (uintptr_t)gate_seq) % 16 != 0) return false;
  return true;
}

int eas_sampler_tc_run(const eas_sampler_cfg* cfg, StepArgs a, float* s0, float* s1, void* wimg, cudaStream_t st) {
  EAS_REQUIRE((uintptr_t)wimg % 16 == 0, EAS_E_ALIGN);
  sampler_tc_pack_weights<<<32, 256, 0, st>>>(a.w, reinterpret_cast<uint8_t*>(wimg));
  EAS_LAUNCH_CHECK();
  const TcGeo g = pick_geo(cfg->W);
  auto kern = cfg->in_dtype == EAS_I32 ? sampler_tc_step_kernel<true> : sampler_tc_step_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  const int64_t total_rows = (int64_t)cfg->B * g.ns * cfg->H;
  // one persistent CTA per SM; tiny problems get fewer CTAs (>= 8 strip rows each)
  int64_t grid = (total_rows + 7) / 8;
  if (grid > EAS_NUM_SMS) grid = EAS_NUM_SMS;
  if (grid < 1) grid = 1;
  for (int t = 0; t < cfg->Tm; ++t) {
    a.t = t;
    a.s_prev = (t & 1) ? s0 : s1;  // step t reads what step t-1 wrote
    a.s_next = (t & 1) ? s1 : s0;
    kern<<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, st>>>(a, g, reinterpret_cast<const uint8_t*>(wimg));
    EAS_LAUNCH_CHECK();
  }
  return EAS_OK;
}

}  // namespace eas_sampler
