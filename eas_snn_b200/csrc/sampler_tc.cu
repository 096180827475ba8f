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
//   * fp32-equivalent accuracy on 16-bit tensor cores: every real operand (counts, hidden
//     activations, weights x 2^8) is split into TWO fp16 planes hi + lo (22 mantissa bits; spikes are
//     exact), the three product terms hi*hi, lo*hi, hi*lo are accumulated in fp32 in TMEM, the
//     hi*lo term in its own accumulator columns (concatenated along N, which is free: a
//     shared-memory-operand MMA of M=128, K=16 costs 64 cycles for any N <= 128, measured with
//     scripts/umma_rate_probe.cu) and added in the epilogue.  Magnitudes >= 65504 cannot be held in
//     fp16: they raise a flag in the workspace and the caller re-runs the step sequence on the
//     FP32-pipe kernel (never seen on event data; keeps the operator total).
// One persistent CTA per SM walks vertical strips of <= 116 pixels (QPR <= 31 quads per row incl.
// halo) top to bottom as a rolling pipeline over 128-quad tiles:
//   producers (4 warps): counts + previous spikes -> fp16 hi/lo planes -> X0 ring (zero padded)
//   MMA (1 thread)     : layer 1 (X0 -> D1 in TMEM), layer 2 (X1 -> D2 in TMEM)
//   epilogue 1 (4 warps): D1 + bias, ReLU, image mask, hi/lo split -> X1 ring
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
constexpr int X1_BYTES = 2 * X1_PLANE;
// B tiles (SWIZZLE_64B, 64 B rows = 32 fp16 of K = two MMA K-steps):
//   layer 1, per ky: IN [32 rows][32] (2 KB), GATE [32 rows][32] (2 KB)
//   layer 2, per (ky, h): A [32 rows][32] (2 KB: w_hi | w_lo), B [16 rows][32] (1 KB: w_hi)
constexpr int W1_KY = 4096, W2_KYH = 3072;
constexpr int OFF_W2 = 5 * W1_KY;
constexpr int WB_BYTES = OFF_W2 + 10 * W2_KYH;             // 51200
constexpr int WIMG_BYTES = WB_BYTES + 64;                  // + 12 bias floats
constexpr float W_SCALE = 256.0f, W_UNSCALE = 1.0f / 256.0f;  // keeps the weight lo plane a normal fp16
constexpr float F16_MAX = 65504.0f;
constexpr int OFF_X0 = WB_BYTES;
constexpr int OFF_X1 = OFF_X0 + X0_BYTES;
constexpr int OFF_BAR = OFF_X1 + X1_BYTES;
constexpr int NBAR = 2 * S0 + 2 * S1 + 8;
constexpr int OFF_MISC = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_MISC + 128 + 1024;          // + slack for the 1024 B alignment
// warps 0-3 epilogue 1, 4-7 / 8-11 epilogue 2 (even / odd tiles), 12-14 MMA issuers, 15-18 producers
constexpr int W_E2 = 4, W_MMA = 12, W_PROD = 15;
constexpr int NUM_THREADS = 19 * 32;
constexpr uint32_t SPIN_LIMIT = 1u << 26;  // a few seconds

// TMEM columns: D1 = [in: x*w_hi (16) | x_hi*w_lo (16)][gate: same] x 2 buffers (64 apart);
// D2 = [x_hi*w_hi (16) | x_hi*w_lo (16) | x_lo*w_hi (16)] x 2 buffers (64 apart).  The partial sums
// are added in fp32 by the epilogues.
constexpr uint32_t TM_D1 = 0, TM_D2 = 128, TM_COLS = 256;

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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
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

// (a, b) -> packed fp16 pairs hi, lo with a ~= hi.x + lo.x (22 mantissa bits); `ovf` is raised when
// a magnitude does not fit fp16.
__device__ __forceinline__ void split2_pair(float a, float b, uint32_t& hi, uint32_t& lo, bool& ovf) {
  ovf = ovf || !(fabsf(a) < F16_MAX) || !(fabsf(b) < F16_MAX);
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack2_f16(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- weight image -------------------------------------------------------------------------------
// B tiles in their SWIZZLE_64B shared-memory layout (image base 1024 B aligned), then 12 bias
// floats.  Row n of a tile = output column n of the MMA; k = position inside the 32-element row.
//   layer 1, IN(ky):   k = h*16 + c*8 + p*2 + ci, c in {ev_hi, ev_lo}; n = part*16 + jpx*4 + co
//                      part 0: w_hi for both c;  part 1: w_lo for c = ev_hi, 0 for c = ev_lo
//            GATE(ky): k = h*16 + c*8 + p*2 + ci, c in {spk, zero pad}; part 0: g_hi, part 1: g_lo
//   layer 2, A(ky,h):  k = p*8 + ci (8 hidden channels: input stack 0-3, gate stack 4-7);
//                      n = part*16 + jpx*4 + co, part 0: w_hi, part 1: w_lo;   B(ky,h): w_hi only
// with tap = 4*h + p - jpx (window pixel minus output pixel), weights pre-scaled by 2^8.
__device__ __forceinline__ float f16_plane(float w, int plane) {
  const float ws = w * W_SCALE;
  const float hi = __half2float(__float2half_rn(ws));
  return plane == 0 ? hi : __half2float(__float2half_rn(ws - hi));
}

__global__ void sampler_tc_pack_weights(const eas_sampler_weights w, uint8_t* img, int* flag) {
  constexpr int K = 5;
  const int n1 = 5 * 2 * 32 * 32, n2 = 10 * 48 * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n1 + n2; idx += gridDim.x * blockDim.x) {
    float val = 0.0f;
    int off;
    if (idx < n1) {
      const int ky = idx / 2048, r = idx % 2048;
      const int gate = r / 1024, n = (r / 32) % 32, k = r % 32;
      const int part = n >> 4, jpx = (n >> 2) & 3, co = n & 3;
      const int h = k >> 4, c = (k >> 3) & 1, p = (k & 7) >> 1, ci = k & 1, tap = 4 * h + p - jpx;
      if (tap >= 0 && tap < K) {
        if (!gate) {
          if (!(part == 1 && c == 1)) val = f16_plane(w.in_w0[((co * 2 + ci) * K + ky) * K + tap], part);
        } else if (c == 0) {
          val = f16_plane(w.gate_w0[((co * 2 + ci) * K + ky) * K + tap], part);
        }
      }
      off = ky * W1_KY + gate * 2048 + n * 64 + (((k >> 3) ^ ((n >> 1) & 3)) << 4) + (k & 7) * 2;
    } else {
      const int i2 = idx - n1;
      const int kyh = i2 / 1536, r = i2 % 1536;
      const int ky = kyh >> 1, h = kyh & 1;
      const int n = r / 32, k = r % 32;          // n 0..31: tile A, 32..47: tile B
      const int nn = n < 32 ? n : n - 32;
      const int part = n < 32 ? (nn >> 4) : 0, jpx = (nn >> 2) & 3, co = nn & 3;
      const int p = k >> 3, ci = k & 7, tap = 4 * h + p - jpx;
      if (tap >= 0 && tap < K) {
        const float wv = ci < 4 ? w.in_w1[((co * 4 + ci) * K + ky) * K + tap]
                                : w.gate_w1[((co * 4 + (ci - 4)) * K + ky) * K + tap];
        val = f16_plane(wv, part);
      }
      off = OFF_W2 + kyh * W2_KYH + (n < 32 ? 0 : 2048) + nn * 64 + (((k >> 3) ^ ((nn >> 1) & 3)) << 4) + (k & 7) * 2;
    }
    *reinterpret_cast<__half*>(img + off) = __float2half_rn(val);
  }
  if (blockIdx.x == 0 && threadIdx.x < 4) {
    float* b = reinterpret_cast<float*>(img + WB_BYTES);
    b[threadIdx.x] = w.in_b1[threadIdx.x] + w.gate_b1[threadIdx.x];
    b[4 + threadIdx.x] = w.in_b0[threadIdx.x];
    b[8 + threadIdx.x] = w.gate_b0[threadIdx.x];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) flag[0] = 0, flag[1] = 0;   // fall-back flag, fall-back grid barrier
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
sampler_tc_step_kernel(const StepArgs a, const TcGeo g, const uint8_t* __restrict__ wimg, int* __restrict__ ovf_flag) {
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

  // (through a shuffle: provably warp-uniform, so the role loops' bookkeeping stays on the uniform datapath)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int QPR = g.QPR;
  const bool first = a.t == 0, last = a.t == a.Tm - 1;
  const int tm = a.Tm - 1 - a.t;  // newest micro-bin first (embedding.py:155-156)
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t BHW2 = (int64_t)a.B * 2 * HW;

  // ---- prologue: barriers, TMEM, weight image -> shared -------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < S0; ++i) mbar_init(x0_full + i, 4), mbar_init(x0_empty + i, 1);
    for (int i = 0; i < S1; ++i) mbar_init(x1_full + i, 4), mbar_init(x1_empty + i, 2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(d1_full + i, 1), mbar_init(d1_empty + i, 4);
      mbar_init(d2_full + i, 2), mbar_init(d2_empty + i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
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

  if (warp >= W_PROD) {
    // ===================== producers: counts + previous spikes -> X0 ring =====================
    const int q = (warp - W_PROD) * 32 + lane;  // quad of the tile this thread fills
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
        // chunk layout: element = px*2 + ch; chunks [ev_hi | ev_lo | spikes | zero]
        uint32_t hi[4], lo[4], sp[4];
        bool ovf = false;
        split2_pair(cur[0].x, cur[1].x, hi[0], lo[0], ovf);
        split2_pair(cur[0].y, cur[1].y, hi[1], lo[1], ovf);
        split2_pair(cur[0].z, cur[1].z, hi[2], lo[2], ovf);
        split2_pair(cur[0].w, cur[1].w, hi[3], lo[3], ovf);
        if (ovf) *ovf_flag = 1;
        sp[0] = pack2_f16(cur[2].x, cur[3].x), sp[1] = pack2_f16(cur[2].y, cur[3].y);
        sp[2] = pack2_f16(cur[2].z, cur[3].z), sp[3] = pack2_f16(cur[2].w, cur[3].w);
        const int slot = gt % S0;
        mbar_wait(x0_empty + slot, ((gt / S0) & 1) ^ 1);
        const int pos = slot * TILE + q;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          if (rep == 1 && slot != 0) break;
          const uint32_t row = sX0 + (uint32_t)(pos + rep * RS0) * ROWB;
          const uint32_t sw = (row >> 7) & 3;
          st_shared_v4(row + ((0 ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(row + ((1 ^ sw) << 4), lo[0], lo[1], lo[2], lo[3]);
          st_shared_v4(row + ((2 ^ sw) << 4), sp[0], sp[1], sp[2], sp[3]);
          st_shared_v4(row + ((3 ^ sw) << 4), 0u, 0u, 0u, 0u);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(x0_full + slot);
      }
    }
  } else if (warp >= W_MMA) {
    // ===================== MMA issuers =====================
    // Three issuing warps (layer 1; layer 2 hi plane; layer 2 lo plane): issuing is what bounds the
    // tensor pipe for such small MMAs, and the three accumulator chains are independent.  Each warp
    // walks its (warp-uniform) schedule so that descriptors live in uniform registers; only the
    // elected lane issues tcgen05.mma / tcgen05.commit.
    const bool leader = elect_one();
    // fp16 x fp16 -> fp32, M = 128, N = 32 / 16
    constexpr uint32_t idesc32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    constexpr uint32_t idesc16 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    // descriptor + (byte offset >> 4) moves the start address (all buffers sit below 256 KB)
    const uint64_t wdesc = sw64_desc(sWB);
    SegIter it(a, g);
    Seg s;
    if (warp == W_MMA) {
      // ---------- layer 1: X0 -> D1 ----------
      const uint64_t x0desc = sw64_desc(sX0);
      int g0 = 0, g1 = 0;  // global tile counters at the segment start
      int w0 = 0;          // X0 tiles already waited for (global)
      while (it.next(s)) {
        int pos = (g0 % S0) * TILE;  // ring row of tile 0, advanced by TILE per tile
        for (int i = 0; i < s.nt1; ++i) {
          const int gi = g1 + i, buf = gi & 1;
          const int need0 = g0 + min(i + 2, s.nt0);
          while (w0 < need0) mbar_wait(x0_full + (w0 % S0), (w0 / S0) & 1), ++w0;
          mbar_wait(d1_empty + buf, ((gi >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_in = tmem_base + TM_D1 + buf * 64, d_gate = d_in + 32;
          // rolled over ky (small code); the ring position advances by QPR rows per filter row and
          // wraps without a division
          int pk = pos;
          uint64_t bd = wdesc;  // IN tile of ky (GATE tile 2048 B further)
#pragma unroll 1
          for (int ky = 0; ky < 5; ++ky) {
            const int pk1 = pk + 1 == RS0 ? 0 : pk + 1;
            const uint64_t a0 = x0desc + (uint64_t)(pk * (ROWB >> 4)), a1 = x0desc + (uint64_t)(pk1 * (ROWB >> 4));
            const uint32_t acc = ky == 0 ? 0u : 1u;
            if (leader) {
              // counts (hi | lo chunks) x IN tile, K-step h; spikes (| zero pad) x GATE tile (zero on step 0)
              tc_mma(d_in, a0, bd, idesc32, acc);
              tc_mma(d_in, a1, bd + 2, idesc32, 1u);
              if (!first) {
                tc_mma(d_gate, a0 + 2, bd + (2048 >> 4), idesc32, acc);
                tc_mma(d_gate, a1 + 2, bd + ((2048 >> 4) + 2), idesc32, 1u);
              }
            }
            bd += (uint64_t)(W1_KY >> 4);
            pk += QPR;
            if (pk >= RS0) pk -= RS0;
          }
          if (leader) {
            tc_commit(d1_full + buf);
            tc_commit(x0_empty + ((g0 + i) % S0));
            if (i == s.nt1 - 1)
              for (int k = s.nt1; k < s.nt0; ++k) tc_commit(x0_empty + ((g0 + k) % S0));
          }
          pos += TILE;
          if (pos >= RS0) pos -= RS0;
        }
        g0 += s.nt0, g1 += s.nt1;
      }
    } else {
      // ---------- layer 2, plane `pl` of X1 -> D2: x_hi * [w_hi | w_lo] -> columns 0..31 (tile A),
      //            x_lo * w_hi -> columns 32..47 (tile B) ----------
      const int pl = warp - W_MMA - 1;
      const uint64_t x1desc = sw64_desc(sX1) + (uint64_t)(pl * (X1_PLANE >> 4));
      const uint64_t wd2 = wdesc + (uint64_t)((OFF_W2 + pl * 2048) >> 4);
      const uint32_t idesc = pl == 0 ? idesc32 : idesc16;
      int g1 = 0, g2 = 0, w1 = 0;
      while (it.next(s)) {
        int pos = (g1 % S1) * TILE;
        for (int j = 0; j < s.nt2; ++j) {
          const int gj = g2 + j, buf = gj & 1;
          const int need1 = g1 + min(j + 2, s.nt1);
          while (w1 < need1) mbar_wait(x1_full + (w1 % S1), (w1 / S1) & 1), ++w1;
          mbar_wait(d2_empty + buf, ((gj >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d2 = tmem_base + TM_D2 + buf * 64 + pl * 32;
          int pk = pos;
          uint64_t bd = wd2;  // tile of (ky, h = 0); h = 1 is W2_KYH further
#pragma unroll 1
          for (int ky = 0; ky < 5; ++ky) {
            const int pk1 = pk + 1 == RS1 ? 0 : pk + 1;
            const uint64_t a0 = x1desc + (uint64_t)(pk * (ROWB >> 4)), a1 = x1desc + (uint64_t)(pk1 * (ROWB >> 4));
            if (leader) {
              tc_mma(d2, a0, bd, idesc, ky == 0 ? 0u : 1u);
              tc_mma(d2, a0 + 2, bd + 2, idesc, 1u);
              tc_mma(d2, a1, bd + (W2_KYH >> 4), idesc, 1u);
              tc_mma(d2, a1 + 2, bd + ((W2_KYH >> 4) + 2), idesc, 1u);
            }
            bd += (uint64_t)((2 * W2_KYH) >> 4);
            pk += QPR;
            if (pk >= RS1) pk -= RS1;
          }
          if (leader) {
            tc_commit(d2_full + buf);
            tc_commit(x1_empty + ((g1 + j) % S1));
            if (j == s.nt2 - 1)
              for (int k = s.nt2; k < s.nt1; ++k) tc_commit(x1_empty + ((g1 + k) % S1));
          }
          pos += TILE;
          if (pos >= RS1) pos -= RS1;
        }
        g1 += s.nt1, g2 += s.nt2;
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue 1: D1 -> bias, ReLU, mask, split -> X1 ring =====================
    const int q = warp * 32 + lane;
    SegIter it(a, g);
    Seg s;
    int gt = 0;
    while (it.next(s)) {
      const int nq1 = (s.nrows + 4) * QPR;    // hidden-layer quads this segment needs
      for (int i = 0; i < s.nt1; ++i, ++gt) {
        const int buf = gt & 1, slot = gt % S1;
        const int f = i * TILE + q;
        const int r1 = f / QPR, m = f - r1 * QPR;
        const int y1 = s.ya - 2 + r1, x1 = s.xs - 2 + 4 * m;
        // quads past nq1 pad the last tile: their windows reach X0 rows nobody wrote in this launch (stale shared
        // memory), so they are forced to zero like the out-of-image ones -- otherwise garbage there can raise the
        // fp16-range flag and send a perfectly good forward down the FP32 fall-back
        const bool row_in = f < nq1 && (unsigned)y1 < (unsigned)a.H;
        mbar_wait(d1_full + buf, (gt >> 1) & 1);
        tc_fence_after();
        // D1 columns: [x*w_hi (16) | x_hi*w_lo (16)] for the input stack, then the same for the gate stack
        uint32_t din[16], dinl[16], dg[16], dgl[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + TM_D1 + buf * 64;
        tmem_ld16(taddr, din);
        tmem_ld16(taddr + 16, dinl);
        if (!first) tmem_ld16(taddr + 32, dg), tmem_ld16(taddr + 48, dgl);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d1_empty + buf);
        uint32_t ph[4][4], pl[4][4];  // [px][4 x (2 channels)]
        bool ovf = false;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const bool in_img = row_in && (unsigned)(x1 + p) < (unsigned)a.W;
          float hv[8];
#pragma unroll
          for (int co = 0; co < 4; ++co) {
            const float ci = (__uint_as_float(din[p * 4 + co]) + __uint_as_float(dinl[p * 4 + co])) * W_UNSCALE;
            const float cg = first ? 0.0f
                                   : (__uint_as_float(dg[p * 4 + co]) + __uint_as_float(dgl[p * 4 + co])) * W_UNSCALE;
            const float vi = fmaxf(ci + sh_b[4 + co], 0.0f);
            const float vg = fmaxf(cg + sh_b[8 + co], 0.0f);
            hv[co] = in_img ? vi : 0.0f;
            hv[4 + co] = in_img ? vg : 0.0f;
          }
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) split2_pair(hv[2 * c2], hv[2 * c2 + 1], ph[p][c2], pl[p][c2], ovf);
        }
        if (ovf) *ovf_flag = 1;
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
            st_shared_v4(ca + X1_PLANE, pl[p][0], pl[p][1], pl[p][2], pl[p][3]);
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
    // Two groups of four warps take alternate tiles (= alternate D2 buffers): one warp per scheduler
    // runs this dependent arithmetic at a few cycles per instruction, too slow for one group alone.
    const int wq = warp & 3, grp = (warp - W_E2) >> 2;
    const int q = wq * 32 + lane;
    SegIter it(a, g);
    Seg s;
    int gt = 0;
    while (it.next(s)) {
      const int nq2 = s.nrows * QPR;
      for (int j = 0; j < s.nt2; ++j, ++gt) {
        const int buf = gt & 1;
        if (buf != grp) continue;
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
        uint32_t d[16], dl[16], dx[16];  // x_hi*w_hi | x_hi*w_lo | x_lo*w_hi
        const uint32_t t2 = tmem_base + ((uint32_t)(wq * 32) << 16) + TM_D2 + buf * 64;
        tmem_ld16(t2, d);
        tmem_ld16(t2 + 16, dl);
        tmem_ld16(t2 + 32, dx);
        tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < 16; ++n)
          d[n] = __float_as_uint(__uint_as_float(d[n]) + (__uint_as_float(dl[n]) + __uint_as_float(dx[n])));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2_empty + buf);
        if (!valid) continue;
#pragma unroll 1   // rolled: halves the code of this role (the four roles share one instruction cache)
        for (int c = 0; c < 2; ++c) {
          const float bg = sh_b[c], bc = sh_b[2 + c];
          const int64_t base = base0 + c * HW;
          __align__(16) float vm4[4], ac4[4], v4[4], g4[4], s4[4], o4[4];
          __align__(8) uint16_t m4[4];
          *reinterpret_cast<float4*>(vm4) = c ? vmq[1] : vmq[0];
          *reinterpret_cast<float4*>(ac4) = c ? acq[1] : acq[0];
          *reinterpret_cast<uint2*>(m4) = c ? mtq[1] : mtq[0];
          float* outp = a.out + base;  // plane k at outp + k*BHW2
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const float gpre = __uint_as_float(c ? d[px * 4 + 1] : d[px * 4]) * W_UNSCALE;
            const float cpre = __uint_as_float(c ? d[px * 4 + 3] : d[px * 4 + 2]) * W_UNSCALE;
            const float gate = __fdividef(1.0f, 1.0f + __expf(-(gpre + bg)));
            const float cur = cpre + bc;
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
  if (warp == W_MMA) {
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

size_t eas_sampler_tc_wimg_bytes() { return (size_t)WIMG_BYTES + 64; }  // + the fp16-range flag

const int* eas_sampler_tc_flag(const void* wimg) {
  return reinterpret_cast<const int*>(reinterpret_cast<const uint8_t*>(wimg) + WIMG_BYTES);
}

bool eas_sampler_tc_supported(const eas_sampler_cfg* c, const void* events, const float* out, const float* v_seq,
                              const float* gate_seq) {
  if (c->depth != 2 || c->ksize != 5) return false;
  if (c->W % 4 != 0) return false;  // quads are loaded / stored as 16 B vectors
  if (((uintptr_t)events | (uintptr_t)out | (uintptr_t)v_seq | (uintptr_t)gate_seq) % 16 != 0) return false;
  return true;
}

int eas_sampler_tc_run(const eas_sampler_cfg* cfg, StepArgs a, float* s0, float* s1, void* wimg, cudaStream_t st) {
  EAS_REQUIRE((uintptr_t)wimg % 16 == 0, EAS_E_ALIGN);
  int* flag = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(wimg) + WIMG_BYTES);
  sampler_tc_pack_weights<<<32, 256, 0, st>>>(a.w, reinterpret_cast<uint8_t*>(wimg), flag);
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
    kern<<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, st>>>(a, g, reinterpret_cast<const uint8_t*>(wimg), flag);
    EAS_LAUNCH_CHECK();
  }
  return EAS_OK;
}

}  // namespace eas_sampler
