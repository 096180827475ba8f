// (a-2) Adaptive event sampler, forward, second tensor-core kernel ("row-folded Toeplitz").
// Replaces AdaptiveRSNNEmbedding.forward / update (yolox/models/embedding.py:132-226) with the
// Rectangle spike function (activation.py:17-30) for depth 2, kernel 5 -- the published
// configuration -- on integer-valued inputs (event counts: exact in one fp16 plane).
//
// What limits the first kernel (sampler_tc.cu): a shared-memory-operand tcgen05.mma of M = 128,
// K = 16 costs 64 cycles whatever N <= 128 is (scripts/umma_rate_probe.cu), and it spent 60 such
// MMAs of N = 32 / 48 per 512 pixels.  This kernel fills N:
//   * a matrix row is still one QUAD (4 px, all channels) and the x taps are still a block-Toeplitz
//     B operand over an 8-pixel window (two K halves h = 0, 1 = quad rows m, m + 1);
//   * but the N axis now carries FOUR output image rows dy = 0..3 (x 4 px x channels): an M tile is
//     128 quads of a ROW GROUP (4 consecutive image rows), and the MMA for input row i = dy + ky
//     (0..7) of the group feeds every dy with tap row ky = i - dy in one instruction.  Rows are
//     stored in 4 interleaved planes (image row mod 4), so the operand of input row i is plane
//     i & 3 shifted by (i >> 2) * QPR + h quad rows -- again a plain start-address shift;
//   * the B operand of input row i is a WINDOW of 1..4 consecutive 32-row blocks of ONE array
//     [W4 | W3 | W2 | W1 | W0] (N = 32..128, accumulator columns shifted by dy_lo * 32), so the
//     weights stay 40 KB of shared memory;
//   * 8 input rows per 4 output rows instead of 5 per 1: 96 MMAs per 2048 pixels instead of 240.
// Precision as before: weights x 2^8 as fp16 hi + lo planes, hidden activations as fp16 hi + lo
// planes, fp32 accumulation in TMEM; the partial products now land in the SAME accumulator columns
// (layer 1) or in two column halves that the epilogue adds (layer 2).  The counts must be exactly
// representable in fp16 (integers <= 2048): anything else raises the workspace flag and the caller
// falls back (sampler_fwd.cu), like magnitudes beyond the fp16 range.
//
// State between the per-step launches is compact: vm, acc (f32, updated in place), seg | t_last in
// one byte, spikes as one byte per (channel, quad), double buffered: 9.5 B per element instead of 18.
//
// One persistent CTA per SM, 21 warps:
//   producers (4)  : counts + previous spike bits -> fp16 [ev | spk] rows -> X0 ring (4 row planes); one tile
//                    quarter each, the loads of the next tile in flight while the current one is converted
//   MMA (1 thread) : layer 1 (X0 -> D1) and layer 2 (X1 hi / lo -> D2), accumulators in TMEM, issued in the fixed
//                    interleave ... L2(k), L1(k + 3), L2(k + 1) ... so that a layer-1 tile fills the hand-over
//                    between two layer-2 tiles
//   epilogues (16) : every warp = (TMEM lane quarter, row pair, group) does both epilogues in the same interleave:
//                    epilogue 1 of hidden tile k + 2 (D1 * 2^-8 + bias, ReLU, image mask, hi / lo split -> X1 ring;
//                    group = conv stack), then epilogue 2 of output tile k (D2 -> sigmoid gate, membrane update,
//                    threshold / reset, spike-triggered read-out, state write-back: statement for statement
//                    sampler_step_kernel; group = sampler channel)
// Rings hold two 128-position tiles per plane plus a 32-row mirror of tile slot 0, filled in 32-position
// chunks (the operand of a tile reaches QPR + 1 <= 32 positions into the next one).  Weight rows (= accumulator
// columns) are ordered so that each epilogue group reads one contiguous column run.
//
// What bounds it (measured, profiles/r2_*): the N = 128 MMAs run at the tensor core's full math rate AND at the
// full shared-memory operand bandwidth (4 KB of A + 4 KB of B per 64 cycles = 128 B / clock), and the
// element-wise work of the two epilogues (~24 k warp instructions per 2048-pixel tile) costs about as many issue
// slots as the MMAs cost cycles; the kernel alternates between the two rather than overlapping them fully.
#include "sampler_common.cuh"
#include "hist_u8.cuh"

namespace eas_sampler {
namespace {

constexpr int TILE = 128;        // positions (MMA rows) per tile
constexpr int RING = 2 * TILE;   // ring positions per plane
constexpr int ROWS = RING + 32;  // + mirror of the first chunk
constexpr int MAX_QPR = 31;      // operand reach QPR + 1 <= 32
constexpr int X0_ROWB = 32, X1_ROWB = 64;
constexpr int X0_PLANE = ROWS * X0_ROWB;   // 9216: one row plane (image row & 3)
constexpr int X1_DY = ROWS * X1_ROWB;      // 18432
constexpr int X1_PLANE = 4 * X1_DY;        // hi / lo plane
// weight image: [L1 pass a | L1 pass b | L2 h=0 | L2 h=1], each 5 blocks x 32 rows x 64 B (SW64)
constexpr int WBLK = 32 * 64;
constexpr int WARR = 5 * WBLK;             // 10240
constexpr int OFF_W1A = 0, OFF_W1B = WARR, OFF_W2 = 2 * WARR;
constexpr int WB_BYTES = 4 * WARR;         // 40960
constexpr int WIMG_BYTES = WB_BYTES + 64;  // + 12 bias floats
constexpr float W_SCALE = 256.0f, W_UNSCALE = 1.0f / 256.0f;
constexpr float F16_MAX = 65504.0f;
constexpr int OFF_X0 = WB_BYTES;                    // 40960
constexpr int OFF_X1 = OFF_X0 + 4 * X0_PLANE;       // 77824
constexpr int OFF_BAR = OFF_X1 + 2 * X1_PLANE;      // 225280
constexpr int NBAR = 8 + 2 + 8 + 2 + 8;
constexpr int OFF_MISC = OFF_BAR + NBAR * 8;
constexpr int SMEM_BYTES = OFF_MISC + 128 + 1024;   // + slack for the 1024 B alignment
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
// warps 0-15 epilogues (quarter x row pair x stack / channel), 16 MMA, 17-20 producers (one tile quarter each)
constexpr int W_MMA = 16, W_PROD = 17;
constexpr int NUM_THREADS = 21 * 32;
constexpr uint32_t SPIN_LIMIT = 1u << 26;
constexpr uint32_t TM_D1 = 0, TM_D2 = 256, TM_COLS = 512;
constexpr uint32_t ISSUE_ORDER = 0x76521043u;  // input rows 3, 4 first: they touch every dy (accumulate = 0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifdef EAS_TC2_DIAG
// Development aid (-DEAS_TC2_DIAG): a wait that times out records (barrier offset, parity, warp, block) and releases
// every other wait, so that a dead-locked launch ends and eas_debug_tc2_diag() can tell who was stuck.
__device__ unsigned int g_diag[64];
__device__ volatile int g_abort;
#endif
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
#ifdef EAS_TC2_DIAG
    if (g_abort) break;
    if (++spins > (1u << 18)) {
      if ((threadIdx.x & 31) == 0) {
        const unsigned int slot = atomicAdd(&g_diag[0], 1u);
        if (slot < 20) {
          g_diag[1 + 3 * slot] = addr & 0xfff;          // barrier offset inside the barrier block
          g_diag[2 + 3 * slot] = (parity << 16) | (threadIdx.x >> 5);
          g_diag[3 + 3 * slot] = blockIdx.x;
        }
      }
      g_abort = 1;
      break;
    }
#else
    if (++spins > SPIN_LIMIT) __trap();  // watchdog: trap instead of hanging the GPU
#endif
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
// K-major operand descriptors, base_offset 0: the swizzle phase comes from the absolute address, so
// the start may sit on any row (scripts/umma_shift_probe.cu, umma_r4_probe.cu).
__device__ __forceinline__ uint64_t sw64_desc(uint32_t saddr) {   // 64 B rows, 8-row groups 512 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ uint64_t sw32_desc(uint32_t saddr) {   // 32 B rows, 8-row groups 256 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(256u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack2_f16(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// (a, b) >= 0 -> packed fp16 pairs hi, lo with a ~= hi.x + lo.x (22 mantissa bits)
__device__ __forceinline__ void split2_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// predicated 4-byte store: no branch, so that the per-pixel update chains stay one basic block (ILP across pixels)
__device__ __forceinline__ void st_global_f32_if(float* p, float v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.u32 q, %2, 0;\n\t"
      "@q st.global.f32 [%0], %1;\n\t}"
      ::"l"(p), "f"(v), "r"((uint32_t)pred)
      : "memory");
}
__device__ __forceinline__ uint32_t ld_stream_u8(const uint8_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// Development aid (compiled in with -DEAS_TC2_TRACE only): CTA 0 logs (role, event, tile, clock) so that the
// pipeline's timeline can be read back with eas_debug_tc2_trace() -- see scripts/tc2_timeline.py.
#ifdef EAS_TC2_TRACE
constexpr int TRACE_PER_ROLE = 2048;
__device__ unsigned long long g_trace[5 * TRACE_PER_ROLE];
__device__ int g_trace_n[5];
// the record index lives in a register of the tracing warp (tr_n): a trace point is a clock read + one store
#define TC2_TRACE_DECL int tr_n[5] = {0, 0, 0, 0, 0};
#define TC2_TRACE(role, ev, tile)                                                                       \
  do {                                                                                                  \
    if (blockIdx.x == 0 && lane == 0 && a.t == 1 && tr_n[role] < TRACE_PER_ROLE) {                      \
      g_trace[(role) * TRACE_PER_ROLE + tr_n[role]] =                                                         \
          ((unsigned long long)(ev) << 56) | ((unsigned long long)((tile) & 0xffff) << 40) |            \
          ((unsigned long long)clock64() & 0xffffffffffull);                                             \
      g_trace_n[role] = ++tr_n[role];                                                                       \
    }                                                                                                   \
  } while (0)
#else
#define TC2_TRACE_DECL
#define TC2_TRACE(role, ev, tile) do {} while (0)
#endif

// ---- weight image -------------------------------------------------------------------------------
// Four arrays of 5 blocks (block u holds tap row ky = 4 - u) x 32 rows x 32 fp16 in the SWIZZLE_64B
// shared-memory layout (image base 1024 B aligned), then 12 bias floats.  tap = 4 * h + p - jpx is the
// x tap of window pixel p of K half h for output pixel jpx; weights pre-scaled by 2^8.  Rows (= accumulator
// columns) are ordered so that each epilogue group reads one contiguous 16-column run:
//   layer 1 (pass a: hi plane, pass b: lo plane): row n = st * 16 + jpx * 4 + hc (st = 0 input stack,
//     1 gate stack; hc = its hidden channel); k = h * 16 + c * 8 + p * 2 + ci with c = 0 counts, c = 1 previous
//     spikes; input-stack rows read the counts chunk, gate-stack rows the spike chunk, the rest is zero.
//   layer 2 (array per K half h): row n = part * 16 + c * 8 + jpx * 2 + t (part 0: hi plane, 1: lo plane;
//     c = sampler channel, t = 0 gate / 1 current, i.e. output channel co = 2 * t + c, embedding.py:172-174);
//     k = st * 16 + p * 4 + hc: the hidden activation rows hold [stack][pixel][channel].
__device__ __forceinline__ float f16_plane(float w, int plane) {
  const float ws = w * W_SCALE;
  const float hi = __half2float(__float2half_rn(ws));
  return plane == 0 ? hi : __half2float(__float2half_rn(ws - hi));
}

__global__ void sampler_tc2_pack_weights(const eas_sampler_weights w, uint8_t* img, int* flag) {
  constexpr int K = 5;
  const int per = 5 * 32 * 32;  // elements per array
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 4 * per; idx += gridDim.x * blockDim.x) {
    const int arr = idx / per, r = idx % per;
    const int u = r / 1024, n = (r / 32) % 32, k = r % 32;
    const int ky = 4 - u;
    float val = 0.0f;
    if (arr < 2) {
      const int st = n >> 4, jpx = (n >> 2) & 3, hc = n & 3;
      const int h = k >> 4, c = (k >> 3) & 1, p = (k & 7) >> 1, ci = k & 1, tap = 4 * h + p - jpx;
      if (tap >= 0 && tap < K && c == st) {
        const float* w0 = st == 0 ? w.in_w0 : w.gate_w0;
        val = f16_plane(w0[((hc * 2 + ci) * K + ky) * K + tap], arr);
      }
    } else {
      const int h = arr - 2;
      const int part = n >> 4, c = (n >> 3) & 1, jpx = (n >> 1) & 3, t = n & 1, co = 2 * t + c;
      const int st = k >> 4, p = (k >> 2) & 3, hc = k & 3, tap = 4 * h + p - jpx;
      if (tap >= 0 && tap < K) {
        const float* w1 = st == 0 ? w.in_w1 : w.gate_w1;
        val = f16_plane(w1[((co * 4 + hc) * K + ky) * K + tap], part);
      }
    }
    const int row = u * 32 + n;
    const int off = arr * WARR + row * 64 + (((k >> 3) ^ ((row >> 1) & 3)) << 4) + (k & 7) * 2;
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
struct Tc2Geo {
  int ns;    // strips per image row
  int TW;    // strip width in pixels (multiple of 4)
  int QPR;   // quads per strip row incl. 2 halo quads = TW/4 + 2
};

// Compact state of the tensor-core kernel (lives in the caller's workspace between the launches).
struct Tc2State {
  uint8_t* meta;          // [B][2][H][W]: seg | (t_last + 1) << 4
  const uint8_t* sb_prev; // [B][2][H][W/4]: spikes of the previous step, one byte per (channel, quad), bit px
  uint8_t* sb_next;
};

struct Seg {
  int b, xs, ya, nrows;
  int P0, P1, P2;      // positions (row groups x QPR) of the X0 / hidden / output planes
  int nt0, nt1, nt2;   // tiles
};

// CTA's share of (image, strip, row) space, cut into segments that stay inside one strip.  32-bit
// arithmetic (the host checks B * ns * H < 2^31): a 64-bit division is ~100 instructions in each role.
struct SegIter {
  uint32_t r, r_end;
  int H, ns, TW, QPR;
  __device__ SegIter(const StepArgs& a, const Tc2Geo& g) {
    const uint64_t total = (uint64_t)a.B * g.ns * a.H;
    r = (uint32_t)(total * blockIdx.x / gridDim.x);
    r_end = (uint32_t)(total * (blockIdx.x + 1) / gridDim.x);
    H = a.H, ns = g.ns, TW = g.TW, QPR = g.QPR;
  }
  __device__ bool next(Seg& s) {
    if (r >= r_end) return false;
    const uint32_t unit = r / (uint32_t)H;
    s.ya = (int)(r - unit * (uint32_t)H);
    s.b = (int)(unit / (uint32_t)ns);
    s.xs = (int)(unit - (uint32_t)s.b * (uint32_t)ns) * TW;
    const int left = (int)(r_end - r);
    s.nrows = left < H - s.ya ? left : H - s.ya;
    const int n2 = (s.nrows + 3) >> 2;          // output row groups; hidden n2 + 1, input n2 + 2
    s.P2 = n2 * QPR, s.P1 = s.P2 + QPR, s.P0 = s.P1 + QPR;
    s.nt0 = (s.P0 + TILE - 1) / TILE;
    s.nt1 = (s.P1 + TILE - 1) / TILE;
    s.nt2 = (s.P2 + TILE - 1) / TILE;
    r += (uint32_t)s.nrows;
    return true;
  }
};

// One X0 row of a producer thread: counts of the quad (both channels) + its spike byte -> fp16 [ev | spk]
// (chunk 0: counts, element = px * 2 + ch; chunk 1: spikes, same order).  Returns true when a count is
// not exact in one fp16 plane.  Not inlined: four call sites per unit, and the roles share the I-cache.
template <bool kInt>
__device__ __noinline__ bool x0_fill_row(uint4 c0, uint4 c1, uint32_t spk, uint32_t row, uint32_t sw, uint32_t mirror) {
  // spk = nibble of channel 0 | nibble of channel 1 << 4
  float e0[4], e1[4];
  if (kInt) {
    e0[0] = (float)(int)c0.x, e0[1] = (float)(int)c0.y, e0[2] = (float)(int)c0.z, e0[3] = (float)(int)c0.w;
    e1[0] = (float)(int)c1.x, e1[1] = (float)(int)c1.y, e1[2] = (float)(int)c1.z, e1[3] = (float)(int)c1.w;
  } else {
    e0[0] = __uint_as_float(c0.x), e0[1] = __uint_as_float(c0.y), e0[2] = __uint_as_float(c0.z), e0[3] = __uint_as_float(c0.w);
    e1[0] = __uint_as_float(c1.x), e1[1] = __uint_as_float(c1.y), e1[2] = __uint_as_float(c1.z), e1[3] = __uint_as_float(c1.w);
  }
  uint32_t hv[4], sv[4];
  bool bad = false;
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const __half2 h = __floats2half2_rn(e0[px], e1[px]);
    const float2 back = __half22float2(h);
    bad = bad || back.x != e0[px] || back.y != e1[px];  // not exact (or NaN)
    hv[px] = *reinterpret_cast<const uint32_t*>(&h);
    // spike bits c * 4 + px -> fp16 1.0 (0x3C00) in lanes (ch 0, ch 1)
    sv[px] = (((spk >> px) & 1u) * 0x3C00u) | (((spk >> (4 + px)) & 1u) * 0x3C000000u);
  }
  st_shared_v4(row + ((0u ^ sw) << 4), hv[0], hv[1], hv[2], hv[3]);
  st_shared_v4(row + ((1u ^ sw) << 4), sv[0], sv[1], sv[2], sv[3]);
  if (mirror) {  // ring rows 0..31 of tile slot 0 are mirrored behind the ring (same swizzle phase: 256 rows further)
    st_shared_v4(row + RING * X0_ROWB + ((0u ^ sw) << 4), hv[0], hv[1], hv[2], hv[3]);
    st_shared_v4(row + RING * X0_ROWB + ((1u ^ sw) << 4), sv[0], sv[1], sv[2], sv[3]);
  }
  return bad;
}

// exact counts of the saturated bytes of one quad (both channels) of the compact histogram
__device__ __noinline__ void sat_fix(uint4& c0, uint4& c1, const uint32_t* __restrict__ tail, uint32_t q0, uint32_t HW) {
  if (c0.x == kHistU8Sat) c0.x = hist_u8_lookup(tail, q0);
  if (c0.y == kHistU8Sat) c0.y = hist_u8_lookup(tail, q0 + 1);
  if (c0.z == kHistU8Sat) c0.z = hist_u8_lookup(tail, q0 + 2);
  if (c0.w == kHistU8Sat) c0.w = hist_u8_lookup(tail, q0 + 3);
  if (c1.x == kHistU8Sat) c1.x = hist_u8_lookup(tail, q0 + HW);
  if (c1.y == kHistU8Sat) c1.y = hist_u8_lookup(tail, q0 + HW + 1);
  if (c1.z == kHistU8Sat) c1.z = hist_u8_lookup(tail, q0 + HW + 2);
  if (c1.w == kHistU8Sat) c1.w = hist_u8_lookup(tail, q0 + HW + 3);
}

// kIn: 0 = fp32 counts, 1 = int32 counts, 2 = the compact byte histogram (hist_u8.cuh).
// kSimple: the published read-out (sum, Ts = 1, no ReLU on the frames, no saved sequences) -- seg is one bit
// and t_last is not needed, which halves the update code of epilogue 2.
// kStep (kSimple only): 0 first step, 1 middle, 2 last, 3 first and last (Tm == 1): compile-time first / last
// remove a third of the update's predicate logic; -1 = decided at run time.
template <int kIn, bool kSimple, int kStep>
__global__ void __launch_bounds__(NUM_THREADS, 1)
sampler_tc2_step_kernel(const StepArgs a, const Tc2State st, const Tc2Geo g, const uint8_t* __restrict__ wimg,
                        int* __restrict__ ovf_flag) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next step's grid may be scheduled behind this one
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* x0_full = bars;                 // [8] chunk = tile slot * 4 + quarter
  uint64_t* x0_empty = x0_full + 8;         // [2] tile slot
  uint64_t* x1_full = x0_empty + 2;         // [8]
  uint64_t* x1_empty = x1_full + 8;         // [2]
  uint64_t* d1_full = x1_empty + 2;         // [2]
  uint64_t* d1_empty = d1_full + 2;
  uint64_t* d2_full = d1_empty + 2;
  uint64_t* d2_empty = d2_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_MISC);
  float* sh_b = reinterpret_cast<float*>(smem + OFF_MISC + 16);  // 12 floats
  const uint32_t sX0 = smem_u32(smem + OFF_X0), sX1 = smem_u32(smem + OFF_X1), sWB = smem_u32(smem);

  // (through a shuffle: provably warp-uniform, so the role loops' bookkeeping stays on the uniform datapath)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int QPR = g.QPR;
  const bool first = kStep < 0 ? a.t == 0 : (kStep == 0 || kStep == 3);
  const bool last = kStep < 0 ? a.t == a.Tm - 1 : kStep >= 2;
  const int tm = a.Tm - 1 - a.t;  // newest micro-bin first (embedding.py:155-156)
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t BHW2 = (int64_t)a.B * 2 * HW;
  const int WQ = a.W >> 2;

  // ---- prologue: barriers, TMEM, weight image -> shared -------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(x0_full + i, 1), mbar_init(x1_full + i, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(x0_empty + i, 1), mbar_init(x1_empty + i, 1);
      mbar_init(d1_full + i, 1), mbar_init(d1_empty + i, 16);
      mbar_init(d2_full + i, 1), mbar_init(d2_empty + i, 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // programmatic dependent launch: the grid was allowed to start before the previous step (or the weight pack) had
  // finished; nothing above touched global memory, everything below may read what that grid wrote
  asm volatile("griddepcontrol.wait;" ::: "memory");
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
  TC2_TRACE_DECL

  if (warp >= W_PROD) {
    // ===================== producers: counts + previous spike bits -> X0 ring =====================
    // unit = (tile, quarter): warp pw fills quarter pw of every tile; a thread fills position
    // (row group, quad) of all four row planes.  The loads of the next unit are in flight while the
    // current one is converted.
    const int pw = warp - W_PROD;
    const uint32_t* sat_tail = reinterpret_cast<const uint32_t*>(
        reinterpret_cast<const uint8_t*>(a.events) + hist_u8_tail_offset((size_t)a.B * a.Tm * 2 * HW));
    (void)sat_tail;
    SegIter it(a, g);
    Seg s;
    int gt = 0;  // global X0 tile counter at the segment start
    while (it.next(s)) {
      // a thread's unit: one quad of both channels in each of the 4 row planes -- 16 B per (row, channel) of fp32 /
      // int32 counts, 4 B of the compact byte histogram
      using V = typename std::conditional<kIn == 2, uint32_t, uint4>::type;
      constexpr int EB = kIn == 2 ? 1 : 4;   // bytes per count
      const int64_t img_bin0 = ((int64_t)s.b * a.Tm + tm) * 2 * HW;
      const uint8_t* ev_img = reinterpret_cast<const uint8_t*>(a.events) + img_bin0 * EB;
      const uint8_t* sp_img = st.sb_prev + (int64_t)s.b * 2 * a.H * WQ;
      V cur[4][2], nxt[4][2] = {};
      uint32_t csp[4], nsp[4] = {};
      auto load_unit = [&](int u, V (&v)[4][2], uint32_t (&sp)[4]) {
        const int p = u * TILE + pw * 32 + lane;
        const int grp = p / QPR, m = p - grp * QPR;
        const int x = s.xs - 4 + 4 * m;
        const bool col_ok = p < s.P0 && (unsigned)x < (unsigned)a.W;
#pragma unroll
        for (int rho = 0; rho < 4; ++rho) {
          const int y = s.ya - 4 + 4 * grp + rho;
          const bool ok = col_ok && (unsigned)y < (unsigned)a.H;
          const int64_t off = ok ? (int64_t)y * a.W + x : 0;
          if constexpr (kIn == 2) {
            v[rho][0] = ok ? ld_stream_u32(reinterpret_cast<const uint32_t*>(ev_img + off)) : 0u;
            v[rho][1] = ok ? ld_stream_u32(reinterpret_cast<const uint32_t*>(ev_img + HW + off)) : 0u;
          } else {
            v[rho][0] = ok ? ld_stream_u4(reinterpret_cast<const uint4*>(ev_img + off * 4)) : make_uint4(0u, 0u, 0u, 0u);
            v[rho][1] = ok ? ld_stream_u4(reinterpret_cast<const uint4*>(ev_img + (HW + off) * 4)) : make_uint4(0u, 0u, 0u, 0u);
          }
          sp[rho] = (ok && !first) ? ((ld_cg_u8(sp_img + y * WQ + (x >> 2)) & 15u) |
                                      (ld_cg_u8(sp_img + (a.H + y) * WQ + (x >> 2)) << 4))
                                   : 0u;
        }
      };
      const int nu = s.nt0;
      // software pipeline with ONE inlined copy of the loads: iteration u issues unit u and converts unit u - 1
#pragma unroll 1
      for (int uu = 0; uu <= nu; ++uu) {
#pragma unroll
        for (int rho = 0; rho < 4; ++rho) cur[rho][0] = nxt[rho][0], cur[rho][1] = nxt[rho][1], csp[rho] = nsp[rho];
        if (uu < nu) load_unit(uu, nxt, nsp);
        if (uu == 0) continue;
        const int u = uu - 1;
        const int gi = gt + u, slot = gi & 1, qt = pw;
        if (pw == 0) TC2_TRACE(0, 0, gi);
        mbar_wait(x0_empty + slot, ((gi >> 1) & 1) ^ 1);
        if (pw == 0) TC2_TRACE(0, 1, gi);
        const int pos = slot * TILE + qt * 32 + lane;
        const uint32_t row0 = sX0 + (uint32_t)pos * X0_ROWB, sw = ((uint32_t)pos >> 2) & 1u;
        const uint32_t mirror = (slot == 0 && qt == 0) ? 1u : 0u;
        bool bad = false;
#pragma unroll
        for (int rho = 0; rho < 4; ++rho) {
          if constexpr (kIn == 2) {
            const uint32_t w0 = cur[rho][0], w1 = cur[rho][1];
            uint4 c0 = make_uint4(w0 & 0xffu, (w0 >> 8) & 0xffu, (w0 >> 16) & 0xffu, w0 >> 24);
            uint4 c1 = make_uint4(w1 & 0xffu, (w1 >> 8) & 0xffu, (w1 >> 16) & 0xffu, w1 >> 24);
            if (__vcmpeq4(w0, 0xffffffffu) | __vcmpeq4(w1, 0xffffffffu)) {
              // a saturated byte (rare): its exact count is in the list behind the histogram
              const int p = u * TILE + pw * 32 + lane;
              const int grp = p / QPR, m = p - grp * QPR;
              const uint32_t q0 = (uint32_t)(img_bin0 + (int64_t)(s.ya - 4 + 4 * grp + rho) * a.W + (s.xs - 4 + 4 * m));
              sat_fix(c0, c1, sat_tail, q0, (uint32_t)HW);
            }
            bad |= x0_fill_row<true>(c0, c1, csp[rho], row0 + (uint32_t)rho * X0_PLANE, sw, mirror);
          } else {
            bad |= x0_fill_row<kIn == 1>(cur[rho][0], cur[rho][1], csp[rho], row0 + (uint32_t)rho * X0_PLANE, sw, mirror);
          }
        }
        if (bad) *ovf_flag = 1;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(x0_full + slot * 4 + qt);
        if (pw == 0) TC2_TRACE(0, 2, gi);
      }
      gt += s.nt0;
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer: layer 1 (X0 -> D1) and layer 2 (X1 hi / lo -> D2) =====================
    // ONE warp issues both layers in a fixed interleave: ... L2(k), L1(k + 3), L2(k + 1), L1(k + 4) ...
    // (tile numbers in the X1 / D1 sequence).  L2(k + 1) cannot start before epilogue 1 has written the first
    // chunk of X1 tile k + 2, and that slot is only free once L2(k) has completed: the layer-1 tile queued
    // behind L2(k) keeps the tensor pipe busy during that hand-over.  Two independent issuers interleaved the
    // two layers at random and left the pipe idle there.
    const bool leader = elect_one();
    constexpr uint32_t idesc0 = (1u << 4) | ((uint32_t)(TILE >> 4) << 24);  // fp16 x fp16 -> fp32, M = 128
    const uint64_t wa = sw64_desc(sWB + OFF_W1A), wb = sw64_desc(sWB + OFF_W1B), w2 = sw64_desc(sWB + OFF_W2);
    const uint64_t x0d = sw32_desc(sX0), x1d = sw64_desc(sX1);
    // ---- layer-1 stream ----
    SegIter it1(a, g);
    Seg s1;
    int k1 = 0, g0 = 0, g1a = 0, w0 = 0;  // tile in segment; global X0 / D1 tile counters at the segment start; X0 chunks waited
    bool more1 = it1.next(s1);
    auto issue_l1 = [&]() {
      const int k = k1, gi = g1a + k, buf = gi & 1;
      const int slot = (g0 + k) & 1;
      TC2_TRACE(1, 0, gi);
      // chunks of this tile and the first chunk of the next (the operands reach QPR + 1 positions ahead)
      const int need = (g0 + min(k + 1, s1.nt0 - 1)) * 4 + (k + 1 < s1.nt0 ? 1 : 4);
      while (w0 < need) mbar_wait(x0_full + (w0 & 7), (w0 >> 3) & 1), ++w0;
      TC2_TRACE(1, 1, gi);
      mbar_wait(d1_empty + buf, ((gi >> 1) & 1) ^ 1);
      TC2_TRACE(1, 2, gi);
      tc_fence_after();
      const uint32_t d1 = tmem_base + TM_D1 + buf * 128;
      const uint64_t abase = x0d + (uint64_t)((slot * TILE * X0_ROWB) >> 4);
      // The accumulation in TMEM truncates: the small (w_lo) terms go first, while the accumulator is still
      // small, so that only the 10 w_hi MMAs per output truncate at full magnitude.
#pragma unroll 1
      for (int n = 0; n < 16; ++n) {
        const int i = (ISSUE_ORDER >> (4 * (n & 7))) & 7;
        const int dy_lo = i > 4 ? i - 4 : 0, dy_hi = i < 3 ? i : 3;
        const int sb = 4 - (i - dy_lo);
        const uint32_t idesc = idesc0 | ((uint32_t)(dy_hi - dy_lo + 1) << 19);  // N = 32 * blocks
        const uint32_t dcol = d1 + dy_lo * 32;
        const uint64_t arow = abase + (uint64_t)(((i & 3) * X0_PLANE + (i >> 2) * QPR * X0_ROWB) >> 4);
        const uint64_t bw = (n < 8 ? wb : wa) + (uint64_t)((sb * WBLK) >> 4);
        if (leader) {
          tc_mma(dcol, arow, bw, idesc, n == 0 ? 0u : 1u);   // h = 0
          tc_mma(dcol, arow + 2, bw + 2, idesc, 1u);          // h = 1: next quad row, second K half
        }
      }
      TC2_TRACE(1, 3, gi);
      if (leader) {
        tc_commit(d1_full + buf);
        tc_commit(x0_empty + slot);
      }
      if (++k1 == s1.nt1) {
        // A trailing X0 tile (only its first chunk was an operand) is freed here -- but only once ALL its chunks
        // have been written: a producer that has not yet passed its wait for this slot would otherwise see the
        // slot's barrier two phases ahead (parity aliasing) and wait forever.
        while (w0 < (g0 + s1.nt0) * 4) mbar_wait(x0_full + (w0 & 7), (w0 >> 3) & 1), ++w0;
        if (leader)
          for (int kk = s1.nt1; kk < s1.nt0; ++kk) tc_commit(x0_empty + ((g0 + kk) & 1));
        g0 += s1.nt0, g1a += s1.nt1, k1 = 0;
        more1 = it1.next(s1);
      }
    };
    // ---- layer-2 stream ----
    SegIter it2(a, g);
    Seg s2;
    int k2 = 0, g1b = 0, g2 = 0, w1 = 0;
    bool more2 = it2.next(s2);
    int issued1 = 0;
    while (more2) {
      while (more1 && issued1 < g1b + k2 + 3) issue_l1(), ++issued1;
      {
        const int k = k2, gj = g2 + k, buf = gj & 1;
        const int slot = (g1b + k) & 1;
        TC2_TRACE(2, 0, gj);
        const int need = (g1b + min(k + 1, s2.nt1 - 1)) * 4 + (k + 1 < s2.nt1 ? 1 : 4);
        while (w1 < need) mbar_wait(x1_full + (w1 & 7), (w1 >> 3) & 1), ++w1;
        TC2_TRACE(2, 1, gj);
        mbar_wait(d2_empty + buf, ((gj >> 1) & 1) ^ 1);
        TC2_TRACE(2, 2, gj);
        tc_fence_after();
        const uint32_t d2 = tmem_base + TM_D2 + buf * 128;
        const uint64_t abase = x1d + (uint64_t)((slot * TILE * X1_ROWB) >> 4);
        // lo plane first (see layer 1): 20 truncating accumulations at full magnitude per output, not 40
#pragma unroll 1
        for (int n = 0; n < 16; ++n) {
          const int i = (ISSUE_ORDER >> (4 * (n & 7))) & 7;
          const int dy_lo = i > 4 ? i - 4 : 0, dy_hi = i < 3 ? i : 3;
          const int sb = 4 - (i - dy_lo);
          const uint32_t idesc = idesc0 | ((uint32_t)(dy_hi - dy_lo + 1) << 19);
          const uint32_t dcol = d2 + dy_lo * 32;
          const uint64_t ap = abase + (uint64_t)(((n < 8 ? X1_PLANE : 0) + (i & 3) * X1_DY + (i >> 2) * QPR * X1_ROWB) >> 4);
          const uint64_t bw = w2 + (uint64_t)((sb * WBLK) >> 4);
          if (leader) {
            tc_mma(dcol, ap, bw, idesc, n == 0 ? 0u : 1u);                         // h = 0, px 0-1
            tc_mma(dcol, ap + 2, bw + 2, idesc, 1u);                                // h = 0, px 2-3
            tc_mma(dcol, ap + 4, bw + (WARR >> 4), idesc, 1u);                      // h = 1 (next quad row)
            tc_mma(dcol, ap + 6, bw + ((WARR >> 4) + 2), idesc, 1u);
          }
        }
        TC2_TRACE(2, 3, gj);
        if (leader) {
          tc_commit(d2_full + buf);
          tc_commit(x1_empty + slot);
        }
        if (++k2 == s2.nt2) {
          // trailing X1 tile: freed only after all of its chunks were written (see the X0 ring above)
          while (w1 < (g1b + s2.nt1) * 4) mbar_wait(x1_full + (w1 & 7), (w1 >> 3) & 1), ++w1;
          if (leader)
            for (int kk = s2.nt2; kk < s2.nt1; ++kk) tc_commit(x1_empty + ((g1b + kk) & 1));
          g1b += s2.nt1, g2 += s2.nt2, k2 = 0;
          more2 = it2.next(s2);
        }
      }
    }
  } else {
    // ===================== epilogue warps (16): both epilogues, interleaved like the MMA issuer =====================
    // warp = (quarter qt = TMEM lanes, row pair hf, group gp).  Per round a warp converts its share of hidden tile
    // k + 2 (epilogue 1: stack gp, rows 2 hf, 2 hf + 1 -> X1 ring; that is what L2(k + 1) is waiting for) and then
    // updates its share of output tile k (epilogue 2: channel gp, rows 2 hf, 2 hf + 1).  Sixteen warps run the same
    // two small loop bodies: separate epilogue-1 / epilogue-2 warps left one set idle 70 % of the time while the
    // other was the bottleneck, and the roles share the instruction caches.
    const int qt = warp & 3, hf = (warp >> 2) & 1, gp = warp >> 3;
    const float b0 = sh_b[4 + gp * 4], b1 = sh_b[5 + gp * 4], b2 = sh_b[6 + gp * 4], b3 = sh_b[7 + gp * 4];
    const float bg = sh_b[gp], bc = sh_b[2 + gp];
    const int iHW = a.H * a.W;  // per-image offsets fit 32 bits (host-checked)
    const float thresh = a.thresh, vreset = a.vreset;
    const bool hard = a.hard_reset != 0, wz = a.write_zero != 0;
    float hmax = 0.0f;

    // ---- epilogue 1 stream: D1 -> bias, ReLU, mask, hi / lo split -> X1 ring ----
    SegIter it1(a, g);
    Seg s1;
    int k1 = 0, gt1 = 0;
    bool more1 = it1.next(s1);
    auto do_e1 = [&]() {
      const int gt = gt1, buf = gt & 1, slot = gt & 1;
      const int p = k1 * TILE + qt * 32 + lane;
      const int grp = p / QPR, m = p - grp * QPR;
      const int x1 = s1.xs - 2 + 4 * m;
      const bool pos_in = p < s1.P1;
      // D1 is ready long before the X1 slot is free (layer 2 is the long stage): the first row is converted
      // while waiting for the slot, so that only its stores and the second row follow the hand-over
      if (warp == 0) TC2_TRACE(3, 0, gt);
      mbar_wait(d1_full + buf, (gt >> 1) & 1);
      if (warp == 0) TC2_TRACE(3, 1, gt);
      tc_fence_after();
      const int pos = slot * TILE + qt * 32 + lane;
      const bool mirror = slot == 0 && qt == 0;
      const uint32_t sw = ((uint32_t)pos >> 1) & 3u;  // the mirror (256 rows further) has the same phase
      // row layout [stack][px][hidden channel]: this stack's 32 B = chunks 2 gp, 2 gp + 1
      const uint32_t row0 = sX1 + (uint32_t)pos * X1_ROWB;
      const uint32_t ca0 = row0 + (((uint32_t)(2 * gp) ^ sw) << 4), ca1 = row0 + (((uint32_t)(2 * gp + 1) ^ sw) << 4);
#pragma unroll 1
      for (int d = 0; d < 2; ++d) {
        const int dy = 2 * hf + d;
        uint32_t acc[16];
        tmem_ld16(tmem_base + ((uint32_t)(qt * 32) << 16) + TM_D1 + buf * 128 + dy * 32 + gp * 16, acc);
        tmem_ld_wait();
        if (d == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d1_empty + buf);
        }
        const int y1 = s1.ya - 2 + 4 * grp + dy;
        const bool row_in = pos_in && (unsigned)y1 < (unsigned)a.H;
        uint32_t hi[8], lo[8];  // [px][2 x (2 channels)]
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const bool in_img = row_in && (unsigned)(x1 + px) < (unsigned)a.W;
          float h0 = fmaxf(fmaf(__uint_as_float(acc[px * 4 + 0]), W_UNSCALE, b0), 0.0f);
          float h1 = fmaxf(fmaf(__uint_as_float(acc[px * 4 + 1]), W_UNSCALE, b1), 0.0f);
          float h2 = fmaxf(fmaf(__uint_as_float(acc[px * 4 + 2]), W_UNSCALE, b2), 0.0f);
          float h3 = fmaxf(fmaf(__uint_as_float(acc[px * 4 + 3]), W_UNSCALE, b3), 0.0f);
          h0 = in_img ? h0 : 0.0f, h1 = in_img ? h1 : 0.0f, h2 = in_img ? h2 : 0.0f, h3 = in_img ? h3 : 0.0f;
          hmax = fmaxf(fmaxf(hmax, fmaxf(h0, h1)), fmaxf(h2, h3));
          split2_pair(h0, h1, hi[2 * px], lo[2 * px]);
          split2_pair(h2, h3, hi[2 * px + 1], lo[2 * px + 1]);
        }
        if (d == 0) {
          mbar_wait(x1_empty + slot, ((gt >> 1) & 1) ^ 1);
          if (warp == 0) TC2_TRACE(3, 2, gt);
        }
        const uint32_t dyo = (uint32_t)dy * X1_DY;
        st_shared_v4(ca0 + dyo, hi[0], hi[1], hi[2], hi[3]);
        st_shared_v4(ca1 + dyo, hi[4], hi[5], hi[6], hi[7]);
        st_shared_v4(ca0 + dyo + X1_PLANE, lo[0], lo[1], lo[2], lo[3]);
        st_shared_v4(ca1 + dyo + X1_PLANE, lo[4], lo[5], lo[6], lo[7]);
        if (mirror) {
          st_shared_v4(ca0 + dyo + RING * X1_ROWB, hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(ca1 + dyo + RING * X1_ROWB, hi[4], hi[5], hi[6], hi[7]);
          st_shared_v4(ca0 + dyo + RING * X1_ROWB + X1_PLANE, lo[0], lo[1], lo[2], lo[3]);
          st_shared_v4(ca1 + dyo + RING * X1_ROWB + X1_PLANE, lo[4], lo[5], lo[6], lo[7]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(x1_full + slot * 4 + qt);
      if (warp == 0) TC2_TRACE(3, 3, gt);
      ++gt1;
      if (++k1 == s1.nt1) k1 = 0, more1 = it1.next(s1);
    };

    // ---- epilogue 2 stream: D2 -> membrane update + spike-triggered read-out ----
    // Same arithmetic, statement for statement, as sampler_step_kernel (embedding.py:132-139, 177-217).
    SegIter it2(a, g);
    Seg s2;
    int k2 = 0, gt2 = 0, g1b = 0;  // tile in segment, global D2 tile counter, X1 tiles before this segment
    bool more2 = it2.next(s2);
    int done1 = 0;
    while (more2) {
      const int64_t img = ((int64_t)s2.b * 2 + gp) * HW;
      float* vm_b = a.vm + img;
      float* acc_b = a.acc + img;
      uint8_t* meta_b = st.meta + img;
      float* out_b = a.out + img;
      uint8_t* sb_b = st.sb_next + ((int64_t)s2.b * 2 + gp) * a.H * WQ;
      // state of this warp's two rows of the tile: in flight while the hidden tile is converted
      float4 vq0 = make_float4(0.f, 0.f, 0.f, 0.f), aq0 = vq0, vq1 = vq0, aq1 = vq0;
      uint32_t mt0 = 0u, mt1 = 0u;
      int off0 = -1, off1 = -1;  // y * W + x of the quad, -1 = nothing to do
      {
        const int p = k2 * TILE + qt * 32 + lane;
        const int grp = p / QPR, m = p - grp * QPR;
        const int gx = s2.xs + 4 * m;
        const int r2 = 4 * grp + 2 * hf;
        const bool col_ok = p < s2.P2 && m < QPR - 2 && gx < a.W;
        if (col_ok && r2 < s2.nrows) off0 = (s2.ya + r2) * a.W + gx;
        if (col_ok && r2 + 1 < s2.nrows) off1 = (s2.ya + r2 + 1) * a.W + gx;
        if (!first) {
          if (off0 >= 0) {
            vq0 = ld_cg_f4(reinterpret_cast<const float4*>(vm_b + off0));
            aq0 = ld_cg_f4(reinterpret_cast<const float4*>(acc_b + off0));
            mt0 = ld_cg_u32(reinterpret_cast<const uint32_t*>(meta_b + off0));
          }
          if (off1 >= 0) {
            vq1 = ld_cg_f4(reinterpret_cast<const float4*>(vm_b + off1));
            aq1 = ld_cg_f4(reinterpret_cast<const float4*>(acc_b + off1));
            mt1 = ld_cg_u32(reinterpret_cast<const uint32_t*>(meta_b + off1));
          }
        }
      }
      while (more1 && done1 < g1b + k2 + 3) do_e1(), ++done1;
      const int gk = gt2, buf = gk & 1;
      if (warp == 0) TC2_TRACE(4, 0, gk);
      mbar_wait(d2_full + buf, (gk >> 1) & 1);
      if (warp == 0) TC2_TRACE(4, 1, gk);
      tc_fence_after();
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int dy = 2 * hf + d;
        float pre[8];  // [px][gate, current], still x 2^8
        {
          uint32_t dh[8], dl[8];  // (x_hi + x_lo) * w_hi | (x_hi + x_lo) * w_lo
          const uint32_t t2 = tmem_base + ((uint32_t)(qt * 32) << 16) + TM_D2 + buf * 128 + dy * 32 + gp * 8;
          tmem_ld8(t2, dh);
          tmem_ld8(t2 + 16, dl);
          tmem_ld_wait();
#pragma unroll
          for (int n = 0; n < 8; ++n) pre[n] = __uint_as_float(dh[n]) + __uint_as_float(dl[n]);
        }
        if (d == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2_empty + buf);
          if (warp == 0) TC2_TRACE(4, 2, gk);
        }
        const int off = d ? off1 : off0;
        if (off >= 0) {
          const float4 vq = d ? vq1 : vq0, aq = d ? aq1 : aq0;
          const uint32_t mt = d ? mt1 : mt0;
          uint32_t spk_bits = 0u, mt_out = 0u;
          float vm4[4] = {vq.x, vq.y, vq.z, vq.w};
          float ac4[4] = {aq.x, aq.y, aq.z, aq.w};
          float o4[4];
          float* outp = out_b + off;  // plane k at outp + k*BHW2
          if (kSimple) {
            // straight-line code (selects and predicated stores only): the four pixel chains interleave
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              const float gate = __fdividef(1.0f, 1.0f + __expf(-(pre[2 * px] * W_UNSCALE + bg)));
              const float cur = pre[2 * px + 1] * W_UNSCALE + bc;
              const bool fired = ((mt >> (8 * px)) & 1u) != 0u;
              const float v = __fadd_rn(__fmul_rn(gate, vm4[px]), cur);
              const float vmt = __fsub_rn(v, thresh);
              const bool sp = vmt > 0.0f;
              const float vm = sp ? (hard ? vreset : vmt) : v;
              const float acs = __fadd_rn(ac4[px], v);
              const bool vld = sp && !fired;
              const bool tail = last && !sp && !fired && !wz;  // residual of a pixel that never fired
              const float ac = sp ? 0.0f : acs;
              // first step: one zero-initialising vector store (Tm == 1: the tail is still inside it);
              // later steps: scattered stores at the first spike / the tail
              o4[px] = vld ? acs : (tail ? ac : 0.0f);
              st_global_f32_if(outp + px, vld ? acs : ac, !first && (vld || tail));
              vm4[px] = vm, ac4[px] = ac;
              spk_bits |= (sp ? 1u : 0u) << px;
              mt_out |= ((fired || vld) ? 1u : 0u) << (8 * px);
            }
            if (first) *reinterpret_cast<float4*>(outp) = make_float4(o4[0], o4[1], o4[2], o4[3]);
          } else {
            float v4[4], g4[4];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              const float gate = __fdividef(1.0f, 1.0f + __expf(-(pre[2 * px] * W_UNSCALE + bg)));
              const float cur = pre[2 * px + 1] * W_UNSCALE + bc;
              int seg = (mt >> (8 * px)) & 0xf;
              int tl = (int)((mt >> (8 * px + 4)) & 0xf) - 1;
              const float v = __fadd_rn(__fmul_rn(gate, vm4[px]), cur);
              const bool sp = __fsub_rn(v, thresh) > 0.0f;
              const float vm = sp ? (hard ? vreset : __fsub_rn(v, thresh)) : v;
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
              if (last && !sp && seg < a.Ts && !wz) {
                float tv = a.readout == EAS_READOUT_SUM ? ac : vm;
                if (a.readout == EAS_READOUT_AVG) tv = ac / (float)(a.Tm - 1 - tl);
                if (a.use_abs) tv = fmaxf(tv, 0.0f);
                if (first && seg == 0) o4[px] = tv;  // Tm == 1: still inside the zero-initialising store
                else outp[(int64_t)seg * BHW2 + px] = tv;
              }
              vm4[px] = vm, ac4[px] = ac, v4[px] = v, g4[px] = gate;
              spk_bits |= (sp ? 1u : 0u) << px;
              mt_out |= (uint32_t)(seg | ((tl + 1) << 4)) << (8 * px);
            }
            if (first) {
              *reinterpret_cast<float4*>(outp) = make_float4(o4[0], o4[1], o4[2], o4[3]);
              for (int kk = 1; kk < a.Ts; ++kk)
                *reinterpret_cast<float4*>(outp + kk * BHW2) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (a.v_seq) {
              const int64_t se = (int64_t)a.t * BHW2 + img + off;
              *reinterpret_cast<float4*>(a.v_seq + se) = make_float4(v4[0], v4[1], v4[2], v4[3]);
              *reinterpret_cast<float4*>(a.gate_seq + se) = make_float4(g4[0], g4[1], g4[2], g4[3]);
            }
          }
          if (!last) {
            *reinterpret_cast<float4*>(vm_b + off) = make_float4(vm4[0], vm4[1], vm4[2], vm4[3]);
            *reinterpret_cast<float4*>(acc_b + off) = make_float4(ac4[0], ac4[1], ac4[2], ac4[3]);
            *reinterpret_cast<uint32_t*>(meta_b + off) = mt_out;
            sb_b[off >> 2] = (uint8_t)spk_bits;
          }
        }
      }
      if (warp == 0) TC2_TRACE(4, 3, gk);
      ++gt2;
      if (++k2 == s2.nt2) g1b += s2.nt1, k2 = 0, more2 = it2.next(s2);
    }
    if (!(hmax < F16_MAX)) *ovf_flag = 1;  // a hidden activation left the fp16 range
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS) : "memory");
  }
}

Tc2Geo pick_geo(int W) {
  Tc2Geo g;
  const int max_tw = 4 * (MAX_QPR - 2);
  g.ns = (W + max_tw - 1) / max_tw;
  g.TW = ((W + g.ns - 1) / g.ns + 3) / 4 * 4;
  g.QPR = g.TW / 4 + 2;
  return g;
}

}  // namespace

size_t eas_sampler_tc2_wimg_bytes() { return (size_t)WIMG_BYTES + 64; }  // + the fall-back flag

const int* eas_sampler_tc2_flag(const void* wimg) {
  return reinterpret_cast<const int*>(reinterpret_cast<const uint8_t*>(wimg) + WIMG_BYTES);
}

bool eas_sampler_tc2_supported(const eas_sampler_cfg* c, const void* events, const float* out, const float* v_seq,
                               const float* gate_seq) {
  if (c->depth != 2 || c->ksize != 5) return false;
  if (c->W % 4 != 0) return false;          // quads are loaded / stored as 16 B vectors
  if (c->Ts > 15 || c->Tm > 14) return false;  // seg | t_last + 1 share one byte
  if (((uintptr_t)events | (uintptr_t)out | (uintptr_t)v_seq | (uintptr_t)gate_seq) % 16 != 0) return false;
  return true;
}

// meta8: B*2*H*W bytes; sb0 / sb1: B*2*H*W/4 bytes each (all inside the caller's workspace)
int eas_sampler_tc2_run(const eas_sampler_cfg* cfg, StepArgs a, uint8_t* meta8, uint8_t* sb0, uint8_t* sb1, void* wimg,
                        cudaStream_t st) {
  EAS_REQUIRE((uintptr_t)wimg % 16 == 0 && (uintptr_t)meta8 % 4 == 0, EAS_E_ALIGN);
  int* flag = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(wimg) + WIMG_BYTES);
  // (a plain launch: the pack WRITES into the caller's workspace without reading anything first, so it must not start
  // while the kernel before it may still be using memory the stream-ordered allocator has recycled into that workspace)
  sampler_tc2_pack_weights<<<40, 256, 0, st>>>(a.w, reinterpret_cast<uint8_t*>(wimg), flag);
  EAS_LAUNCH_CHECK();
  const Tc2Geo g = pick_geo(cfg->W);
  EAS_REQUIRE((int64_t)cfg->B * g.ns * cfg->H < (1ll << 31) && (int64_t)cfg->H * cfg->W * 2 < (1ll << 31), EAS_E_SHAPE);
  const bool simple = cfg->readout == EAS_READOUT_SUM && cfg->Ts == 1 && !cfg->use_abs && a.v_seq == nullptr;
  using kern_t = void (*)(const StepArgs, const Tc2State, const Tc2Geo, const uint8_t*, int*);
  const int ki = cfg->in_dtype == EAS_I32 ? 1 : (cfg->in_dtype == EAS_U8 ? 2 : 0);
  kern_t kerns[4];  // by step kind: first, middle, last, first and last
#define TC2_PICK(S, K) (ki == 2 ? sampler_tc2_step_kernel<2, S, K> : ki == 1 ? sampler_tc2_step_kernel<1, S, K> \
                                                                            : sampler_tc2_step_kernel<0, S, K>)
  if (simple) {
    kerns[0] = TC2_PICK(true, 0);
    kerns[1] = TC2_PICK(true, 1);
    kerns[2] = TC2_PICK(true, 2);
    kerns[3] = TC2_PICK(true, 3);
  } else {
    kerns[0] = kerns[1] = kerns[2] = kerns[3] = TC2_PICK(false, -1);
  }
#undef TC2_PICK
  for (int i = 0; i < 4; ++i) {
    cudaError_t e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t total_rows = (int64_t)cfg->B * g.ns * cfg->H;
  // one persistent CTA per SM; tiny problems get fewer CTAs (>= 8 strip rows each)
  int64_t grid = (total_rows + 7) / 8;
  if (grid > EAS_NUM_SMS) grid = EAS_NUM_SMS;
  if (grid < 1) grid = 1;
  Tc2State s2;
  s2.meta = meta8;
  for (int t = 0; t < cfg->Tm; ++t) {
    a.t = t;
    s2.sb_prev = (t & 1) ? sb0 : sb1;  // step t reads what step t-1 wrote
    s2.sb_next = (t & 1) ? sb1 : sb0;
    const int kind = (t == 0 ? (cfg->Tm == 1 ? 3 : 0) : (t == cfg->Tm - 1 ? 2 : 1));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid), lc.blockDim = dim3(NUM_THREADS), lc.dynamicSmemBytes = SMEM_BYTES, lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr, lc.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&lc, kerns[kind], a, s2, g, reinterpret_cast<const uint8_t*>(wimg), flag);
    if (le != cudaSuccess) return (int)le;
  }
  return EAS_OK;
}

}  // namespace eas_sampler

#ifdef EAS_TC2_DIAG
extern "C" int eas_debug_tc2_diag(unsigned int* out) {
  using namespace eas_sampler;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_diag, sizeof(unsigned int) * 64);
  unsigned int zero[64] = {0};
  int z = 0;
  cudaMemcpyToSymbol(g_diag, zero, sizeof(zero));
  cudaMemcpyToSymbol(g_abort, &z, sizeof(int));
  return (int)(OFF_BAR & 0xfff);
}
#endif

#ifdef EAS_TC2_TRACE
// development aid: copies the trace (5 roles x TRACE_PER_ROLE records, then 5 counts) to the host and resets it
extern "C" int eas_debug_tc2_trace(unsigned long long* rec, int* counts) {
  using namespace eas_sampler;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(rec, g_trace, sizeof(unsigned long long) * 5 * TRACE_PER_ROLE);
  cudaMemcpyFromSymbol(counts, g_trace_n, sizeof(int) * 5);
  int zero[5] = {0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(g_trace_n, zero, sizeof(zero));
  return TRACE_PER_ROLE;
}
#endif
