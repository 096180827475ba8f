// (a-5) conv -> folded BatchNorm -> multi-step PLIF on the 5th-gen tensor cores (inference).
// Replaces BaseConv.forward after convert_to_spiking: SeqToANNContainer(Conv2d) -> BatchNorm2d('m')
// -> ParametricLIFNode (yolox/models/network_blocks.py:52-53, yolox/utils/utils_snn.py:25-53), BN
// folded as in yolox/utils/model_utils.py:61-75.
//
// Implicit GEMM, one CTA per (128-pixel, BLOCK_N-channel) output tile:
//   M = 128 output pixels = NB images x TH rows x TW cols of ONE time step,
//   N = BLOCK_N output channels, K = taps x Cin walked in 64-channel blocks (one 128 B swizzle row).
//   A (activations, channels-last fp16) arrives by TMA as a 5-D box [1 plane][NB][TH][TW][64 ch]: the
//     box origin is shifted per filter tap, out-of-image rows/cols/channels are zero filled by the
//     TMA unit (= conv padding), stride-2 convs use the tensor map's element strides.  The box lands
//     in shared memory exactly in the K-major SWIZZLE_128B layout tcgen05 wants.
//   B (weights [split][Cout][tap][Cin] fp16) arrives by TMA as [BLOCK_N][1][64].
//   D: one fp32 accumulator per time step in TMEM (T x BLOCK_N columns), tcgen05.mma issued by one
//     thread.  fp32-equivalent accuracy on the 16-bit tensor cores comes from splitting the folded fp32
//     weight, scaled per output channel by a power of two (w_unscale undoes it in the epilogue), into
//     two fp16 planes hi + lo (22 mantissa bits); spikes / SEW sums are small integers, exact in fp16,
//     so A needs no split except for real-valued inputs (n_xsplit = 2 planes).  Two planes instead of
//     bf16's three = 2/3 of the MMAs, and every MMA costs a fixed 64 cycles here.
//   Epilogue (4 warps, one TMEM lane = one pixel each): tcgen05.ld the T accumulators of a 16-channel
//     chunk, add the folded BN shift, run the LIF recurrence over t with v in registers, store fp16
//     spikes channels-last.  The conv output and the membrane potential never touch HBM.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// Two mbarrier rings: B tiles are loaded once per K block and reused by all T time steps / input
// planes, A tiles stream through a deeper ring.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include "lif.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int MAX_WSPLIT = 2;
constexpr int N_ISSUERS = 3;        // MMA issuing warps (time steps t = w, w + 3, ...)
constexpr int NUM_THREADS = (1 + N_ISSUERS + 8) * 32;   // TMA producer, MMA issuers, 2 x 4 epilogue warps
constexpr uint32_t SPIN_LIMIT = 1u << 26;  // watchdog (a few seconds): trap instead of hanging the GPU

struct ConvArgs {
  int T, Tx, B, Ho, Wo, Cin, Cout, ksize, stride, pad;
  int n_wsplit, n_xsplit;
  int NB, TH, TW;
  int sa_ring;            // A ring slots in use: a multiple of the slots per K block (Tx * n_xsplit), so a group never wraps
  int a_group;            // 1: ONE TMA box [Tx * n_xsplit planes] per K block into consecutive slots (one full barrier)
  int ksplit;             // single-accumulator layers: 2 = two issuer warps on alternate B stages, two accumulators
  int TWp;                // tap-reuse mode: tile width incl. the 2 halo columns (rows of the tile = TH x TWp)
  int tiles_w, tiles_h, tiles_b, tiles_n;
  int out_ld, out_mode, res_ld;
  const __half* residual;   // SEW shortcut added to the spikes (network_blocks.py:99-103) or null
  float vth, vreset;
  int hard_reset, decay_input;
  const float* bias;
  const float* unscale;   // per output channel, multiplies the accumulator (null = 1)
  const float* plif_w;
  void* out;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 x fp16 -> fp32, issued by ONE thread for the whole CTA.
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate));   // volatile: ordered with the
  // surrounding mbarrier waits / commits (all asm volatile); no memory clobber, so loop invariants stay in registers
}
// K-major swizzled shared-memory operand descriptor.  One K block is one swizzle row of BK fp16
// (128 / 64 / 32 B), 8-row groups are 8 rows apart (1024 / 512 / 256 B).
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
  constexpr uint64_t layout = BK == 64 ? 2 : BK == 32 ? 4 : 6;  // SWIZZLE_128B / 64B / 32B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);       // start address
  d |= (uint64_t)((8u * BK * 2u) >> 4) << 32;     // stride byte offset
  d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
  d |= layout << 61;
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

__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// Tap-reuse mode (3x3, stride 1): an A slot holds the whole haloed input tile [(TH+2) x TWp pixels][BK] once;
// the nine taps are nine start-address shifts (ky*TWp + kx rows) of the same operand.  TWp <= 64.
constexpr int REUSE_ROWS = BLOCK_M + 2 * 64 + 8;

template <int BLOCK_N, int BK, bool REUSE = false>
struct SmemLayout {
  static constexpr int A_BYTES = (REUSE ? REUSE_ROWS : BLOCK_M) * BK * 2;
  static constexpr int B_BYTES = BLOCK_N * BK * 2;
  // one persistent CTA per SM: deep rings (A <= 128 KB, B <= 64 KB; 128-channel tiles: A <= 102 KB, B <= 96 KB)
  static constexpr int SA = BLOCK_N > 64 ? 6 : (REUSE ? 8 : 9);   // 9 = three K-block groups of T = 3 time steps
  static constexpr int B_BUDGET = (BLOCK_N > 64 ? 96 : 64) * 1024;
  static constexpr int SB = (B_BUDGET / (MAX_WSPLIT * B_BYTES)) > 8 ? 8 : (B_BUDGET / (MAX_WSPLIT * B_BYTES));
  static_assert(SB >= 2, "B ring too shallow");
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + SA * A_BYTES;
  static constexpr int OFF_BAR = OFF_B + SB * MAX_WSPLIT * B_BYTES;
  static constexpr int OFF_BIAS = OFF_BAR + 512;
  static constexpr int TOTAL = OFF_BIAS + 4 * BLOCK_N * 4 + 1024;  // (bias + unscale) x 2 groups, + alignment slack
};

struct TileCoord {
  int n0, wo0, ho0, b0;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvArgs& a, int tile, int block_n) {
  TileCoord c;
  const int tn = tile % a.tiles_n;
  tile /= a.tiles_n;
  const int tw = tile % a.tiles_w;
  tile /= a.tiles_w;
  const int th = tile % a.tiles_h;
  const int tb = tile / a.tiles_h;
  c.n0 = tn * block_n, c.wo0 = tw * a.TW, c.ho0 = th * a.TH, c.b0 = tb * a.NB;
  return c;
}

// Persistent kernel, one CTA per SM, tiles blockIdx.x, blockIdx.x + gridDim.x, ... (N tiles of the same
// pixels are neighbours, so co-running CTAs share their A tiles in L2).  Warp 0 = TMA producer, warp 1 =
// MMA issuer, warps 2-5 / 6-9 = two epilogue groups taking alternate tiles; the T accumulators are
// double buffered in TMEM so that the main loop of tile i+1 overlaps the epilogue of tile i.
template <int BLOCK_N, int TMAX, int BK, bool REUSE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_bn_plif_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap,
                    const ConvArgs a) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next layer may set itself up beside this one
  using L = SmemLayout<BLOCK_N, BK, REUSE>;
  constexpr int SA = L::SA, SB = L::SB, A_BYTES = L::A_BYTES, BLOCK_K = BK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem + L::OFF_A;
  uint8_t* sB = smem + L::OFF_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* fullA = bars;               // [SA]
  uint64_t* emptyA = bars + SA;         // [SA]
  uint64_t* fullB = bars + 2 * SA;      // [SB]
  uint64_t* emptyB = bars + 2 * SA + SB;
  uint64_t* accum_full = bars + 2 * SA + 2 * SB;    // [2]
  uint64_t* accum_empty = accum_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_empty + 2);
  float* sBiasAll = reinterpret_cast<float*>(smem + L::OFF_BIAS);

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops' ring / descriptor
  // arithmetic on the uniform datapath (no R2UR moves in front of every UTCHMMA)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int Tacc = a.Tx;  // accumulators: one per distinct input time step
  constexpr int KS = TMAX == 1 ? 2 : 1;             // K-split accumulators of the single-time-step layers
  constexpr uint32_t kBufCols = KS * TMAX * BLOCK_N;   // one accumulator set
  constexpr uint32_t kCols = 2 * kBufCols <= 32 ? 32 : 2 * kBufCols <= 64 ? 64 : 2 * kBufCols <= 128 ? 128
                             : 2 * kBufCols <= 256 ? 256 : 512;
  static_assert(2 * kBufCols <= 512, "double-buffered accumulators must fit TMEM");
  const int ncb = (a.Cin + BLOCK_K - 1) / BLOCK_K;
  const int taps = a.ksize * a.ksize;
  const int nkb = taps * ncb;
  const int n_tiles = a.tiles_w * a.tiles_h * a.tiles_b * a.tiles_n;
  const bool ks2 = TMAX == 1 && a.ksplit == 2;
  const int n_issuers = ks2 ? 2 : (Tacc < N_ISSUERS ? Tacc : N_ISSUERS);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    // K-split: a B stage belongs to one issuer; in tap-reuse mode both issuers read every A slot
    for (int i = 0; i < SA; ++i) mbar_init(fullA + i, 1), mbar_init(emptyA + i, (ks2 && REUSE) ? 2 : 1);
    for (int i = 0; i < SB; ++i) mbar_init(fullB + i, 1), mbar_init(emptyB + i, ks2 ? 1 : n_issuers);
    for (int i = 0; i < 2; ++i) mbar_init(accum_full + i, n_issuers), mbar_init(accum_empty + i, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // (also the first MMA issuer)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: this grid may have started while the layer before it was still running (its CTAs
  // take SMs that one leaves idle -- small maps, batch 1); everything above touched no global memory.  From here on
  // the previous grid has completed and its stores are visible.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    // The single producer thread is the scarce resource of the short-K layers (ncu: it never waits for a free slot
    // while the issuers wait for data), so it issues as few TMA instructions as possible: the weight planes hi / lo of
    // a K block are ONE 4-D box, and the activation tiles of all time steps / input planes of a K block are ONE 5-D
    // box landing in consecutive ring slots (slot order = plane-major, like the tensor).
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const int per_kb = Tacc * a.n_xsplit;
    const int SAR = a.sa_ring;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const TileCoord tc = tile_coord(a, tile, BLOCK_N);
      if constexpr (REUSE) {
        // channel block outer: the haloed tile of every time step once, then the nine weight taps
        const uint32_t box_bytes = (uint32_t)((a.TH + 2) * a.TWp * BLOCK_K * 2);
        for (int cb = 0; cb < ncb; ++cb) {
          for (int g = 0; g < per_kb; ++g) {       // slot g = plane g of the tensor (i * Tx + t)
            mbar_wait(emptyA + sa + g, pa ^ 1);
            mbar_expect_tx(fullA + sa + g, box_bytes);
            tma_load_5d(sA + (sa + g) * A_BYTES, &xmap, fullA + sa + g, cb * BLOCK_K, tc.wo0 - 1, tc.ho0 - 1, tc.b0, g);
          }
          sa += per_kb;
          if (sa == SAR) sa = 0, pa ^= 1;
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(emptyB + sb, pb ^ 1);
            mbar_expect_tx(fullB + sb, (uint32_t)(a.n_wsplit * L::B_BYTES));
            tma_load_4d(sB + sb * MAX_WSPLIT * L::B_BYTES, &wmap, fullB + sb, cb * BLOCK_K, tap, tc.n0, 0);
            if (++sb == SB) sb = 0, pb ^= 1;
          }
        }
        continue;
      }
      for (int kb = 0; kb < nkb; ++kb) {
        const int tap = kb / ncb, cb = kb - tap * ncb;
        const int ky = tap / a.ksize, kx = tap - ky * a.ksize;
        mbar_wait(emptyB + sb, pb ^ 1);
        mbar_expect_tx(fullB + sb, (uint32_t)(a.n_wsplit * L::B_BYTES));
        tma_load_4d(sB + sb * MAX_WSPLIT * L::B_BYTES, &wmap, fullB + sb, cb * BLOCK_K, tap, tc.n0, 0);
        if (++sb == SB) sb = 0, pb ^= 1;
        const int cx = tc.wo0 * a.stride + kx - a.pad, cy = tc.ho0 * a.stride + ky - a.pad;
        if (a.a_group) {
          for (int g = 0; g < per_kb; ++g) mbar_wait(emptyA + sa + g, pa ^ 1);
          mbar_expect_tx(fullA + sa, (uint32_t)(per_kb * A_BYTES));
          tma_load_5d(sA + sa * A_BYTES, &xmap, fullA + sa, cb * BLOCK_K, cx, cy, tc.b0, 0);
        } else {
          for (int g = 0; g < per_kb; ++g) {
            int st = sa + g;
            uint32_t pt = pa;
            while (st >= SAR) st -= SAR, pt ^= 1;
            mbar_wait(emptyA + st, pt ^ 1);
            mbar_expect_tx(fullA + st, (uint32_t)A_BYTES);
            tma_load_5d(sA + st * A_BYTES, &xmap, fullA + st, cb * BLOCK_K, cx, cy, tc.b0, g);
          }
        }
        sa += per_kb;
        while (sa >= SAR) sa -= SAR, pa ^= 1;
      }
    }
  } else if (warp >= 1 && warp <= N_ISSUERS) {
    // ===================== MMA issuers =====================
    // One issuing warp per time step (t = warp-1, + N_ISSUERS): the T accumulator chains are independent and issuing,
    // not the tensor pipe, is what bounds these small-N MMAs.  Each warp walks the
    // (warp-uniform) loop nest so that descriptors live in uniform registers; only its elected lane issues
    // tcgen05.mma / tcgen05.commit.
    const int wi = warp - 1;
    if (ks2) {
      // ---- single time step (ANN layers, broadcast first spiking conv): one chain would leave the issue rate of
      // ONE warp as the bound, so two warps take alternate B stages (K blocks / taps) into their own accumulators,
      // which the epilogue adds ----
      if constexpr (TMAX == 1) {
        if (wi < 2) {
          const bool leader = elect_one();
          constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
          const uint64_t a_desc0 = make_kmajor_desc<BK>(smem_u32(sA)), b_desc0 = make_kmajor_desc<BK>(smem_u32(sB));
          const int nx = a.n_xsplit, nw = a.n_wsplit;
          int sa = 0, sb = 0;
          uint32_t pa = 0, pb = 0;
          int it = 0;
          for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(accum_empty + buf, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)buf * kBufCols + (uint32_t)(wi * BLOCK_N);
            uint32_t accf = 0;
            int bcount = 0;
            const int n_outer = REUSE ? ncb : nkb;
            for (int kb = 0; kb < n_outer; ++kb) {
              const int n_taps = REUSE ? 9 : 1;
              const int first_tap = ((bcount & 1) == wi) ? 0 : 1;
              const int last_tap = (((bcount + n_taps - 1) & 1) == wi) ? n_taps - 1 : n_taps - 2;
              for (int tap = 0; tap < n_taps; ++tap, ++bcount) {
                if ((bcount & 1) == wi) {
                  mbar_wait(fullB + sb, pb);
                  const uint64_t bdesc = b_desc0 + (uint64_t)((sb * MAX_WSPLIT * L::B_BYTES) >> 4);
                  uint32_t shift = 0;
                  if constexpr (REUSE) {
                    const int ky = tap / 3, kx = tap - ky * 3;
                    shift = (uint32_t)(((ky * a.TWp + kx) * BLOCK_K * 2) >> 4);
                  }
#pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    if (i < nx) {
                      const int st = sa + i;                      // (a group never wraps: sa_ring % per_kb == 0)
                      if (!REUSE || tap == first_tap) {
                        mbar_wait(fullA + (a.a_group ? sa : st), pa);   // grouped load: one barrier per K block
                        tc_fence_after();
                      }
                      const uint64_t adesc = a_desc0 + (uint64_t)((st * A_BYTES) >> 4) + (uint64_t)shift;
                      if (leader) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                          if (j + i < nw) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k) {
                              tc_mma_f16(d0, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(((j * L::B_BYTES) >> 4) + k * 2),
                                         idesc, accf);
                              accf = 1u;
                            }
                          }
                        }
                        if (!REUSE || tap == last_tap) tc_commit(emptyA + st);
                      }
                    }
                  }
                  if (leader) tc_commit(emptyB + sb);
                }
                if (++sb == SB) sb = 0, pb ^= 1;
              }
              sa += nx;
              if (sa == a.sa_ring) sa = 0, pa ^= 1;
            }
            if (leader) tc_commit(accum_full + buf);
          }
        }
      }
    } else if (wi < n_issuers) {
      const bool leader = elect_one();
      // D = f32 (bit 4), A = B = f16 (format fields 0), K-major operands
      constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      const uint64_t a_desc0 = make_kmajor_desc<BK>(smem_u32(sA)), b_desc0 = make_kmajor_desc<BK>(smem_u32(sB));
      const int nx = a.n_xsplit, nw = a.n_wsplit;
      const int per_kb = Tacc * nx;           // A slots consumed per K block (per channel block in tap-reuse mode)
      int sa = 0, sb = 0;                     // ring positions at the start of the current K block
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(accum_empty + buf, ((it >> 1) & 1) ^ 1);   // the epilogue of tile it-2 has drained this buffer
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)buf * kBufCols;
        const int n_outer = REUSE ? ncb : nkb;
        for (int kb = 0; kb < n_outer; ++kb) {
          const int n_taps = REUSE ? 9 : 1;
          for (int tap = 0; tap < n_taps; ++tap) {
            mbar_wait(fullB + sb, pb);
            const uint64_t bdesc = b_desc0 + (uint64_t)((sb * MAX_WSPLIT * L::B_BYTES) >> 4);
            uint32_t shift = 0;
            if constexpr (REUSE) {
              const int ky = tap / 3, kx = tap - ky * 3;
              shift = (uint32_t)(((ky * a.TWp + kx) * BLOCK_K * 2) >> 4);   // rows -> descriptor units
            }
#pragma unroll
            for (int tt = 0; tt < (TMAX + N_ISSUERS - 1) / N_ISSUERS; ++tt) {
              const int t = wi + N_ISSUERS * tt;
              if (t < Tacc) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  if (i < nx) {
                    int st = sa + i * Tacc + t;                   // plane-major, like the tensor
                    uint32_t pt = pa;
                    while (st >= a.sa_ring) st -= a.sa_ring, pt ^= 1;   // (only the slot-by-slot fall-back wraps)
                    if (tap == 0) {
                      mbar_wait(fullA + (a.a_group ? sa : st), pt);   // grouped load: one barrier per K block
                      tc_fence_after();
                    }
                    const uint64_t adesc = a_desc0 + (uint64_t)((st * A_BYTES) >> 4) + (uint64_t)shift;
                    if (leader) {
                      // product terms a_i * w_j with i + j < n_wsplit (the dropped ones are below fp32 rounding)
#pragma unroll
                      for (int j = 0; j < 2; ++j) {
                        if (j + i < nw) {
#pragma unroll
                          for (int k = 0; k < BLOCK_K / 16; ++k)
                            tc_mma_f16(d0 + (uint32_t)(t * BLOCK_N), adesc + (uint64_t)(k * 2),
                                       bdesc + (uint64_t)(((j * L::B_BYTES) >> 4) + k * 2), idesc,
                                       (kb > 0 || tap > 0 || i > 0 || j > 0 || k > 0) ? 1u : 0u);
                        }
                      }
                      if (tap == n_taps - 1) tc_commit(emptyA + st);  // frees the A slot once its MMAs have read it
                    }
                  }
                }
              }
            }
            if (leader) tc_commit(emptyB + sb);
            if (++sb == SB) sb = 0, pb ^= 1;
          }
          sa += per_kb;
          while (sa >= a.sa_ring) sa -= a.sa_ring, pa ^= 1;
        }
        if (leader) tc_commit(accum_full + buf);
      }
    }
  } else if (warp > N_ISSUERS) {
    // ===================== epilogue: bias + LIF over t + store (two groups, alternate tiles) =====================
    const int ew = warp - 1 - N_ISSUERS;     // epilogue warp 0..7
    const int grp = ew >> 2;                 // 0: even tiles of this CTA, 1: odd tiles
    const int lg = warp & 3;                 // TMEM lane group this warp may read (any 4 consecutive warps cover all)
    const int m = lg * 32 + lane;            // pixel of the tile
    const int gtid = (ew & 3) * 32 + lane;
    float* sBias = sBiasAll + grp * 2 * BLOCK_N;
    float* sUnscale = sBias + BLOCK_N;
    // row m of the tile -> (image, row, column); in tap-reuse mode rows run over the haloed width TWp and
    // the last two columns of every row (and the rows past TH) are garbage positions
    int nb, ph, pw;
    bool row_ok = true;
    if constexpr (REUSE) {
      nb = 0, ph = m / a.TWp, pw = m - ph * a.TWp;
      row_ok = ph < a.TH && pw < a.TW;
    } else {
      nb = m / (a.TH * a.TW);
      const int rem = m - nb * (a.TH * a.TW);
      ph = rem / a.TW, pw = rem - ph * a.TW;
    }
    const LifDyn d = make_lif(a.plif_w ? *a.plif_w : 0.0f, a.vth, a.hard_reset, a.vreset, a.decay_input);
    // the neuron the reference builds (soft reset, decay_input=False, utils_snn.py:44-53): 5 instructions per step
    const bool fast_lif = a.out_mode == EAS_CONV_OUT_SPIKES && !a.hard_reset && !a.decay_input;
    const int64_t step = (int64_t)a.B * a.Ho * a.Wo;                   // pixels per time step
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const TileCoord tc = tile_coord(a, tile, BLOCK_N);
      const int n0 = tc.n0;
      const int b = tc.b0 + nb, ho = tc.ho0 + ph, wo = tc.wo0 + pw;
      const bool valid = row_ok && b < a.B && ho < a.Ho && wo < a.Wo;
      const int64_t pix = ((int64_t)b * a.Ho + ho) * a.Wo + wo;          // within one time step
      // the group's previous tile is done with sBias (every thread passed this barrier after its last read)
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      for (int i = gtid; i < BLOCK_N; i += 128) {
        const float b_ = (n0 + i < a.Cout) ? a.bias[n0 + i] : 0.0f;
        const float u_ = (a.unscale && n0 + i < a.Cout) ? a.unscale[n0 + i] : 1.0f;
        // fast neuron path: work in the accumulator's scale.  u_ is a power of two, so dividing the bias and the
        // threshold by it commutes with every rounding below: bit-identical potentials / spikes, one multiply less
        sBias[i] = fast_lif ? __fdiv_rn(b_, u_) : b_;
        sUnscale[i] = fast_lif ? __fdiv_rn(a.vth, u_) : u_;
      }
      const bool res_vec = a.residual != nullptr && valid && a.out_mode == EAS_CONV_OUT_SPIKES &&
                           (a.res_ld & 7) == 0 && (n0 & 7) == 0;
      constexpr bool kPrefRes = TMAX == 3 || TMAX == 4;   // the shortcut of every time step is prefetched (T <= 4 layers)
      constexpr int RT = kPrefRes ? TMAX : 1;
      uint4 rres[RT][2];
      auto load_res = [&](int c16, uint4 (&r)[RT][2]) {
        const int ch0 = n0 + c16 * 16;
#pragma unroll
        for (int t = 0; t < RT; ++t) {
          if (kPrefRes && t < a.T && res_vec && ch0 + 16 <= a.Cout) {
            const uint4* rp = reinterpret_cast<const uint4*>(a.residual + ((int64_t)t * step + pix) * a.res_ld + ch0);
            r[t][0] = ld_cg_u4(rp), r[t][1] = ld_cg_u4(rp + 1);
          }
        }
      };
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // sBias visible to the group
      mbar_wait(accum_full + grp, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)grp * kBufCols;
      // live 16-channel chunks of this N tile (warp-uniform)
      const int NCH = min(BLOCK_N / 16, (a.Cout - n0 + 15) / 16);
#pragma unroll 1
      for (int c16 = 0; c16 < NCH; ++c16) {
        const int ch0 = n0 + c16 * 16;
        load_res(c16, rres);    // SEW shortcut of this chunk: in flight during the TMEM reads and the first LIF steps
        uint32_t acc[TMAX][16];
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
          if (t < Tacc) tmem_ld16(tbase + (uint32_t)(t * BLOCK_N + c16 * 16), acc[t]);
        if constexpr (TMAX == 1) {
          if (ks2) {   // K-split: the two partial sums of the single time step
            uint32_t acc2[16];
            tmem_ld16(tbase + (uint32_t)(BLOCK_N + c16 * 16), acc2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              acc[0][j] = __float_as_uint(__fadd_rn(__uint_as_float(acc[0][j]), __uint_as_float(acc2[j])));
          }
        }
        tmem_ld_wait();
        if (c16 == NCH - 1) {   // last TMEM read of this tile: hand the accumulator buffer back (once per warp)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accum_empty + grp);
        }
        if (valid) {
          const int nch = min(16, a.Cout - ch0);
          if (a.out_mode == EAS_CONV_OUT_SPIKES) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = lif_v_init(d);
            __half* outp = reinterpret_cast<__half*>(a.out);
            // unrolled over t so that acc[t][j] stays in registers (a runtime t would spill the tile to local memory)
#pragma unroll
            for (int t = 0; t < (TMAX == 1 ? 8 : TMAX); ++t) {
              if (t >= a.T) break;
              float sp[16];
              if (fast_lif) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float xin = __uint_as_float(TMAX == 1 ? acc[0][j] : (Tacc == 1 ? acc[0][j] : acc[t < TMAX ? t : 0][j]));
                  const float vth_s = sUnscale[c16 * 16 + j];                       // vth / unscale
                  const float h = __fadd_rn(__fmul_rn(v[j], d.k), __fadd_rn(xin, sBias[c16 * 16 + j]));
                  sp[j] = h >= vth_s ? 1.0f : 0.0f;
                  v[j] = __fmaf_rn(-sp[j], vth_s, h);                               // h - s * vth, one rounding
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float xin = __uint_as_float(TMAX == 1 ? acc[0][j] : (Tacc == 1 ? acc[0][j] : acc[t < TMAX ? t : 0][j]));
                  const float h = lif_charge(d, v[j], __fadd_rn(__fmul_rn(xin, sUnscale[c16 * 16 + j]), sBias[c16 * 16 + j]));
                  sp[j] = lif_fire(d, h);
                  v[j] = lif_reset(d, h, sp[j]);
                }
              }
              __half* dst = outp + ((int64_t)t * step + pix) * a.out_ld + ch0;
              if (a.residual) {  // SEW add: y = spikes + x (small integers, exact in fp16)
                if (kPrefRes && res_vec && nch == 16) {
                  const __half* rb0 = reinterpret_cast<const __half*>(&rres[t < RT ? t : 0][0]);
                  const __half* rb1 = reinterpret_cast<const __half*>(&rres[t < RT ? t : 0][1]);
#pragma unroll
                  for (int j = 0; j < 8; ++j) sp[j] += __half2float(rb0[j]), sp[8 + j] += __half2float(rb1[j]);
                } else {
                  const __half* rp = a.residual + ((int64_t)t * step + pix) * a.res_ld + ch0;
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (j < nch) sp[j] += __half2float(rp[j]);
                }
              }
              if (nch == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                uint4 q0 = make_uint4(pack_f16(sp[0], sp[1]), pack_f16(sp[2], sp[3]), pack_f16(sp[4], sp[5]),
                                      pack_f16(sp[6], sp[7]));
                uint4 q1 = make_uint4(pack_f16(sp[8], sp[9]), pack_f16(sp[10], sp[11]), pack_f16(sp[12], sp[13]),
                                      pack_f16(sp[14], sp[15]));
                reinterpret_cast<uint4*>(dst)[0] = q0;
                reinterpret_cast<uint4*>(dst)[1] = q1;
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (j < nch) dst[j] = __float2half_rn(sp[j]);
              }
            }
          } else {
#pragma unroll
            for (int t = 0; t < TMAX; ++t) {
              if (t < Tacc) {
                if (a.out_mode == EAS_CONV_OUT_PREACT) {
                  float* dst = reinterpret_cast<float*>(a.out) + ((int64_t)t * step + pix) * a.out_ld + ch0;
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (j < nch)
                      dst[j] = __fadd_rn(__fmul_rn(__uint_as_float(acc[t][j]), sUnscale[c16 * 16 + j]), sBias[c16 * 16 + j]);
                } else {  // SiLU, written as two fp16 planes hi + lo whose sum is the fp32 value to 2^-22
                  __half* outp = reinterpret_cast<__half*>(a.out);
                  const int64_t plane = (int64_t)Tacc * step * a.out_ld;
                  __half* dst = outp + ((int64_t)t * step + pix) * a.out_ld + ch0;
                  float y[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const float xv = __fadd_rn(__fmul_rn(__uint_as_float(acc[t][j]), sUnscale[c16 * 16 + j]), sBias[c16 * 16 + j]);
                    // SiLU = x / (1 + e^-x) on the SFU (ex2.approx + rcp.approx: ~2 ulp, far inside the 22-bit planes)
                    y[j] = fminf(fmaxf(__fdividef(xv, 1.0f + __expf(-xv)), -65504.0f), 65504.0f);
                  }
                  if (nch == 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && ((plane & 7) == 0)) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const __half2 h2 = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
                      const float2 hf = __half22float2(h2);
                      const __half2 l2 = __floats2half2_rn(y[2 * j] - hf.x, y[2 * j + 1] - hf.y);
                      hi[j] = *reinterpret_cast<const uint32_t*>(&h2), lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                    reinterpret_cast<uint4*>(dst)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    reinterpret_cast<uint4*>(dst)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    reinterpret_cast<uint4*>(dst + plane)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    reinterpret_cast<uint4*>(dst + plane)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      if (j < nch) {
                        const __half hi = __float2half_rn(y[j]);
                        dst[j] = hi, dst[plane + j] = __float2half_rn(y[j] - __half2float(hi));
                      }
                    }
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = []() -> PFN_cuTensorMapEncodeTiled_v12000 {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

int check_conv(const eas_conv_cfg* c) {
  EAS_REQUIRE(c, EAS_E_NULL);
  EAS_REQUIRE(c->T >= 1 && c->T <= 8 && (c->Tx == c->T || c->Tx == 1), EAS_E_SHAPE);
  EAS_REQUIRE(c->B >= 1 && c->H >= 1 && c->W >= 1 && c->Cin >= 8 && c->Cout >= 1, EAS_E_SHAPE);
  EAS_REQUIRE(c->Cin % 8 == 0, EAS_E_SHAPE);  // TMA global strides are multiples of 16 B
  EAS_REQUIRE((c->ksize == 1 || c->ksize == 3) && (c->stride == 1 || c->stride == 2), EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->n_wsplit >= 1 && c->n_wsplit <= 2 && c->n_xsplit >= 1 && c->n_xsplit <= 2, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->out_mode >= EAS_CONV_OUT_SPIKES && c->out_mode <= EAS_CONV_OUT_SILU2, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(c->x_ld == 0 || (c->x_ld >= c->Cin && c->x_ld % 8 == 0), EAS_E_SHAPE);
  EAS_REQUIRE(c->out_ld == 0 || c->out_ld >= c->Cout, EAS_E_SHAPE);
  EAS_REQUIRE(c->res_ld == 0 || c->res_ld >= c->Cout, EAS_E_SHAPE);
  EAS_REQUIRE(!c->residual || c->out_mode == EAS_CONV_OUT_SPIKES, EAS_E_UNSUPPORTED);
  return EAS_OK;
}

// M-tile decomposition: NB x TH x TW = 128 (powers of two) covering B x Ho x Wo with the least padding.
void pick_tile(int B, int Ho, int Wo, int stride, int* NB, int* TH, int* TW) {
  int64_t best = -1;
  for (int tw = 1; tw <= 128; tw *= 2) {
    if (tw * stride > 256) break;
    for (int th = 1; th * tw <= 128; th *= 2) {
      if (th * stride > 256) break;
      const int nb = 128 / (tw * th);
      const int64_t cost = eas_ceil_div(Wo, tw) * eas_ceil_div(Ho, th) * eas_ceil_div(B, nb);
      // prefer wide rows on ties (longer contiguous TMA rows)
      if (best < 0 || cost < best || (cost == best && tw > *TW)) best = cost, *NB = nb, *TH = th, *TW = tw;
    }
  }
}

template <int BLOCK_N, int TMAX, int BK, bool REUSE = false>
int launch_conv(const CUtensorMap& xmap, const CUtensorMap& wmap, const ConvArgs& a, int64_t grid, cudaStream_t st) {
  using L = SmemLayout<BLOCK_N, BK, REUSE>;
  auto kern = conv_bn_plif_kernel<BLOCK_N, TMAX, BK, REUSE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(NUM_THREADS), cfg.dynamicSmemBytes = L::TOTAL, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, xmap, wmap, a);
  if (e != cudaSuccess) return (int)e;
  return EAS_OK;
}

template <int BLOCK_N, int BK>
int launch_conv_t(int Tacc, const CUtensorMap& xmap, const CUtensorMap& wmap, const ConvArgs& a, int64_t grid,
                  cudaStream_t st) {
  // two accumulator sets of Tacc x BLOCK_N columns must fit the 512 TMEM columns
  if (Tacc <= 1) return launch_conv<BLOCK_N, 1, BK>(xmap, wmap, a, grid, st);
  // T = 3 is the reference configuration: its own instantiation keeps 16 accumulator + 8 shortcut registers less live
  // than the T <= 4 one (the epilogue sits at the 168-register cap of a 384-thread CTA, and spills cost more than MMAs)
  if (Tacc <= 3) return launch_conv<BLOCK_N, 3, BK>(xmap, wmap, a, grid, st);
  if (Tacc <= 4) return launch_conv<BLOCK_N, 4, BK>(xmap, wmap, a, grid, st);
  if constexpr (BLOCK_N <= 32) return launch_conv<BLOCK_N, 8, BK>(xmap, wmap, a, grid, st);
  return EAS_E_UNSUPPORTED;
}

template <int BLOCK_N>
int launch_conv_reuse(int Tacc, const CUtensorMap& xmap, const CUtensorMap& wmap, const ConvArgs& a, int64_t grid,
                      cudaStream_t st) {
  if (Tacc <= 1) return launch_conv<BLOCK_N, 1, 32, true>(xmap, wmap, a, grid, st);
  if (Tacc <= 3) return launch_conv<BLOCK_N, 3, 32, true>(xmap, wmap, a, grid, st);
  if (Tacc <= 4) return launch_conv<BLOCK_N, 4, 32, true>(xmap, wmap, a, grid, st);
  if constexpr (BLOCK_N <= 32) return launch_conv<BLOCK_N, 8, 32, true>(xmap, wmap, a, grid, st);
  return EAS_E_UNSUPPORTED;
}

// Tap-reuse tile for a 3x3 stride-1 layer: TH x (TW + 2) <= 128 rows with the best useful fraction.
double pick_reuse_tile(int Ho, int Wo, int* TH, int* TW) {
  double best = 0.0;
  for (int tw = 4; tw <= Wo && tw <= 62; ++tw) {
    const int th_max = 128 / (tw + 2);
    for (int th = 1; th <= th_max && th <= Ho; ++th) {
      const double eff = (double)Ho * Wo / ((double)eas_ceil_div(Ho, th) * eas_ceil_div(Wo, tw) * 128.0);
      if (eff > best) best = eff, *TH = th, *TW = tw;
    }
  }
  return best;
}

// K block = one swizzle row of 64 / 32 / 16 channels: least padded K plus a per-block overhead.
int pick_bk(int Cin) {
  int best = 64, best_cost = 1 << 30;
  for (int bk : {64, 32, 16}) {
    const int nb = (Cin + bk - 1) / bk;
    const int cost = nb * bk + 8 * nb;
    if (cost < best_cost) best_cost = cost, best = bk;
  }
  return best;
}

}  // namespace

extern "C" size_t eas_conv_bn_plif_ws_bytes(const eas_conv_cfg*) { return 0; }

extern "C" int eas_conv_bn_plif_fwd(const eas_conv_cfg* c, const void* x, const void* w_planes, const float* bias,
                                    const float* plif_w, void* out, void* /*ws*/, size_t /*ws_bytes*/, void* stream) {
  int rc = check_conv(c);
  if (rc) return rc;
  EAS_REQUIRE(x && w_planes && bias && out, EAS_E_NULL);
  EAS_REQUIRE(c->out_mode != EAS_CONV_OUT_SPIKES || plif_w, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)w_planes % 16 == 0, EAS_E_ALIGN);
  auto encode = get_encode_fn();
  EAS_REQUIRE(encode != nullptr, EAS_E_UNSUPPORTED);

  const int pad = (c->ksize - 1) / 2;
  const int Ho = (c->H + 2 * pad - c->ksize) / c->stride + 1;
  const int Wo = (c->W + 2 * pad - c->ksize) / c->stride + 1;
  const int x_ld = c->x_ld ? c->x_ld : c->Cin;
  const int out_ld = c->out_ld ? c->out_ld : c->Cout;
  ConvArgs a{};
  a.T = c->T, a.Tx = c->Tx, a.B = c->B, a.Ho = Ho, a.Wo = Wo, a.Cin = c->Cin, a.Cout = c->Cout;
  a.ksize = c->ksize, a.stride = c->stride, a.pad = pad, a.n_wsplit = c->n_wsplit, a.n_xsplit = c->n_xsplit;
  pick_tile(c->B, Ho, Wo, c->stride, &a.NB, &a.TH, &a.TW);
  // 3x3 stride-1 layers on spike inputs: load every haloed input tile once and realise the nine taps as
  // descriptor shifts (9x less activation traffic from L2), when the padded tiles waste < 30 % of the MMA rows
  bool reuse = false;
  // (every plane of every time step of a channel block must sit in the A ring at once: Tx * n_xsplit <= 6)
  if (c->ksize == 3 && c->stride == 1 && (c->n_xsplit == 1 || c->Tx == 1) && c->Cin >= 32) {
    int th = 0, tw = 0;
    if (pick_reuse_tile(Ho, Wo, &th, &tw) >= 0.70) reuse = true, a.NB = 1, a.TH = th, a.TW = tw, a.TWp = tw + 2;
  }
  const int BK = reuse ? 32 : pick_bk(c->Cin);
  const int taps_ = c->ksize * c->ksize;
  const int nkb_ = taps_ * (int)eas_ceil_div(c->Cin, BK);
  // 64 output channels per tile; 32 for thin layers and for more than 4 distinct time steps (TMEM)
  // a single accumulator (Tx == 1: the ANN layers and the broadcast first spiking conv) leaves TMEM room for 128
  // (an MMA of N <= 128 costs the same 64 cycles, so half as many of them); kept to grids of >= 1 wave
  int BLOCK_N = (c->Cout <= 32 || c->Tx > 4) ? 32 : 64;
  if (c->Tx == 1 && c->Cout > 64) {
    const int64_t t128 = eas_ceil_div(Wo, a.TW) * eas_ceil_div(Ho, a.TH) * eas_ceil_div(c->B, a.NB) *
                         eas_ceil_div(c->Cout, 128);
    if (t128 >= EAS_NUM_SMS) BLOCK_N = 128;
  }
  a.tiles_w = (int)eas_ceil_div(Wo, a.TW), a.tiles_h = (int)eas_ceil_div(Ho, a.TH);
  a.tiles_b = (int)eas_ceil_div(c->B, a.NB), a.tiles_n = (int)eas_ceil_div(c->Cout, BLOCK_N);
  a.out_ld = out_ld, a.out_mode = c->out_mode;
  a.residual = (const __half*)c->residual, a.res_ld = c->res_ld ? c->res_ld : c->Cout;
  a.vth = c->v_threshold, a.vreset = c->v_reset, a.hard_reset = c->hard_reset, a.decay_input = c->decay_input;
  a.bias = bias, a.unscale = c->w_unscale, a.plif_w = plif_w, a.out = out;
  {
    const int per_kb = c->Tx * c->n_xsplit;
    const int SA_T = BLOCK_N > 64 ? 6 : (reuse ? 8 : 9);              // SmemLayout::SA of the kernel picked below
    EAS_REQUIRE(!reuse || per_kb <= SA_T, EAS_E_UNSUPPORTED);         // tap reuse keeps every plane of a channel block
    a.a_group = (!reuse && per_kb <= SA_T) ? 1 : 0;      // (more planes than slots: slot-by-slot loads)
    a.sa_ring = per_kb <= SA_T ? SA_T / per_kb * per_kb : SA_T;       // whole groups, so that a group never wraps
  }
  a.ksplit = (c->Tx == 1 && (reuse ? 9 * (int)eas_ceil_div(c->Cin, BK) : nkb_) >= 2) ? 2 : 1;
  const int64_t n_tiles = (int64_t)a.tiles_w * a.tiles_h * a.tiles_b * a.tiles_n;
  EAS_REQUIRE(n_tiles > 0 && n_tiles < (1ll << 31), EAS_E_SHAPE);
  const int64_t grid = n_tiles < EAS_NUM_SMS ? n_tiles : EAS_NUM_SMS;   // persistent: one CTA per SM

  // activations: [plane*Tx][B][H][W][x_ld] fp16, innermost first for the tensor map
  CUtensorMap xmap, wmap;
  const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  {
    cuuint64_t dims[5] = {(cuuint64_t)c->Cin, (cuuint64_t)c->W, (cuuint64_t)c->H, (cuuint64_t)c->B,
                          (cuuint64_t)(c->n_xsplit * c->Tx)};
    cuuint64_t strides[4] = {(cuuint64_t)x_ld * 2, (cuuint64_t)c->W * x_ld * 2, (cuuint64_t)c->H * c->W * x_ld * 2,
                             (cuuint64_t)c->B * c->H * c->W * x_ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)(a.TW * c->stride), (cuuint32_t)(a.TH * c->stride),
                         (cuuint32_t)a.NB, (cuuint32_t)(a.a_group ? c->Tx * c->n_xsplit : 1)};
    if (reuse) box[1] = (cuuint32_t)a.TWp, box[2] = (cuuint32_t)(a.TH + 2);
    cuuint32_t estr[5] = {1, (cuuint32_t)c->stride, (cuuint32_t)c->stride, 1, 1};
    CUresult r = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return EAS_E_SHAPE;
  }
  {
    // weights [n_wsplit][Cout][taps][Cin]: the hi / lo planes of an N tile are one box {BK, 1, BLOCK_N, n_wsplit}
    const int taps = c->ksize * c->ksize;
    cuuint64_t dims[4] = {(cuuint64_t)c->Cin, (cuuint64_t)taps, (cuuint64_t)c->Cout, (cuuint64_t)c->n_wsplit};
    cuuint64_t strides[3] = {(cuuint64_t)c->Cin * 2, (cuuint64_t)taps * c->Cin * 2,
                             (cuuint64_t)c->Cout * taps * c->Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, 1, (cuuint32_t)BLOCK_N, (cuuint32_t)c->n_wsplit};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(w_planes), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return EAS_E_SHAPE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Tacc = c->Tx;
#define EAS_CONV_BK(BN_)                                                              \
  (BK == 64 ? launch_conv_t<BN_, 64>(Tacc, xmap, wmap, a, grid, st)                   \
   : BK == 32 ? launch_conv_t<BN_, 32>(Tacc, xmap, wmap, a, grid, st)                 \
              : launch_conv_t<BN_, 16>(Tacc, xmap, wmap, a, grid, st))
  if (BLOCK_N == 128) {   // Tacc == 1 only
    if (reuse) return launch_conv<128, 1, 32, true>(xmap, wmap, a, grid, st);
    return BK == 64 ? launch_conv<128, 1, 64>(xmap, wmap, a, grid, st)
           : BK == 32 ? launch_conv<128, 1, 32>(xmap, wmap, a, grid, st)
                      : launch_conv<128, 1, 16>(xmap, wmap, a, grid, st);
  }
  if (reuse) return BLOCK_N == 32 ? launch_conv_reuse<32>(Tacc, xmap, wmap, a, grid, st)
                                  : launch_conv_reuse<64>(Tacc, xmap, wmap, a, grid, st);
  if (BLOCK_N == 32) return EAS_CONV_BK(32);
  return EAS_CONV_BK(64);
#undef EAS_CONV_BK
}
