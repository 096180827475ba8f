// (a-1) Event binning: time-sorted (x, y, t, p) windows -> per-pixel micro-bin histograms.
// Replaces GEN1Dataset.slice_events + agrregate('micro_sum') (yolox/data/datasets/gen1.py:313-360).
//
// Data layout in HBM: events are SoA (x:i16, y:i16, t:i64, p:u8), B windows back to back with
// offsets[B+1]; output hist[B][Tm][2][H][W] int32.
//
// Two kernels run per call:
//   1. bin_bounds_kernel: per window, tw = (t_last - t_first)/Tm and Tm+1 lower_bound searches on
//      the sorted timestamps (one warp per search, 32-ary).  After this nobody reads t again:
//      membership of an event in a micro-bin is an index-range test, exactly like the reference's
//      searchsorted slicing.
//   2. one of
//      bin_hist_smem_kernel  ("tiles"): a CTA takes (window, micro-bin, polarity, row slab) work items
//         from an atomic counter, counts a slab in shared memory as packed 16-bit lanes (<= 72 KB, so
//         3 CTAs per SM overlap their zero / scan / write phases) and writes it once with 16 B
//         stores: no global atomics, no pre-zeroing, HBM traffic = 5 B/event (re-read from L2 by
//         the other slabs) + 4 B/bin (1 B/bin for the compact byte histogram, hist_u8.cuh).  Chunks of
//         <= 65535 events make 16-bit overflow impossible.  The kernel is issue bound (~20 SASS
//         instructions per scanned event), so the variant that big batches get holds BOTH polarities of a
//         slab per item (2 x 73 KB, one 1024-thread CTA per SM): half as many scans of every event.
//      bin_hist_global_kernel ("reds"): event-parallel, 8 events per thread with 16 B loads,
//         red.global.add.u32 into the (L2-resident when it fits) histogram after a memset.  Used for
//         frames that do not fit in shared memory and for very long windows.
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "hist_u8.cuh"

namespace {

// Process-wide: where compact binning calls report lost counts (eas_hist_u8_set_sticky); may be pinned host memory.
int32_t* g_sticky = nullptr;

// Where the events come from.  SoaSrc: the reference's events_struct de-interleaved (x, y, t, p) with
// windows back to back (offsets[B+1]).  DatSrc: raw 8-byte PSEE .dat Event2D records
// {u32 t; u32 x:14 | y:14 << 14 | p << 28} (dat_events_tools.py:24, 46-51) with one [first, last) record
// range per window (eas_dat_windows).
struct SoaSrc {
  const int16_t* __restrict__ x;
  const int16_t* __restrict__ y;
  const int64_t* __restrict__ t;
  const uint8_t* __restrict__ p;
  const int64_t* __restrict__ offsets;
  int64_t n;       // events in the arrays (vector loads never run past it)
  // (8-event vector loads: the ABI requires 16 B aligned x / y and 8 B aligned p)
  static constexpr int VW = 8;
  // events [i8, i8 + 8), i8 % 8 == 0 and i8 + 8 <= n: one 16 B load of x and of y, one 8 B load of p
  __device__ __forceinline__ void loadv(int64_t i8, int (&xs)[8], int (&ys)[8], int (&cs)[8]) const {
    const uint4 xv = ld_stream_u4(reinterpret_cast<const uint4*>(x + i8));
    const uint4 yv = ld_stream_u4(reinterpret_cast<const uint4*>(y + i8));
    const uint2 pv = ld_stream_u2(reinterpret_cast<const uint2*>(p + i8));
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w}, pw[2] = {pv.x, pv.y};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xs[j] = (int)(int16_t)(xw[j >> 1] >> ((j & 1) * 16));
      ys[j] = (int)(int16_t)(yw[j >> 1] >> ((j & 1) * 16));
      cs[j] = ((pw[j >> 2] >> ((j & 3) * 8)) & 0xffu) != 0u;
    }
  }
  __device__ __forceinline__ int64_t begin(int64_t b) const { return offsets[b]; }
  __device__ __forceinline__ int64_t end(int64_t b) const { return offsets[b + 1]; }
  __device__ __forceinline__ int64_t time(int64_t i) const { return t[i]; }
  __device__ __forceinline__ void xyc(int64_t i, int& xi, int& yi, int& ci) const {
    xi = x[i], yi = y[i], ci = p[i] != 0;
  }
};
struct DatSrc {
  const uint2* __restrict__ rec;
  const int64_t* __restrict__ ranges;
  int64_t n;       // records in the array
  // (records are read one 4-byte (x, y, p) word at a time: 16-byte loads of record pairs also drag the timestamps
  // through L2 and measured slower, 0.111 vs 0.085 ms per batch)
  static constexpr int VW = 1;
  __device__ __forceinline__ int64_t begin(int64_t b) const { return ranges[2 * b]; }
  __device__ __forceinline__ int64_t end(int64_t b) const { return ranges[2 * b + 1]; }
  __device__ __forceinline__ int64_t time(int64_t i) const { return (int64_t)rec[i].x; }
  __device__ __forceinline__ void xyc(int64_t i, int& xi, int& yi, int& ci) const {
    const uint32_t w = rec[i].y;
    xi = (int)(w & 16383u), yi = (int)((w >> 14) & 16383u), ci = (int)((w >> 28) & 1u);
  }
};

constexpr int kSmemThreads = 512;
constexpr int kSlabMaxBytes = 72 * 1024;   // 3 CTAs/SM: phases of different CTAs overlap
constexpr int kChunk = 65535;

// One warp per (window, boundary): 32-ary search on the sorted timestamps (4 dependent loads for
// 1e5 events instead of 17).  Also resets the work counter of the tile kernel.
template <typename SRC>
__global__ void __launch_bounds__(128)
bin_bounds_kernel(const SRC src, int64_t B, int Tm, int64_t* __restrict__ bounds,
                  unsigned int* __restrict__ work_counter, uint32_t* __restrict__ sat_tail, int32_t* sticky) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the histogram kernel may be scheduled behind this one
  const int lane = threadIdx.x & 31;
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gid == 0 && lane == 0) *work_counter = 0u;
  if (gid == 0 && lane < 2 && sat_tail) sat_tail[lane] = 0u;   // compact output: empty saturation list
  if (gid == 0 && lane == 2 && sat_tail) *reinterpret_cast<int32_t**>(sat_tail + 2) = sticky;
  if (gid >= B * (Tm + 1)) return;
  const int64_t b = gid / (Tm + 1);
  const int k = (int)(gid - b * (Tm + 1));
  const int64_t s = src.begin(b), e = src.end(b);
  if (e <= s) {
    if (lane == 0) bounds[gid] = s;
    return;
  }
  const int64_t t0 = src.time(s);
  const int64_t tw = (src.time(e - 1) - t0) / Tm;  // sorted => non-negative => trunc == floor
  const int64_t key = t0 + (int64_t)k * tw;
  int64_t lo = s, hi = e;  // answer (first index with t[i] >= key) is in [lo, hi]
  while (hi - lo > 0) {
    const int64_t len = hi - lo;
    const int64_t step = (len + 31) / 32;  // probes at lo + (lane+1)*step - 1
    const int64_t pi = lo + (int64_t)(lane + 1) * step - 1;
    const bool less = pi < hi ? (src.time(pi) < key) : false;  // out-of-range probes count as ">= key"
    const unsigned m = __ballot_sync(0xffffffffu, less);
    const int nless = __popc(m);  // sorted => the lanes with t < key are a prefix
    const int64_t nlo = lo + (int64_t)nless * step;
    int64_t nhi = lo + (int64_t)(nless + 1) * step - 1;
    if (nhi > hi) nhi = hi;
    lo = nlo > hi ? hi : nlo;
    hi = nhi;
  }
  if (lane == 0) bounds[gid] = lo;
}

// ---- strategy "tiles" ---------------------------------------------------------------------
// work item = (window b, micro-bin k, polarity c, row slab): counted in shared memory as packed
// 16-bit lanes, written once.  Items are handed out through an atomic counter.
template <typename OUT_T> struct Cvt;
template <> struct Cvt<int32_t> {
  static __device__ __forceinline__ uint32_t enc(uint32_t c) { return c; }
  static __device__ __forceinline__ uint32_t add(uint32_t old, uint32_t c) { return old + c; }
};
template <> struct Cvt<float> {  // counts < 2^24 are exact in fp32
  static __device__ __forceinline__ uint32_t enc(uint32_t c) { return __float_as_uint((float)c); }
  static __device__ __forceinline__ uint32_t add(uint32_t old, uint32_t c) {
    return __float_as_uint(__uint_as_float(old) + (float)c);
  }
};

// BOTH: one item = (window, micro-bin, slab) with the counters of BOTH polarities in shared memory (2 x the slab, one
// 1024-thread CTA per SM): every event is scanned by n_slabs items instead of 2 * n_slabs, and the polarity test becomes
// a plane offset.
template <typename OUT_T, typename SRC, int NT, bool BOTH>
__global__ void __launch_bounds__(NT, BOTH ? 1 : 3)
bin_hist_smem_kernel(const SRC src, const int64_t* __restrict__ bounds, int64_t n_items,
                     int H, int W, int Tm, int n_slabs, int slab_rows, uint32_t* __restrict__ hist,
                     unsigned int* __restrict__ work_counter, uint32_t* __restrict__ sat_tail) {
  constexpr bool kU8 = std::is_same<OUT_T, uint8_t>::value;
  // programmatic dependent launch: this grid may have been scheduled before the bounds kernel finished -- wait for it
  // before reading anything
  asm volatile("griddepcontrol.wait;" ::: "memory");
  extern __shared__ __align__(16) uint32_t cnt[];  // ceil(slab_rows*W/2) words, two 16-bit counters per word
  __shared__ unsigned int sh_item;
  const int64_t HW = (int64_t)H * W;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sh_item = atomicAdd(work_counter, 1u);
    __syncthreads();
    const int64_t item = sh_item;
    if (item >= n_items) break;
    const int slab = (int)(item % n_slabs);
    const int64_t bkc0 = item / n_slabs;  // (b*Tm + k)*2 + c, or b*Tm + k with both polarities in one item
    const int c = BOTH ? 0 : (int)(bkc0 & 1);
    const int64_t bk = BOTH ? bkc0 : (bkc0 >> 1);
    const int64_t b = bk / Tm;
    const int k = (int)(bk - b * Tm);
    // (coherent loads: this grid may have been scheduled before the bounds kernel finished, see ld_cg_* in common.cuh)
    const int64_t s = __ldcg(bounds + b * (Tm + 1) + k), e = __ldcg(bounds + b * (Tm + 1) + k + 1);
    const int y_lo = slab * slab_rows;
    const int rows = min(slab_rows, H - y_lo);
    const int npix = rows * W;
    const int nwords = (npix + 1) >> 1;
    const int nwords4 = (nwords + 3) & ~3;
    constexpr int NP = BOTH ? 2 : 1;     // polarity planes held by this item
    const bool vec_ok = (npix & 3) == 0 && ((((int64_t)y_lo * W) & 3) == 0) && ((HW & 3) == 0);
    bool first = true;
    for (int64_t cs = s; first || cs < e; cs += kChunk) {
      for (int w = threadIdx.x * 4; w < NP * nwords4; w += NT * 4)
        *reinterpret_cast<uint4*>(cnt + w) = make_uint4(0u, 0u, 0u, 0u);
      __syncthreads();
      const int64_t ce = (e - cs > kChunk) ? cs + kChunk : e;
      // the scan is issue bound (ncu: ALU pipe 60 %, ~50 thread instructions per event): chunk-relative 32-bit
      // indices and unsigned range tests keep the per-event work to a dozen instructions
      const uint32_t uW = (uint32_t)W, urows = (uint32_t)rows, uc = (uint32_t)c;
      // One predicated red.shared per event, no branch: the scan is issue bound and a branch per event costs the
      // reconvergence pair on top of the test; the shared-memory base is one 32-bit register instead of a generic
      // pointer that the compiler re-derives (S2UR / ULEA) at every use.
      const uint32_t plane_bytes = (uint32_t)nwords4 * 4u;
      uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared(cnt);
      asm volatile("mov.u32 %0, %0;" : "+r"(cnt_s));   // opaque: one register, not a constant to rematerialise per event
      auto count = [&](int xi, int yi, int ci) {
        const uint32_t yr = (uint32_t)(yi - y_lo);
        const uint32_t pix = yr * uW + (uint32_t)xi;
        const uint32_t ok = (BOTH ? 1u : (uint32_t)((uint32_t)ci == uc)) & ((uint32_t)xi < uW) & (yr < urows);
        const uint32_t plane = BOTH ? (uint32_t)ci * plane_bytes : 0u;
        asm volatile(
            "{\n\t.reg .pred q;\n\t"
            "setp.ne.u32 q, %2, 0;\n\t"
            "@q red.shared.add.u32 [%0], %1;\n\t}"
            ::"r"(cnt_s + plane + ((pix + pix) & ~3u)), "r"((pix & 1u) * 0xffffu + 1u), "r"(ok)
            : "memory");
      };
      const int n_chunk = (int)(ce - cs);   // <= 65535
      if constexpr (SRC::VW > 1) {
        // SRC::VW events per thread and iteration from 16-byte aligned groups (the scan is load-latency bound: ncu showed
        // 43 % of the kernel's stall samples behind the 2-byte / 1-byte event loads); ragged ends are masked
        constexpr int VW = SRC::VW;
        const int64_t g0 = cs & ~(int64_t)(VW - 1);      // aligned start of the chunk's first group
        const int lead = (int)(cs - g0);                 // events of that group before the chunk
        const int64_t avail = src.n - g0;               // events from g0 on that whole groups cover (clamped: int)
        const int n_full = (int)((avail < (1 << 20) ? avail : (int64_t)(1 << 20)) / VW * VW);
#pragma unroll 2
        for (int o = (int)threadIdx.x * VW; o < lead + n_chunk; o += NT * VW) {
          if (o + VW <= n_full) {
            int xs[VW], ys[VW], cc[VW];
            src.loadv(g0 + o, xs, ys, cc);
#pragma unroll
            for (int j = 0; j < VW; ++j)
              if ((uint32_t)(o + j - lead) < (uint32_t)n_chunk) count(xs[j], ys[j], cc[j]);
          } else {
            for (int i = o > lead ? o : lead; i < lead + n_chunk; ++i) {   // the last, partial group of the arrays
              int xi, yi, ci;
              src.xyc(g0 + i, xi, yi, ci);
              count(xi, yi, ci);
            }
          }
        }
      } else {
#pragma unroll 4
        for (int i = (int)threadIdx.x; i < n_chunk; i += NT) {
          int xi, yi, ci;
          src.xyc(cs + i, xi, yi, ci);
          count(xi, yi, ci);
        }
      }
      __syncthreads();
      for (int cpl = 0; cpl < NP; ++cpl) {
        const uint32_t* __restrict__ cn = cnt + cpl * nwords4;
        const int64_t bkc = BOTH ? bk * 2 + cpl : bkc0;
        uint32_t* __restrict__ out = hist + bkc * HW + (int64_t)y_lo * W;
        (void)out;
        if constexpr (kU8) {
          // compact output: one byte per bin, the exact count of a saturated bin goes to the side list (hist_u8.cuh)
          const int64_t base = bkc * HW + (int64_t)y_lo * W;
          uint8_t* __restrict__ out8 = reinterpret_cast<uint8_t*>(hist) + base;
          const uint32_t idx0 = (uint32_t)base;
          if (first && (npix & 15) == 0 && (base & 15) == 0) {
            // 8 words = 16 counters -> one 16 B store
            for (int w = threadIdx.x * 8; w < nwords; w += NT * 8) {
              const uint4 va = *reinterpret_cast<const uint4*>(cn + w), vb = *reinterpret_cast<const uint4*>(cn + w + 4);
              const uint32_t cw[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
              uint32_t o[4];
  #pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint32_t c0 = cw[2 * j] & 0xffffu, c1 = cw[2 * j] >> 16, c2 = cw[2 * j + 1] & 0xffffu, c3 = cw[2 * j + 1] >> 16;
                if ((c0 | c1 | c2 | c3) >= kHistU8Sat) {   // (a superset of "one of them saturates")
                  const uint32_t q = idx0 + 2u * (uint32_t)w + 4u * (uint32_t)j;
                  c0 = hist_u8_enc(sat_tail, q, c0), c1 = hist_u8_enc(sat_tail, q + 1, c1);
                  c2 = hist_u8_enc(sat_tail, q + 2, c2), c3 = hist_u8_enc(sat_tail, q + 3, c3);
                }
                o[j] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
              }
              st_stream_u4(reinterpret_cast<uint4*>(out8 + 2 * w), make_uint4(o[0], o[1], o[2], o[3]));
            }
          } else {
            for (int q = threadIdx.x; q < npix; q += NT) {
              const uint32_t v = (cn[q >> 1] >> ((q & 1) << 4)) & 0xffffu;
              if (first) {
                out8[q] = (uint8_t)hist_u8_enc(sat_tail, idx0 + (uint32_t)q, v);
              } else if (v) {   // a further chunk of a very long micro-bin (same thread as the first chunk's write)
                const uint32_t old = out8[q];
                if (old < kHistU8Sat) {
                  out8[q] = (uint8_t)hist_u8_enc(sat_tail, idx0 + (uint32_t)q, old + v);
                } else {
                  uint32_t n = sat_tail[0];
                  if (n > (uint32_t)EAS_HIST_U8_SAT_CAP) n = (uint32_t)EAS_HIST_U8_SAT_CAP;
                  uint2* ent = reinterpret_cast<uint2*>(sat_tail + 4);
                  for (uint32_t i = 0; i < n; ++i)
                    if (ent[i].x == idx0 + (uint32_t)q) {
                      ent[i].y += v;
                      break;
                    }
                }
              }
            }
          }
        } else if (vec_ok) {
          // 2 words = 4 counters -> one 16 B store
          for (int w = threadIdx.x * 2; w < nwords; w += NT * 2) {
            const uint2 v = *reinterpret_cast<const uint2*>(cn + w);
            const uint32_t c0 = v.x & 0xffffu, c1 = v.x >> 16, c2 = v.y & 0xffffu, c3 = v.y >> 16;
            uint4* dst = reinterpret_cast<uint4*>(out + 2 * w);
            uint4 o;
            if (first) {
              o = make_uint4(Cvt<OUT_T>::enc(c0), Cvt<OUT_T>::enc(c1), Cvt<OUT_T>::enc(c2), Cvt<OUT_T>::enc(c3));
            } else {
              const uint4 old = *dst;
              o = make_uint4(Cvt<OUT_T>::add(old.x, c0), Cvt<OUT_T>::add(old.y, c1), Cvt<OUT_T>::add(old.z, c2),
                             Cvt<OUT_T>::add(old.w, c3));
            }
            st_stream_u4(dst, o);
          }
        } else {
          for (int q = threadIdx.x; q < npix; q += NT) {
            const uint32_t v = (cn[q >> 1] >> ((q & 1) << 4)) & 0xffffu;
            out[q] = first ? Cvt<OUT_T>::enc(v) : Cvt<OUT_T>::add(out[q], v);
          }
        }
      }
      first = false;
      if (cs + kChunk < e) __syncthreads();
    }
  }
}

// ---- strategy "reds" ----------------------------------------------------------------------
constexpr int kEpt = 8;  // events per thread per iteration (one 16 B load of x and of y)

template <typename OUT_T>
__global__ void __launch_bounds__(256)
bin_hist_global_kernel(const int16_t* __restrict__ x, const int16_t* __restrict__ y,
                       const uint8_t* __restrict__ p, const int64_t* __restrict__ offsets,
                       const int64_t* __restrict__ bounds, int64_t B, int64_t n, int H, int W, int Tm,
                       OUT_T* __restrict__ hist) {
  const int64_t HW = (int64_t)H * W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * kEpt;
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kEpt; i0 < n; i0 += stride) {
    // window of the first event: largest b with offsets[b] <= i0
    int64_t lo = 0, hi = B;  // invariant: offsets[lo] <= i0 < offsets[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (offsets[mid] <= i0) lo = mid;
      else hi = mid;
    }
    int64_t b = lo;
    int64_t off_next = offsets[b + 1];
    const int64_t* bnd = bounds + b * (Tm + 1);
    int k = 0;
    __align__(16) int16_t xs[kEpt];
    __align__(16) int16_t ys[kEpt];
    __align__(8) uint8_t ps[kEpt];
    const int cnt = (n - i0 >= kEpt) ? kEpt : (int)(n - i0);
    if (cnt == kEpt) {
      *reinterpret_cast<uint4*>(xs) = ld_stream_u4(reinterpret_cast<const uint4*>(x + i0));
      *reinterpret_cast<uint4*>(ys) = ld_stream_u4(reinterpret_cast<const uint4*>(y + i0));
      *reinterpret_cast<uint2*>(ps) = ld_stream_u2(reinterpret_cast<const uint2*>(p + i0));
    } else {
#pragma unroll
      for (int j = 0; j < kEpt; ++j)
        if (j < cnt) xs[j] = x[i0 + j], ys[j] = y[i0 + j], ps[j] = p[i0 + j];
    }
#pragma unroll
    for (int j = 0; j < kEpt; ++j) {
      if (j >= cnt) break;
      const int64_t i = i0 + j;
      while (i >= off_next && b + 1 < B) {  // also skips empty windows
        ++b;
        off_next = offsets[b + 1];
        bnd += Tm + 1;
        k = 0;
      }
      if (i >= off_next) break;  // past the last window
      while (k < Tm && i >= bnd[k + 1]) ++k;
      const int xi = xs[j], yi = ys[j];
      if (k < Tm && i >= bnd[0] && (unsigned)xi < (unsigned)W && (unsigned)yi < (unsigned)H) {
        const int c = ps[j] != 0;
        atomicAdd(hist + ((b * Tm + k) * 2 + c) * HW + (int64_t)yi * W + xi, (OUT_T)1);
      }
    }
  }
}

// Event-parallel counting for windows given as arbitrary record ranges (the .dat path): every
// micro-bin segment is cut into chunks of kRangeChunk events that are dealt round-robin to the CTAs.
constexpr int kRangeChunk = 4096;

template <typename OUT_T, typename SRC>
__global__ void __launch_bounds__(256)
bin_hist_ranges_kernel(const SRC src, const int64_t* __restrict__ bounds, int64_t B, int H, int W, int Tm,
                       OUT_T* __restrict__ hist) {
  const int64_t HW = (int64_t)H * W;
  const int64_t n_seg = B * Tm;
  for (int64_t seg = 0; seg < n_seg; ++seg) {
    const int64_t b = seg / Tm;
    const int k = (int)(seg - b * Tm);
    const int64_t s = bounds[b * (Tm + 1) + k], e = bounds[b * (Tm + 1) + k + 1];
    const int64_t n_chunks = (e - s + kRangeChunk - 1) / kRangeChunk;
    OUT_T* __restrict__ h = hist + seg * 2 * HW;
    // chunk j of segment seg belongs to CTA (j + 7*seg) mod gridDim.x
    int64_t j = ((int64_t)blockIdx.x - 7 * seg) % (int64_t)gridDim.x;
    if (j < 0) j += gridDim.x;
    for (; j < n_chunks; j += gridDim.x) {
      const int64_t cs = s + j * kRangeChunk;
      const int64_t ce = cs + kRangeChunk < e ? cs + kRangeChunk : e;
      for (int64_t i = cs + threadIdx.x; i < ce; i += 256) {
        int xi, yi, ci;
        src.xyc(i, xi, yi, ci);
        if ((unsigned)xi < (unsigned)W && (unsigned)yi < (unsigned)H) atomicAdd(h + ci * HW + (int64_t)yi * W + xi, (OUT_T)1);
      }
    }
  }
}

// tile geometry shared by both front doors
struct SlabGeo {
  int n_slabs, slab_rows;
  size_t smem;
  bool fits;
};
SlabGeo slab_geo(int H, int W) {
  SlabGeo g;
  const int64_t HW = (int64_t)H * W;
  g.n_slabs = (int)eas_ceil_div(HW * 2, kSlabMaxBytes);
  if (g.n_slabs > H) g.n_slabs = H;
  g.slab_rows = (int)eas_ceil_div(H, g.n_slabs);
  g.n_slabs = (int)eas_ceil_div(H, g.slab_rows);
  g.smem = (size_t)((((int64_t)g.slab_rows * W + 1) / 2 + 3) / 4 * 4) * 4;
  g.fits = g.smem <= (size_t)kSlabMaxBytes + 4096 && g.n_slabs <= 8;
  return g;
}

template <typename OUT_T, typename SRC>
int launch_tiles(const SRC& src, const int64_t* bounds, int64_t n_items, int H, int W, int Tm, const SlabGeo& g,
                 void* hist, unsigned int* counter, cudaStream_t stream, uint32_t* sat_tail = nullptr, int force = 0) {
  // both polarities in one item (half as many scans) when two slabs fit one CTA's shared memory and there are enough
  // (window, micro-bin, slab) items for every SM; EAS_BIN_PLANES=1 keeps the plane-per-item kernel (A/B runs)
  // force: 3 = the pair kernel whenever it fits, 4 = never
  const bool both = 2 * g.smem <= 200 * 1024 && force != 4 && getenv("EAS_BIN_PLANES") == nullptr &&
                    (force == 3 || n_items / 2 >= 2 * EAS_NUM_SMS);
  if (force == 3 && !both) return EAS_E_UNSUPPORTED;
  if (both) {
    auto kb = bin_hist_smem_kernel<OUT_T, SRC, 1024, true>;
    cudaError_t eb = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * g.smem));
    if (eb != cudaSuccess) return (int)eb;
    cudaLaunchConfig_t lb = {};
    int64_t gb = EAS_NUM_SMS;
    if (gb > n_items / 2) gb = n_items / 2;
    lb.gridDim = dim3((unsigned)gb), lb.blockDim = dim3(1024), lb.dynamicSmemBytes = 2 * g.smem, lb.stream = stream;
    cudaLaunchAttribute ab[1];
    ab[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    ab[0].val.programmaticStreamSerializationAllowed = 1;
    lb.attrs = ab, lb.numAttrs = 1;
    eb = cudaLaunchKernelEx(&lb, kb, src, bounds, n_items / 2, H, W, Tm, g.n_slabs, g.slab_rows, (uint32_t*)hist, counter,
                            sat_tail);
    return eb == cudaSuccess ? EAS_OK : (int)eb;
  }
  auto kern = bin_hist_smem_kernel<OUT_T, SRC, kSmemThreads, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return (int)e;
  const int per_sm = (int)((220 * 1024) / (g.smem + 1024));
  int64_t grid = (int64_t)EAS_NUM_SMS * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
  if (grid > n_items) grid = n_items;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid), lc.blockDim = dim3(kSmemThreads), lc.dynamicSmemBytes = g.smem, lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr, lc.numAttrs = 1;
  e = cudaLaunchKernelEx(&lc, kern, src, bounds, n_items, H, W, Tm, g.n_slabs, g.slab_rows, (uint32_t*)hist, counter,
                         sat_tail);
  if (e != cudaSuccess) return (int)e;
  return EAS_OK;
}

// ---- .dat window search (one thread per labelled timestamp) ---------------------------------
// first record in [lo, hi) with t >= key
__device__ __forceinline__ int64_t dat_lower_bound(const uint2* __restrict__ rec, int64_t lo, int64_t hi, int64_t key) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)rec[mid].x < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void dat_windows_kernel(const uint2* __restrict__ rec, int64_t n, const int64_t* __restrict__ t_label,
                                   int64_t B, int64_t win_lo, int64_t win_hi, int max_backoff,
                                   int64_t* __restrict__ ranges) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t delta = win_hi - win_lo;
  const int64_t total = n > 0 ? (int64_t)rec[n - 1].x : 0;   // PSEELoader.total_time()
  int64_t cur = t_label[b] + win_lo;
  int64_t lo = 0, hi = 0;
  for (int trigger = 0;; ++trigger) {
    // ---- PSEELoader.seek_time(cur), psee_loader.py:196-238 ----
    int64_t cursor, now;
    bool done;
    if (cur > total) {
      cursor = n, now = total + 1, done = true;
    } else if (cur <= 0) {
      cursor = 0, now = 0, done = false;
    } else {
      int64_t low = 0, high = n;
      bool hit = false;
      while (high - low > 100000) {                 // term_criterion
        const int64_t middle = (low + high) / 2;
        const int64_t mid = (int64_t)rec[middle].x;
        if (mid > cur) high = middle;
        else if (mid < cur) low = middle + 1;
        else {                                       // the probe read left the file cursor one event further
          cursor = middle + 1, hit = true;
          break;
        }
      }
      if (!hit) cursor = dat_lower_bound(rec, low, high, cur);
      now = cur, done = cursor >= n;
    }
    // ---- PSEELoader.load_delta_t(delta), psee_loader.py:128-171 ----
    lo = hi = cursor;
    if (!done && cursor < n) hi = dat_lower_bound(rec, cursor, n, now + delta);
    // ---- GEN1Dataset.search_events zero_trigger loop, gen1.py:223-232 ----
    if (hi > lo || trigger > max_backoff) break;
    cur -= delta;
  }
  ranges[2 * b] = lo;
  ranges[2 * b + 1] = hi;
}

}  // namespace

extern "C" int eas_dat_windows(const void* rec, int64_t n_rec, const int64_t* t_label, int64_t B, int64_t win_lo,
                               int64_t win_hi, int32_t max_backoff, int64_t* ranges, void* stream) {
  EAS_REQUIRE(n_rec >= 0 && B >= 0 && win_hi > win_lo && max_backoff >= 0, EAS_E_SHAPE);
  if (B == 0) return EAS_OK;
  EAS_REQUIRE(t_label && ranges && (rec || n_rec == 0), EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)rec % 8 == 0, EAS_E_ALIGN);
  dat_windows_kernel<<<(unsigned)eas_ceil_div(B, 64), 64, 0, (cudaStream_t)stream>>>(
      (const uint2*)rec, n_rec, t_label, B, win_lo, win_hi, max_backoff, ranges);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

extern "C" size_t eas_bin_dat_ws_bytes(int64_t B, int Tm) { return eas_bin_events_ws_bytes(B, Tm); }

extern "C" int eas_bin_dat(const void* rec, int64_t n_rec, const int64_t* ranges, int64_t B, int H, int W, int Tm,
                           void* hist, void* ws, size_t ws_bytes, void* stream_, int strategy, int out_dtype) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EAS_REQUIRE(B >= 0 && n_rec >= 0, EAS_E_SHAPE);
  EAS_REQUIRE(H > 0 && W > 0 && H <= 16384 && W <= 16384 && Tm > 0 && Tm <= 1024, EAS_E_SHAPE);
  EAS_REQUIRE((int64_t)H * W < (1ll << 30), EAS_E_SHAPE);
  EAS_REQUIRE(strategy >= 0 && strategy <= 4, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(out_dtype == EAS_I32 || out_dtype == EAS_F32 || out_dtype == EAS_U8, EAS_E_UNSUPPORTED);
  if (B == 0) return EAS_OK;
  EAS_REQUIRE(ranges && hist && ws && (rec || n_rec == 0), EAS_E_NULL);
  EAS_REQUIRE(ws_bytes >= eas_bin_dat_ws_bytes(B, Tm), EAS_E_WORKSPACE);
  EAS_REQUIRE(((uintptr_t)rec % 8 == 0) && ((uintptr_t)hist % 16 == 0) && ((uintptr_t)ws % 8 == 0), EAS_E_ALIGN);
  int64_t* bounds = (int64_t*)ws;
  unsigned int* counter =
      (unsigned int*)((char*)ws + eas_align_up((size_t)B * (size_t)(Tm + 1) * sizeof(int64_t), 256));
  const DatSrc src{(const uint2*)rec, ranges, n_rec};
  const int64_t nb = B * (Tm + 1);
  const int64_t nbins = B * Tm * 2 * (int64_t)H * W;
  uint32_t* sat_tail = nullptr;
  if (out_dtype == EAS_U8) {
    EAS_REQUIRE(nbins < (1ll << 32), EAS_E_SHAPE);
    sat_tail = (uint32_t*)((char*)hist + hist_u8_tail_offset((size_t)nbins));
  }
  bin_bounds_kernel<DatSrc><<<(unsigned)eas_ceil_div(nb * 32, 128), 128, 0, stream>>>(src, B, Tm, bounds, counter,
                                                                                    sat_tail, g_sticky);
  EAS_LAUNCH_CHECK();
  const SlabGeo g = slab_geo(H, W);
  const int64_t n_items = B * Tm * 2 * g.n_slabs;
  // window lengths live on the device: "auto" takes the write-once tiles whenever the frame fits them
  // (event windows of tens of ms), the event-parallel kernel otherwise
  if (strategy == 0) strategy = (g.fits && (n_items >= EAS_NUM_SMS || out_dtype == EAS_U8)) ? 2 : 1;
  if (strategy >= 2) {
    EAS_REQUIRE(g.fits, EAS_E_UNSUPPORTED);
    if (out_dtype == EAS_U8)
      return launch_tiles<uint8_t>(src, bounds, n_items, H, W, Tm, g, hist, counter, stream, sat_tail, strategy);
    return out_dtype == EAS_F32 ? launch_tiles<float>(src, bounds, n_items, H, W, Tm, g, hist, counter, stream, nullptr, strategy)
                                : launch_tiles<int32_t>(src, bounds, n_items, H, W, Tm, g, hist, counter, stream, nullptr, strategy);
  }
  EAS_REQUIRE(out_dtype != EAS_U8, EAS_E_UNSUPPORTED);   // the byte form is written by the tiles kernel only
  const int64_t HW = (int64_t)H * W;
  cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)B * Tm * 2 * HW * sizeof(int32_t), stream);
  if (e != cudaSuccess) return (int)e;
  const unsigned grid = EAS_NUM_SMS * 8;
  if (out_dtype == EAS_F32)
    bin_hist_ranges_kernel<float><<<grid, 256, 0, stream>>>(src, bounds, B, H, W, Tm, (float*)hist);
  else
    bin_hist_ranges_kernel<int32_t><<<grid, 256, 0, stream>>>(src, bounds, B, H, W, Tm, (int32_t*)hist);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

extern "C" size_t eas_bin_events_ws_bytes(int64_t B, int Tm) {
  if (B <= 0 || Tm <= 0) return 0;
  // bounds[B][Tm+1] int64 + one work counter
  return eas_align_up((size_t)B * (size_t)(Tm + 1) * sizeof(int64_t), 256) + 256;
}

extern "C" int eas_bin_events_ex(const int16_t* x, const int16_t* y, const int64_t* t, const uint8_t* p,
                                 const int64_t* offsets, int64_t B, int64_t n_events, int H, int W, int Tm,
                                 void* hist, void* ws, size_t ws_bytes, void* stream_, int strategy,
                                 int out_dtype) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EAS_REQUIRE(B >= 0 && n_events >= 0, EAS_E_SHAPE);
  EAS_REQUIRE(H > 0 && W > 0 && Tm > 0 && Tm <= 1024, EAS_E_SHAPE);
  EAS_REQUIRE((int64_t)H * W < (1ll << 30), EAS_E_SHAPE);
  EAS_REQUIRE(strategy >= 0 && strategy <= 4, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(out_dtype == EAS_I32 || out_dtype == EAS_F32 || out_dtype == EAS_U8, EAS_E_UNSUPPORTED);
  if (B == 0) return EAS_OK;
  EAS_REQUIRE(offsets && hist && ws, EAS_E_NULL);
  EAS_REQUIRE(n_events == 0 || (x && y && t && p), EAS_E_NULL);
  EAS_REQUIRE(ws_bytes >= eas_bin_events_ws_bytes(B, Tm), EAS_E_WORKSPACE);
  EAS_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)p % 8 == 0) &&
                  ((uintptr_t)hist % 16 == 0) && ((uintptr_t)ws % 8 == 0),
              EAS_E_ALIGN);
  int64_t* bounds = (int64_t*)ws;
  unsigned int* counter =
      (unsigned int*)((char*)ws + eas_align_up((size_t)B * (size_t)(Tm + 1) * sizeof(int64_t), 256));
  const int64_t HW = (int64_t)H * W;
  const int64_t nb = B * (Tm + 1);
  const SoaSrc src{x, y, t, p, offsets, n_events};
  uint32_t* sat_tail = nullptr;
  if (out_dtype == EAS_U8) {
    EAS_REQUIRE(B * Tm * 2 * HW < (1ll << 32), EAS_E_SHAPE);
    sat_tail = (uint32_t*)((char*)hist + hist_u8_tail_offset((size_t)(B * Tm * 2 * HW)));
  }
  bin_bounds_kernel<SoaSrc><<<(unsigned)eas_ceil_div(nb * 32, 128), 128, 0, stream>>>(src, B, Tm, bounds, counter,
                                                                                    sat_tail, g_sticky);
  EAS_LAUNCH_CHECK();

  // row slabs so that one slab of 16-bit counters fits the per-CTA shared memory budget
  const SlabGeo g = slab_geo(H, W);
  const bool fits = g.fits;
  const int64_t n_items = B * Tm * 2 * g.n_slabs;
  if (strategy == 0) {
    // tiles: every event of a segment is scanned by 2*n_slabs CTAs (from L2); only worth it while
    // the write-once output dominates, i.e. for short windows; long windows go event-parallel.
    const bool enough = n_items >= EAS_NUM_SMS && n_events / (B * Tm) <= (1 << 17);
    strategy = (fits && (enough || out_dtype == EAS_U8)) ? 2 : 1;
  }
  if (strategy >= 2) {
    EAS_REQUIRE(fits, EAS_E_UNSUPPORTED);
    if (out_dtype == EAS_U8)
      return launch_tiles<uint8_t>(src, bounds, n_items, H, W, Tm, g, hist, counter, stream, sat_tail, strategy);
    return out_dtype == EAS_F32 ? launch_tiles<float>(src, bounds, n_items, H, W, Tm, g, hist, counter, stream, nullptr, strategy)
                                : launch_tiles<int32_t>(src, bounds, n_items, H, W, Tm, g, hist, counter, stream, nullptr, strategy);
  } else {
    EAS_REQUIRE(out_dtype != EAS_U8, EAS_E_UNSUPPORTED);   // the byte form is written by the tiles kernel only
    cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)B * Tm * 2 * HW * sizeof(int32_t), stream);
    if (e != cudaSuccess) return (int)e;
    if (n_events > 0) {
      const int64_t threads = eas_ceil_div(n_events, kEpt);
      int64_t blocks = eas_ceil_div(threads, 256);
      const int64_t cap = (int64_t)EAS_NUM_SMS * 8 * 4;  // 8 resident CTAs/SM, a few waves
      if (blocks > cap) blocks = cap;
      if (out_dtype == EAS_F32)
        bin_hist_global_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(x, y, p, offsets, bounds, B, n_events,
                                                                            H, W, Tm, (float*)hist);
      else
        bin_hist_global_kernel<int32_t><<<(unsigned)blocks, 256, 0, stream>>>(x, y, p, offsets, bounds, B, n_events,
                                                                              H, W, Tm, (int32_t*)hist);
      EAS_LAUNCH_CHECK();
    }
  }
  return EAS_OK;
}

extern "C" int eas_bin_events(const int16_t* x, const int16_t* y, const int64_t* t, const uint8_t* p,
                              const int64_t* offsets, int64_t B, int64_t n_events, int H, int W, int Tm,
                              int32_t* hist, void* ws, size_t ws_bytes, void* stream) {
  return eas_bin_events_ex(x, y, t, p, offsets, B, n_events, H, W, Tm, hist, ws, ws_bytes, stream, 0, EAS_I32);
}

// ---- compact histogram (hist_u8.cuh) -> dense counts ----------------------------------------------
namespace {
template <typename OUT_T>
__global__ void __launch_bounds__(256)
hist_u8_expand_kernel(const uint8_t* __restrict__ h, int64_t nbins, OUT_T* __restrict__ out,
                      const int* __restrict__ run_if) {
  if (run_if && *run_if == 0) return;
  const uint32_t* tail = reinterpret_cast<const uint32_t*>(h + hist_u8_tail_offset((size_t)nbins));
  const int64_t n4 = nbins >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t w = reinterpret_cast<const uint32_t*>(h)[i];
    uint32_t c[4] = {w & 0xffu, (w >> 8) & 0xffu, (w >> 16) & 0xffu, w >> 24};
    if (__vcmpeq4(w, 0xffffffffu)) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c[j] == kHistU8Sat) c[j] = hist_u8_lookup(tail, (uint32_t)(4 * i + j));
    }
    if constexpr (std::is_same<OUT_T, float>::value)
      reinterpret_cast<float4*>(out)[i] = make_float4((float)c[0], (float)c[1], (float)c[2], (float)c[3]);
    else
      reinterpret_cast<int4*>(out)[i] = make_int4((int)c[0], (int)c[1], (int)c[2], (int)c[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (nbins & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    uint32_t c = h[i];
    if (c == kHistU8Sat) c = hist_u8_lookup(tail, (uint32_t)i);
    out[i] = (OUT_T)c;
  }
}
}  // namespace

int eas_hist_u8_expand_if(const void* hist_u8, int64_t nbins, void* out, int out_dtype, const int* run_if,
                          cudaStream_t stream) {
  const unsigned grid = EAS_NUM_SMS * 8;
  if (out_dtype == EAS_F32)
    hist_u8_expand_kernel<float><<<grid, 256, 0, stream>>>((const uint8_t*)hist_u8, nbins, (float*)out, run_if);
  else
    hist_u8_expand_kernel<int32_t><<<grid, 256, 0, stream>>>((const uint8_t*)hist_u8, nbins, (int32_t*)out, run_if);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

namespace {
__global__ void hist_u8_report_kernel(const uint8_t* __restrict__ h, int64_t nbins, volatile int32_t* sticky) {
  const uint32_t* tail = reinterpret_cast<const uint32_t*>(h + hist_u8_tail_offset((size_t)nbins));
  if (tail[1] != 0u) *sticky = 1;
}
}  // namespace

extern "C" int eas_hist_u8_report(const void* hist_u8, int64_t B, int Tm, int H, int W, int32_t* sticky_flag,
                                  void* stream) {
  EAS_REQUIRE(B >= 0 && Tm > 0 && H > 0 && W > 0, EAS_E_SHAPE);
  if (B == 0) return EAS_OK;
  EAS_REQUIRE(hist_u8 && sticky_flag, EAS_E_NULL);
  hist_u8_report_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const uint8_t*)hist_u8, B * Tm * 2 * (int64_t)H * W,
                                                         sticky_flag);
  EAS_LAUNCH_CHECK();
  return EAS_OK;
}

extern "C" int eas_hist_u8_set_sticky(int32_t* sticky_flag) {
  g_sticky = sticky_flag;
  return EAS_OK;
}

extern "C" size_t eas_hist_u8_bytes(int64_t B, int Tm, int H, int W) {
  if (B <= 0 || Tm <= 0 || H <= 0 || W <= 0) return 0;
  return hist_u8_bytes((size_t)B * (size_t)Tm * 2 * (size_t)H * (size_t)W);
}

extern "C" int eas_hist_u8_expand(const void* hist_u8, int64_t B, int Tm, int H, int W, void* out, int out_dtype,
                                  void* stream) {
  EAS_REQUIRE(B >= 0 && Tm > 0 && H > 0 && W > 0, EAS_E_SHAPE);
  EAS_REQUIRE(out_dtype == EAS_I32 || out_dtype == EAS_F32, EAS_E_UNSUPPORTED);
  if (B == 0) return EAS_OK;
  EAS_REQUIRE(hist_u8 && out, EAS_E_NULL);
  EAS_REQUIRE((uintptr_t)hist_u8 % 16 == 0 && (uintptr_t)out % 16 == 0, EAS_E_ALIGN);
  const int64_t nbins = B * Tm * 2 * (int64_t)H * W;
  EAS_REQUIRE(nbins < (1ll << 32), EAS_E_SHAPE);
  return eas_hist_u8_expand_if(hist_u8, nbins, out, out_dtype, nullptr, (cudaStream_t)stream);
}
