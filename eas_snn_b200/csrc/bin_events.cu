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
//         the other slabs) + 4 B/bin.  Chunks of <= 65535 events make 16-bit overflow impossible.
//      bin_hist_global_kernel ("reds"): event-parallel, 8 events per thread with 16 B loads,
//         red.global.add.u32 into the (L2-resident when it fits) histogram after a memset.  Used for
//         frames that do not fit in shared memory and for very long windows.
#include "common.cuh"

namespace {

constexpr int kSmemThreads = 512;
constexpr int kSlabMaxBytes = 72 * 1024;   // 3 CTAs/SM: phases of different CTAs overlap
constexpr int kChunk = 65535;

// One warp per (window, boundary): 32-ary search on the sorted timestamps (4 dependent loads for
// 1e5 events instead of 17).  Also resets the work counter of the tile kernel.
__global__ void __launch_bounds__(128)
bin_bounds_kernel(const int64_t* __restrict__ t, const int64_t* __restrict__ offsets, int64_t B, int Tm,
                  int64_t* __restrict__ bounds, unsigned int* __restrict__ work_counter) {
  const int lane = threadIdx.x & 31;
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gid == 0 && lane == 0) *work_counter = 0u;
  if (gid >= B * (Tm + 1)) return;
  const int64_t b = gid / (Tm + 1);
  const int k = (int)(gid - b * (Tm + 1));
  const int64_t s = offsets[b], e = offsets[b + 1];
  if (e <= s) {
    if (lane == 0) bounds[gid] = s;
    return;
  }
  const int64_t t0 = t[s];
  const int64_t tw = (t[e - 1] - t0) / Tm;  // sorted => non-negative => trunc == floor
  const int64_t key = t0 + (int64_t)k * tw;
  int64_t lo = s, hi = e;  // answer (first index with t[i] >= key) is in [lo, hi]
  while (hi - lo > 0) {
    const int64_t len = hi - lo;
    const int64_t step = (len + 31) / 32;  // probes at lo + (lane+1)*step - 1
    const int64_t pi = lo + (int64_t)(lane + 1) * step - 1;
    const bool less = pi < hi ? (t[pi] < key) : false;  // out-of-range probes count as ">= key"
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

template <typename OUT_T>
__global__ void __launch_bounds__(kSmemThreads)
bin_hist_smem_kernel(const int16_t* __restrict__ x, const int16_t* __restrict__ y,
                     const uint8_t* __restrict__ p, const int64_t* __restrict__ bounds, int64_t n_items,
                     int H, int W, int Tm, int n_slabs, int slab_rows, uint32_t* __restrict__ hist,
                     unsigned int* __restrict__ work_counter) {
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
    const int64_t bkc = item / n_slabs;  // (b*Tm + k)*2 + c
    const int c = (int)(bkc & 1);
    const int64_t bk = bkc >> 1;
    const int64_t b = bk / Tm;
    const int k = (int)(bk - b * Tm);
    const int64_t s = bounds[b * (Tm + 1) + k], e = bounds[b * (Tm + 1) + k + 1];
    const int y_lo = slab * slab_rows;
    const int rows = min(slab_rows, H - y_lo);
    const int npix = rows * W;
    const int nwords = (npix + 1) >> 1;
    const int nwords4 = (nwords + 3) & ~3;
    uint32_t* __restrict__ out = hist + bkc * HW + (int64_t)y_lo * W;
    const bool vec_ok = (npix & 3) == 0 && ((((int64_t)y_lo * W) & 3) == 0) && ((HW & 3) == 0);
    bool first = true;
    for (int64_t cs = s; first || cs < e; cs += kChunk) {
      for (int w = threadIdx.x * 4; w < nwords4; w += kSmemThreads * 4)
        *reinterpret_cast<uint4*>(cnt + w) = make_uint4(0u, 0u, 0u, 0u);
      __syncthreads();
      const int64_t ce = (e - cs > kChunk) ? cs + kChunk : e;
#pragma unroll 4
      for (int64_t i = cs + threadIdx.x; i < ce; i += kSmemThreads) {
        const int yi = (int)y[i] - y_lo;
        const int ci = p[i] != 0;
        const int xi = x[i];
        if (ci == c && (unsigned)xi < (unsigned)W && (unsigned)yi < (unsigned)rows) {
          const int pix = yi * W + xi;
          atomicAdd(cnt + (pix >> 1), 1u << ((pix & 1) << 4));
        }
      }
      __syncthreads();
      if (vec_ok) {
        // 2 words = 4 counters -> one 16 B store
        for (int w = threadIdx.x * 2; w < nwords; w += kSmemThreads * 2) {
          const uint2 v = *reinterpret_cast<const uint2*>(cnt + w);
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
        for (int q = threadIdx.x; q < npix; q += kSmemThreads) {
          const uint32_t v = (cnt[q >> 1] >> ((q & 1) << 4)) & 0xffffu;
          out[q] = first ? Cvt<OUT_T>::enc(v) : Cvt<OUT_T>::add(out[q], v);
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

}  // namespace

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
  EAS_REQUIRE(strategy >= 0 && strategy <= 2, EAS_E_UNSUPPORTED);
  EAS_REQUIRE(out_dtype == EAS_I32 || out_dtype == EAS_F32, EAS_E_UNSUPPORTED);
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
  bin_bounds_kernel<<<(unsigned)eas_ceil_div(nb * 32, 128), 128, 0, stream>>>(t, offsets, B, Tm, bounds, counter);
  EAS_LAUNCH_CHECK();

  // row slabs so that one slab of 16-bit counters fits the per-CTA shared memory budget
  int n_slabs = (int)eas_ceil_div(HW * 2, kSlabMaxBytes);
  if (n_slabs > H) n_slabs = H;
  const int slab_rows = (int)eas_ceil_div(H, n_slabs);
  n_slabs = (int)eas_ceil_div(H, slab_rows);
  const size_t smem = (size_t)((((int64_t)slab_rows * W + 1) / 2 + 3) / 4 * 4) * 4;
  const bool fits = smem <= (size_t)kSlabMaxBytes + 4096 && n_slabs <= 8;
  const int64_t n_items = B * Tm * 2 * n_slabs;
  if (strategy == 0) {
    // tiles: every event of a segment is scanned by 2*n_slabs CTAs (from L2); only worth it while
    // the write-once output dominates, i.e. for short windows; long windows go event-parallel.
    const bool enough = n_items >= EAS_NUM_SMS && n_events / (B * Tm) <= (1 << 17);
    strategy = (fits && enough) ? 2 : 1;
  }
  if (strategy == 2) {
    EAS_REQUIRE(fits, EAS_E_UNSUPPORTED);
    auto kern = out_dtype == EAS_F32 ? bin_hist_smem_kernel<float> : bin_hist_smem_kernel<int32_t>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int per_sm = (int)((220 * 1024) / (smem + 1024));
    int64_t grid = (int64_t)EAS_NUM_SMS * (per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
    if (grid > n_items) grid = n_items;
    kern<<<(unsigned)grid, kSmemThreads, smem, stream>>>(x, y, p, bounds, n_items, H, W, Tm, n_slabs, slab_rows,
                                                         (uint32_t*)hist, counter);
    EAS_LAUNCH_CHECK();
  } else {
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
