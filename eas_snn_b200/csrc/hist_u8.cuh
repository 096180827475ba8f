// Compact histogram (out_dtype / in_dtype EAS_U8): one byte per bin + an exact side list for the rare
// saturated bins, in ONE buffer:
//   bytes [0, nbins)                 min(count, 255); 255 = "saturated: the count is in the list"
//   at eas_align_up(nbins, 256)      uint32 n_sat   entries appended (may exceed the capacity)
//                                    uint32 lost    1 when an entry did not fit: the histogram is NOT exact any more
//                                    uint64 sticky  address of a caller-registered int32 (eas_hist_u8_set_sticky) that is set to 1
//                                                   together with `lost`, or 0
//   then EAS_HIST_U8_SAT_CAP         uint2 {bin index, count >= 255}
// The reference's histogram is an int64 / float tensor (gen1.py:330-360); on event data almost every count is a
// single digit, so the byte form carries the same information in a quarter of the fp32 bytes -- what the binning
// kernel has to write and the sampler has to read.
#pragma once
#include <cstdint>
#include "common.cuh"
#include "../../include/eas_b200.h"

constexpr uint32_t kHistU8Sat = 255u;

__host__ __device__ static inline size_t hist_u8_tail_offset(size_t nbins) { return (nbins + 255) / 256 * 256; }
static inline size_t hist_u8_bytes(size_t nbins) {
  return hist_u8_tail_offset(nbins) + 16 + (size_t)EAS_HIST_U8_SAT_CAP * 8;
}

// binning side: remember the exact count of a saturated bin
__device__ __forceinline__ void hist_u8_append(uint32_t* __restrict__ tail, uint32_t idx, uint32_t count) {
  const uint32_t slot = atomicAdd(tail, 1u);
  if (slot < (uint32_t)EAS_HIST_U8_SAT_CAP)
    reinterpret_cast<uint2*>(tail + 4)[slot] = make_uint2(idx, count);
  else {
    tail[1] = 1u;
    int32_t* sticky = *reinterpret_cast<int32_t* const*>(tail + 2);   // pinned host int of the caller, if registered
    if (sticky) *reinterpret_cast<volatile int32_t*>(sticky) = 1;
  }
}
__device__ __forceinline__ uint32_t hist_u8_enc(uint32_t* __restrict__ tail, uint32_t idx, uint32_t count) {
  if (count < kHistU8Sat) return count;
  hist_u8_append(tail, idx, count);
  return kHistU8Sat;
}
// consumer side: the exact count of a bin whose byte is 255 (255 itself when its entry was lost)
static __device__ __noinline__ uint32_t hist_u8_lookup(const uint32_t* __restrict__ tail, uint32_t idx) {
  uint32_t n = tail[0];
  if (n > (uint32_t)EAS_HIST_U8_SAT_CAP) n = (uint32_t)EAS_HIST_U8_SAT_CAP;
  const uint2* e = reinterpret_cast<const uint2*>(tail + 4);
  for (uint32_t i = 0; i < n; ++i)
    if (e[i].x == idx) return e[i].y;
  return kHistU8Sat;
}

// grid-stride expansion of a compact histogram into dense fp32 counts (callable from inside a kernel)
static __device__ __forceinline__ void hist_u8_expand_f32(const uint8_t* __restrict__ h, int64_t nbins, float* __restrict__ out) {
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
    reinterpret_cast<float4*>(out)[i] = make_float4((float)c[0], (float)c[1], (float)c[2], (float)c[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (nbins & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    uint32_t c = h[i];
    if (c == kHistU8Sat) c = hist_u8_lookup(tail, (uint32_t)i);
    out[i] = (float)c;
  }
}

// dense counts (EAS_F32 / EAS_I32) of a compact histogram; a no-op unless *run_if != 0 when run_if is given (bin_events.cu)
int eas_hist_u8_expand_if(const void* hist_u8, int64_t nbins, void* out, int out_dtype, const int* run_if,
                          cudaStream_t stream);
