// Shared helpers for the sm_100a kernels of libeas_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/eas_b200.h"

#define EAS_NUM_SMS 148  // B200: 2 dies x 74 SMs

#define EAS_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

// Launch check: returns the positive cudaError_t of a failed launch (no sync).
#define EAS_LAUNCH_CHECK()                       \
  do {                                           \
    cudaError_t _e = cudaGetLastError();         \
    if (_e != cudaSuccess) return (int)_e;       \
  } while (0)

static inline int64_t eas_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t eas_align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ float eas_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// Streaming (read-once) loads / write-once stores that do not pollute L1.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream_u4(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// Coherent (L2) loads for data the PREVIOUS kernel of the stream wrote, in kernels launched with the programmatic
// dependent-launch attribute: such a grid starts before that kernel has finished, which is outside the contract of
// the read-only (.nc) path even though every dependent access comes after griddepcontrol.wait.
__device__ __forceinline__ uint4 ld_cg_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_cg_u8(const uint8_t* p) {
  uint32_t r;
  asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

