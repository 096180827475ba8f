// Device helpers shared by the sampler forward and backward kernels: cp.async wrappers, the packed
// FFMA2 register-tiled KxK convolution, and the tile geometry.
#pragma once
#include "common.cuh"

namespace eas_sampler {

constexpr int ru4(int a) { return (a + 3) / 4 * 4; }

// Arguments of one sampler step (shared by the FP32-pipe kernel and the tensor-core kernel).
struct StepArgs {
  const void* events;     // [B][Tm][2][H][W]
  const float* s_prev;    // [B][2][H][W]
  float* s_next;
  float* vm;
  float* acc;
  uint16_t* meta;         // seg | (t_last + 1) << 8
  float* out;             // [Ts][B][2][H][W]
  float* v_seq;           // [Tm][B][2][H][W] or null
  float* gate_seq;
  eas_sampler_weights w;
  int B, H, W, Tm, Ts;
  int t;                  // sampler step (0 = newest micro-bin)
  int readout, hard_reset, write_zero, use_abs;
  float vreset, thresh;
  const int* run_if;      // FP32-pipe kernel only: when set, the launch is a no-op unless *run_if != 0
  // FP32-pipe kernel as ONE cooperative launch over steps [t, t + t_count) (the predicated fall-back behind the
  // tensor-core kernels: one idle launch instead of Tm + 1): s0 / s1 = the two spike buffers (step t reads what
  // step t - 1 wrote), grid_bar = a zeroed device word for the barrier between the steps, expand_src = compact byte
  // histogram to expand into `events` first (or null)
  int t_count;
  float* s0;
  float* s1;
  unsigned int* grid_bar;
  const void* expand_src;
};


// Tensor-core path (sampler_tc.cu): depth 2, k 5 on tcgen05.  `wimg` = eas_sampler_tc_wimg_bytes()
// of workspace holding the pre-packed weight tiles (built once per forward call).
size_t eas_sampler_tc_wimg_bytes();
// device flag raised by the tensor-core kernels when an operand did not fit fp16 (|x| >= 65504)
const int* eas_sampler_tc_flag(const void* wimg);
bool eas_sampler_tc_supported(const eas_sampler_cfg* c, const void* events, const float* out, const float* v_seq,
                              const float* gate_seq);
int eas_sampler_tc_run(const eas_sampler_cfg* cfg, StepArgs a, float* s0, float* s1, void* wimg, cudaStream_t st);

// Row-folded tensor-core path (sampler_tc2.cu): depth 2, k 5, inputs exact in one fp16 plane (event
// counts).  Compact state: meta8 = B*2*H*W bytes, sb0 / sb1 = B*2*H*W/4 spike bytes each; vm / acc as in
// StepArgs.  The flag is raised when an input is not exactly representable or a magnitude leaves the
// fp16 range: the caller then recomputes the forward on another kernel.
size_t eas_sampler_tc2_wimg_bytes();
const int* eas_sampler_tc2_flag(const void* wimg);
bool eas_sampler_tc2_supported(const eas_sampler_cfg* c, const void* events, const float* out, const float* v_seq,
                               const float* gate_seq);
int eas_sampler_tc2_run(const eas_sampler_cfg* cfg, StepArgs a, uint8_t* meta8, uint8_t* sb0, uint8_t* sb1, void* wimg,
                        cudaStream_t st);

__device__ __forceinline__ void cp_async_16(void* smem, const void* g, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* g, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* g, bool pred) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Packed fp32 FMA (Blackwell FFMA2): {d.lo, d.hi} += a * {b.lo, b.hi}; `a` is a scalar that the
// assembler encodes as a broadcast operand, so one issue slot does two IEEE fmas.
__device__ __forceinline__ void ffma2_bcast(unsigned long long& d, float a, float b0, float b1) {
  unsigned long long av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(av), "l"(bv));
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// acc[co][px] += sum_{ci,ky,kx} src[ci][ky][px+kx] * w[ci][ky][kx][co]      (CO == 4)
// The (ci, ky) loop is rolled (small code: the whole kernel must stay inside the instruction cache)
// and software pipelined: the next input row is fetched from shared memory while the current one
// feeds K*CO*PX fmas, issued as FFMA2 over output-channel pairs (accp[co/2][px] = {co, co+1}).
template <int CI, int CO, int K, int PX>
__device__ __forceinline__ void conv_acc(const float* __restrict__ src, int ch_stride, int row_stride,
                                         const float* __restrict__ wsm, unsigned long long (&accp)[CO / 2][PX]) {
  static_assert(CO == 4, "weights are fetched as float4");
  constexpr int NIN = PX + K - 1;
  constexpr int NV = (NIN + 3) / 4;
  float4 nxt[NV];
  {
    const float4* rowp = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int v = 0; v < NV; ++v) nxt[v] = rowp[v];
  }
#pragma unroll 1
  for (int it = 0; it < CI * K; ++it) {
    float in[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v)
      in[4 * v + 0] = nxt[v].x, in[4 * v + 1] = nxt[v].y, in[4 * v + 2] = nxt[v].z, in[4 * v + 3] = nxt[v].w;
    if (it + 1 < CI * K) {
      const int ci = (it + 1) / K, ky = (it + 1) - ci * K;
      const float4* rowp = reinterpret_cast<const float4*>(src + ci * ch_stride + ky * row_stride);
#pragma unroll
      for (int v = 0; v < NV; ++v) nxt[v] = rowp[v];
    }
    const float4* wp = reinterpret_cast<const float4*>(wsm + it * K * CO);
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 wq = wp[kx];
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        ffma2_bcast(accp[0][px], in[px + kx], wq.x, wq.y);
        ffma2_bcast(accp[1][px], in[px + kx], wq.z, wq.w);
      }
    }
  }
}

template <int K, int DEPTH, int TH, int TW>
struct Geo {
  static constexpr int R = K / 2;
  static constexpr int HALO = R * DEPTH;
  static constexpr int PX = 4;
  static constexpr int NT = TH * TW / PX;
  // second-layer input (h1 for depth 2, the raw tile for depth 1)
  static constexpr int HR = TH + 2 * R;
  static constexpr int HC = ru4(TW + 2 * R);
  static constexpr int HS = HC + 4;
  // first-layer input (depth 2 only)
  static constexpr int IR = HR + 2 * R;
  static constexpr int IC = HC + 2 * R;
  static constexpr int IS = ru4(IC) + 4;
  // the tile that is filled from global memory
  static constexpr int LR = DEPTH == 2 ? IR : HR;
  static constexpr int LC = DEPTH == 2 ? IC : HC;
  static constexpr int LS = DEPTH == 2 ? IS : HS;
  static constexpr int W1 = 2 * K * K * 4;   // one first-layer stack [2][K][K][4]
  static constexpr int W2 = (DEPTH == 2 ? 8 : 4) * K * K * 4;
  static constexpr int SM_H = (DEPTH == 2 ? 8 : 4) * HR * HS + 16;
  static constexpr int SM_I = DEPTH == 2 ? 4 * IR * IS + 16 : 0;
  static constexpr int SM_W = W2 + (DEPTH == 2 ? 2 * W1 : 0) + 16;
  static constexpr int SM_ST = 2 * TH * TW;  // per f32 state array
  // floats: conv buffers + weights + vm + acc + meta (u16 -> half the floats)
  static constexpr size_t SMEM = sizeof(float) * (size_t)(SM_H + SM_I + SM_W + 2 * SM_ST + SM_ST / 2);
};

}  // namespace eas_sampler
