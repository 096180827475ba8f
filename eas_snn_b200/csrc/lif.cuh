// LIF dynamics shared by the standalone PLIF kernels (plif.cu) and the conv epilogue
// (conv_bn_plif.cu).  Restates spikingjelly 0.0.0.0.14 ParametricLIFNode.neuronal_charge /
// neuronal_fire / jit_soft_reset / jit_hard_reset with separate IEEE mul and add, like eager PyTorch.
#pragma once
#include "common.cuh"

struct LifDyn {
  float sw, k, vth, vr, vr_eff;
  bool hard, decay_in;
};

__device__ __forceinline__ LifDyn make_lif(float w, float vth, int hard_reset, float v_reset, int decay_input) {
  LifDyn d;
  d.sw = eas_sigmoid(w);
  d.k = 1.0f - d.sw;
  d.vth = vth;
  d.hard = hard_reset != 0;
  d.vr = v_reset;
  d.vr_eff = d.hard ? v_reset : 0.0f;
  d.decay_in = decay_input != 0;
  return d;
}
__device__ __forceinline__ float lif_charge(const LifDyn& d, float v, float x) {
  if (!d.decay_in) {
    if (d.vr_eff == 0.0f) return __fadd_rn(__fmul_rn(v, d.k), x);
    return __fadd_rn(__fsub_rn(v, __fmul_rn(__fsub_rn(v, d.vr_eff), d.sw)), x);
  }
  if (d.vr_eff == 0.0f) return __fadd_rn(v, __fmul_rn(__fsub_rn(x, v), d.sw));
  return __fadd_rn(v, __fmul_rn(__fsub_rn(x, __fsub_rn(v, d.vr_eff)), d.sw));
}
__device__ __forceinline__ float lif_fire(const LifDyn& d, float h) {
  return __fsub_rn(h, d.vth) >= 0.0f ? 1.0f : 0.0f;
}
__device__ __forceinline__ float lif_reset(const LifDyn& d, float h, float s) {
  if (d.hard) return s != 0.0f ? d.vr : h;
  return __fsub_rn(h, __fmul_rn(s, d.vth));
}
__device__ __forceinline__ float lif_v_init(const LifDyn& d) { return d.hard ? d.vr : 0.0f; }
