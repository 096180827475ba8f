// Entry points declared in include/eas_b200.h whose kernels are not built yet.
#include "common.cuh"

extern "C" size_t eas_sampler_bwd_ws_bytes(const eas_sampler_cfg*) { return 0; }
extern "C" int eas_sampler_bwd(const eas_sampler_cfg*, const void*, const eas_sampler_weights*, const float*,
                               const float*, const float*, const eas_sampler_grads*, float*, void*, size_t, void*) {
  return EAS_E_UNSUPPORTED;
}
