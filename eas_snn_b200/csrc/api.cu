// Version / error-string entry points of libeas_b200.so.
#include "common.cuh"

extern "C" int eas_abi_version(void) { return EAS_ABI_VERSION; }

extern "C" const char* eas_error_string(int code) {
  switch (code) {
    case EAS_OK: return "ok";
    case EAS_E_NULL: return "EAS_E_NULL: a required pointer is NULL";
    case EAS_E_SHAPE: return "EAS_E_SHAPE: a dimension is out of the supported range";
    case EAS_E_UNSUPPORTED: return "EAS_E_UNSUPPORTED: flag / dtype / kernel-size combination not built";
    case EAS_E_WORKSPACE: return "EAS_E_WORKSPACE: workspace too small";
    case EAS_E_ALIGN: return "EAS_E_ALIGN: pointer not aligned as required";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown eas error";
}
