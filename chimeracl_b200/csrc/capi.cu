// chimera-b200: library-level entry points (version, error strings).
#include "common.cuh"
#include "../../include/chimera_b200.h"

extern "C" {

int chb_version(void) { return 100; }  // 0.1.0

const char* chb_error_string(int code) {
  if (code == CHB_OK) return "ok";
  if (code == CHB_ERR_ARG) return "chimera_b200: invalid argument";
  if (code == CHB_ERR_WORKSPACE) return "chimera_b200: workspace too small";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "chimera_b200: unknown error";
}

}  // extern "C"
