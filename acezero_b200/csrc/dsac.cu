#include "common.cuh"
extern "C" size_t acez_dsac_workspace_bytes(int, int, int, int) { return 0; }
extern "C" int acez_dsac_forward_rgb_batch(const float*, int, int, int, const float*, const float*, const float*,
                                           const acez_dsac_params*, const int*, float*, int*, const acez_dsac_debug*,
                                           void*, size_t, acez_stream_t) {
  acez::set_error("dsac: not implemented yet");
  return ACEZ_ERR_UNSUPPORTED;
}
