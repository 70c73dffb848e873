// ACE encoder (reference ace_network.py:14-59) — placeholder entry points until the implicit-GEMM kernels land.
#include "common.cuh"

extern "C" size_t acez_encoder_workspace_bytes(int, int, int) { return 0; }
extern "C" int acez_encoder_plan_create(const float* const*, int, int, int, void*, size_t, acez_stream_t,
                                        acez_encoder_plan**) {
  acez::set_error("encoder: not implemented yet");
  return ACEZ_ERR_UNSUPPORTED;
}
extern "C" void acez_encoder_plan_destroy(acez_encoder_plan*) {}
extern "C" int acez_encoder_out_hw(int H, int W, int* h8, int* w8) {
  // three stride-2 3x3 convs with padding 1: n -> floor((n - 1) / 2) + 1
  auto down = [](int n) { return (n - 1) / 2 + 1; };
  if (h8) *h8 = down(down(down(H)));
  if (w8) *w8 = down(down(down(W)));
  return ACEZ_OK;
}
extern "C" int acez_encoder_forward(acez_encoder_plan*, const void*, int, int, int, int, void*, acez_stream_t) {
  acez::set_error("encoder: not implemented yet");
  return ACEZ_ERR_UNSUPPORTED;
}
