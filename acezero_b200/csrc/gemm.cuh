// tcgen05 / TMA GEMM used by the ACE head (1x1-conv MLP, ace_network.py:62-149 of the reference) and, through the
// implicit-GEMM front end in encoder.cu, by the encoder convolutions (ace_network.py:14-59).
//
//   D[z][M,N] = A[z] * B[z]         fp16 operands, fp32 accumulation in TMEM
//
// Operand storage (both handled with SWIZZLE_128B shared-memory tiles filled by TMA):
//   K-major  A: memory [M, K] row-major (K contiguous)      MN-major A: memory [K, M] row-major (M contiguous)
//   K-major  B: memory [N, K] row-major (K contiguous)      MN-major B: memory [K, N] row-major (N contiguous)
// The three head passes map onto it without any transposed copies:
//   forward  Y = X W^T         : A = X  (K-major),  B = W  (K-major)
//   dgrad    dX = dZ W         : A = dZ (K-major),  B = W  (MN-major, contraction over the out-channel rows of W)
//   wgrad    dW = dZ^T X       : A = dZ (MN-major), B = X  (MN-major, contraction over the batch rows)
#pragma once
#include "common.cuh"

namespace acez {

enum GemmEpi : int {
  EPI_FWD = 0,    // out = fp16(act(acc + fp16(bias)));  out2 = fp16(resid + out)            (optional pieces)
  EPI_DGRAD = 1,  // v = acc (+ addend); out2 = fp16(v) (optional); out = mask > 0 ? fp16(v) : 0; non-finite flag
  EPI_WGRAD = 2,  // out32[z] = acc (fp32); optional bias-gradient column (sum over the contraction of A)
};

// Implicit-GEMM convolution front end (encoder, ace_network.py:26-39): the A tile of a k-block is one 4-D TMA box of an
// NHWC activation tensor: 16 x 8 output pixels x 64 input channels at filter tap (ky, kx); zero padding comes from
// TMA out-of-bounds fill, stride 2 from the tensor map's element strides. Output rows are NHWC pixels.
struct ConvGeom {
  int enabled;
  int cin_blocks;  // Cin / 64
  int ksize;       // 3 (1x1 convolutions run as plain GEMMs)
  int stride, pad;
  int Ho, Wo;      // output extent
  int tiles_x, tiles_y;
};
static constexpr int kConvTileW = 16, kConvTileH = 8;  // 128 output pixels per CTA tile

struct GemmArgs {
  ConvGeom conv;
  int M, N;       // logical output extent; rows >= M / cols >= N are not stored
  int k_blocks;   // contraction length / 64
  // UMMA descriptor constants (bytes); set by gemm_prepare, exposed so the GPU test can probe alternatives
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;
  // fp16 epilogues
  const float* bias;     // [N] fp32 master bias, rounded to fp16 before the add (autocast semantics); nullable
  const __half* resid;   // [M,ldo] nullable
  const __half* mask;    // [M,ldo] EPI_DGRAD: ReLU mask source (post-activation output of the producing layer)
  const __half* addend;  // [M,ldo] EPI_DGRAD: skip-path gradient added before masking; nullable
  __half* out;           // [M,ldo]
  __half* out2;          // [M,ldo] nullable
  int ldo;
  int relu;
  int* nonfinite;        // EPI_DGRAD: set to 1 if a stored value is inf/nan; nullable
  // fp32 epilogue
  float* out32;
  long long out32_zstride;
  int ldo32;
  float* bias_grad;      // [z][M] nullable: column sum over the contraction dimension of A (dZ^T 1)
  long long bias_grad_zstride;
  long long* dbg_clock;  // nullable: 8 clock64 stamps per CTA (profiling probe)
  int st_out, st_out2, ld_op;  // set by gemm_finalize: which of tmOut / tmOut2 / tmOp are live
};

struct GemmLaunch {
  CUtensorMap tmA, tmB;
  CUtensorMap tmOut, tmOut2, tmOp;  // fp16 epilogues: staged TMA stores / prefetched operand tile (gemm_finalize)
  GemmArgs args;
  int bn;        // 64, 128 or 256
  int a_mn, b_mn;
  int epi;
  int batch;
};

struct GemmProblem {
  const __half* A;
  const __half* B;
  int a_mn, b_mn;             // 0 = K-major, 1 = MN-major
  int M, N, K;
  int batch;                  // >= 1
  long long a_zstride, b_zstride;  // elements
  int lda, ldb;               // leading dimension (elements) of the 2-D operand as stored
  int bn;                     // 0 = choose
  int epi;
};

// Fills L->tmA/tmB and the descriptor constants; the caller then fills the epilogue pointers in L->args.
int gemm_prepare(GemmLaunch* L, const GemmProblem& p);
// After the epilogue pointers are set: encode the output / operand tensor maps (fp16 epilogues).
int gemm_finalize(GemmLaunch* L);
// pdl: launch with the programmatic-dependent-launch attribute (only valid when the stream predecessor is a kernel)
int gemm_launch(const GemmLaunch& L, cudaStream_t stream, bool pdl = true);

}  // namespace acez
