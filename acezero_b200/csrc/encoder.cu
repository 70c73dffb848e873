// ACE encoder (reference ace_network.py:14-59, weights ace_encoder_pretrained.pt) on sm_100a.
//
//   conv1 1->32 3x3 s1      : CUDA-core direct convolution (K = 9 is too small for the tensor core), fused bias + ReLU,
//                             NHWC fp16 output padded to 64 channels (upper 32 = 0) so that conv2 is a full 64-wide k-block
//   conv2..conv4, res*_conv1/3 (3x3): tcgen05 implicit GEMM (gemm.cu conv front end): A tiles are 4-D TMA boxes of the
//                             NHWC activation (zero padding = TMA OOB fill, stride 2 = tensor-map element strides),
//                             B = weights packed once to fp16 [Cout, 9*Cin] in (tap, cin) order
//   1x1 convolutions        : plain GEMMs on the NHWC rows
//   `res = res + x` (ace_network.py:51) and `res2_skip(res) + x` (:57) are fused into the producing GEMM's epilogue.
// Activations stay NHWC fp16; the output [n, h/8, w/8, 512] is exactly the row order `normalize_shape`
// (ace_trainer.py:399-401) produces per image, so the patch buffer can be filled with row gathers.
// Arithmetic as under CUDA autocast in the reference: fp16 operands, fp32 accumulation, fp16 bias, fp16 outputs.
#include <vector>

#include "gemm.cuh"

namespace acez {

static inline size_t align_up_e(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int down2(int n) { return (n - 1) / 2 + 1; }  // 3x3, stride 2, pad 1

struct ConvSpec {
  int cin, cout, k, stride;
  int cin_pad;  // channels of the stored input activation (conv2 reads the 64-padded conv1 output)
};
// index = position in the state dict: conv1..conv4, res1_conv1..3, res2_conv1..3, res2_skip
static const ConvSpec kSpecs[11] = {
    {1, 32, 3, 1, 1},      {32, 64, 3, 2, 64},    {64, 128, 3, 2, 64},   {128, 256, 3, 2, 128},
    {256, 256, 3, 1, 256}, {256, 256, 1, 1, 256}, {256, 256, 3, 1, 256}, {256, 512, 3, 1, 256},
    {512, 512, 1, 1, 512}, {512, 512, 3, 1, 512}, {256, 512, 1, 1, 256},
};

// OIHW fp32 -> [Cout][k*k][cin_pad] fp16 (tap-major, channel-minor; zero for padded channels)
__global__ void pack_weights_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin, int cin_pad,
                                    int kk) {
  const size_t n = (size_t)cout * kk * cin_pad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin_pad);
    const int tap = (int)((i / cin_pad) % kk);
    const int o = (int)(i / ((size_t)cin_pad * kk));
    out[i] = (c < cin) ? __float2half_rn(w[((size_t)o * cin + c) * kk + tap]) : __float2half_rn(0.f);
  }
}

// conv1: one thread per output pixel, 32 output channels, input [n,1,H,W] (fp16 or fp32), output NHWC [n,H,W,64]
template <typename TIn>
__global__ void __launch_bounds__(256) conv1_kernel(const TIn* __restrict__ img, const float* __restrict__ w /*[32][9]*/,
                                                    const float* __restrict__ b, __half* __restrict__ out, int n, int H,
                                                    int W) {
  __shared__ float sw[32 * 9];
  __shared__ float sb[32];
  for (int i = threadIdx.x; i < 288; i += 256) sw[i] = __half2float(__float2half_rn(w[i]));  // autocast: fp16 weights
  if (threadIdx.x < 32) sb[threadIdx.x] = __half2float(__float2half_rn(b[threadIdx.x]));
  __syncthreads();
  const size_t total = (size_t)n * H * W;
  const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= total) return;
  const int x = (int)(p % W), y = (int)((p / W) % H);
  const size_t base = p - (size_t)y * W - x;  // image offset (single channel)
  float v[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      float t = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) t = (float)img[base + (size_t)yy * W + xx];
      v[ky * 3 + kx] = __half2float(__float2half_rn(t));  // autocast: fp16 input
    }
  uint4 o[8];
  __half2* oh = reinterpret_cast<__half2*>(o);
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      a0 = fmaf(v[t], sw[c * 9 + t], a0);
      a1 = fmaf(v[t], sw[(c + 1) * 9 + t], a1);
    }
    oh[c / 2] = __floats2half2_rn(fmaxf(a0 + sb[c], 0.f), fmaxf(a1 + sb[c + 1], 0.f));
  }
#pragma unroll
  for (int k = 16; k < 32; ++k) oh[k] = __floats2half2_rn(0.f, 0.f);
  uint4* dst = reinterpret_cast<uint4*>(out + p * 64);
#pragma unroll
  for (int k = 0; k < 8; ++k) dst[k] = o[k];
}

}  // namespace acez

struct acez_encoder_plan {
  int max_n, max_h, max_w;
  const float* w[11];
  const float* b[11];
  __half* wp[11];   // packed fp16 weights (index 0 unused)
  __half* act[10];  // activation buffers
  int prepared_n, prepared_h, prepared_w;
  std::vector<acez::GemmLaunch> launches;  // layers 1..10 in execution order
  void* prepared_out;
};

namespace acez {

struct EncDims {
  int h[4], w[4];  // resolution after conv1 (full), conv2, conv3, conv4
};
static EncDims enc_dims(int H, int W) {
  EncDims d;
  d.h[0] = H; d.w[0] = W;
  for (int i = 1; i < 4; ++i) { d.h[i] = down2(d.h[i - 1]); d.w[i] = down2(d.w[i - 1]); }
  return d;
}

// activation buffers: 0 conv1 out (64 ch), 1 conv2 out (64), 2 conv3 out (128), 3 res (256), 4 t1 (256), 5 t2 (256),
// 6 res' = res + x (256), 7 u1 (512), 8 u2 (512), 9 x10 (512)
static const int kActCh[10] = {64, 64, 128, 256, 256, 256, 256, 512, 512, 512};
static const int kActLvl[10] = {0, 1, 2, 3, 3, 3, 3, 3, 3, 3};

struct EncLayout {
  size_t wp[11], act[10], total;
};
static EncLayout enc_layout(int n, int H, int W) {
  EncLayout lo{};
  size_t off = 0;
  for (int l = 1; l < 11; ++l) {
    lo.wp[l] = off;
    off = align_up_e(off + (size_t)kSpecs[l].cout * kSpecs[l].k * kSpecs[l].k * kSpecs[l].cin_pad * 2, 1024);
  }
  const EncDims d = enc_dims(H, W);
  for (int a = 0; a < 10; ++a) {
    lo.act[a] = off;
    off = align_up_e(off + (size_t)n * d.h[kActLvl[a]] * d.w[kActLvl[a]] * kActCh[a] * 2, 1024);
  }
  lo.total = off;
  return lo;
}

static int conv_launch_prepare(GemmLaunch* L, const __half* in, int n, int Hin, int Win, int cin_pad, const __half* wp,
                               int cout, int k, int stride, int Ho, int Wo) {
  GemmProblem p{};
  p.A = in;  // replaced by the 4-D map for 3x3
  p.B = wp;
  p.a_mn = 0; p.b_mn = 0;
  p.M = n * Ho * Wo;
  p.N = cout;
  p.K = k * k * cin_pad;
  p.batch = 1;
  p.lda = (k == 3) ? p.K : cin_pad;  // 3x3: the 2-D map built here is replaced by the 4-D NHWC map below
  p.ldb = p.K;
  p.bn = 0;
  p.epi = EPI_FWD;
  int rc = gemm_prepare(L, p);
  if (rc) return rc;
  if (k == 3) {
    uint64_t dims[4] = {(uint64_t)cin_pad, (uint64_t)Win, (uint64_t)Hin, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)cin_pad * 2, (uint64_t)Win * cin_pad * 2, (uint64_t)Hin * Win * cin_pad * 2};
    // "to load N elements along a strided dimension, boxDim must be N * elementStride"
    uint32_t box[4] = {64, (uint32_t)(kConvTileW * stride), (uint32_t)(kConvTileH * stride), 1};
    uint32_t estr[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    rc = make_tensor_map(&L->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, in, dims, strides, box, estr,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    ConvGeom& cg = L->args.conv;
    cg.enabled = 1;
    cg.cin_blocks = cin_pad / 64;
    cg.ksize = 3;
    cg.stride = stride;
    cg.pad = 1;
    cg.Ho = Ho; cg.Wo = Wo;
    cg.tiles_x = (Wo + kConvTileW - 1) / kConvTileW;
    cg.tiles_y = (Ho + kConvTileH - 1) / kConvTileH;
    L->batch = n;
  }
  L->args.ldo = cout;
  L->args.relu = 1;
  return ACEZ_OK;
}

static int encoder_prepare(acez_encoder_plan* e, int n, int H, int W, void* out) {
  if (e->prepared_n == n && e->prepared_h == H && e->prepared_w == W && e->prepared_out == out) return ACEZ_OK;
  const EncDims d = enc_dims(H, W);
  e->launches.assign(10, GemmLaunch{});
  auto L = [&](int i) -> GemmLaunch* { return &e->launches[i]; };
  int rc;
  const int h8 = d.h[3], w8 = d.w[3];
  // conv2: act0 -> act1
  if ((rc = conv_launch_prepare(L(0), e->act[0], n, d.h[0], d.w[0], 64, e->wp[1], 64, 3, 2, d.h[1], d.w[1]))) return rc;
  L(0)->args.bias = e->b[1]; L(0)->args.out = e->act[1];
  // conv3: act1 -> act2
  if ((rc = conv_launch_prepare(L(1), e->act[1], n, d.h[1], d.w[1], 64, e->wp[2], 128, 3, 2, d.h[2], d.w[2]))) return rc;
  L(1)->args.bias = e->b[2]; L(1)->args.out = e->act[2];
  // conv4: act2 -> res (act3)
  if ((rc = conv_launch_prepare(L(2), e->act[2], n, d.h[2], d.w[2], 128, e->wp[3], 256, 3, 2, h8, w8))) return rc;
  L(2)->args.bias = e->b[3]; L(2)->args.out = e->act[3];
  // res1_conv1 3x3: res -> t1
  if ((rc = conv_launch_prepare(L(3), e->act[3], n, h8, w8, 256, e->wp[4], 256, 3, 1, h8, w8))) return rc;
  L(3)->args.bias = e->b[4]; L(3)->args.out = e->act[4];
  // res1_conv2 1x1: t1 -> t2
  if ((rc = conv_launch_prepare(L(4), e->act[4], n, h8, w8, 256, e->wp[5], 256, 1, 1, h8, w8))) return rc;
  L(4)->args.bias = e->b[5]; L(4)->args.out = e->act[5];
  // res1_conv3 3x3: t2 -> x ; res' = res + x  (ace_network.py:49-51)
  if ((rc = conv_launch_prepare(L(5), e->act[5], n, h8, w8, 256, e->wp[6], 256, 3, 1, h8, w8))) return rc;
  L(5)->args.bias = e->b[6]; L(5)->args.out = nullptr; L(5)->args.resid = e->act[3]; L(5)->args.out2 = e->act[6];
  // res2_conv1 3x3 256->512: res' -> u1
  if ((rc = conv_launch_prepare(L(6), e->act[6], n, h8, w8, 256, e->wp[7], 512, 3, 1, h8, w8))) return rc;
  L(6)->args.bias = e->b[7]; L(6)->args.out = e->act[7];
  // res2_conv2 1x1: u1 -> u2
  if ((rc = conv_launch_prepare(L(7), e->act[7], n, h8, w8, 512, e->wp[8], 512, 1, 1, h8, w8))) return rc;
  L(7)->args.bias = e->b[8]; L(7)->args.out = e->act[8];
  // res2_conv3 3x3: u2 -> x10
  if ((rc = conv_launch_prepare(L(8), e->act[8], n, h8, w8, 512, e->wp[9], 512, 3, 1, h8, w8))) return rc;
  L(8)->args.bias = e->b[9]; L(8)->args.out = e->act[9];
  // res2_skip 1x1 on res' (no ReLU) + x10 -> features (ace_network.py:57)
  if ((rc = conv_launch_prepare(L(9), e->act[6], n, h8, w8, 256, e->wp[10], 512, 1, 1, h8, w8))) return rc;
  L(9)->args.bias = e->b[10]; L(9)->args.relu = 0; L(9)->args.out = nullptr; L(9)->args.resid = e->act[9];
  L(9)->args.out2 = reinterpret_cast<__half*>(out);
  for (int i = 0; i < 10; ++i)
    if ((rc = gemm_finalize(L(i)))) return rc;
  e->prepared_n = n; e->prepared_h = H; e->prepared_w = W; e->prepared_out = out;
  return ACEZ_OK;
}

}  // namespace acez

using namespace acez;

extern "C" size_t acez_encoder_workspace_bytes(int max_n, int max_h, int max_w) {
  if (max_n < 1 || max_h < 8 || max_w < 8) return 0;
  return enc_layout(max_n, max_h, max_w).total + 1024;
}

extern "C" int acez_encoder_out_hw(int H, int W, int* h8, int* w8) {
  const EncDims d = enc_dims(H, W);
  if (h8) *h8 = d.h[3];
  if (w8) *w8 = d.w[3];
  return ACEZ_OK;
}

extern "C" int acez_encoder_plan_create(const float* const* weights, int max_n, int max_h, int max_w, void* workspace,
                                        size_t workspace_bytes, acez_stream_t stream, acez_encoder_plan** out) {
  ACEZ_REQUIRE(weights && workspace && out, "encoder_plan_create: null argument");
  ACEZ_REQUIRE(max_n >= 1 && max_h >= 8 && max_w >= 8, "encoder_plan_create: bad capacity");
  ACEZ_REQUIRE(workspace_bytes >= acez_encoder_workspace_bytes(max_n, max_h, max_w), "encoder_plan_create: workspace too small");
  int rc = acez_device_check();
  if (rc) return rc;
  for (int i = 0; i < 22; ++i) ACEZ_REQUIRE(weights[i] != nullptr, "encoder_plan_create: weight tensor %d is null", i);
  acez_encoder_plan* e = new acez_encoder_plan();
  e->max_n = max_n; e->max_h = max_h; e->max_w = max_w;
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up_e(reinterpret_cast<uintptr_t>(workspace), 1024));
  const EncLayout lo = enc_layout(max_n, max_h, max_w);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  for (int l = 0; l < 11; ++l) {
    e->w[l] = weights[2 * l];
    e->b[l] = weights[2 * l + 1];
    e->wp[l] = nullptr;
    if (l >= 1) {
      e->wp[l] = reinterpret_cast<__half*>(base + lo.wp[l]);
      const ConvSpec& sp = kSpecs[l];
      pack_weights_kernel<<<256, 256, 0, s>>>(e->w[l], e->wp[l], sp.cout, sp.cin, sp.cin_pad, sp.k * sp.k);
      ACEZ_CUDA(cudaGetLastError());
    }
  }
  for (int a = 0; a < 10; ++a) e->act[a] = reinterpret_cast<__half*>(base + lo.act[a]);
  e->prepared_n = e->prepared_h = e->prepared_w = -1;
  e->prepared_out = nullptr;
  *out = e;
  return ACEZ_OK;
}

extern "C" void acez_encoder_plan_destroy(acez_encoder_plan* plan) { delete plan; }

extern "C" int acez_encoder_forward(acez_encoder_plan* e, const void* image, int image_is_fp16, int n, int H, int W,
                                    void* features_out, acez_stream_t stream) {
  ACEZ_REQUIRE(e && image && features_out, "encoder_forward: null argument");
  ACEZ_REQUIRE(n >= 1 && H >= 8 && W >= 8, "encoder_forward: bad shape n=%d H=%d W=%d", n, H, W);
  // the activation buffers were laid out for (max_n, max_h, max_w): any (n, H, W) whose every level fits is fine
  {
    // every activation level of (n, H, W) must fit the buffer laid out for (max_n, max_h, max_w)
    const EncDims d = enc_dims(H, W), dm = enc_dims(e->max_h, e->max_w);
    bool fits = true;
    for (int lvl = 0; lvl < 4; ++lvl)
      fits &= (size_t)n * d.h[lvl] * d.w[lvl] <= (size_t)e->max_n * dm.h[lvl] * dm.w[lvl];
    ACEZ_REQUIRE(fits, "encoder_forward: n=%d H=%d W=%d exceeds the plan capacity (%d, %d, %d)", n, H, W, e->max_n,
                 e->max_h, e->max_w);
  }
  int rc = acez_device_check();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  rc = encoder_prepare(e, n, H, W, features_out);
  if (rc) return rc;
  const size_t px = (size_t)n * H * W;
  const unsigned grid = (unsigned)((px + 255) / 256);
  if (image_is_fp16)
    conv1_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(image), e->w[0], e->b[0], e->act[0], n, H, W);
  else
    conv1_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(image), e->w[0], e->b[0], e->act[0], n, H, W);
  ACEZ_CUDA(cudaGetLastError());
  for (int i = 0; i < 10; ++i) {
    rc = gemm_launch(e->launches[i], s);
    if (rc) return rc;
  }
  return ACEZ_OK;
}
