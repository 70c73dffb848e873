// ACE head (reference ace_network.py:62-149) forward / backward on sm_100a, plus the fused "tail" kernel that
// joins fc3, the homogeneous de-normalisation (ace_network.py:139-147) and the reprojection loss + backward
// (ace_trainer.py:521-613) into one pass over the last hidden activation.
//
// Data layout in HBM (all caller-owned; see acez_head_workspace_bytes):
//   params / grads : flat fp32, per hidden layer W[512,512] then b[512]; fc3 W[C3,512], b[C3] last
//   W16            : [L][512][512] fp16 shadow of the hidden-layer weights (autocast's cast of the conv weights)
//   W3h            : [4][512] fp16 shadow of fc3
//   ACT            : [L+1][max_rows][512] fp16; ACT[l] is the input of hidden layer l, ACT[L] feeds fc3
//   XTRA           : [nres][max_rows][512] fp16 post-ReLU output of the last conv of each residual block (ReLU mask)
//   DZ             : [L][max_rows][512] fp16 gradient w.r.t. the pre-activation of each hidden layer (x grad_scale)
//   GRES           : [max_rows][512] fp16 running skip-path gradient
//   MASKB          : [L+1][max_rows][64 B] one bit per channel: ReLU masks written by the fused forward chain and read
//                    by the fused dgrad chain (head_chain.cu); MASKB[l] belongs to ACT[l]
#include <stdlib.h>

#include <vector>

#include "gemm.cuh"
#include "gemm2cta.cuh"
#include "head_chain.cuh"
#include "repro_loss.cuh"

namespace acez {

static constexpr int kC = 512;  // head width, hard-coded in the reference (ace_network.py:76)
static constexpr size_t kLayerStride = (size_t)kC * kC + kC;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace acez

struct acez_head_plan {
  acez_head_config cfg;
  int L, nres, C3;
  size_t n_params;
  float* params;
  float* grads;
  __half* W16;
  __half* W3h;
  __half* ACT;
  __half* XTRA;
  __half* DZ;
  __half* GRES;
  uint8_t* MASKB;
  float* G3;
  float* FC3PART;
  float* BLKPART;
  unsigned int* BLKCOUNT;
  bool counters_zeroed;
  size_t act_stride;  // max_rows * 512
  int prepared_rows;
  int prepared_training;
  std::vector<acez::GemmLaunch> fwd;
  std::vector<acez::GemmLaunch> dgrad;
  acez::GemmLaunch wgrad;                    // all layers in one launch (grid.z = layer)
  std::vector<acez::GemmLaunch> wgrad_layer;  // one launch per layer, run on a side stream under the dgrad chain
  // fused layer chains (head_chain.cu): one launch for all hidden layers of a pass
  int use_chain;
  acez::ChainLaunch chain_fwd, chain_bwd;
  int use_wgrad2;              // ACEZ_WGRAD_2CTA=1 (experimental): batched weight gradient on cta_group::2 tiles (gemm2cta.cu)
  acez::Gemm2Launch wgrad2;
  cudaStream_t side_stream;
  cudaEvent_t ev_dz[32];
  cudaEvent_t ev_join;
  bool side_ready;
  int overlap_wgrad;
  int fc3_overlap;     // ACEZ_FC3_OVERLAP=1 (experimental): the fc3 weight-gradient kernels run on the side stream, under the dgrad chain
  bool fc3_pending;    // a join with the side stream is due at the end of the backward
};

namespace acez {

struct HeadLayout {
  size_t w16, w3h, act, xtra, dz, gres, maskb, g3, fc3part, blkpart, total;
};

static HeadLayout head_layout(const acez_head_config& cfg) {
  const int nres = cfg.num_res_blocks;
  const int L = 3 * nres + 2;
  const size_t rows = (size_t)cfg.max_rows;
  HeadLayout o{};
  size_t off = 0;
  o.w16 = off; off = align_up(off + (size_t)L * kC * kC * 2, 1024);
  o.w3h = off; off = align_up(off + 4 * kC * 2, 1024);
  o.act = off; off = align_up(off + (size_t)(L + 1) * rows * kC * 2, 1024);
  if (cfg.training) {
    o.xtra = off; off = align_up(off + (size_t)nres * rows * kC * 2, 1024);
    o.dz = off; off = align_up(off + (size_t)L * rows * kC * 2, 1024);
    o.gres = off; off = align_up(off + rows * kC * 2, 1024);
    o.maskb = off; off = align_up(off + (size_t)(L + 1) * rows * 64, 1024);
    o.g3 = off; off = align_up(off + rows * 4 * sizeof(float), 1024);
    o.fc3part = off; off = align_up(off + ((rows + 31) / 32) * (size_t)(4 * kC + 4) * sizeof(float), 1024);
  }
  o.blkpart = off; off = align_up(off + 4096 * 8 * sizeof(float) + 256, 1024);  // tail per-block partials + counter
  o.total = off;
  return o;
}

// Launch with the programmatic-dependent-launch attribute: the kernel may be scheduled while its predecessor in the
// stream still runs; every kernel launched this way starts with pdl_wait() (griddepcontrol.wait) before it touches
// global memory and calls pdl_launch_dependents() so that ITS successor can be scheduled early in turn.
template <typename... KArgs, typename... Args>
static int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // The attribute is only safe when the stream predecessor is a kernel: griddepcontrol.wait does not order against a
  // preceding memcpy / memset / cross-stream event, so the FIRST kernel of every C-ABI call is launched plainly.
  cfg.numAttrs = pdl ? 1 : 0;
  ACEZ_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
  return ACEZ_OK;
}

// ----------------------------------------------------------------------------------------------
// small kernels
// ----------------------------------------------------------------------------------------------
__global__ void cast_weights_kernel(const float* __restrict__ params, __half* __restrict__ W16,
                                    __half* __restrict__ W3h, int L, int C3) {
  const size_t n_hidden = (size_t)L * kC * kC;
  const size_t n_total = n_hidden + (size_t)4 * kC;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n_hidden) {
      const size_t l = i / ((size_t)kC * kC), r = i % ((size_t)kC * kC);
      W16[i] = __float2half_rn(params[l * kLayerStride + r]);
    } else {
      const size_t r = i - n_hidden;  // [4][512]
      const size_t row = r / kC;
      W3h[r] = (row < (size_t)C3) ? __float2half_rn(params[(size_t)L * kLayerStride + r]) : __float2half_rn(0.f);
    }
  }
}

__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ idx, int rows,
                                   int row_bytes, uint8_t* __restrict__ dst) {
  // one warp per row, 16-byte vectors when the row size allows
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const uint8_t* s = src + (size_t)idx[warp] * row_bytes;
  uint8_t* d = dst + (size_t)warp * row_bytes;
  if ((row_bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    for (int o = lane * 16; o < row_bytes; o += 512) *reinterpret_cast<uint4*>(d + o) = *reinterpret_cast<const uint4*>(s + o);
  } else if ((row_bytes & 3) == 0) {
    for (int o = lane * 4; o < row_bytes; o += 128) *reinterpret_cast<uint32_t*>(d + o) = *reinterpret_cast<const uint32_t*>(s + o);
  } else {
    for (int o = lane * 2; o < row_bytes; o += 64) *reinterpret_cast<uint16_t*>(d + o) = *reinterpret_cast<const uint16_t*>(s + o);
  }
}

// All arrays of the patch buffer in one launch: blockIdx.y selects the array (reference ace_trainer.py:485-494
// issues 8 index kernels + 8 H2D index copies per iteration).
struct MultiGather {
  const uint8_t* src[8];
  uint8_t* dst[8];
  int row_bytes[8];
};
__global__ void gather_rows_multi_kernel(const MultiGather g, const int64_t* __restrict__ idx, int rows, int n_arrays) {
  pdl_wait();
  pdl_launch_dependents();
  // one warp per batch row, all arrays: the 1 KB feature row moves as 2 x 16 B per lane, the small arrays as words
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const size_t src_row = (size_t)idx[warp];
  for (int a = 0; a < n_arrays; ++a) {
    const int rb = g.row_bytes[a];
    const uint8_t* s = g.src[a] + src_row * rb;
    uint8_t* d = g.dst[a] + (size_t)warp * rb;
    if ((rb & 15) == 0 && ((reinterpret_cast<uintptr_t>(g.src[a]) | reinterpret_cast<uintptr_t>(g.dst[a])) & 15) == 0) {
      for (int o = lane * 16; o < rb; o += 512) *reinterpret_cast<uint4*>(d + o) = *reinterpret_cast<const uint4*>(s + o);
    } else if ((rb & 3) == 0) {
      for (int o = lane * 4; o < rb; o += 128) *reinterpret_cast<uint32_t*>(d + o) = *reinterpret_cast<const uint32_t*>(s + o);
    } else {
      for (int o = lane * 2; o < rb; o += 64) *reinterpret_cast<uint16_t*>(d + o) = *reinterpret_cast<const uint16_t*>(s + o);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// tail kernel: fc3 + homogeneous + (loss + backward into DZ[L-1]); one warp per row, lane owns columns
// [16*lane, 16*lane+16) with its slice of the fc3 weights held in registers. The gradient w.r.t. the 4 fc3 outputs is
// written to G3 [rows,4]; fc3_wgrad_kernel turns it into dW3 / db3.
// ----------------------------------------------------------------------------------------------
struct TailArgs {
  int rows, C3, use_homogeneous, training;
  float mean[3], h_beta, max_inv_scale, min_inv_scale;
  const __half* x;      // ACT[L]  [rows,512]
  const __half* W3h;    // [4,512] fp16
  const float* b3;      // [C3] fp32
  float* sc_out;        // [rows,3] nullable
  // training
  acez_loss_params lp;
  const float* grad_scale_dev;  // nullable: overrides lp.grad_scale (GradScaler state lives on the device)
  const float* loss_weight_dev; // nullable: overrides lp.loss_weight (per-iteration dyntanh weight, graph-stable)
  const float* tpx; const float* Pin; const float* A; const float* T; const float* K; const float* Kinv; const float* G;
  float* d_P; float* d_Kdiag;
  const float* d_sc_in; // training == 2: gradient w.r.t. the scene coordinates supplied by the caller (autograd)
  __half* dz;           // DZ[L-1] [rows,512]
  float* g3;            // [rows,4] gradient w.r.t. the fc3 outputs (fp16-rounded values)
  float* stats;         // [4]: written (not accumulated) by the last block to finish
  int* nonfinite;       // written (not OR-ed) by the last block: later kernels of the iteration OR into it
  float* blk_part;      // [gridDim.x][8] per-block partial sums
  unsigned int* blk_count;  // self-resetting completion counter
};

static constexpr int kTailThreads = 256;

__global__ void __launch_bounds__(kTailThreads) head_tail_kernel(const TailArgs a) {
#include "head_tail_body.inc"
}
// ACEZ_TAIL_OCC2=1 (experimental, round 2): the same body capped at 128 registers (184 otherwise; a few hundred bytes of
// spills) so that two CTAs = 16 warps fit per SM and the whole grid is resident in one wave (round-1 profile: 12.7 % of the
// warp slots active, two waves of 148 CTAs).
__global__ void __launch_bounds__(kTailThreads, 2) head_tail_kernel_occ2(const TailArgs a) {
#include "head_tail_body.inc"
}

// dW3[j][c] = sum_rows G3[row][j] * x8[row][c], db3[j] = sum_rows G3[row][j].
// Stage 1: one block per 32-row slab; thread (cg, rs) owns 8 columns (one 16-byte load per row) of every 4th row, all
// loads of a thread are independent; slab partials go to global. Stage 2: sum the slab partials (coalesced), write the
// gradient, and fold the GradScaler overflow check for these values into the same pass.
static constexpr int kFc3Threads = 256;
static constexpr int kFc3Rows = 32;
__global__ void __launch_bounds__(kFc3Threads) fc3_wgrad_partial_kernel(const __half* __restrict__ x,
                                                                       const float* __restrict__ g3, int rows,
                                                                       float* __restrict__ part /*[nblk][2052]*/) {
  __shared__ float sAcc[4][64][33];
  __shared__ float sB[4][4];
  pdl_wait();
  pdl_launch_dependents();
  const int t = threadIdx.x, cg = t & 63, rs = t >> 6;
  const int r0 = blockIdx.x * kFc3Rows, r1 = min(rows, r0 + kFc3Rows);
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < kFc3Rows / 4; ++i) {
    const int r = r0 + rs + 4 * i;
    if (r < r1) {
      const uint4 xv = *reinterpret_cast<const uint4*>(x + (size_t)r * kC + 8 * cg);
      const float4 g = *reinterpret_cast<const float4*>(g3 + 4 * (size_t)r);
      const __half2* xh = reinterpret_cast<const __half2*>(&xv);
      const float gj[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 f = __half22float2(xh[c]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[j][2 * c] = fmaf(gj[j], f.x, acc[j][2 * c]);
          acc[j][2 * c + 1] = fmaf(gj[j], f.y, acc[j][2 * c + 1]);
        }
      }
      if (cg == 0) { accb[0] += g.x; accb[1] += g.y; accb[2] += g.z; accb[3] += g.w; }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) sAcc[rs][cg][j * 8 + c] = acc[j][c];
  if (cg == 0)
    for (int j = 0; j < 4; ++j) sB[rs][j] = accb[j];
  __syncthreads();
  float* out = part + (size_t)blockIdx.x * (4 * kC + 4);
  for (int idx = t; idx < 4 * kC; idx += kFc3Threads) {
    const int j = idx / kC, col = idx % kC;
    const int g = col >> 3, c = col & 7;
    out[idx] = sAcc[0][g][j * 8 + c] + sAcc[1][g][j * 8 + c] + sAcc[2][g][j * 8 + c] + sAcc[3][g][j * 8 + c];
  }
  if (t < 4) out[4 * kC + t] = sB[0][t] + sB[1][t] + sB[2][t] + sB[3][t];
}

__global__ void fc3_reduce_kernel(const float* __restrict__ part, int nblk, int C3, float* __restrict__ gW3,
                                  float* __restrict__ gb3, int* __restrict__ nonfinite) {
  pdl_wait();
  pdl_launch_dependents();
  // 8 threads per output element (strided over the slab partials), 3 shuffles to combine
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gt >> 3, sub = gt & 7;
  const int total = 4 * kC + 4;
  float s = 0.f;
  if (idx < total)
    for (int b = sub; b < nblk; b += 8) s += part[(size_t)b * total + idx];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  bool bad = false;
  if (idx < total && sub == 0) {
    if (idx < 4 * kC) {
      if (idx < C3 * kC) { gW3[idx] = s; bad = !isfinite(s) || fabsf(s) > 65504.f; }
    } else if (idx - 4 * kC < C3) {
      gb3[idx - 4 * kC] = s;
      bad = !isfinite(s) || fabsf(s) > 65504.f;
    }
  }
  if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0 && nonfinite != nullptr) atomicOr(nonfinite, 1);
}

// ----------------------------------------------------------------------------------------------
// GradScaler inf check + AdamW + GradScaler.update, all on the device (no host sync)
//   scaler_state: [0] scale S, [1] growth tracker, [2] optimizer step count t
// ----------------------------------------------------------------------------------------------
__global__ void grad_check_kernel(const float* __restrict__ g, size_t n, int* __restrict__ found_inf) {
  pdl_wait();
  pdl_launch_dependents();
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    // under autocast the weight gradient is materialised in fp16: |g| > 65504 overflows to inf there
    bad |= !isfinite(v) || fabsf(v) > 65504.f;
  }
  if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) atomicOr(found_inf, 1);
}

__device__ __forceinline__ void scaler_update(float* st, int found, int use_scaler) {
  // torch.cuda.amp.GradScaler.update(): backoff 0.5 on inf, growth x2 every 2000 clean steps
  if (use_scaler && found) {
    st[0] *= 0.5f;
    st[1] = 0.f;
  } else {
    st[2] += 1.f;
    if (use_scaler) {
      st[1] += 1.f;
      if (st[1] >= 2000.f) { st[0] *= 2.f; st[1] = 0.f; }
    }
  }
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, size_t n, const float* __restrict__ hyper,
                             float* __restrict__ scaler_state, const int* __restrict__ found_inf,
                             int use_scaler, __half* __restrict__ W16, __half* __restrict__ W3h, int L, int C3) {
  pdl_wait();
  pdl_launch_dependents();
  const int found = use_scaler ? *found_inf : 0;
  if (!found) {  // GradScaler.step skips optimizer.step() on inf/nan
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
    const float inv_scale = use_scaler ? 1.f / scaler_state[0] : 1.f;
    const float t = scaler_state[2] + 1.f;  // this step's index (torch: state['step'] += 1 before use)
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float step_size = lr / bc1;
    const float bc2_sqrt = sqrtf(bc2);
    auto update = [&](float gi, float& pi, float& mi, float& vi) {
      if (use_scaler) gi = __half2float(__float2half_rn(gi));  // fp16 weight gradient of the autocast conv
      gi *= inv_scale;                                          // GradScaler.unscale_
      pi *= (1.f - lr * wd);                                    // decoupled weight decay (torch adamw)
      mi = mi + (1.f - b1) * (gi - mi);                         // exp_avg.lerp_(grad, 1 - beta1)
      vi = b2 * vi + (1.f - b2) * gi * gi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      pi -= step_size * (mi / denom);
    };
    const size_t n4 = n / 4;
    const size_t wsz = (size_t)kC * kC;
    const size_t tstride = (size_t)gridDim.x * blockDim.x;
    for (size_t q0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q0 < n4; q0 += 2 * tstride) {
     // two independent groups per trip: 8 loads in flight per thread
     float4 G[2], P[2], M[2], V[2];
#pragma unroll
     for (int u = 0; u < 2; ++u) {
       const size_t q = q0 + u * tstride;
       if (q < n4) {
         G[u] = reinterpret_cast<const float4*>(g)[q]; P[u] = reinterpret_cast<float4*>(p)[q];
         M[u] = reinterpret_cast<float4*>(m)[q]; V[u] = reinterpret_cast<float4*>(v)[q];
       }
     }
#pragma unroll
     for (int u = 0; u < 2; ++u) {
      const size_t q = q0 + u * tstride;
      if (q >= n4) continue;
      const float4 g4 = G[u];
      float4 p4 = P[u], m4 = M[u], v4 = V[u];
      update(g4.x, p4.x, m4.x, v4.x);
      update(g4.y, p4.y, m4.y, v4.y);
      update(g4.z, p4.z, m4.z, v4.z);
      update(g4.w, p4.w, m4.w, v4.w);
      reinterpret_cast<float4*>(p)[q] = p4;
      reinterpret_cast<float4*>(m)[q] = m4;
      reinterpret_cast<float4*>(v)[q] = v4;
      if (W16 != nullptr) {  // refresh the fp16 shadow the next forward reads (layer strides are multiples of 4)
        const size_t i = 4 * q, l = i / kLayerStride, r = i % kLayerStride;
        __half2 lo = __floats2half2_rn(p4.x, p4.y), hi = __floats2half2_rn(p4.z, p4.w);
        uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        if (l < (size_t)L) {
          if (r < wsz) *reinterpret_cast<uint2*>(W16 + l * wsz + r) = pk;
        } else if (r + 3 < (size_t)C3 * kC) {
          *reinterpret_cast<uint2*>(W3h + r) = pk;
        } else {
          const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
          for (int k = 0; k < 4; ++k)
            if (r + k < (size_t)C3 * kC) W3h[r + k] = __float2half_rn(pv[k]);
        }
      }
     }
    }
    for (size_t i = 4 * n4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      float pi = p[i], mi = m[i], vi = v[i];
      update(g[i], pi, mi, vi);
      p[i] = pi; m[i] = mi; v[i] = vi;
      if (W16 != nullptr) {
        const size_t l = i / kLayerStride, r = i % kLayerStride;
        if (l >= (size_t)L && r < (size_t)C3 * kC) W3h[r] = __float2half_rn(pi);
      }
    }
  }
  // the last block to finish applies GradScaler.update() (scale / growth tracker / step count); the completion
  // counter lives in scaler_state[3] (bit pattern, starts at 0) and resets itself
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int* cnt = reinterpret_cast<unsigned int*>(scaler_state + 3);
    if (atomicAdd(cnt, 1u) == gridDim.x - 1) {
      scaler_update(scaler_state, found, use_scaler);
      *cnt = 0u;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// plan
// ----------------------------------------------------------------------------------------------
static int head_prepare(acez_head_plan* h, int rows, int training) {
  if (h->prepared_rows == rows && h->prepared_training >= training) return ACEZ_OK;
  const int L = h->L;
  h->fwd.assign(L, GemmLaunch{});
  for (int l = 0; l < L; ++l) {
    GemmProblem p{};
    p.A = h->ACT + (size_t)l * h->act_stride;
    p.B = h->W16 + (size_t)l * kC * kC;
    p.a_mn = 0; p.b_mn = 0;
    p.M = rows; p.N = kC; p.K = kC; p.batch = 1;
    p.lda = kC; p.ldb = kC;
    p.bn = 0;
    p.epi = EPI_FWD;
    int rc = gemm_prepare(&h->fwd[l], p);
    if (rc) return rc;
    GemmArgs& a = h->fwd[l].args;
    a.bias = h->params + (size_t)l * kLayerStride + (size_t)kC * kC;
    a.relu = 1;
    a.ldo = kC;
    const bool res_end = (l % 3 == 2) && (l < 3 * h->nres);
    if (res_end) {
      const int k = l / 3;
      // x -> XTRA[k] (ReLU mask for the backward), res = ACT[l-2] + x -> ACT[l+1]   (ace_network.py:126,133)
      a.out = (h->XTRA != nullptr) ? h->XTRA + (size_t)k * h->act_stride : nullptr;  // nullptr: x is not kept
      a.resid = h->ACT + (size_t)(l - 2) * h->act_stride;
      a.out2 = h->ACT + (size_t)(l + 1) * h->act_stride;
    } else {
      a.out = h->ACT + (size_t)(l + 1) * h->act_stride;
    }
    rc = gemm_finalize(&h->fwd[l]);
    if (rc) return rc;
  }
  if (training) {
    h->dgrad.assign(L, GemmLaunch{});
    for (int l = L - 1; l >= 1; --l) {
      // gradient w.r.t. ACT[l] = DZ[l] * W_l, then through the ReLU of the layer that produced ACT[l]
      GemmProblem p{};
      p.A = h->DZ + (size_t)l * h->act_stride;
      p.B = h->W16 + (size_t)l * kC * kC;  // [out, in] row-major: contraction over rows -> MN-major B
      p.a_mn = 0; p.b_mn = 1;
      p.M = rows; p.N = kC; p.K = kC; p.batch = 1;
      p.lda = kC; p.ldb = kC;
      p.bn = 0;
      p.epi = EPI_DGRAD;
      int rc = gemm_prepare(&h->dgrad[l], p);
      if (rc) return rc;
      GemmArgs& a = h->dgrad[l].args;
      a.ldo = kC;
      a.out = h->DZ + (size_t)(l - 1) * h->act_stride;
      a.nonfinite = nullptr;  // patched per call
      const bool is_res = (l % 3 == 0) && (l <= 3 * h->nres);
      if (is_res) {
        const int k = l / 3;  // ACT[l] = res_k = res_{k-1} + x_{l-1}
        a.mask = h->XTRA + (size_t)(k - 1) * h->act_stride;
        a.addend = (k < h->nres) ? h->GRES : nullptr;   // skip gradient from res_{k+1}
        a.out2 = (k >= 2) ? h->GRES : nullptr;          // res_{k-1} needs it (res_0 = features has no grad)
      } else {
        a.mask = h->ACT + (size_t)l * h->act_stride;
      }
      rc = gemm_finalize(&h->dgrad[l]);
      if (rc) return rc;
    }
    // all weight gradients in one launch: grid.z = layer, no split-K, plain fp32 stores
    GemmProblem p{};
    p.A = h->DZ; p.B = h->ACT;
    p.a_mn = 1; p.b_mn = 1;
    p.M = kC; p.N = kC; p.K = (rows + 63) / 64 * 64; p.batch = L;
    p.a_zstride = (long long)h->act_stride; p.b_zstride = (long long)h->act_stride;
    p.lda = kC; p.ldb = kC;
    {
      const char* e = getenv("ACEZ_WGRAD_BN");
      // measured (round 1, warm graph replays): 128 x 128 tiles / 128 CTAs: 220 us per iteration, 128 x 256 / 64 CTAs: 231 us
      p.bn = (e != nullptr && atoi(e) == 256) ? 256 : 128;
    }
    p.epi = EPI_WGRAD;
    int rc = gemm_prepare(&h->wgrad, p);
    if (rc) return rc;
    // contraction rows beyond `rows` must read as zero: rebuild the maps with the true row count so TMA zero-fills
    {
      uint64_t dims[3] = {(uint64_t)kC, (uint64_t)rows, (uint64_t)L};
      uint64_t strides[2] = {(uint64_t)kC * 2, (uint64_t)h->act_stride * 2};
      uint32_t box[3] = {64, 64, 1};
      rc = make_tensor_map(&h->wgrad.tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, h->DZ, dims, strides, box, nullptr,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      rc = make_tensor_map(&h->wgrad.tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, h->ACT, dims, strides, box, nullptr,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    if (h->use_wgrad2) {
      // the same operands (MN-major DZ / ACT, rows beyond `rows` zero-filled by TMA) on 256 x bn tiles of CTA pairs
      Gemm2Launch& W = h->wgrad2;
      W = Gemm2Launch{};
      W.tmA = h->wgrad.tmA;
      W.tmB = h->wgrad.tmB;
      W.batch = L;
      W.a_mn = W.b_mn = 1;
      {
        const char* e = getenv("ACEZ_WGRAD_2CTA_BN");
        W.bn = (e != nullptr && atoi(e) == 256) ? 256 : 128;  // 128: 64 pairs = 128 CTAs without split-K
      }
      Gemm2Args& g = W.args;
      g.M = kC; g.N = kC; g.k_blocks = (rows + 63) / 64;
      g.tiles_n = kC / W.bn;
      g.out32 = h->grads; g.out32_zstride = (long long)kLayerStride; g.ldo32 = kC;
      g.bias_grad = h->grads + (size_t)kC * kC; g.bias_grad_zstride = (long long)kLayerStride;
      g.a_lbo = 8192; g.a_sbo = 1024; g.a_kstep = 2048;
      g.b_lbo = 8192; g.b_sbo = 1024; g.b_kstep = 2048;
    }
    GemmArgs& a = h->wgrad.args;
    a.out32 = h->grads;
    a.out32_zstride = (long long)kLayerStride;
    a.ldo32 = kC;
    a.bias_grad = h->grads + (size_t)kC * kC;
    a.bias_grad_zstride = (long long)kLayerStride;
    rc = gemm_finalize(&h->wgrad);
    if (rc) return rc;
    // per-layer variant (128 x 128 tiles, 16 CTAs per layer): small enough to run on the SMs the 80-CTA dgrad kernels
    // leave idle, so the weight gradients are computed concurrently with the dgrad chain on a second stream
    h->wgrad_layer.assign(L, GemmLaunch{});
    for (int l = 0; l < L; ++l) {
      GemmProblem q = p;
      q.A = h->DZ + (size_t)l * h->act_stride;
      q.B = h->ACT + (size_t)l * h->act_stride;
      q.batch = 1;
      q.bn = 128;
      rc = gemm_prepare(&h->wgrad_layer[l], q);
      if (rc) return rc;
      uint64_t dims[3] = {(uint64_t)kC, (uint64_t)rows, 1};
      uint64_t strides[2] = {(uint64_t)kC * 2, (uint64_t)h->act_stride * 2};
      uint32_t box[3] = {64, 64, 1};
      rc = make_tensor_map(&h->wgrad_layer[l].tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, q.A, dims, strides, box, nullptr,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      rc = make_tensor_map(&h->wgrad_layer[l].tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, q.B, dims, strides, box, nullptr,
                           CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      GemmArgs& b = h->wgrad_layer[l].args;
      b.out32 = h->grads + (size_t)l * kLayerStride;
      b.out32_zstride = 0;
      b.ldo32 = kC;
      b.bias_grad = h->grads + (size_t)l * kLayerStride + (size_t)kC * kC;
      b.bias_grad_zstride = 0;
      rc = gemm_finalize(&h->wgrad_layer[l]);
      if (rc) return rc;
    }
  }
  if (h->use_chain) {
    // ---- fused forward chain: step l = hidden layer l ----
    int rc = chain_prepare(&h->chain_fwd, CHAIN_FWD, h->ACT, h->W16, L, h->ACT, (long long)h->act_stride, L + 1, rows);
    if (rc) return rc;
    ChainArgs& f = h->chain_fwd.args;
    f.n_steps = L;
    f.flags |= kChainFlagResInit;
    const size_t mask_stride = (size_t)h->cfg.max_rows * 64;
    for (int l = 0; l < L; ++l) {
      ChainStep st{};
      st.w_layer = l;
      st.relu = 1;
      st.bias = h->params + (size_t)l * kLayerStride + (size_t)kC * kC;
      st.out_slot = (h->XTRA != nullptr || l == L - 1) ? l + 1 : -1;  // inference plans keep the tiles on chip
      st.res_add = ((l % 3 == 2) && (l < 3 * h->nres)) ? 1 : 0;        // res_{k+1} = res_k + x (ace_network.py:126,133)
      st.mask_out = (h->MASKB != nullptr) ? h->MASKB + (size_t)(l + 1) * mask_stride : nullptr;
      f.step[l] = st;
    }
    if (training) {
      // ---- fused dgrad chain: step s handles layer l = L-1-s (gradient w.r.t. ACT[l], through the ReLU below) ----
      rc = chain_prepare(&h->chain_bwd, CHAIN_DGRAD, h->DZ + (size_t)(L - 1) * h->act_stride, h->W16, L, h->DZ,
                         (long long)h->act_stride, L, rows);
      if (rc) return rc;
      ChainArgs& b = h->chain_bwd.args;
      b.n_steps = L - 1;
      for (int l = L - 1; l >= 1; --l) {
        ChainStep st{};
        st.w_layer = l;
        st.out_slot = l - 1;
        st.mask_in = h->MASKB + (size_t)l * mask_stride;  // (x > 0) of the layer that produced ACT[l]
        const bool is_res = (l % 3 == 0) && (l <= 3 * h->nres);
        if (is_res) {
          const int k = l / 3;              // ACT[l] = res_k = res_{k-1} + x_{l-1}
          st.res_add = (k < h->nres) ? 1 : 0;   // + skip gradient from res_{k+1}
          st.res_save = (k >= 2) ? 1 : 0;       // res_{k-1} needs it (res_0 = features has no gradient)
        }
        b.step[L - 1 - l] = st;
      }
    }
  }
  h->prepared_rows = rows;
  h->prepared_training = training;
  return ACEZ_OK;
}

static void fill_tail_common(const acez_head_plan* h, int rows, TailArgs& t) {
  t.rows = rows;
  t.C3 = h->C3;
  t.use_homogeneous = h->cfg.use_homogeneous;
  for (int i = 0; i < 3; ++i) t.mean[i] = h->cfg.mean[i];
  t.h_beta = h->cfg.h_beta;
  t.max_inv_scale = h->cfg.max_inv_scale;
  t.min_inv_scale = h->cfg.min_inv_scale;
  t.x = h->ACT + (size_t)h->L * h->act_stride;
  t.W3h = h->W3h;
  t.b3 = h->params + (size_t)h->L * kLayerStride + (size_t)h->C3 * kC;
}

static int tail_grid(int rows) {
  const int per_block = kTailThreads / 32;
  int g = (rows + per_block - 1) / per_block;
  // CTAs per SM (default 2; ACEZ_TAIL_BLOCKS_PER_SM to probe): fewer rows per warp shorten the serial per-row chain, more
  // blocks lengthen the last-block reduction over the per-block partials (4096 slots)
  static const int per_sm = [] {
    const char* e = getenv("ACEZ_TAIL_BLOCKS_PER_SM");
    const int v = e != nullptr ? atoi(e) : 2;
    return v >= 1 && v <= 8 ? v : 2;
  }();
  const int cap = per_sm * sm_count() < 4096 ? per_sm * sm_count() : 4096;
  return g < cap ? (g < 1 ? 1 : g) : cap;
}

static int ensure_side_stream(acez_head_plan* h) {
  if (!h->side_ready) {
    ACEZ_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    for (int i = 0; i < h->L; ++i) ACEZ_CUDA(cudaEventCreateWithFlags(&h->ev_dz[i], cudaEventDisableTiming));
    ACEZ_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->side_ready = true;
  }
  return ACEZ_OK;
}

// tail (+ fc3 gradient) launch sequence shared by the training entries
static int launch_tail(acez_head_plan* h, TailArgs& t, int rows, cudaStream_t s, int* nonfinite, bool with_fc3_grad,
                       bool pdl) {
  if (!h->counters_zeroed) {  // once per plan: the completion counter is self-resetting afterwards
    ACEZ_CUDA(cudaMemsetAsync(h->BLKCOUNT, 0, 256, s));
    h->counters_zeroed = true;
    pdl = false;  // predecessor is a memset
  }
  t.blk_part = h->BLKPART;
  t.blk_count = h->BLKCOUNT;
  static const bool occ2 = [] {
    const char* e = getenv("ACEZ_TAIL_OCC2");
    return e != nullptr && atoi(e) != 0;
  }();
  int rc = occ2 ? launch_pdl(head_tail_kernel_occ2, dim3(tail_grid(rows)), dim3(kTailThreads), 0, s, pdl, t)
                : launch_pdl(head_tail_kernel, dim3(tail_grid(rows)), dim3(kTailThreads), 0, s, pdl, t);
  if (rc) return rc;
  if (with_fc3_grad) {
    const int nblk = (rows + kFc3Rows - 1) / kFc3Rows;
    float* gW3 = h->grads + (size_t)h->L * kLayerStride;
    cudaStream_t fs = s;
    bool first_pdl = true;
    if (h->fc3_overlap) {
      // fork: the two fc3 gradient kernels (10 us, 160 + 65 small CTAs) only need G3 and ACT[L] from the tail; they run on
      // the SMs the 80-CTA dgrad chain leaves idle and are joined at the end of launch_backward_gemms
      rc = ensure_side_stream(h);
      if (rc) return rc;
      ACEZ_CUDA(cudaEventRecord(h->ev_dz[0], s));
      ACEZ_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_dz[0], 0));
      fs = h->side_stream;
      first_pdl = false;  // predecessor on that stream is an event wait, not a kernel
      h->fc3_pending = true;
    }
    rc = launch_pdl(fc3_wgrad_partial_kernel, dim3(nblk), dim3(kFc3Threads), 0, fs, first_pdl, t.x, (const float*)h->G3, rows, h->FC3PART);
    if (rc) return rc;
    const int total = 4 * kC + 4;
    rc = launch_pdl(fc3_reduce_kernel, dim3((total * 8 + 255) / 256), dim3(256), 0, fs, true, (const float*)h->FC3PART, nblk, h->C3, gW3,
                    gW3 + (size_t)h->C3 * kC, nonfinite);
    if (rc) return rc;
  }
  return ACEZ_OK;
}

// dgrad chain + weight gradients. Overlapped mode: layer l's wgrad is enqueued on the plan's side stream as soon as
// DZ[l] exists (event after the kernel that produced it) and runs on the SMs the dgrad kernels leave idle; the main
// stream joins at the end. Works eagerly and under stream capture (fork / join through events).
static int launch_backward_gemms(acez_head_plan* h, cudaStream_t s, int* nonfinite) {
  const int L = h->L;
  if (h->use_chain && L >= 2) {
    h->chain_bwd.args.nonfinite = nonfinite;
    int rc = chain_launch(h->chain_bwd, s, /*pdl=*/!h->fc3_pending);  // predecessor: fc3_reduce_kernel (or the side-stream fork)
    if (rc) return rc;
    if (h->use_wgrad2) {
      h->wgrad2.args.nonfinite = nonfinite;
      rc = gemm2_launch(h->wgrad2, s, /*pdl=*/true);
    } else {
      h->wgrad.args.nonfinite = nonfinite;
      rc = gemm_launch(h->wgrad, s);
    }
    if (rc) return rc;
    if (h->fc3_pending) {  // join the fc3 gradient kernels forked in launch_tail
      ACEZ_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
      ACEZ_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
      h->fc3_pending = false;
    }
    return ACEZ_OK;
  }
  if (!h->overlap_wgrad) {
    for (int l = L - 1; l >= 1; --l) {
      h->dgrad[l].args.nonfinite = nonfinite;
      int rc = gemm_launch(h->dgrad[l], s);
      if (rc) return rc;
    }
    h->wgrad.args.nonfinite = nonfinite;  // fp16-overflow / inf check of the weight gradients in the epilogue
    return gemm_launch(h->wgrad, s);
  }
  if (!h->side_ready) {
    ACEZ_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    for (int i = 0; i < L; ++i) ACEZ_CUDA(cudaEventCreateWithFlags(&h->ev_dz[i], cudaEventDisableTiming));
    ACEZ_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->side_ready = true;
  }
  cudaStream_t side = h->side_stream;
  for (int l = L - 1; l >= 0; --l) {
    // DZ[l] is complete here (tail for l = L-1, dgrad l+1 otherwise)
    ACEZ_CUDA(cudaEventRecord(h->ev_dz[l], s));
    ACEZ_CUDA(cudaStreamWaitEvent(side, h->ev_dz[l], 0));
    h->wgrad_layer[l].args.nonfinite = nonfinite;
    int rc = gemm_launch(h->wgrad_layer[l], side, /*pdl=*/false);
    if (rc) return rc;
    if (l >= 1) {
      h->dgrad[l].args.nonfinite = nonfinite;
      rc = gemm_launch(h->dgrad[l], s, /*pdl=*/false);  // an event record sits between consecutive dgrad kernels
      if (rc) return rc;
    }
  }
  ACEZ_CUDA(cudaEventRecord(h->ev_join, side));
  ACEZ_CUDA(cudaStreamWaitEvent(s, h->ev_join, 0));
  return ACEZ_OK;
}

}  // namespace acez

// ----------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------
using namespace acez;

extern "C" size_t acez_head_param_count(const acez_head_config* cfg) {
  if (!cfg || cfg->num_res_blocks < 1) return 0;
  const int L = 3 * cfg->num_res_blocks + 2;
  const int C3 = cfg->use_homogeneous ? 4 : 3;
  return (size_t)L * kLayerStride + (size_t)C3 * kC + C3;
}

extern "C" size_t acez_head_workspace_bytes(const acez_head_config* cfg) {
  if (!cfg || cfg->num_res_blocks < 1 || cfg->max_rows < 1) return 0;
  return head_layout(*cfg).total + 1024;
}

extern "C" int acez_head_plan_create(const acez_head_config* cfg, float* params, float* grads, void* workspace,
                                     size_t workspace_bytes, acez_head_plan** out) {
  ACEZ_REQUIRE(cfg && params && workspace && out, "head_plan_create: null argument");
  ACEZ_REQUIRE(cfg->num_res_blocks >= 1 && cfg->num_res_blocks <= 16, "head_plan_create: num_res_blocks out of range");
  ACEZ_REQUIRE(cfg->max_rows >= 1, "head_plan_create: max_rows must be positive");
  ACEZ_REQUIRE(!cfg->training || grads != nullptr, "head_plan_create: training plan needs a gradient buffer");
  ACEZ_REQUIRE(workspace_bytes >= acez_head_workspace_bytes(cfg), "head_plan_create: workspace too small (%zu < %zu)",
               workspace_bytes, acez_head_workspace_bytes(cfg));
  acez_head_plan* h = new acez_head_plan();
  h->cfg = *cfg;
  h->nres = cfg->num_res_blocks;
  h->L = 3 * h->nres + 2;
  h->C3 = cfg->use_homogeneous ? 4 : 3;
  h->n_params = acez_head_param_count(cfg);
  h->params = params;
  h->grads = grads;
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  const HeadLayout lo = head_layout(*cfg);
  h->W16 = reinterpret_cast<__half*>(base + lo.w16);
  h->W3h = reinterpret_cast<__half*>(base + lo.w3h);
  h->ACT = reinterpret_cast<__half*>(base + lo.act);
  h->XTRA = cfg->training ? reinterpret_cast<__half*>(base + lo.xtra) : nullptr;
  h->DZ = cfg->training ? reinterpret_cast<__half*>(base + lo.dz) : nullptr;
  h->GRES = cfg->training ? reinterpret_cast<__half*>(base + lo.gres) : nullptr;
  h->MASKB = cfg->training ? base + lo.maskb : nullptr;
  h->G3 = cfg->training ? reinterpret_cast<float*>(base + lo.g3) : nullptr;
  h->FC3PART = cfg->training ? reinterpret_cast<float*>(base + lo.fc3part) : nullptr;
  h->BLKPART = reinterpret_cast<float*>(base + lo.blkpart);
  h->BLKCOUNT = reinterpret_cast<unsigned int*>(base + lo.blkpart + 4096 * 8 * sizeof(float));
  h->counters_zeroed = false;
  h->side_ready = false;
  {
    const char* e = getenv("ACEZ_WGRAD_OVERLAP");
    // measured on B200 (round 1): the 16-CTA per-layer kernels are bound by the per-SM L2 ingest rate (~20 us each)
    // and become the critical path (393 us / iteration vs 216 us batched), so the batched launch stays the default
    h->overlap_wgrad = (e == nullptr) ? 0 : atoi(e);
  }
  {
    // Default: all hidden layers of the forward / dgrad pass in one cluster kernel each (head_chain.cu); measured on
    // B200 (round 1, b = 5120): 0.212 ms per training iteration against 0.232 ms with one GEMM launch per layer.
    // ACEZ_HEAD_CHAIN=0 selects the per-layer tcgen05 GEMM path (also used when the head is deeper than the chain holds).
    const char* e = getenv("ACEZ_HEAD_CHAIN");
    h->use_chain = ((e == nullptr || atoi(e) != 0) && h->L <= kChainMaxSteps) ? 1 : 0;
  }
  {
    const char* e = getenv("ACEZ_WGRAD_2CTA");
    h->use_wgrad2 = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    const char* f = getenv("ACEZ_FC3_OVERLAP");   // only with the fused chain (the join sits in its branch of the backward)
    h->fc3_overlap = (f != nullptr && atoi(f) != 0 && h->use_chain) ? 1 : 0;
    h->fc3_pending = false;
  }
  h->act_stride = (size_t)cfg->max_rows * kC;
  h->prepared_rows = -1;
  h->prepared_training = 0;
  *out = h;
  return ACEZ_OK;
}

extern "C" void acez_head_plan_destroy(acez_head_plan* plan) {
  if (plan == nullptr) return;
  if (plan->side_ready) {
    cudaStreamDestroy(plan->side_stream);
    for (int i = 0; i < plan->L; ++i) cudaEventDestroy(plan->ev_dz[i]);
    cudaEventDestroy(plan->ev_join);
  }
  delete plan;
}

extern "C" int acez_head_sync_weights(acez_head_plan* h, acez_stream_t stream) {
  ACEZ_REQUIRE(h != nullptr, "head_sync_weights: null plan");
  int rc = acez_device_check();
  if (rc) return rc;
  cast_weights_kernel<<<4 * sm_count(), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(h->params, h->W16, h->W3h,
                                                                                          h->L, h->C3);
  ACEZ_CUDA(cudaGetLastError());
  return ACEZ_OK;
}

extern "C" void* acez_head_input_ptr(acez_head_plan* h) { return h ? h->ACT : nullptr; }

extern "C" int acez_head_plan_fused_chain(const acez_head_plan* h) { return (h != nullptr && h->use_chain) ? 1 : 0; }

extern "C" int acez_debug_chain_clocks(long long* host_out, size_t max_slots, int* n_ctas) {
  return chain_debug_read(host_out, max_slots, n_ctas);
}

static int head_run_forward(acez_head_plan* h, const void* features, int rows, int training, cudaStream_t s) {
  ACEZ_REQUIRE(rows >= 1 && rows <= h->cfg.max_rows, "head: rows=%d outside [1, %d]", rows, h->cfg.max_rows);
  int rc = head_prepare(h, rows, training);
  if (rc) return rc;
  if (features != nullptr && features != h->ACT)
    ACEZ_CUDA(cudaMemcpyAsync(h->ACT, features, (size_t)rows * kC * 2, cudaMemcpyDeviceToDevice, s));
  if (h->use_chain) return chain_launch(h->chain_fwd, s);
  for (int l = 0; l < h->L; ++l) {
    rc = gemm_launch(h->fwd[l], s, /*pdl=*/l > 0);  // the first kernel of the call follows a copy / foreign work
    if (rc) return rc;
  }
  return ACEZ_OK;
}

extern "C" int acez_head_forward(acez_head_plan* h, const void* features, int rows, float* sc_out,
                                 acez_stream_t stream) {
  ACEZ_REQUIRE(h != nullptr, "head_forward: null plan");
  int rc = acez_device_check();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  rc = head_run_forward(h, features, rows, 0, s);
  if (rc) return rc;
  if (sc_out == nullptr) return ACEZ_OK;  // GEMM chain only (profiling)
  TailArgs t{};
  fill_tail_common(h, rows, t);
  t.training = 0;
  t.sc_out = sc_out;
  return launch_tail(h, t, rows, s, nullptr, false, true);
}

extern "C" int acez_head_train_fwd_bwd(acez_head_plan* h, int rows, const acez_loss_params* lp,
                                       const acez_train_batch* b, float* stats, int* nonfinite,
                                       acez_stream_t stream) {
  ACEZ_REQUIRE(h && lp && b && stats && nonfinite, "head_train_fwd_bwd: null argument");
  ACEZ_REQUIRE(h->cfg.training && h->grads, "head_train_fwd_bwd: plan was not created for training");
  ACEZ_REQUIRE(b->target_px_b2 && b->K_b33 && b->Kinv_b33, "head_train_fwd_bwd: missing batch tensors");
  ACEZ_REQUIRE(b->P_b34 || (b->aug_inv_b34 && b->pose_inv_b44), "head_train_fwd_bwd: need P_b34 or aug_inv+pose_inv");
  ACEZ_REQUIRE(!lp->use_depth || b->target_crds_b3, "head_train_fwd_bwd: use_depth needs target_crds_b3");
  ACEZ_REQUIRE(lp->divisor > 0, "head_train_fwd_bwd: divisor must be positive");
  int rc = acez_device_check();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  rc = head_run_forward(h, b->features, rows, 1, s);
  if (rc) return rc;
  const int L = h->L;
  TailArgs t{};
  fill_tail_common(h, rows, t);
  t.training = 1;
  t.sc_out = b->sc_out_b3;
  t.lp = *lp;
  t.tpx = b->target_px_b2; t.Pin = b->P_b34; t.A = b->aug_inv_b34; t.T = b->pose_inv_b44;
  t.K = b->K_b33; t.Kinv = b->Kinv_b33; t.G = b->target_crds_b3;
  t.d_P = b->d_P_b34; t.d_Kdiag = b->d_Kdiag_b2;
  t.grad_scale_dev = b->grad_scale_dev;
  t.loss_weight_dev = b->loss_weight_dev;
  t.dz = h->DZ + (size_t)(L - 1) * h->act_stride;
  t.g3 = h->G3;
  t.stats = stats;
  t.nonfinite = nonfinite;
  rc = launch_tail(h, t, rows, s, nonfinite, true, true);
  if (rc) return rc;
  return launch_backward_gemms(h, s, nonfinite);
}

extern "C" int acez_head_backward(acez_head_plan* h, int rows, const float* d_sc_b3, int* nonfinite,
                                  acez_stream_t stream) {
  ACEZ_REQUIRE(h && d_sc_b3 && nonfinite, "head_backward: null argument");
  ACEZ_REQUIRE(h->cfg.training && h->grads, "head_backward: plan was not created for training");
  ACEZ_REQUIRE(h->prepared_rows == rows && h->prepared_training >= 1,
               "head_backward: call acez_head_forward_train with the same row count first");
  int rc = acez_device_check();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int L = h->L;
  TailArgs t{};
  fill_tail_common(h, rows, t);
  t.training = 2;
  t.lp.grad_scale = 1.f; t.lp.divisor = 1;
  t.d_sc_in = d_sc_b3;
  t.dz = h->DZ + (size_t)(L - 1) * h->act_stride;
  t.g3 = h->G3;
  t.stats = nullptr;
  t.nonfinite = nonfinite;
  rc = launch_tail(h, t, rows, s, nonfinite, true, false);  // first kernel of this call
  if (rc) return rc;
  return launch_backward_gemms(h, s, nonfinite);
}

extern "C" int acez_head_forward_train(acez_head_plan* h, const void* features, int rows, float* sc_out,
                                       acez_stream_t stream) {
  ACEZ_REQUIRE(h != nullptr && sc_out != nullptr, "head_forward_train: null argument");
  ACEZ_REQUIRE(h->cfg.training && h->grads, "head_forward_train: plan was not created for training");
  int rc = acez_device_check();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  rc = head_run_forward(h, features, rows, 1, s);
  if (rc) return rc;
  TailArgs t{};
  fill_tail_common(h, rows, t);
  t.training = 0;
  t.sc_out = sc_out;
  return launch_tail(h, t, rows, s, nullptr, false, true);
}

extern "C" int acez_gather_rows(const void* src, const int64_t* idx, int rows, int row_bytes, void* dst,
                                acez_stream_t stream) {
  ACEZ_REQUIRE(src && idx && dst && rows >= 0 && row_bytes > 0 && (row_bytes & 1) == 0, "gather_rows: bad arguments");
  int rc = acez_device_check();
  if (rc) return rc;
  if (rows == 0) return ACEZ_OK;
  const int threads = 256;
  const int grid = (rows * 32 + threads - 1) / threads;
  gather_rows_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint8_t*>(src), idx, rows, row_bytes, reinterpret_cast<uint8_t*>(dst));
  ACEZ_CUDA(cudaGetLastError());
  return ACEZ_OK;
}

extern "C" int acez_gather_rows_multi(const void* const* srcs, void* const* dsts, const int* row_bytes, int n_arrays,
                                      const int64_t* idx, int rows, acez_stream_t stream) {
  ACEZ_REQUIRE(srcs && dsts && row_bytes && idx && n_arrays >= 1 && n_arrays <= 8 && rows >= 0,
               "gather_rows_multi: bad arguments");
  MultiGather g{};
  for (int a = 0; a < n_arrays; ++a) {
    ACEZ_REQUIRE(srcs[a] && dsts[a] && row_bytes[a] > 0 && (row_bytes[a] & 1) == 0, "gather_rows_multi: bad array %d", a);
    g.src[a] = reinterpret_cast<const uint8_t*>(srcs[a]);
    g.dst[a] = reinterpret_cast<uint8_t*>(dsts[a]);
    g.row_bytes[a] = row_bytes[a];
  }
  int rc = acez_device_check();
  if (rc) return rc;
  if (rows == 0) return ACEZ_OK;
  const int threads = 256;
  dim3 grid((rows * 32 + threads - 1) / threads);
  return launch_pdl(gather_rows_multi_kernel, grid, dim3(threads), 0, reinterpret_cast<cudaStream_t>(stream), false, g, idx, rows,
                    n_arrays);
}

extern "C" int acez_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n,
                               const float* hyper_dev, float* scaler_state_dev, int* found_inf_dev, int use_scaler,
                               acez_head_plan* plan, acez_stream_t stream) {
  ACEZ_REQUIRE(params && grads && exp_avg && exp_avg_sq && hyper_dev && scaler_state_dev && found_inf_dev,
               "adamw_step: null argument");
  ACEZ_REQUIRE(plan == nullptr || (plan->params == params && plan->n_params == n),
               "adamw_step: plan does not own this parameter buffer");
  int rc = acez_device_check();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = 8 * sm_count();
  if (use_scaler == 1 || use_scaler == 3) {  // 2 = the caller's flag already covers every gradient (acez_head_train_fwd_bwd does)
    // 3 (data parallel, experimental): one more element behind the gradient is checked too - the slot in which the ranks'
    // local GradScaler flags travelled through the all-reduce (+inf when any rank overflowed)
    rc = launch_pdl(grad_check_kernel, dim3(grid), dim3(256), 0, s, false, grads, n + (use_scaler == 3 ? 1 : 0), found_inf_dev);
    if (rc) return rc;
  }
  rc = launch_pdl(adamw_kernel, dim3(grid), dim3(256), 0, s, false, params, grads, exp_avg, exp_avg_sq, n, hyper_dev,
                  scaler_state_dev, (const int*)found_inf_dev, use_scaler ? 1 : 0, plan ? plan->W16 : (__half*)nullptr,
                  plan ? plan->W3h : (__half*)nullptr, plan ? plan->L : 0, plan ? plan->C3 : 0);
  if (rc) return rc;
  return ACEZ_OK;
}
