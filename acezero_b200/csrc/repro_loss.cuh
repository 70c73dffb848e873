// Per-row reprojection loss + analytic backward, shared by the stand-alone kernel (repro_loss.cu) and the fused
// head tail kernel (head.cu). Follows reference ace_trainer.py:530-613 and ace_loss.py:39-90 step by step, fp32.
#pragma once
#include "common.cuh"

namespace acez {

struct RowLoss {
  float loss;     // un-normalised per-row loss term
  float gX[3];    // dL/dX   (already multiplied by grad_scale / divisor)
  float gc[3];    // dL/dc   (camera coordinates), same scaling; zero for depth-mode invalid rows
  float gK00, gK11;
  bool valid;
  bool inlier;
};

// P = A(3x4) * T(4x4)   (ace_trainer.py:530, torch.bmm)
__device__ __forceinline__ void compose_pose(const float* __restrict__ A, const float* __restrict__ T, float P[12]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fmaf(A[i * 4 + k], T[k * 4 + j], s);
      P[i * 4 + j] = s;
    }
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__device__ __forceinline__ void repro_row(const acez_loss_params& lp, const float X[3], const float P[12],
                                          const float K[9], const float Kinv[9], float tx, float ty,
                                          const float* G /* nullable: GT coords */, RowLoss& o) {
  const float gs = lp.grad_scale / (float)lp.divisor;
  // camera coordinates  (ace_trainer.py:533)
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = fmaf(P[i * 4 + 0], X[0], fmaf(P[i * 4 + 1], X[1], fmaf(P[i * 4 + 2], X[2], P[i * 4 + 3])));
  // homogeneous pixel (ace_trainer.py:538/540)
  float p[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = fmaf(K[i * 3 + 0], c[0], fmaf(K[i * 3 + 1], c[1], K[i * 3 + 2] * c[2]));
  const bool z_pass = p[2] >= lp.depth_min;       // clamp_(min) passes gradient where input >= min
  const float z = fmaxf(p[2], lp.depth_min);      // :545
  const float u = p[0] / z, v = p[1] / z;         // :548
  const float ex = u - tx, ey = v - ty;           // :551
  const float r = fabsf(ex) + fabsf(ey);          // :552 (L1 norm)
  bool invalid = (c[2] < lp.depth_min) | (r > lp.hard_clamp) | (c[2] > lp.depth_max);  // :558-565
  bool avail = false;
  float dist = 0.f;
  if (lp.use_depth && G != nullptr) {             // :567-574
    const float dx = G[0] - X[0], dy = G[1] - X[1], dz = G[2] - X[2];
    dist = sqrtf(dx * dx + dy * dy + dz * dz);
    avail = (fabsf(G[0]) + fabsf(G[1]) + fabsf(G[2])) > 0.00001f;
    invalid |= (dist > 0.1f) & avail;
  }
  o.valid = !invalid;
  o.inlier = o.valid && (r < lp.inlier_px);       // :585
  o.gK00 = 0.f; o.gK11 = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) { o.gX[i] = 0.f; o.gc[i] = 0.f; }
  if (o.valid) {
    float g_r;
    const float w = lp.loss_weight;
    if (lp.loss_type == ACEZ_LOSS_TANH) {         // ace_loss.py:7-8,53-69
      const float t = tanhf(r / w);
      o.loss = w * t;
      g_r = 1.f - t * t;
    } else if (r <= w) {                          // ace_loss.py:72-90, small-error branch
      o.loss = r;
      g_r = 1.f;
    } else if (lp.loss_type == ACEZ_LOSS_L1) {
      o.loss = 0.f;
      g_r = 0.f;
    } else if (lp.loss_type == ACEZ_LOSS_L1_SQRT) {
      const float s = sqrtf(w * r);
      o.loss = s;
      g_r = 0.5f * w / s;
    } else {
      o.loss = logf(1.f + w * r);
      g_r = w / (1.f + w * r);
    }
    g_r *= gs;
    const float gu = g_r * sgn(ex), gv = g_r * sgn(ey);
    float gp[3];
    gp[0] = gu / z;
    gp[1] = gv / z;
    gp[2] = z_pass ? -(gu * p[0] + gv * p[1]) / (z * z) : 0.f;
    o.gK00 = gp[0] * c[0];
    o.gK11 = gp[1] * c[1];
#pragma unroll
    for (int j = 0; j < 3; ++j) o.gc[j] = K[0 * 3 + j] * gp[0] + K[1 * 3 + j] * gp[1] + K[2 * 3 + j] * gp[2];
  } else if (!lp.use_depth) {
    // constant-depth proxy target (ace_trainer.py:592-600)
    float l = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float t = lp.depth_target * fmaf(Kinv[i * 3 + 0], tx, fmaf(Kinv[i * 3 + 1], ty, Kinv[i * 3 + 2]));
      const float d = t - c[i];
      l += fabsf(d);
      o.gc[i] = -sgn(d) * gs;
    }
    o.loss = l;
  } else {
    // GT-coordinate loss on invalid rows that have a target (ace_trainer.py:602-609)
    if (avail) {
      o.loss = dist;
      if (dist > 0.f) {
        const float inv = gs / dist;
        o.gX[0] = (X[0] - G[0]) * inv;
        o.gX[1] = (X[1] - G[1]) * inv;
        o.gX[2] = (X[2] - G[2]) * inv;
      }
    } else {
      o.loss = 0.f;
    }
  }
  // dL/dX += R^T dL/dc
#pragma unroll
  for (int j = 0; j < 3; ++j) o.gX[j] += P[0 * 4 + j] * o.gc[0] + P[1 * 4 + j] * o.gc[1] + P[2 * 4 + j] * o.gc[2];
}

}  // namespace acez
