// Point-cloud export metrics (reference ace_vis_util.py:431-592, get_point_cloud_from_network): for every cell of a batch of
// predicted scene-coordinate maps, in ONE pass: the L1 reprojection error against the cell's pixel under the mapping pose
// (:489-503, same clamp of the projective depth at 0.1), the camera depth (:528) and the local scene-coordinate gradient
// max(|dX/dx|, |dX/dy|) with the reference's reflect padding (:506-515). The reference runs ~15 ATen kernels per image for
// these; the selection logic that follows (threshold ladder, top-k relaxation) works on the three maps.
#include "common.cuh"

namespace acez {

__global__ void pointcloud_metrics_kernel(const float* __restrict__ sc, int n, int h, int w, const float* __restrict__ pose_inv,
                                          const float* __restrict__ K, int sub, float* __restrict__ err, float* __restrict__ grad,
                                          float* __restrict__ depth) {
  const int cells = h * w;
  const long long total = (long long)n * cells;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / cells), c = (int)(i % cells);
    const int x = c % w, y = c / w;
    const float* s = sc + (size_t)img * 3 * cells;
    const float X = s[c], Y = s[cells + c], Z = s[2 * cells + c];
    const float* P = pose_inv + (size_t)img * 12;   // world -> camera, rows of [R | t]
    const float* Km = K + (size_t)img * 9;
    const float cx = P[0] * X + P[1] * Y + P[2] * Z + P[3];
    const float cy = P[4] * X + P[5] * Y + P[6] * Z + P[7];
    const float cz = P[8] * X + P[9] * Y + P[10] * Z + P[11];
    const float px = Km[0] * cx + Km[1] * cy + Km[2] * cz;
    const float py = Km[3] * cx + Km[4] * cy + Km[5] * cz;
    const float pz = fmaxf(Km[6] * cx + Km[7] * cy + Km[8] * cz, 0.1f);   // clamp_(min=0.1), :492
    const float u = px / pz, v = py / pz;
    const float tx = (float)sub * ((float)x + 0.5f), ty = (float)sub * ((float)y + 0.5f);   // ace_util.get_pixel_grid
    err[i] = fabsf(u - tx) + fabsf(v - ty);
    depth[i] = cz;
    // gradient to the left / upper neighbour; the first column / row mirrors the SECOND difference (F.pad(..., 'reflect'))
    auto diff = [&](int c0, int c1) {
      const float dx = s[c0] - s[c1], dy = s[cells + c0] - s[cells + c1], dz = s[2 * cells + c0] - s[2 * cells + c1];
      return sqrtf(dx * dx + dy * dy + dz * dz);
    };
    float gx = 0.f, gy = 0.f;
    if (w >= 3) { const int xx = x >= 1 ? x : 2; gx = diff(y * w + xx, y * w + xx - 1); }
    else if (w == 2) gx = diff(y * w + 1, y * w);
    if (h >= 3) { const int yy = y >= 1 ? y : 2; gy = diff(yy * w + x, (yy - 1) * w + x); }
    else if (h == 2) gy = diff(w + x, x);
    grad[i] = fmaxf(gx, gy);
  }
}

}  // namespace acez

extern "C" int acez_pointcloud_metrics(const float* sc, int n, int h, int w, const float* pose_inv_n34, const float* K_n33,
                                       int subsample, float* err, float* grad, float* depth, acez_stream_t stream) {
  ACEZ_REQUIRE(sc && pose_inv_n34 && K_n33 && err && grad && depth, "pointcloud_metrics: null argument");
  ACEZ_REQUIRE(n >= 0 && h > 0 && w > 0 && subsample >= 1, "pointcloud_metrics: bad shape");
  int rc = acez_device_check();
  if (rc) return rc;
  if (n == 0) return ACEZ_OK;
  const long long total = (long long)n * h * w;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = 16LL * acez::sm_count();
  if (blocks > cap) blocks = cap;
  acez::pointcloud_metrics_kernel<<<(int)blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      sc, n, h, w, pose_inv_n34, K_n33, subsample, err, grad, depth);
  ACEZ_CUDA(cudaGetLastError());
  return ACEZ_OK;
}
