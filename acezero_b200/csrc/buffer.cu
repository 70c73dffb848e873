// Patch-buffer fill (reference ace_trainer.py:381-436): for the S cells `torch.multinomial` drew for one image, write
// all 8 arrays of the training buffer in ONE launch — gather the 512-d feature rows, recompute the pixel targets from
// the cell index (the `get_pixel_grid` table of ace_util.py:7-13 is 8*(x+0.5), 8*(y+0.5)), broadcast the image's
// poses / intrinsics, gather the ground-truth coordinates. The reference does this with ~12 small kernels per image
// (expand / reshape / 8 index ops / 8 slice copies).
#include "common.cuh"

namespace acez {

struct FillArgs {
  const __half* feat;   // [cells, 512] NHWC rows of this image
  const int64_t* idx;   // [S] sampled cells (row-major y * w + x)
  int S, w, cells, subsample;
  const float* mats;    // 50 floats: aug_inv 3x4 (12) | pose_inv 4x4 (16) | K 3x3 (9) | Kinv 3x3 (9) | pad
  const float* crds;    // [3, cells] planar ground-truth coordinates or nullptr
  int pose_idx;
  long long row0;       // first destination row
  __half* d_feat; float* d_px; float* d_aug; float* d_pose; float* d_K; float* d_Kinv; float* d_crds; int16_t* d_pidx;
};

__global__ void buffer_fill_kernel(const FillArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= a.S) return;
  const long long cell = a.idx[warp];
  const long long dst = a.row0 + warp;
  // features: 1 KB row, 32 lanes x 2 x 16 B
  const uint4* s = reinterpret_cast<const uint4*>(a.feat + cell * 512);
  uint4* d = reinterpret_cast<uint4*>(a.d_feat + dst * 512);
  d[lane] = s[lane];
  d[lane + 32] = s[lane + 32];
  const int x = (int)(cell % a.w), y = (int)(cell / a.w);
  if (lane < 2) a.d_px[dst * 2 + lane] = (float)a.subsample * ((lane == 0 ? (float)x : (float)y) + 0.5f);
  if (lane < 12) a.d_aug[dst * 12 + lane] = a.mats[lane];
  if (lane < 16) a.d_pose[dst * 16 + lane] = a.mats[12 + lane];
  if (lane < 9) { a.d_K[dst * 9 + lane] = a.mats[28 + lane]; a.d_Kinv[dst * 9 + lane] = a.mats[37 + lane]; }
  if (lane < 3) a.d_crds[dst * 3 + lane] = a.crds ? a.crds[(long long)lane * a.cells + cell] : 0.f;
  if (lane == 0) a.d_pidx[dst] = (int16_t)a.pose_idx;
}

}  // namespace acez

extern "C" int acez_buffer_fill(const void* feat_rows, const int64_t* sample_idx, int n_samples, int map_w, int cells,
                                int subsample, const float* mats46, const float* target_crds_3hw, int pose_idx,
                                long long row0, void* d_features, float* d_target_px, float* d_aug_inv, float* d_pose_inv,
                                float* d_K, float* d_Kinv, float* d_target_crds, int16_t* d_pose_idx,
                                acez_stream_t stream) {
  using namespace acez;
  ACEZ_REQUIRE(feat_rows && sample_idx && mats46 && d_features && d_target_px && d_aug_inv && d_pose_inv && d_K &&
                   d_Kinv && d_target_crds && d_pose_idx,
               "buffer_fill: null argument");
  ACEZ_REQUIRE(n_samples >= 0 && map_w > 0 && cells > 0 && subsample > 0 && row0 >= 0, "buffer_fill: bad sizes");
  int rc = acez_device_check();
  if (rc) return rc;
  if (n_samples == 0) return ACEZ_OK;
  FillArgs a{};
  a.feat = reinterpret_cast<const __half*>(feat_rows);
  a.idx = sample_idx; a.S = n_samples; a.w = map_w; a.cells = cells; a.subsample = subsample;
  a.mats = mats46; a.crds = target_crds_3hw; a.pose_idx = pose_idx; a.row0 = row0;
  a.d_feat = reinterpret_cast<__half*>(d_features);
  a.d_px = d_target_px; a.d_aug = d_aug_inv; a.d_pose = d_pose_inv; a.d_K = d_K; a.d_Kinv = d_Kinv;
  a.d_crds = d_target_crds; a.d_pidx = d_pose_idx;
  const int threads = 256;
  buffer_fill_kernel<<<(n_samples * 32 + threads - 1) / threads, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ACEZ_CUDA(cudaGetLastError());
  return ACEZ_OK;
}
