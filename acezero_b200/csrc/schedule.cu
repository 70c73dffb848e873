// Device-side training schedule: learning rate, loss weight and the cool-down trigger of the reference's ScheduleACE
// (ace_schedule.py:8-126) and ReproLoss weight (ace_loss.py:53-69), evaluated by ONE thread at the start of every iteration
// inside the iteration's CUDA graph. Nothing about the schedule travels host -> device per iteration any more, and the
// `1cyclepoly` schedule (the ACE0 default, ace_zero.py:105) needs no per-iteration read-back of the inlier fraction: the
// ring of the last 100 batch-inlier fractions, the trigger test `min(ring) > threshold` and the shortened
// `max_iterations = iteration + cooldown_iterations` (ace_schedule.py:72-101) live in device memory; the host polls the state
// with a bounded lag only to know when to stop. Once `iteration >= max_iterations` the kernel publishes lr = 0 (AdamW then
// leaves the weights bit-unchanged: p - 0 * (...)), so iterations the host enqueued beyond the end are no-ops.
#include "schedule.cuh"

namespace acez {

__global__ void schedule_kernel(const acez_schedule_params p, float* __restrict__ st, const float* __restrict__ inlier_count,
                                float* __restrict__ hyper) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  pdl_wait();
  schedule_step_device(p, st, inlier_count, hyper);
}

}  // namespace acez

using namespace acez;

extern "C" int acez_schedule_init(const acez_schedule_params* p, float* state_dev, acez_stream_t stream) {
  ACEZ_REQUIRE(p != nullptr && state_dev != nullptr, "schedule_init: null argument");
  ACEZ_REQUIRE(p->iterations > 0 && p->iterations < (1 << 24), "schedule_init: iterations out of range");
  float h[ACEZ_SCHED_STATE_FLOATS] = {0};
  h[S_MAXIT] = (float)p->iterations;
  ACEZ_CUDA(cudaMemcpyAsync(state_dev, h, sizeof(h), cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream)));
  ACEZ_CUDA(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));   // `h` lives on this stack frame
  return ACEZ_OK;
}

extern "C" int acez_schedule_step(const acez_schedule_params* p, float* state_dev, const float* inlier_count_dev,
                                  float* hyper_dev, acez_stream_t stream) {
  ACEZ_REQUIRE(p && state_dev && inlier_count_dev && hyper_dev, "schedule_step: null argument");
  ACEZ_REQUIRE(p->kind >= ACEZ_SCHED_CONSTANT && p->kind <= ACEZ_SCHED_1CYCLEPOLY, "schedule_step: unknown schedule %d", p->kind);
  ACEZ_REQUIRE(p->batch_global > 0, "schedule_step: batch_global must be positive");
  int rc = acez_device_check();
  if (rc) return rc;
  static bool hinted = false;
  if (!hinted) {   // same shared-memory carve-out as the big kernels of the iteration (see acez_head_plan_create)
    cudaFuncSetAttribute(schedule_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
    hinted = true;
  }
  schedule_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p, state_dev, inlier_count_dev, hyper_dev);
  ACEZ_CUDA(cudaGetLastError());
  return ACEZ_OK;
}
