// Device-side training schedule (see schedule.cu): the per-iteration evaluation as a device function, so that it can run as
// its own one-thread kernel or ride in the first block of the batch-gather kernel.
#pragma once
#include "common.cuh"

namespace acez {

// state layout (floats; integers stored exactly as floats < 2^24)
enum : int { S_ITER = 0, S_STEPS = 1, S_COOL = 2, S_COOL_START = 3, S_MAXIT = 4, S_RING_N = 5, S_RING_POS = 6, S_DONE = 7,
             S_LR = 8, S_LW = 9, S_RING = 16 };

__device__ __forceinline__ double sched_one_cycle(double max_lr, double total, double s) {
  // torch.optim.lr_scheduler.OneCycleLR(max_lr, total_steps, cycle_momentum=False): pct_start 0.3, cos, div 25, final div 1e4
  const double initial = max_lr / 25.0, minimum = initial / 1e4;
  const double up_end = 0.3 * total - 1.0, down_end = total - 1.0;
  if (s > total - 1.0) s = total - 1.0;
  const double pi = 3.14159265358979323846;
  if (s <= up_end) return max_lr + (initial - max_lr) / 2.0 * (cos(pi * (s / up_end)) + 1.0);
  return minimum + (max_lr - minimum) / 2.0 * (cos(pi * ((s - up_end) / (down_end - up_end))) + 1.0);
}

// One thread: books the previous iteration's inlier count, runs check_and_set_cooldown, publishes lr / loss weight, advances
// the iteration counter (ace_schedule.py:72-126, ace_loss.py:53-69).
static __device__ __noinline__ void schedule_step_device(const acez_schedule_params& p, float* __restrict__ st,
                                                  const float* __restrict__ inlier_count, float* __restrict__ hyper) {
  const int it = (int)st[S_ITER];
  int steps = (int)st[S_STEPS];
  // scheduler.step() of the PREVIOUS iteration (ace_schedule.py:115-126): the batch-inlier fraction enters the ring
  if (it > 0 && p.kind != ACEZ_SCHED_CONSTANT) {
    steps += 1;
    if (p.kind == ACEZ_SCHED_1CYCLEPOLY) {
      const int pos = (int)st[S_RING_POS];
      st[S_RING + pos] = inlier_count[0] / (float)p.batch_global;
      st[S_RING_POS] = (float)((pos + 1) % ACEZ_SCHED_RING);
      const int n = (int)st[S_RING_N];
      if (n < ACEZ_SCHED_RING) st[S_RING_N] = (float)(n + 1);
    }
  }
  int in_cool = (int)st[S_COOL];
  int max_it = (int)st[S_MAXIT];
  int cool_start = (int)st[S_COOL_START];
  // check_and_set_cooldown(iteration)   (ace_schedule.py:72-101)
  if (p.kind == ACEZ_SCHED_1CYCLEPOLY && !in_cool && it >= p.warmup_iterations) {
    const bool by_duration = it >= max_it - p.cooldown_iterations;
    const int n = (int)st[S_RING_N];
    float mn = 3.0e38f;
    for (int k = 0; k < n; ++k) mn = fminf(mn, st[S_RING + k]);
    const bool dynamic = n > 0 && mn > p.cooldown_trigger;
    if (by_duration || dynamic) {
      max_it = it + p.cooldown_iterations;
      in_cool = 1;
      cool_start = steps;
    }
  }
  const bool done = it >= max_it;   // ace_trainer.py:509: the step returns without training
  double lr;
  if (p.kind == ACEZ_SCHED_CONSTANT) lr = p.lr_min;
  else if (p.kind == ACEZ_SCHED_CIRCLE) lr = sched_one_cycle(p.lr_max, (double)p.iterations, (double)steps);
  else if (!in_cool) {   // LinearLR warm-up: start_factor = warmup_lr / lr_max over warmup_iterations steps
    const double f0 = (double)p.warmup_lr / (double)p.lr_max;
    const int s = steps < p.warmup_iterations ? steps : p.warmup_iterations;
    lr = (double)p.lr_max * (f0 + (1.0 - f0) * (double)s / (double)p.warmup_iterations);
  } else {               // LinearLR cool-down: 1 -> lr_min / lr_max over cooldown_iterations steps
    int k = steps - cool_start;
    if (k > p.cooldown_iterations) k = p.cooldown_iterations;
    const double f1 = (double)p.lr_min / (double)p.lr_max;
    lr = (double)p.lr_max * (1.0 + (f1 - 1.0) * (double)k / (double)p.cooldown_iterations);
  }
  // loss weight of this iteration (ace_loss.py:53-69)
  double lw = p.soft_clamp;
  if (p.loss_dyntanh) {
    double w = (double)it / (double)p.iterations;
    if (p.loss_schedule_circle) w = 1.0 - sqrt(fmax(0.0, 1.0 - w * w));
    lw = (1.0 - w) * p.soft_clamp + p.soft_clamp_min;
  }
  hyper[0] = done ? 0.f : (float)lr;
  hyper[5] = (float)lw;
  st[S_LR] = (float)lr;
  st[S_LW] = (float)lw;
  st[S_STEPS] = (float)steps;
  st[S_COOL] = (float)in_cool;
  st[S_COOL_START] = (float)cool_start;
  st[S_MAXIT] = (float)max_it;
  st[S_DONE] = done ? 1.f : 0.f;
  if (!done) st[S_ITER] = (float)(it + 1);
}

}  // namespace acez
