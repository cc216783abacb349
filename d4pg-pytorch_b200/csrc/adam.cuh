// Fused Adam + Polyak and the device-side step clock.
#pragma once
#include "common.cuh"

namespace d4pg {

// Per-step scalars produced on the device so a captured CUDA graph needs no host patching.
struct LearnerClock {
  int64_t adam_step;     // post-increment count used by this step's Adam (state['step'])
  int64_t beta_t;        // LinearSchedule.t
  int64_t steps_done;    // Philox counter / bookkeeping
  float beta;            // PER beta for this step's IS weights
  float neg_step_size[2];   // -(lr/bc1) for actor, critic
  float bc2_sqrt;        // sqrt(1 - beta2^step)
  float pad;
};

struct AdamSeg {
  float* p; const float* g; float* m; float* v; float* target; int64_t n;
  float neg_step_size; int clock_slot;        // clock_slot >= 0: read -step_size from the device clock
};
struct AdamArgs {
  AdamSeg seg[2]; int nseg;
  float w1, w2, beta2, eps, bc2_sqrt, tau, one_minus_tau, grad_scale;
  const LearnerClock* clock;                  // optional
};
int launch_adam(const AdamArgs& a, cudaStream_t st);

struct ClockArgs {
  LearnerClock* clock;
  double lr_actor, lr_critic, beta1, beta2;
  double per_beta0, per_beta_final; int64_t per_beta_iters;
};
int launch_clock(const ClockArgs& a, cudaStream_t st);
int launch_loss_reduce(const float* loss_rows, const float* pi_rows, int B, float inv_count, float* out, cudaStream_t st);

}  // namespace d4pg
