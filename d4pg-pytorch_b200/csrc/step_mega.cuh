// Parameters of the persistent step kernel (step_mega.cu).
#pragma once
#include "gemm_ffma.cuh"
#include "replay_dev.cuh"

struct d4pg_replay;

namespace d4pg {

struct GemmBatchLite {           // GemmBatch without the TMA descriptors (FFMA path)
  GemmProblem p[GEMM_MAX_PROBLEMS];
  int n;
  int total_tiles;
};
constexpr int MEGA_MAX_LEVELS = 16;

struct MegaParams {
  SampleArgs sample;
  HeadsArgs heads; int heads_mode;
  GemmBatchLite level[MEGA_MAX_LEVELS]; int n_fwd, n_bwd;
  TreeArgs tree; int do_tree;
  AdamArgs adam;
  LearnerClock* clock;
  unsigned long long* barrier;
  unsigned long long* trace;       // optional phase stamps (D4PG_TC_TRACE)
};

int launch_step_mega(const MegaParams& p, cudaStream_t st);
void learner_sample_args(d4pg_replay* h, int B, int prioritized, const double* uniforms, const int32_t* positions,
                         uint64_t seed, LearnerClock* clock, const ClockParams& cp,
                         int32_t* idx, float* weights, float* s, float* a, double* r, float* s2, uint8_t* d,
                         int ld_obs, int ld_act, SampleArgs& sa);
void tree_update_args(d4pg_replay* h, int B, const int32_t* idx, const float* prio, TreeArgs& a);

}  // namespace d4pg
