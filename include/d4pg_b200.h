/*
 * d4pg_b200.h -- C ABI of libd4pg_sm100.so: the B200 (sm_100a) D4PG learner hot path.
 *
 * The reference (ajgupta93/d4pg-pytorch) is pure Python and has no FFI of its own; its
 * boundary for this path is the Python class API.  Each entry point
 * below replaces the *body* of one reference method; the Python classes in
 * `d4pg-pytorch_b200/` keep the reference signatures and bind these symbols with ctypes
 * (INTEGRATION.md shows the stub).  Citations are relative to /root/reference.
 *
 * Conventions
 *   - every function returns 0 on success or a negative D4PG_E* code; never throws.
 *     `d4pg_last_error()` returns a thread-local message for the last failure.
 *   - the CALLER owns all device memory (plain pointers + element counts); the library owns
 *     only opaque handles made by *_create and freed by *_destroy.
 *   - all work is asynchronous on the given `cudaStream_t` (passed as void*); no hidden
 *     synchronisation, no allocation after *_create.
 *   - a handle is not thread-safe; distinct handles are.
 *   - no torch types cross this boundary.
 */
#ifndef D4PG_B200_H_
#define D4PG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D4PG_OK            0
#define D4PG_EINVAL       -1   /* bad argument */
#define D4PG_ECUDA        -2   /* CUDA runtime error (see d4pg_last_error) */
#define D4PG_ENOTSUP      -3   /* unsupported configuration */
#define D4PG_ENCCL        -4   /* NCCL error / NCCL not loadable */
#define D4PG_ESTATE       -5   /* call out of order */

#define D4PG_HIDDEN      256   /* models.py:18-23,56-62 hard-code 256 hidden units */
#define D4PG_MAX_ATOMS   128

typedef void* d4pg_stream_t;   /* cudaStream_t */

const char* d4pg_last_error(void);
int32_t     d4pg_version(void);            /* 10000*major + 100*minor + patch */
/* sizeof of the structs that cross this ABI by pointer (0 d4pg_learner_config_t, 1 d4pg_learner_buffers_t,
 * 2 d4pg_net_layout_t; -1 otherwise): lets a binding verify that its mirror of the struct is current */
int32_t     d4pg_struct_size(int32_t which);
/* compute capability of the current device as 10*major+minor (100 on B200), or <0 */
int32_t     d4pg_device_sm(void);

/* ---------------------------------------------------------------------------------------
 * Parameter layout.  One flat fp32 buffer per network role, tensors in nn.Module order
 * fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc2_2.weight, fc2_2.bias, fc3.weight, fc3.bias
 * (models.py:18-23 actor, models.py:56-62 critic), nn.Linear row-major [out,in] with the row
 * PITCH rounded up to 4 floats (16-B aligned rows: every operand is TMA- and 128-bit addressable;
 * the pad columns are zero and stay zero), every tensor start aligned to 4 floats.
 * `offsets[8]` / `sizes[8]` (allocated floats, = out*pitch for weights) / `pitch[4]` are in floats.
 * ------------------------------------------------------------------------------------- */
typedef struct {
  int64_t offsets[8];
  int64_t sizes[8];
  int64_t pitch[4];
  int64_t total;      /* padded float count of the network */
} d4pg_net_layout_t;

int32_t d4pg_actor_layout(int32_t obs_dim, int32_t act_dim, d4pg_net_layout_t* out);
int32_t d4pg_critic_layout(int32_t obs_dim, int32_t act_dim, int32_t n_atoms, d4pg_net_layout_t* out);

/* ---------------------------------------------------------------------------------------
 * Fused projection + critic loss + TD proxy + priorities + logit gradients.
 * Replaces DDPG.reproject2 (ddpg.py:142-185, proj_mode 0) or DDPG.reproj_categorical_dist
 * (ddpg.py:122-140, proj_mode 1), the loss / td expressions ddpg.py:217,220-222,253 and the
 * policy-loss head ddpg.py:236-238.  One warp per batch row.
 *
 *   target_logits [B,N] f32  critic_target pre-softmax output for (s', actor_target(s'))
 *   q_logits      [B,N] f32  critic pre-softmax output for (s,a)
 *   pi_logits     [B,N] f32  critic pre-softmax output for (s, actor(s)); may be NULL
 *   rewards       [B]   f64, dones [B] u8
 *   discount            gamma (mode 0) or gamma**n_steps (mode 1)
 * outputs (any may be NULL):
 *   m [B,N] f32 projected target; bins_l/bins_u [B,N] i32 (the integer atom bins);
 *   target_probs, q_probs [B,N] f32; loss_rows [B] = -sum_j m log(q+1e-10);
 *   td [B] = -sum_j m q; prio [B] = |td| + prio_eps; dlogits_q [B,N] = d(mean loss)/d q_logits;
 *   pi_rows [B] = -sum_j softmax(pi)_j z_j; dlogits_pi [B,N] = d(mean pi loss)/d pi_logits.
 *   `grad_scale` multiplies both gradients (1/B for the reference's mean; 1/(B*world) under DP).
 *   `flags`: D4PG_PROJ_TARGET_IS_PROBS = target_logits already holds softmax outputs (the
 *   reference's reproject2(target_z_dist, ...) signature); D4PG_PROJ_Q_IS_PROBS likewise for q.
 * ------------------------------------------------------------------------------------- */
#define D4PG_PROJ_TARGET_IS_PROBS 1
#define D4PG_PROJ_Q_IS_PROBS      2
int32_t d4pg_proj_loss(const float* target_logits, const float* q_logits, const float* pi_logits,
                       const double* rewards, const uint8_t* dones,
                       int32_t B, int32_t N, double v_min, double v_max, double discount,
                       int32_t proj_mode, int32_t flags, double prio_eps, float grad_scale,
                       float* m, int32_t* bins_l, int32_t* bins_u,
                       float* target_probs, float* q_probs,
                       float* loss_rows, float* td, float* prio, float* dlogits_q,
                       float* pi_rows, float* dlogits_pi, d4pg_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Prioritized replay: GPU-resident sum/min segment trees + SoA transition storage.
 * Replaces SegmentTree / SumSegmentTree / MinSegmentTree (prioritized_replay_memory.py:33-162),
 * ReplayBuffer (:164-222) and PrioritizedReplayBuffer (:224-335).
 * Tree layout as the reference: root at 1, leaves at [cap, 2cap), fp32 nodes
 * (NumPy-2 semantics of the reference), cap = next pow2 >= size (:243-245).
 *
 * Caller-owned device buffers handed over at create time (sizes in elements):
 *   sum_tree, min_tree  f32 [2*cap]      obs, obs2 f32 [size*obs_dim]   act f32 [size*act_dim]
 *   rew f64 [size]                       done u8 [size]
 *   scratch i32 [cap]  (last-writer resolution for duplicate indices)
 *   state   f32 [8]    (device-resident scalars: max_priority, ...; opaque)
 * ------------------------------------------------------------------------------------- */
typedef struct d4pg_replay d4pg_replay_t;

int32_t d4pg_replay_capacity(int64_t size, int64_t* cap_out);          /* :243-245 */
int32_t d4pg_replay_create(int64_t size, int32_t obs_dim, int32_t act_dim, double alpha,
                           float* sum_tree, float* min_tree,
                           float* obs, float* act, double* rew, float* obs2, uint8_t* done,
                           int32_t* scratch, float* state,
                           d4pg_stream_t stream, d4pg_replay_t** out);
int32_t d4pg_replay_destroy(d4pg_replay_t* h);
int64_t d4pg_replay_len(const d4pg_replay_t* h);                        /* __len__, :177 */
int64_t d4pg_replay_next_idx(const d4pg_replay_t* h);

/* add() for n transitions already resident on the device (row-major [n,dim]); ring insert at
 * _next_idx, leaf = max_priority**alpha in both trees (:180-187,251-256).  `prioritized`=0
 * skips the trees (uniform Replay.add, replay_memory.py:14-19). */
int32_t d4pg_replay_add(d4pg_replay_t* h, int64_t n, const float* obs, const float* act,
                        const double* rew, const float* obs2, const uint8_t* done,
                        int32_t prioritized, d4pg_stream_t stream);

/* Host-side ingest: register a caller-owned PINNED host staging buffer and a device staging buffer
 * of `bytes` each (>= d4pg_replay_staging_bytes(rows)), then add() n <= rows transitions straight from
 * ordinary host arrays: the five arrays are packed into the pinned buffer, moved with ONE async H2D
 * copy and unpacked by the ring-write kernel.  Stream-ordered.  The buffers are used as TWO slots of bytes/2
 * (d4pg_replay_staging_bytes already counts both), alternating per call, so the host can stage add k+1 while
 * add k still waits on the device; a slot is re-used only after the add that read it has completed. */
int64_t d4pg_replay_staging_bytes(const d4pg_replay_t* h, int64_t rows);
int32_t d4pg_replay_set_staging(d4pg_replay_t* h, void* pinned_host, void* device, int64_t bytes);
int32_t d4pg_replay_add_host(d4pg_replay_t* h, int64_t n, const float* obs, const float* act, const double* rew,
                             const float* obs2, const uint8_t* done, int32_t prioritized, d4pg_stream_t stream);
/* Order stream `then` after everything enqueued so far on stream `first` (event record + wait; no-op if equal).  The
 * host mirror uses it to keep buffer operations issued on the caller's stream and on a learner's ingest stream
 * (d4pg_learner_ingest_stream) in program order. */
int32_t d4pg_replay_order_after(d4pg_replay_t* h, d4pg_stream_t first, d4pg_stream_t then);

/* Device-side ingest with n-step return accumulation at insert (replay_memory.py:38-45).  The arrays hold ONE episode
 * of T consecutive steps, resident on the device; transition i = (s_i, a_i, sum_{k<n} gamma^k r_{i+k}, s'_{i+n-1},
 * done_{i+n-1}) for i <= T-n is inserted (nothing when T < n, like the reference before step n-1).  The return is the
 * reference's left-to-right f64 loop, bit for bit.  `rew_scratch` f64 [T] is caller-owned device scratch.
 * d4pg_nstep_returns is the arithmetic alone: out[i], i <= T-n. */
int32_t d4pg_nstep_returns(const double* rew, int64_t T, int32_t n_steps, double gamma, double* out, d4pg_stream_t stream);
int32_t d4pg_replay_add_nstep(d4pg_replay_t* h, int64_t T, const float* obs, const float* act, const double* rew,
                              const float* obs2, const uint8_t* done, int32_t n_steps, double gamma,
                              double* rew_scratch, int32_t prioritized, d4pg_stream_t stream);

/* Hindsight-experience relabelling on the device (main.py:154-184, "future" strategy) as a gather kernel that produces
 * the rows d4pg_replay_add then inserts.  Episode of T goal-conditioned steps: obs / obs_next f32 [T, obs_dim], goal
 * f64 [T, goal_dim] (desired goal of every step), ag_next f64 [T, goal_dim] (achieved goal of the next state), act f32
 * [T, act_dim], rew f64 [T], done u8 [T].  select u8 [T] / future i32 [T] are the caller's draws
 * (np.random.uniform() < her_ratio; np.random.randint(t, T)), dst_row i32 [T] = exclusive prefix sum of (1 + select).
 * Output rows (state width obs_dim + goal_dim): the original transition of step t, then -- if selected -- its copy with
 * goal' = ag_next[future[t]], reward = -(||ag_next[t] - goal'||_2 > threshold) in f64 (the sparse gym-robotics
 * compute_reward) and done = (reward == 0).  her_action_mode 0 keeps the reference's behaviour of storing the rollout's
 * LAST action with the relabelled copy (main.py:184 uses `action`, not the step's `a`); 1 stores a_t. */
int32_t d4pg_her_relabel(int32_t T, int32_t obs_dim, int32_t goal_dim, int32_t act_dim,
                         const float* obs, const float* obs_next, const double* goal, const double* ag_next,
                         const float* act, const double* rew, const uint8_t* done,
                         const uint8_t* select, const int32_t* future, const int32_t* dst_row,
                         double threshold, int32_t her_action_mode,
                         float* out_s, float* out_a, double* out_r, float* out_s2, uint8_t* out_d,
                         d4pg_stream_t stream);

/* _sample_proportional + IS weights + _encode_sample (:258-313,189-199).
 *   uniforms [B] f64 in [0,1): the reference's random.random() draws; NULL = device Philox
 *   (seed, counter) stream.  mass = u * sum(0,len-1) with the reference's association and
 *   dtype rules (f64 while the tree is pristine, f32 afterwards).
 *   outputs: idx [B] i32, weights [B] f32 (may be NULL), gathered batch s,a,r,s2,done. */
int32_t d4pg_replay_sample(d4pg_replay_t* h, int32_t B, const double* uniforms,
                           uint64_t philox_seed, uint64_t philox_counter, double beta,
                           int32_t* idx, float* weights,
                           float* s, float* a, double* r, float* s2, uint8_t* done,
                           d4pg_stream_t stream);
/* uniform Replay.sample gather for caller-chosen positions (replay_memory.py:61-80) */
int32_t d4pg_replay_gather(d4pg_replay_t* h, int32_t B, const int32_t* idx,
                           float* s, float* a, double* r, float* s2, uint8_t* done,
                           d4pg_stream_t stream);
/* update_priorities (:315-335): leaf = prio**alpha (fp32 pow semantics), duplicates: last
 * writer wins, max_priority = max(max_priority, prio). */
int32_t d4pg_replay_update_priorities(d4pg_replay_t* h, int32_t B, const int32_t* idx,
                                      const float* prio, d4pg_stream_t stream);
/* SumSegmentTree.sum(start,end) / MinSegmentTree.min(start,end) over leaves [start,end) into
 * out[0], out[1] (device f32[2]) with SegmentTree.reduce's association; end<=0 counts from the
 * capacity as the reference's None/negative `end` does (:61-96,122-124,158-162). */
int32_t d4pg_replay_reduce(d4pg_replay_t* h, int64_t start, int64_t end, float* out, d4pg_stream_t stream);
/* SumSegmentTree.find_prefixsum_idx for n caller-supplied masses (:126-149) */
int32_t d4pg_replay_find_prefixsum(d4pg_replay_t* h, int32_t n, const double* masses, int32_t* idx,
                                   d4pg_stream_t stream);
/* raw leaf write + parent recompute for n (idx, value) pairs: SegmentTree.__setitem__ (:98-108) */
int32_t d4pg_replay_set_leaves(d4pg_replay_t* h, int32_t n, const int32_t* idx, const float* sum_vals,
                               const float* min_vals, d4pg_stream_t stream);
/* host-visible bookkeeping the drop-in needs after a host-side restore (stream-ordered like every other mutator) */
int32_t d4pg_replay_set_len(d4pg_replay_t* h, int64_t len, int64_t next_idx, int32_t pristine, d4pg_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Actor / critic forward (inference entry points).  Replace actor.forward (models.py:32-41)
 * and critic.forward (models.py:76-88).  `params` = flat buffer in d4pg_*_layout order.
 * `workspace` f32 [3*B*256] scratch.  precision: 0 fp32 (FFMA), 1 3xTF32 tcgen05 (fp32-accurate), 2 one TF32 tcgen05 pass.
 * ------------------------------------------------------------------------------------- */
int32_t d4pg_actor_forward(const float* params, int32_t obs_dim, int32_t act_dim,
                           const float* s, int32_t B, float* action, float* workspace,
                           int32_t precision, d4pg_stream_t stream);
int32_t d4pg_critic_forward(const float* params, int32_t obs_dim, int32_t act_dim, int32_t n_atoms,
                            const float* s, const float* a, int32_t B, float* probs, float* logits,
                            float* workspace, int32_t precision, d4pg_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused Adam + Polyak.  Replaces SharedAdam / torch.optim.Adam.step (shared_adam.py:3-17,
 * called at ddpg.py:232,244; torch-2.11 single-tensor formula), sync_local_global
 * (ddpg.py:118-120, identity on shared storage) and update_target_parameters (ddpg.py:110-116).
 *   p <- p - (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps);  target <- (1-tau)*target + tau*p
 * `step` is the post-increment step count.  `grad_scale` multiplies g first (DP averaging).
 * ------------------------------------------------------------------------------------- */
int32_t d4pg_adam_polyak(float* p, const float* g, float* m, float* v, float* target, int64_t n,
                         double lr, double beta1, double beta2, double eps, int64_t step,
                         double tau, float grad_scale, d4pg_stream_t stream);
/* update_target_parameters alone (ddpg.py:110-116): target <- (1-tau)*target + tau*src */
int32_t d4pg_polyak(float* target, const float* src, int64_t n, double tau, d4pg_stream_t stream);
/* hard_update (ddpg.py:92-94) / load_state_dict copies */
int32_t d4pg_copy_f32(float* dst, const float* src, int64_t n, d4pg_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * The learner: one DDPG.train() body (ddpg.py:200-255) per d4pg_learner_step call,
 * captured once into a CUDA graph and replayed.
 * ------------------------------------------------------------------------------------- */
typedef struct {
  int32_t obs_dim, act_dim, n_atoms, batch;
  double  v_min, v_max, gamma;
  int32_t n_steps;
  int32_t proj_mode;          /* 0 = reproject2 (live, discount gamma), 1 = n-step (gamma**n) */
  double  tau;
  double  lr_actor, lr_critic, beta1, beta2, adam_eps;
  int32_t prioritized;        /* 1 = PrioritizedReplayBuffer path, 0 = uniform Replay path */
  double  per_beta0, per_beta_final; int64_t per_beta_iters;   /* LinearSchedule, ddpg.py:81-86 */
  double  prio_eps;           /* ddpg.py:87 */
  int32_t precision;          /* 0 exact fp32 FFMA, 1 3xTF32 tcgen05 (hi/lo split, fp32-accurate: meets the 1e-5 parity bar),
                                 2 one TF32 tcgen05 pass (not parity-grade) */
  int32_t sample_mode;        /* 0 = caller uniforms/positions (parity), 1 = device Philox */
  uint64_t philox_seed;
  int32_t world_size;         /* >1: gradients are averaged over ranks before Adam */
  int32_t use_graph;          /* 1 = capture the step into a CUDA graph */
  int32_t loss_flags;         /* corrected-semantics switches, 0 = reference behaviour:
                                 1 = importance-weighted critic CE (the reference samples the weights but
                                     ignores them, ddpg.py:217), 2 = priority = CE_i + eps instead of
                                     |sum_j m_ij q_ij| + eps (ddpg.py:221-222,253), 4 = the actor gradient flows
                                     through the critic AFTER this step's critic update (the reference uses the stale
                                     pre-update local copy, ddpg.py:229-247); needs the tcgen05 chain plan, one GPU */
  int32_t chain;              /* step plan of the MLP passes (batches above 512 rows always use plan 0): 0 = one grouped launch per dependency
                                 level (18 kernels/step); 1 = cluster-fused layer chains: forward passes, dX passes
                                 and all dW are ONE launch each (7 kernels/step; precision 0: FFMA tiles, bit-identical
                                 to plan 0; precision 1/2: tcgen05 tiles, 64-row clusters, pre-packed hi/lo weight images) */
  int32_t prefetch;           /* 1 (sample_mode 1 only): step t samples batch t+1 on a side branch, right after its own
                                 priorities are in the trees, while its backward pass and Adam still run.  Same
                                 Philox counters and the same trees as sampling at the start of step t+1, so results
                                 are identical; any replay mutation by the caller (add / set / update) between two
                                 steps discards the prefetched batch and step t+1 samples again at its start.
                                 With sample_mode 0 (host-drawn uniforms / positions) and use_graph: the HOST pipeline --
                                 d4pg_learner_step_host* samples batch k on the library's ingest stream
                                 (d4pg_learner_ingest_stream), behind the add()s the caller issued on that stream and
                                 gated on step k-1's priority write-back, while step k-1's backward pass, dW and Adam
                                 still run.  Tree operations keep the reference's order update(k-1) -> add(k) ->
                                 sample(k) (ddpg.py:200-255 + main.py's add loop): results are identical */
} d4pg_learner_config_t;

/* Caller-owned device buffers.  P_a / P_c = d4pg_*_layout().total. */
typedef struct {
  float* actor;  float* actor_target;  float* critic;  float* critic_target;
  float* grad_actor;  float* grad_critic;          /* contiguous: grad_critic == grad_actor + P_a */
  float* adam_m_actor; float* adam_v_actor; float* adam_m_critic; float* adam_v_critic;
  double*   uniforms;      /* [B] f64 (sample_mode 0, prioritized) */
  int32_t*  positions;     /* [B] i32 (sample_mode 0, uniform replay) */
  int32_t*  idx;           /* [B] i32 out: sampled indices */
  float*    weights;       /* [B] f32 out: IS weights (unused by the loss) */
  float*    prio;          /* [B] f32 out: new priorities */
  float*    td;            /* [B] f32 out */
  float*    losses;        /* [4]  f32 out: critic loss, actor loss, reserved, reserved */
  float*    workspace;     /* f32 [d4pg_learner_workspace_floats()] */
} d4pg_learner_buffers_t;

typedef struct d4pg_learner d4pg_learner_t;
typedef struct d4pg_comm    d4pg_comm_t;

int64_t d4pg_learner_workspace_floats(const d4pg_learner_config_t* cfg);
int32_t d4pg_learner_create(const d4pg_learner_config_t* cfg, const d4pg_learner_buffers_t* buf,
                            d4pg_replay_t* replay, d4pg_comm_t* comm, d4pg_learner_t** out);
int32_t d4pg_learner_destroy(d4pg_learner_t* h);
/* One gradient step.  Everything is stream-ordered; results land in buf->losses etc. */
int32_t d4pg_learner_step(d4pg_learner_t* h, d4pg_stream_t stream);
/* Host-facing step: everything `DDPG.train()` needs per call in ONE library call.
 *   d4pg_learner_step_host: order the learner stream after `caller_stream`, copy this step's host inputs
 *     (uniforms f64[B] for prioritized replay / positions i32[B] for uniform replay; NULL with device-side sampling)
 *     to the device, run the step on `learner_stream`, order `caller_stream` after it.  The inputs may live in
 *     ordinary host memory and may be reused as soon as the call returns: they are staged in library-owned pinned
 *     buffers (allocated by d4pg_learner_create, double-buffered by step parity, a buffer is rewritten only after the
 *     H2D copy out of it has completed).
 *   d4pg_learner_step_host_mt: the same for prioritized replay, taking the 2*B raw 32-bit MT19937 outputs of
 *     `random.randbytes(8*B)` instead of B doubles: uniform i = ((w[2i] >> 5) * 2^26 + (w[2i+1] >> 6)) / 2^53, which is
 *     CPython's random.random() -- the draws of prioritized_replay_memory.py:262, same generator state afterwards.
 *   d4pg_learner_read_losses: D2H of {critic loss, actor loss, -, -} and wait for it (the step's result). */
int32_t d4pg_learner_step_host(d4pg_learner_t* h, const double* uniforms, const int32_t* positions,
                               d4pg_stream_t caller_stream, d4pg_stream_t learner_stream);
int32_t d4pg_learner_step_host_mt(d4pg_learner_t* h, const uint32_t* mt_words,
                                  d4pg_stream_t caller_stream, d4pg_stream_t learner_stream);
int32_t d4pg_learner_read_losses(d4pg_learner_t* h, float* out4, d4pg_stream_t learner_stream);
/* Every d4pg_learner_step_host* call also queues an async D2H copy of its {critic loss, actor loss, -, -} into a pinned
 * two-slot ring.  d4pg_learner_fetch_losses waits for and returns the result of the most recent host-facing step (lag 0)
 * or of the one before it (lag 1: usually complete already, so a caller can read step k-1 while step k runs). */
int32_t d4pg_learner_fetch_losses(d4pg_learner_t* h, int32_t lag, float* out4);
/* n_steps back-to-back gradient steps without returning to the caller in between (device-side
 * sampling keeps advancing; caller-supplied uniforms/positions would be reused). */
int32_t d4pg_learner_run(d4pg_learner_t* h, int32_t n_steps, d4pg_stream_t stream);
/* Named intermediate (for parity tests): returns device pointer + element count.
 * names: "s","a","r","s2","done","target_logits","q_logits","pi_logits","m","q_probs",
 *        "target_probs","dlogits_q","dlogits_pi","actor_out","loss_rows","pi_rows"
 * 2-D tensors are stored with a row pitch of `*ld` floats (>= the logical width). */
int32_t d4pg_learner_tensor(d4pg_learner_t* h, const char* name, void** ptr, int64_t* count, int32_t* ld);
/* One EAGER (non-graph) step with a CUDA-event pair around every launch; synchronises.
 * ms_out[i] = device time of launch i, names_out[i*name_stride] = its launcher name. */
int32_t d4pg_learner_profile_step(d4pg_learner_t* h, d4pg_stream_t stream, int32_t max_launches,
                                  float* ms_out, char* names_out, int32_t name_stride, int32_t* n_out);
int64_t d4pg_learner_steps_done(const d4pg_learner_t* h);
/* host pipeline (cfg.prefetch with sample_mode 0): the stream buffer adds should be issued on so they overlap the
 * running step (NULL when the pipeline is off).  Owned by the learner.  With the pipeline on, d4pg_learner_step_host*
 * samples on this stream and does NOT wait for caller_stream first (that stream is ordered after the whole previous
 * step, which would serialise the pipeline): a caller that touched the buffer on another stream (add, update_priorities,
 * set_leaves ...) calls d4pg_replay_order_after(replay, that_stream, ingest_stream) before the next step. */
void* d4pg_learner_ingest_stream(const d4pg_learner_t* h);
/* The tcgen05 plans consume pre-split hi/lo weight IMAGES that the library's Adam / Polyak kernel keeps current.  Every
 * CUDA-graph step that samples in the graph re-packs them from the fp32 parameters first (any external write is picked
 * up); the host pipeline's steps (above) re-pack only after this call -- make it whenever actor / critic / target
 * parameters were written from outside the library (state_dict load, hard update, manual edits) since the last step.
 * A new learner starts "changed". */
int32_t d4pg_learner_weights_changed(d4pg_learner_t* h);
int32_t d4pg_learner_kernels_per_step(const d4pg_learner_t* h);
/* restore the optimiser step counters / beta-schedule clock (checkpoint resume) */
int32_t d4pg_learner_set_counters(d4pg_learner_t* h, int64_t adam_step, int64_t beta_t, d4pg_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Data-parallel communicator (one process per GPU).  The reference has no collective (its
 * multi-worker mode is Hogwild over shared CPU memory, main.py:394-405, ddpg.py:104-108);
 * the B200 build is synchronous DP: one all-reduce of the flat [P_a+P_c] gradient per step.
 * NCCL is resolved at run time (dlopen of the torch-bundled libnccl.so.2).
 * ------------------------------------------------------------------------------------- */
int32_t d4pg_comm_unique_id(uint8_t* id128);                       /* ncclGetUniqueId, 128 bytes */
int32_t d4pg_comm_create(const uint8_t* id128, int32_t rank, int32_t world, d4pg_comm_t** out);
int32_t d4pg_comm_destroy(d4pg_comm_t* c);
int32_t d4pg_comm_allreduce_sum(d4pg_comm_t* c, float* buf, int64_t n, d4pg_stream_t stream);
/* Fused gradient all-reduce over peer memory (one node, <= 8 ranks) instead of the NCCL kernel: every rank allocates
 * an exchange block ([2][n_floats] gradient halves + flags) and exports it with CUDA IPC (64-byte handle); after the
 * handles of all ranks were gathered (rank order) every rank maps them.  A learner created with such a communicator
 * writes its dW into its own half, a flag barrier orders the ranks, and the fused Adam kernel sums all ranks' halves
 * (fixed rank order: replicas stay bit-identical) while it updates the parameters. */
int32_t d4pg_comm_peer_alloc(d4pg_comm_t* c, int64_t n_floats, uint8_t* handle64);
int32_t d4pg_comm_peer_open(d4pg_comm_t* c, const uint8_t* all_handles /* world x 64 bytes */);
int32_t d4pg_comm_peer_ready(const d4pg_comm_t* c);
int32_t d4pg_comm_peer_disable(d4pg_comm_t* c);      /* collective decision: fall back to the NCCL all-reduce */

/* In-switch gradient reduction (NVLS): ONE multicast object spans every rank's [2][n] gradient buffer; the fused Adam
 * kernel then reads the sum over all ranks with multimem.ld_reduce (the NVSwitch adds) -- one NVLink hop, n floats inbound
 * per rank whatever the rank count.  Collective setup, driven by the host binding after d4pg_comm_peer_open:
 *   every rank: d4pg_comm_mc_supported;  rank 0: d4pg_comm_mc_create -> POSIX file descriptor, passed to the other ranks
 *   (SCM_RIGHTS over a Unix socket) which d4pg_comm_mc_import it;  every rank: d4pg_comm_mc_add_device;  barrier;  every
 *   rank: d4pg_comm_mc_bind (allocates its buffer, binds it, maps the unicast and the multicast view);  barrier.
 * d4pg_comm_mc_selftest copies `src` (n floats, device) into this rank's buffer and/or writes the switch-reduced sum over
 * all ranks to `out`; the binding uses it to check that every rank reads bit-identical sums before enabling the path. */
int32_t d4pg_comm_mc_supported(d4pg_comm_t* c);
int32_t d4pg_comm_mc_create(d4pg_comm_t* c, int32_t* fd_out);
int32_t d4pg_comm_mc_import(d4pg_comm_t* c, int32_t fd);
int32_t d4pg_comm_mc_add_device(d4pg_comm_t* c);
int32_t d4pg_comm_mc_bind(d4pg_comm_t* c);
int32_t d4pg_comm_mc_ready(const d4pg_comm_t* c);
int32_t d4pg_comm_mc_disable(d4pg_comm_t* c);
int32_t d4pg_comm_mc_selftest(d4pg_comm_t* c, const float* src, float* out, int64_t n, d4pg_stream_t stream);

/* Debug: %globaltimer (ns) phase stamps written by CTA 0 of the most recent tcgen05 GEMM launch when
 * the environment variable D4PG_TC_TRACE is set (out16 = 32 x uint64, host memory). */
int32_t d4pg_debug_tc_trace(unsigned long long* out16);
/* first n (<= 512) stamps of the same buffer: the chain kernels write 6-8 per layer slot of one CTA */
int32_t d4pg_debug_trace_read(unsigned long long* out, int32_t n);
/* Watchdog record of the tcgen05 chain kernels (16 x uint64, host memory): every mbarrier wait inside them is bounded;
 * a wait that times out traps the launch and leaves {1, code | slot<<8 | rank<<16 | parity<<24 | block<<32, aux, ...}
 * in host-mapped memory, readable here even after the failed launch invalidated the context; words [4 + 2k, 5 + 2k] hold
 * the first timed-out wait of kind k = 1..5 (loader: ring buffer free / accumulator done, MMA: weights / A chunk landed,
 * epilogue: accumulator done).  All zero = never fired. */
int32_t d4pg_debug_watchdog(unsigned long long* out16);

#ifdef __cplusplus
}
#endif
#endif /* D4PG_B200_H_ */
