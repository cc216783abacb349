"""`DDPG` with the reference's constructor and method names (ddpg.py:15-255); the body of
`train()` is one `d4pg_learner_step` call into libd4pg_sm100.so (a CUDA graph of hand-written
sm_100a kernels), not Python/NumPy/ATen.

Reference behaviours kept on purpose, each switchable only explicitly:
  * importance weights are sampled but NOT used by the loss                (ddpg.py:217)
  * priority = |sum_j m_ij q_ij| + 1e-6, not a KL/CE                       (ddpg.py:221-222,253)
  * the live projection discounts with gamma even when n_steps > 1         (ddpg.py:155);
    `projection="nstep"` selects the gamma**n variant (ddpg.py:122-140)
  * the actor gradient uses the critic weights from BEFORE this step's critic update
  * Adam betas (0.9, 0.9) come from SharedAdam; DDPG's own lr_actor/lr_critic optimisers are
    constructed but never stepped (ddpg.py:67-68)
"""
import ctypes as C
import random

import numpy as np
import torch

from . import _lib
from .models import actor, critic
from .prioritized_replay_memory import LinearSchedule, PrioritizedReplayBuffer
from .random_process import GaussianNoise
from .replay_memory import Replay
from .shared_adam import SharedAdam
from .utils import default_device


class _Learner(object):
    """Owns the C learner handle and the device buffers handed to it."""

    def __init__(self, ddpg, global_model):
        _lib.require_cuda()
        L = _lib.lib()
        dev = ddpg.device
        g = global_model
        opt_a, opt_c = ddpg.optimizer_global_actor, ddpg.optimizer_global_critic
        lr_a, b1, b2, eps = opt_a.hyper()
        lr_c, b1c, b2c, epsc = opt_c.hyper()
        if (b1, b2, eps) != (b1c, b2c, epsc):
            raise _lib.D4PGError("actor and critic SharedAdam must share betas/eps")
        # local networks alias the global storage (what ddpg.py:104-108,118-120 converge to)
        if g is not ddpg:
            ddpg.actor.adopt_flat(g.actor.flat_params())
            ddpg.critic.adopt_flat(g.critic.flat_params())
        Pa, Pc = ddpg.actor._total, ddpg.critic._total
        self.grads = torch.zeros(Pa + Pc, dtype=torch.float32, device=dev)
        for net, view in ((ddpg.actor, self.grads[:Pa]), (ddpg.critic, self.grads[Pa:])):
            net._flat_grad = view
            net._bind_grads()
        ma, va = opt_a.moments(g.actor)
        mc, vc = opt_c.moments(g.critic)
        B = ddpg.batch_size
        cfg = _lib.LearnerConfig()
        cfg.obs_dim, cfg.act_dim, cfg.n_atoms, cfg.batch = ddpg.obs_dim, ddpg.act_dim, ddpg.n_atoms, B
        cfg.v_min, cfg.v_max, cfg.gamma = float(ddpg.v_min), float(ddpg.v_max), float(ddpg.gamma)
        cfg.n_steps = int(ddpg.n_steps)
        cfg.proj_mode = 1 if ddpg.projection == "nstep" else 0
        cfg.tau = float(ddpg.tau)
        cfg.lr_actor, cfg.lr_critic, cfg.beta1, cfg.beta2, cfg.adam_eps = lr_a, lr_c, b1, b2, eps
        cfg.prioritized = 1 if ddpg.prioritized_replay else 0
        if ddpg.prioritized_replay:
            sch = ddpg.beta_schedule
            cfg.per_beta0, cfg.per_beta_final, cfg.per_beta_iters = sch.initial_p, sch.final_p, sch.schedule_timesteps
            cfg.prio_eps = ddpg.prioritized_replay_eps
        else:
            cfg.per_beta0, cfg.per_beta_final, cfg.per_beta_iters, cfg.prio_eps = 1.0, 1.0, 1, 1e-6
        cfg.precision = {"fp32": 0, "tf32x3": 1, "tf32": 2}[ddpg.precision]
        cfg.sample_mode = 0 if ddpg.sampling == "reference" else 1
        cfg.philox_seed = int(ddpg.philox_seed)
        cfg.world_size = ddpg.comm.world_size if ddpg.comm is not None else 1
        cfg.use_graph = 1 if ddpg.use_graph else 0
        cfg.loss_flags = ((1 if ddpg.importance_weighted else 0) | (2 if ddpg.priority == "ce" else 0) |
                          (4 if ddpg.actor_critic == "post_update" else 0))
        # plan 1: fp32 = FFMA chain tiles; tf32x3 / tf32 = tcgen05 chain tiles (mlp_tc_chain.cu); plan 0: one launch per level
        cfg.chain = {"levels": 0, "cluster": 1, False: 0, True: 1, 0: 0, 1: 1}[ddpg.chain]
        # device sampling: step t samples batch t+1 on a side branch; reference sampling (host-drawn uniforms): the host
        # pipeline -- train() samples batch k on the learner's ingest stream, behind the add()s issued there, while step
        # k-1 still runs (needs the CUDA-graph step)
        cfg.prefetch = 1 if (ddpg.prefetch and (cfg.sample_mode == 1 or ddpg.use_graph)) else 0
        self.cfg = cfg
        nws = L.d4pg_learner_workspace_floats(C.byref(cfg))
        f32 = torch.float32
        self.workspace = torch.zeros(nws, dtype=f32, device=dev)
        self.uniforms = torch.zeros(B, dtype=torch.float64, device=dev)
        self.positions = torch.zeros(B, dtype=torch.int32, device=dev)
        self.idx = torch.zeros(B, dtype=torch.int32, device=dev)
        self.weights = torch.zeros(B, dtype=f32, device=dev)
        self.prio = torch.zeros(B, dtype=f32, device=dev)
        self.td = torch.zeros(B, dtype=f32, device=dev)
        self.losses = torch.zeros(4, dtype=f32, device=dev)
        buf = _lib.LearnerBuffers()
        buf.actor, buf.actor_target = g.actor.flat_params().data_ptr(), ddpg.actor_target.flat_params().data_ptr()
        buf.critic, buf.critic_target = g.critic.flat_params().data_ptr(), ddpg.critic_target.flat_params().data_ptr()
        buf.grad_actor = self.grads.data_ptr()
        buf.grad_critic = self.grads.data_ptr() + 4 * Pa
        buf.adam_m_actor, buf.adam_v_actor = ma.data_ptr(), va.data_ptr()
        buf.adam_m_critic, buf.adam_v_critic = mc.data_ptr(), vc.data_ptr()
        buf.uniforms, buf.positions = self.uniforms.data_ptr(), self.positions.data_ptr()
        buf.idx, buf.weights = self.idx.data_ptr(), self.weights.data_ptr()
        buf.prio, buf.td, buf.losses = self.prio.data_ptr(), self.td.data_ptr(), self.losses.data_ptr()
        buf.workspace = self.workspace.data_ptr()
        self._keep = (ma, va, mc, vc, g.actor.flat_params(), g.critic.flat_params(),
                      ddpg.actor_target.flat_params(), ddpg.critic_target.flat_params())
        store = ddpg.replayBuffer._store
        if store.handle is None:
            raise _lib.D4PGError("train() called before any transition was added to the replay buffer")
        store.flush()
        h = C.c_void_p()
        if ddpg.comm is not None and cfg.world_size > 1:
            ddpg.comm.setup_peers(Pa + Pc)            # fused all-reduce over peer memory (collective call)
        comm = ddpg.comm.handle if ddpg.comm is not None else None
        _lib.check(L.d4pg_learner_create(C.byref(cfg), C.byref(buf), store.handle, comm, C.byref(h)), "d4pg_learner_create")
        self.handle = h
        self.global_model = g
        self.stream = torch.cuda.Stream(device=dev)      # graph capture needs a non-default stream
        self.dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        self.stream_ptr = C.c_void_p(self.stream.cuda_stream)
        self.step_host = L.d4pg_learner_step_host
        self.step_host_mt = L.d4pg_learner_step_host_mt
        self.read_losses = L.d4pg_learner_read_losses
        self.losses_out = (C.c_float * 4)()
        self.fresh_host_step = False          # the most recent step was a train() (its losses are in the pinned ring)
        self._store = store
        # parameter writes torch can see (load_state_dict, hard_update, optimizer steps: in-place ops bump the version counter
        # the views share with the flat buffer) are reported to the library by train(); see DDPG.weights_changed
        self._flats = (g.actor.flat_params(), g.critic.flat_params(), ddpg.actor_target.flat_params(), ddpg.critic_target.flat_params())
        self.seen_versions = None
        self.weights_changed = L.d4pg_learner_weights_changed
        ing = L.d4pg_learner_ingest_stream(h)
        store.attach_ingest_stream(int(ing) if ing else None)   # host pipeline: add_batch_host goes to the ingest stream
        if opt_a.step_count or (ddpg.prioritized_replay and ddpg.beta_schedule.t):
            _lib.check(L.d4pg_learner_set_counters(h, opt_a.step_count,
                                                   ddpg.beta_schedule.t if ddpg.prioritized_replay else 0,
                                                   _lib.stream_ptr()), "d4pg_learner_set_counters")

    def tensor(self, name, dtype=torch.float32):
        """Copy of a named intermediate.  2-D planes come back as [rows, pitch] (pitch >= width)."""
        p, n, ld = C.c_void_p(), C.c_int64(), C.c_int32()
        _lib.check(_lib.lib().d4pg_learner_tensor(self.handle, name.encode(), C.byref(p), C.byref(n), C.byref(ld)),
                   "d4pg_learner_tensor")
        esize = {torch.float32: 4, torch.float64: 8, torch.uint8: 1}[dtype]
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.uint8: "|u1"}[dtype]

        class _Arr(object):
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": typestr, "data": (int(p.value), False),
                                        "version": 2, "strides": (esize,)}
        t = torch.as_tensor(_Arr(), device=self.workspace.device).clone()
        return t.view(-1, ld.value) if ld.value > 1 else t

    def close(self):
        if self.handle is not None:
            if getattr(self, "_store", None) is not None:
                self._store.attach_ingest_stream(None)            # joins the ingest stream first
                self._store = None
            _lib.lib().d4pg_learner_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DDPG:
    replayBuffer = None

    def __init__(self, obs_dim, act_dim, env=None, memory_size=50000, batch_size=64,
                 lr_critic=1e-4, lr_actor=1e-4, gamma=0.99, tau=0.001, prioritized_replay=True,
                 critic_dist_info=None, n_steps=1,
                 # ---- B200 build extensions (keyword-only in spirit; reference callers never pass them)
                 device=None, sampling="reference", projection="reference", precision="fp32",
                 use_graph=True, philox_seed=0, comm=None, chain="cluster", prefetch=True, track_weights=True,
                 importance_weighted=False, priority="reference", actor_critic="reference"):
        self.gamma = gamma
        self.n_steps = n_steps
        self.n_step_gamma = self.gamma ** self.n_steps
        self.batch_size = batch_size
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.memory_size = memory_size
        self.tau = tau
        self.env = env
        self.device = torch.device(device) if device is not None else default_device()
        assert sampling in ("reference", "device") and projection in ("reference", "nstep")
        assert precision in ("fp32", "tf32x3", "tf32")
        self.sampling, self.projection, self.precision = sampling, projection, precision
        self.use_graph, self.philox_seed, self.comm = use_graph, philox_seed, comm
        # step plan of the MLP passes: "cluster" (default) cluster-fused layer chains (exact FFMA tiles for fp32,
        # tcgen05 tiles for tf32x3 / tf32), "levels" one launch per dependency level
        self.chain = chain
        # device-side sampling only: step t already samples batch t+1 behind its own backward pass (identical results)
        self.prefetch = prefetch
        self.track_weights = track_weights
        # corrected-semantics switches (default = the reference's behaviour)
        assert priority in ("reference", "ce") and actor_critic in ("reference", "post_update")
        self.importance_weighted, self.priority = bool(importance_weighted), priority
        # "post_update": the actor gradient flows through the critic AFTER this step's critic update (corrected SURVEY.md
        # H7); "reference": through the stale pre-update copy, as ddpg.py:229-247 does.  tcgen05 chain plan, one GPU.
        self.actor_critic = actor_critic

        self.dist_type = critic_dist_info["type"]
        if self.dist_type != "categorical":
            raise NotImplementedError("only the categorical critic exists (the reference's "
                                      "mixture_of_gaussian branch is a TODO stub, ddpg.py:48-50)")
        self.v_min = critic_dist_info["v_min"]
        self.v_max = critic_dist_info["v_max"]
        self.n_atoms = critic_dist_info["n_atoms"]
        self.delta = (self.v_max - self.v_min) / float(self.n_atoms - 1)
        self.bin_centers = np.array([self.v_min + i * self.delta for i in range(self.n_atoms)]).reshape(-1, 1)

        # networks, built in the reference's order so a seeded RNG yields the same weights (ddpg.py:56-64)
        self.actor = actor(input_size=obs_dim, output_size=act_dim, device=self.device)
        self.actor_target = actor(input_size=obs_dim, output_size=act_dim, device=self.device)
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.critic = critic(state_size=obs_dim, action_size=act_dim, dist_info=critic_dist_info, device=self.device)
        self.critic_target = critic(state_size=obs_dim, action_size=act_dim, dist_info=critic_dist_info, device=self.device)
        self.critic_target.load_state_dict(self.critic.state_dict())

        # constructed for attribute compatibility; like the reference's, never stepped by train()
        self.optimizer_actor = SharedAdam(self.actor.parameters(), lr=lr_actor, betas=(0.9, 0.999))
        self.optimizer_critic = SharedAdam(self.critic.parameters(), lr=lr_critic, betas=(0.9, 0.999))
        self.optimizer_global_actor = None
        self.optimizer_global_critic = None

        self.noise = GaussianNoise(dimension=act_dim, num_epochs=5000)                       # ddpg.py:75

        self.prioritized_replay = prioritized_replay
        if self.prioritized_replay:                                                          # ddpg.py:78-87
            self.replayBuffer = PrioritizedReplayBuffer(self.memory_size, alpha=0.6, obs_dim=obs_dim,
                                                        act_dim=act_dim, device=self.device)
            self.beta_schedule = LinearSchedule(100000, initial_p=0.4, final_p=1.0)
            self.prioritized_replay_eps = 1e-6
        else:
            self.replayBuffer = Replay(self.memory_size, self.env, n_steps=self.n_steps, gamma=self.gamma,
                                       obs_dim=obs_dim, act_dim=act_dim, device=self.device)
        self._learner = None

    # ---- reference plumbing methods ------------------------------------------------------
    def hard_update(self):                                                                   # ddpg.py:92-94
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.critic_target.load_state_dict(self.critic.state_dict())

    def share_memory(self):                                                                  # ddpg.py:96-98
        self.actor.share_memory()
        self.critic.share_memory()

    def assign_global_optimizer(self, optimizer_global_actor, optimizer_global_critic):      # ddpg.py:100-102
        self.optimizer_global_actor = optimizer_global_actor
        self.optimizer_global_critic = optimizer_global_critic
        self._drop_learner()

    def copy_gradients(self, model_local, model_global):                                     # ddpg.py:104-108
        if model_global.flat_params().data_ptr() == model_local.flat_params().data_ptr():
            return                                    # shared storage: gradients are already "global"
        model_global._flat_grad = model_local.flat_grads()
        model_global._bind_grads()

    def weights_changed(self):
        """Report a parameter write the learner cannot see.  train() notices every in-place torch operation on actor /
        critic / target parameters (load_state_dict, hard_update, sync_local_global, `with torch.no_grad(): p.copy_(..)`)
        through the tensors' version counters; writes through `p.data` or raw pointers bypass those counters -- call this
        after them (or construct with track_weights=False: the weight images are then rebuilt on every step)."""
        if self._learner is not None and self._learner.handle is not None:
            self._learner.seen_versions = None

    def update_target_parameters(self):                                                      # ddpg.py:110-116
        _lib.require_cuda()
        self.weights_changed()
        for tgt, src in ((self.actor_target, self.actor), (self.critic_target, self.critic)):
            _lib.check(_lib.lib().d4pg_polyak(_lib.ptr(tgt.flat_params()), _lib.ptr(src.flat_params()),
                                              tgt._total, float(self.tau), _lib.stream_ptr()), "d4pg_polyak")

    def sync_local_global(self, global_model):                                               # ddpg.py:118-120
        self.actor.load_state_dict(global_model.actor.state_dict())
        self.critic.load_state_dict(global_model.critic.state_dict())

    # ---- projections as standalone methods (numpy in / numpy out, computed on the GPU) --------
    def _project(self, target_z_dist, rewards, terminates, mode):
        _lib.require_cuda()
        p = torch.as_tensor(np.ascontiguousarray(target_z_dist, dtype=np.float32)).to(self.device)
        B, N = p.shape
        r = torch.as_tensor(np.asarray(rewards, dtype=np.float64).reshape(-1)).to(self.device)
        d = torch.as_tensor(np.asarray(terminates).reshape(-1).astype(bool).astype(np.uint8)).to(self.device)
        m = torch.empty(B, N, dtype=torch.float32, device=self.device)
        disc = self.n_step_gamma if mode == 1 else self.gamma
        _lib.check(_lib.lib().d4pg_proj_loss(_lib.ptr(p), _lib.ptr(p), None, _lib.ptr(r), _lib.ptr(d), B, N,
                                             float(self.v_min), float(self.v_max), float(disc), mode,
                                             _lib.PROJ_TARGET_IS_PROBS | _lib.PROJ_Q_IS_PROBS, 1e-6, 1.0 / B,
                                             _lib.ptr(m), None, None, None, None, None, None, None, None, None, None,
                                             _lib.stream_ptr()), "d4pg_proj_loss")
        return m.cpu().numpy()

    def reproject2(self, target_z_dist, rewards, terminates):                                # ddpg.py:142-185
        return self._project(target_z_dist, rewards, terminates, 0)

    def reproj_categorical_dist(self, target_z_dist, rewards, terminates):                   # ddpg.py:122-140
        return self._project(target_z_dist, rewards, terminates, 1).astype(np.float64)

    # ---- sampling -------------------------------------------------------------------------
    def sample(self, batch_size=None):                                                       # ddpg.py:187-197
        weights = None
        batch_idxes = None
        if self.prioritized_replay:
            experience = self.replayBuffer.sample(batch_size, beta=self.beta_schedule.value())
            (states, actions, rewards, next_states, terminates, weights, batch_idxes) = experience
        else:
            states, actions, rewards, next_states, terminates = self.replayBuffer.sample(self.batch_size)
        return states, actions, rewards, next_states, terminates, weights, batch_idxes

    # ---- the hot path ----------------------------------------------------------------------
    def _drop_learner(self):
        if self._learner is not None:
            self._learner.close()
            self._learner = None

    def _get_learner(self, global_model):
        if self._learner is None or self._learner.global_model is not global_model:
            self._drop_learner()
            if self.optimizer_global_actor is None or self.optimizer_global_critic is None:
                raise _lib.D4PGError("call assign_global_optimizer(SharedAdam, SharedAdam) before train() "
                                     "(main.py:194)")
            self._learner = _Learner(self, global_model)
        return self._learner

    def train(self, global_model=None):
        """One learner gradient step (ddpg.py:200-255), asynchronous on the learner's stream.
        Results (losses, td, priorities, sampled indices) stay on the device; read them with
        `last_losses()` / `last_batch_info()`."""
        g = global_model if global_model is not None else self
        L = self._learner
        if L is None or L.global_model is not g:
            L = self._get_learner(g)
        store = self.replayBuffer._store
        if store._n_staged:
            store.flush()
        f = L._flats
        v = (f[0]._version, f[1]._version, f[2]._version, f[3]._version) if self.track_weights else None
        if v != L.seen_versions or v is None:
            L.weights_changed(L.handle)
            L.seen_versions = v
        B = self.batch_size
        # one library call: order after the caller's stream, H2D of this step's host inputs, the step's
        # CUDA graph on the learner stream, order the caller's stream after it
        if self.sampling == "reference" and self.prioritized_replay:
            # B x random.random() in order (prioritized_replay_memory.py:262): the 2*B raw MT19937 words are drawn in
            # one call -- same generator state afterwards -- and turned into the doubles by the library
            rc = L.step_host_mt(L.handle, random.randbytes(8 * B), _lib.raw_stream(L.dev_index), L.stream_ptr)
        elif self.sampling == "reference":
            pos = np.ascontiguousarray(self.replayBuffer.sample_positions(B), dtype=np.int32)
            rc = L.step_host(L.handle, None, pos.ctypes.data, _lib.raw_stream(L.dev_index), L.stream_ptr)
        else:
            rc = L.step_host(L.handle, None, None, _lib.raw_stream(L.dev_index), L.stream_ptr)
        if rc:
            _lib.check(rc, "d4pg_learner_step_host")
        L.fresh_host_step = True
        if self.prioritized_replay:
            self.beta_schedule.t += 1
        for opt in (self.optimizer_global_actor, self.optimizer_global_critic):
            opt.step_count += 1

    def train_n(self, n, global_model=None):
        """`n` gradient steps in one C call (device-side sampling only): the learner's CUDA graph is
        replayed back to back with no Python in between."""
        if self.sampling != "device":
            raise _lib.D4PGError("train_n needs sampling='device' (host-drawn uniforms are per-step inputs)")
        g = global_model if global_model is not None else self
        L = self._get_learner(g)
        self.replayBuffer._store.flush()
        L.stream.wait_stream(torch.cuda.current_stream())
        _lib.check(_lib.lib().d4pg_learner_run(L.handle, int(n), C.c_void_p(L.stream.cuda_stream)), "d4pg_learner_run")
        L.fresh_host_step = False
        torch.cuda.current_stream().wait_stream(L.stream)
        if self.prioritized_replay:
            self.beta_schedule.t += n
        for opt in (self.optimizer_global_actor, self.optimizer_global_critic):
            opt.step_count += n

    def last_losses(self, lag=0):
        """(critic_loss, actor_loss) of the most recent train() -- waits for that step's result (a 16-byte D2H copy every
        train() queues).  `lag=1` returns the result of the train() call BEFORE the most recent one instead, which lets a
        training loop read every step's losses without draining the GPU: `train(); losses = last_losses(lag=1)`.
        After train_n / profile_step (no per-step copy queued) the result is read synchronously."""
        L = self._learner
        if L.fresh_host_step:
            rc = _lib.lib().d4pg_learner_fetch_losses(L.handle, int(lag), L.losses_out)
            if rc:
                _lib.check(rc, "d4pg_learner_fetch_losses")
            return L.losses_out[0], L.losses_out[1]
        rc = L.read_losses(L.handle, L.losses_out, L.stream_ptr)       # D2H + wait: the step's result
        if rc:
            _lib.check(rc, "d4pg_learner_read_losses")
        return L.losses_out[0], L.losses_out[1]

    def last_batch_info(self):
        """Device tensors of the most recent step: sampled idx, IS weights, td, new priorities."""
        L = self._learner
        torch.cuda.current_stream().wait_stream(L.stream)
        return dict(idx=L.idx, weights=L.weights, td=L.td, prio=L.prio)

    def debug_tensor(self, name, shape=None, dtype=torch.float32):
        L = self._learner
        L.stream.synchronize()
        t = L.tensor(name, dtype)
        if shape is not None and t.dim() == 2:
            return t[:, :shape[1]].contiguous()          # drop the pad columns of the row pitch
        return t.view(*shape) if shape is not None else t

    def profile_step(self, global_model=None):
        """One eager step with CUDA events around every launch -> [(launcher, ms), ...].
        Counts as a real training step (device-side sampling state advances)."""
        g = global_model if global_model is not None else self
        L = self._get_learner(g)
        self.replayBuffer._store.flush()
        L.stream.wait_stream(torch.cuda.current_stream())
        n, cap, stride = C.c_int32(), 64, 48
        ms = (C.c_float * cap)()
        names = C.create_string_buffer(cap * stride)
        with torch.cuda.stream(L.stream):
            if self.sampling == "reference":
                if self.prioritized_replay:
                    L.uniforms.copy_(torch.tensor([random.random() for _ in range(self.batch_size)], dtype=torch.float64))
                else:
                    L.positions.copy_(torch.as_tensor(np.asarray(self.replayBuffer.sample_positions(self.batch_size), dtype=np.int32)))
            _lib.check(_lib.lib().d4pg_learner_profile_step(L.handle, C.c_void_p(L.stream.cuda_stream), cap, ms, names,
                                                            stride, C.byref(n)), "d4pg_learner_profile_step")
        L.fresh_host_step = False
        if self.prioritized_replay:
            self.beta_schedule.t += 1
        for opt in (self.optimizer_global_actor, self.optimizer_global_critic):
            opt.step_count += 1
        raw = names.raw
        return [(raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode(), float(ms[i])) for i in range(n.value)]

    def kernels_per_step(self):
        return int(_lib.lib().d4pg_learner_kernels_per_step(self._learner.handle)) if self._learner else 0
