"""Prioritized replay with the reference's class names and method signatures
(prioritized_replay_memory.py), backed by GPU-resident storage + segment trees.

  LinearSchedule(schedule_timesteps, final_p, initial_p).value()            :5-29
  SegmentTree / SumSegmentTree / MinSegmentTree(capacity)                    :33-162
  ReplayBuffer(size).add/.sample/__len__                                     :164-222
  PrioritizedReplayBuffer(size, alpha).add/.sample(B, beta)/.update_priorities/__len__   :224-335

Transitions live in SoA device arrays (obs/obs2 f32, act f32, reward f64, done u8) and the
sum/min trees are fp32 device arrays; every operation is a call into libd4pg_sm100.so.
`add()` stages rows in pinned host memory and flushes them with one H2D copy + one kernel
before anything reads the buffer, so the observable behaviour is the reference's.
Seeded-index parity: `sample()` draws its B uniforms from Python's global `random`
exactly as `_sample_proportional` does (:262), so `random.seed(k)` reproduces the
reference's indices.
"""
import ctypes as C
import random

import numpy as np
import torch

from . import _lib
from .utils import default_device


class LinearSchedule(object):
    """Linear interpolation initial_p -> final_p over `schedule_timesteps` calls; `value()`
    post-increments its clock (:25-29)."""

    def __init__(self, schedule_timesteps, final_p, initial_p=1.0):
        self.schedule_timesteps = schedule_timesteps
        self.final_p = final_p
        self.initial_p = initial_p
        self.t = 0

    def value(self):
        fraction = min(float(self.t) / self.schedule_timesteps, 1.0)
        self.t += 1
        return self.initial_p + fraction * (self.final_p - self.initial_p)


class _DeviceReplay(object):
    """Device storage + trees + the C handle.  Allocated lazily on the first add (the
    reference constructors do not know obs/act dims)."""

    STAGE_ROWS = 4096

    def __init__(self, size, alpha, prioritized, obs_dim=None, act_dim=None, device=None):
        self.size = int(size)
        self.alpha = float(alpha)
        self.prioritized = bool(prioritized)
        self.device = torch.device(device) if device is not None else None
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.handle = None
        self._n_staged = 0
        self._len = 0
        self._next_idx = 0
        # host pipeline (a learner's ingest stream, ddpg.py): add_batch_host is issued there; every other device
        # operation runs on the caller's stream -- the two are kept in program order by events, only when they interleave
        self._ingest_stream = None
        self._ing_dirty = False
        self._cs_dirty = False
        if obs_dim is not None and act_dim is not None and torch.cuda.is_available():
            self._allocate(obs_dim, act_dim)

    # -- allocation --------------------------------------------------------------------
    def _allocate(self, obs_dim, act_dim):
        _lib.require_cuda()
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        dev = self.device or default_device()
        self.device = dev
        cap = C.c_int64()
        _lib.check(_lib.lib().d4pg_replay_capacity(self.size, C.byref(cap)), "d4pg_replay_capacity")
        self.capacity = int(cap.value)
        f32, n = torch.float32, self.size
        self.sum_tree = torch.empty(2 * self.capacity, dtype=f32, device=dev)
        self.min_tree = torch.empty(2 * self.capacity, dtype=f32, device=dev)
        self.obs = torch.zeros(n, self.obs_dim, dtype=f32, device=dev)
        self.obs2 = torch.zeros(n, self.obs_dim, dtype=f32, device=dev)
        self.act = torch.zeros(n, self.act_dim, dtype=f32, device=dev)
        self.rew = torch.zeros(n, dtype=torch.float64, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.scratch = torch.empty(self.capacity, dtype=torch.int32, device=dev)
        self.state = torch.zeros(8, dtype=f32, device=dev)
        h = C.c_void_p()
        _lib.check(_lib.lib().d4pg_replay_create(self.size, self.obs_dim, self.act_dim, self.alpha,
                                                 _lib.ptr(self.sum_tree), _lib.ptr(self.min_tree),
                                                 _lib.ptr(self.obs), _lib.ptr(self.act), _lib.ptr(self.rew),
                                                 _lib.ptr(self.obs2), _lib.ptr(self.done), _lib.ptr(self.scratch),
                                                 _lib.ptr(self.state), _lib.stream_ptr(), C.byref(h)),
                   "d4pg_replay_create")
        self.handle = h
        R = self.STAGE_ROWS
        pin = dict(pin_memory=True)
        self._st_obs = torch.empty(R, self.obs_dim, dtype=f32, **pin)
        self._st_obs2 = torch.empty(R, self.obs_dim, dtype=f32, **pin)
        self._st_act = torch.empty(R, self.act_dim, dtype=f32, **pin)
        self._st_rew = torch.empty(R, dtype=torch.float64, **pin)
        self._st_done = torch.empty(R, dtype=torch.uint8, **pin)
        self._np = [t.numpy() for t in (self._st_obs, self._st_act, self._st_rew, self._st_obs2, self._st_done)]
        self._dev_stage = [torch.empty_like(t, device=dev) for t in
                           (self._st_obs, self._st_act, self._st_rew, self._st_obs2, self._st_done)]

    def __del__(self):
        try:
            if self.handle is not None:
                h, self.handle = self.handle, None       # a learner collected later (reference cycles) must not touch it
                self._ingest_stream = None
                _lib.lib().d4pg_replay_destroy(h)
        except Exception:
            pass

    # -- streams -----------------------------------------------------------------------
    def attach_ingest_stream(self, raw_stream):
        """raw cudaStream_t (int) of the learner's ingest stream, or None to detach (joins it first)."""
        if self._ingest_stream is not None and self.handle is not None:
            self._join_ingest()
        self._ingest_stream = raw_stream
        self._ing_dirty = False
        self._cs_dirty = raw_stream is not None          # whatever the caller's stream did so far comes first

    def _join_ingest(self):
        """The caller's stream is about to touch the buffer: order it after the ingest stream's adds."""
        if self._ingest_stream is not None and self.handle is not None:
            if self._ing_dirty:
                _lib.check(_lib.lib().d4pg_replay_order_after(self.handle, C.c_void_p(self._ingest_stream), _lib.stream_ptr()),
                           "d4pg_replay_order_after")
                self._ing_dirty = False
            self._cs_dirty = True

    def _ingest_ptr(self):
        """Stream of a host add: the ingest stream when attached (ordered after the caller's earlier buffer operations)."""
        ing = self._ingest_stream
        if ing is None:
            return _lib.raw_stream()
        if self._cs_dirty:
            _lib.check(_lib.lib().d4pg_replay_order_after(self.handle, _lib.stream_ptr(), C.c_void_p(ing)),
                       "d4pg_replay_order_after")
            self._cs_dirty = False
        self._ing_dirty = True
        return ing

    # -- ingest ------------------------------------------------------------------------
    def add(self, s, a, r, s2, done):
        s = np.asarray(s, dtype=np.float32).reshape(-1)
        a = np.asarray(a, dtype=np.float32).reshape(-1)
        if self.handle is None:
            self._allocate(s.shape[0], a.shape[0])
        i = self._n_staged
        o, ac, rw, o2, dn = self._np
        o[i] = s
        ac[i] = a
        rw[i] = float(r)
        o2[i] = np.asarray(s2, dtype=np.float32).reshape(-1)
        dn[i] = 1 if done else 0
        self._n_staged = i + 1
        self._len = min(self.size, self._len + 1)
        if self._n_staged == self.STAGE_ROWS or self._n_staged == self.size:
            self.flush()

    def _pack_layout(self, n):
        """Byte offsets of (obs, act, rew, obs2, done) for n rows packed into one staging buffer."""
        S, A = self.obs_dim * 4, self.act_dim * 4
        o_obs = 0
        o_obs2 = o_obs + n * S
        o_act = o_obs2 + n * S
        o_rew = (o_act + n * A + 15) & ~15
        o_done = o_rew + n * 8
        return o_obs, o_act, o_rew, o_obs2, o_done, (o_done + n + 15) & ~15

    def add_batch_host(self, s, a, r, s2, done):
        """Fast ingest of n <= STAGE_ROWS host transitions through ONE library call
        (`d4pg_replay_add_host`): packed into a pinned staging buffer, one async H2D copy, ring +
        tree kernels."""
        if (torch.is_tensor(s) and self.handle is not None and getattr(self, "_pack_host", None) is not None
                and s.dtype == torch.float32 and s.dim() == 2 and 0 < s.shape[0] <= min(self.STAGE_ROWS, self.size)
                and not self._n_staged and torch.is_tensor(a) and torch.is_tensor(r) and torch.is_tensor(s2)
                and torch.is_tensor(done) and a.dtype == torch.float32 and r.dtype == torch.float64
                and s2.dtype == torch.float32 and done.dtype in (torch.bool, torch.uint8)
                and s.is_contiguous() and a.is_contiguous() and r.is_contiguous() and s2.is_contiguous()
                and done.is_contiguous()):
            # host tensors of the right types (e.g. slices of a pinned rollout buffer): no numpy round trip
            n = s.shape[0]
            rc = _lib.lib().d4pg_replay_add_host(self.handle, n, s.data_ptr(), a.data_ptr(), r.data_ptr(), s2.data_ptr(),
                                                 done.data_ptr(), 1 if self.prioritized else 0, self._ingest_ptr())
            if rc:
                _lib.check(rc, "d4pg_replay_add_host")
            self._next_idx = (self._next_idx + n) % self.size
            self._len = min(self.size, self._len + n)
            return
        s = np.ascontiguousarray(s, dtype=np.float32)
        n = s.shape[0] if s.ndim == 2 else 1
        a = np.ascontiguousarray(a, dtype=np.float32)
        if self.handle is None:
            self._allocate(s.reshape(n, -1).shape[1], a.reshape(n, -1).shape[1])
        if n > self.STAGE_ROWS or n > self.size:
            return self.add_batch(s, a, r, s2, done)
        if self._n_staged:
            self.flush()
        L = _lib.lib()
        if getattr(self, "_pack_host", None) is None:
            nbytes = int(L.d4pg_replay_staging_bytes(self.handle, self.STAGE_ROWS))
            self._pack_host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self._pack_dev = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            _lib.check(L.d4pg_replay_set_staging(self.handle, _lib.ptr(self._pack_host), _lib.ptr(self._pack_dev), nbytes),
                       "d4pg_replay_set_staging")
        r = np.ascontiguousarray(r, dtype=np.float64)
        s2 = np.ascontiguousarray(s2, dtype=np.float32)
        d = np.ascontiguousarray(done)
        if d.dtype != np.uint8:
            d = d.astype(np.uint8)
        rc = L.d4pg_replay_add_host(self.handle, n, s.ctypes.data, a.ctypes.data, r.ctypes.data, s2.ctypes.data,
                                    d.ctypes.data, 1 if self.prioritized else 0, self._ingest_ptr())
        if rc:
            _lib.check(rc, "d4pg_replay_add_host")
        self._next_idx = (self._next_idx + n) % self.size
        self._len = min(self.size, self._len + n)

    def add_batch(self, s, a, r, s2, done):
        """Vectorised ingest of n transitions (host numpy / CPU or CUDA tensors)."""
        self.flush()
        s = torch.as_tensor(s, dtype=torch.float32)
        a = torch.as_tensor(a, dtype=torch.float32)
        if s.dim() == 1:
            s = s.view(1, -1)
        n = s.shape[0]
        a = a.reshape(n, -1)
        if self.handle is None:
            self._allocate(s.shape[1], a.shape[1])
        dev = self.device
        nb = not s.is_cuda
        done_t = torch.as_tensor(done)
        args = [s.to(dev, non_blocking=nb).contiguous(), a.to(dev, non_blocking=nb).contiguous(),
                torch.as_tensor(r, dtype=torch.float64).reshape(n).to(dev, non_blocking=nb).contiguous(),
                torch.as_tensor(s2, dtype=torch.float32).reshape(n, -1).to(dev, non_blocking=nb).contiguous(),
                done_t.reshape(n).to(torch.uint8).to(dev, non_blocking=nb).contiguous()]
        for lo in range(0, n, self.size):
            hi = min(n, lo + self.size)
            self._add_device(hi - lo, [t[lo:hi] for t in args])

    def _episode_to_device(self, s, a, r, s2, done):
        self.flush()
        s = torch.as_tensor(s, dtype=torch.float32)
        if s.dim() == 1:
            s = s.view(1, -1)
        T = s.shape[0]
        a = torch.as_tensor(a, dtype=torch.float32).reshape(T, -1)
        if self.handle is None:
            self._allocate(s.shape[1], a.shape[1])
        dev = self.device
        return T, [s.to(dev).contiguous(), a.to(dev).contiguous(),
                   torch.as_tensor(r, dtype=torch.float64).reshape(T).to(dev).contiguous(),
                   torch.as_tensor(s2, dtype=torch.float32).reshape(T, -1).to(dev).contiguous(),
                   torch.as_tensor(done).reshape(T).to(torch.uint8).to(dev).contiguous()]

    def add_episode_nstep(self, s, a, r, s2, done, n_steps, gamma):
        """One episode of T consecutive steps with the n-step return accumulated ON THE DEVICE at insert
        (replay_memory.py:38-45): transition i = (s_i, a_i, sum_k gamma^k r_{i+k}, s'_{i+n-1}, done_{i+n-1})."""
        T, args = self._episode_to_device(s, a, r, s2, done)
        n_steps = int(n_steps)
        if T < n_steps:
            return 0
        m = T - n_steps + 1
        if m > self.size:
            raise _lib.D4PGError("add_episode_nstep: episode longer than the buffer")
        scratch = torch.empty(T, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().d4pg_replay_add_nstep(self.handle, T, *[_lib.ptr(t) for t in args], n_steps, float(gamma),
                                                    _lib.ptr(scratch), 1 if self.prioritized else 0, _lib.stream_ptr()),
                   "d4pg_replay_add_nstep")
        self._len = int(_lib.lib().d4pg_replay_len(self.handle))
        self._next_idx = int(_lib.lib().d4pg_replay_next_idx(self.handle))
        return m

    def add_her_episode(self, obs, obs_next, goal, ag_next, act, rew, done, her_ratio=0.8, threshold=0.05,
                        her_action="reference", rng=None):
        """Hindsight relabelling of one goal-conditioned episode on the device (main.py:154-184): every transition is
        stored, and with probability `her_ratio` also a copy whose goal is the achieved goal of a uniformly chosen
        FUTURE step (reward recomputed as the sparse -(distance > threshold), done = reward == 0).  The random draws
        are made on the host in the reference's order (np.random.uniform() then np.random.randint(t, T) per step)."""
        rng = np.random if rng is None else rng
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        T, So = obs.shape
        goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(T, -1)
        G = goal.shape[1]
        act = np.ascontiguousarray(act, dtype=np.float32).reshape(T, -1)
        A = act.shape[1]
        select = np.zeros(T, dtype=np.uint8)
        future = np.arange(T, dtype=np.int32)
        for t in range(T):
            if rng.uniform() < her_ratio:                                                   # main.py:166
                select[t] = 1
                future[t] = rng.randint(t, T)                                               # main.py:170
        counts = 1 + select.astype(np.int64)
        dst = (np.cumsum(counts) - counts).astype(np.int32)
        n_out = int(counts.sum())
        if self.handle is None:
            self._allocate(So + G, A)
        if n_out > self.size:
            raise _lib.D4PGError("add_her_episode: episode longer than the buffer")
        self.flush()
        dev = self.device
        up = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x, dtype=dt)).to(dev)
        ins = [up(obs, np.float32), up(np.asarray(obs_next).reshape(T, So), np.float32), up(goal, np.float64),
               up(np.asarray(ag_next).reshape(T, G), np.float64), up(act, np.float32), up(np.asarray(rew).reshape(T), np.float64),
               up(np.asarray(done).reshape(T).astype(np.uint8), np.uint8), up(select, np.uint8), up(future, np.int32),
               up(dst, np.int32)]
        outs = [torch.empty(n_out, So + G, dtype=torch.float32, device=dev), torch.empty(n_out, A, dtype=torch.float32, device=dev),
                torch.empty(n_out, dtype=torch.float64, device=dev), torch.empty(n_out, So + G, dtype=torch.float32, device=dev),
                torch.empty(n_out, dtype=torch.uint8, device=dev)]
        _lib.check(_lib.lib().d4pg_her_relabel(T, So, G, A, *[_lib.ptr(t) for t in ins], float(threshold),
                                               0 if her_action == "reference" else 1, *[_lib.ptr(t) for t in outs],
                                               _lib.stream_ptr()), "d4pg_her_relabel")
        self._add_device(n_out, outs)
        return n_out

    def _add_device(self, n, tensors):
        _lib.check(_lib.lib().d4pg_replay_add(self.handle, n, *[_lib.ptr(t) for t in tensors],
                                              1 if self.prioritized else 0, _lib.stream_ptr()), "d4pg_replay_add")
        self._len = int(_lib.lib().d4pg_replay_len(self.handle))
        self._next_idx = int(_lib.lib().d4pg_replay_next_idx(self.handle))

    def flush(self):
        self._join_ingest()
        n = self._n_staged
        if n == 0:
            return
        host = (self._st_obs, self._st_act, self._st_rew, self._st_obs2, self._st_done)
        for d, h in zip(self._dev_stage, host):
            d[:n].copy_(h[:n], non_blocking=True)
        self._add_device(n, [d[:n] for d in self._dev_stage])
        # the pinned staging rows may be overwritten by the next add(): wait for the copies
        torch.cuda.current_stream().synchronize()
        self._n_staged = 0

    def __len__(self):
        return self._len

    # -- read paths -------------------------------------------------------------------
    def _batch_buffers(self, B):
        dev, f32 = self.device, torch.float32
        return dict(idx=torch.empty(B, dtype=torch.int32, device=dev),
                    w=torch.empty(B, dtype=f32, device=dev),
                    s=torch.empty(B, self.obs_dim, dtype=f32, device=dev),
                    a=torch.empty(B, self.act_dim, dtype=f32, device=dev),
                    r=torch.empty(B, dtype=torch.float64, device=dev),
                    s2=torch.empty(B, self.obs_dim, dtype=f32, device=dev),
                    d=torch.empty(B, dtype=torch.uint8, device=dev))

    def sample_proportional(self, B, beta, uniforms=None, philox=None):
        """Device tensors (idx i32, weights f32, s, a, r f64, s2, done u8)."""
        self.flush()
        if self.handle is None:
            raise _lib.D4PGError("sample() on an empty replay buffer")
        o = self._batch_buffers(B)
        u_dev = None
        seed, ctr = 0, 0
        if philox is not None:
            seed, ctr = philox
        else:
            if uniforms is None:
                uniforms = [random.random() for _ in range(B)]          # :262, global `random`
            u_dev = torch.tensor(np.asarray(uniforms, dtype=np.float64)).to(self.device)
        _lib.check(_lib.lib().d4pg_replay_sample(self.handle, B, _lib.ptr(u_dev), seed, ctr, float(beta),
                                                 _lib.ptr(o["idx"]), _lib.ptr(o["w"]), _lib.ptr(o["s"]), _lib.ptr(o["a"]),
                                                 _lib.ptr(o["r"]), _lib.ptr(o["s2"]), _lib.ptr(o["d"]), _lib.stream_ptr()),
                   "d4pg_replay_sample")
        return o

    def gather(self, positions):
        self.flush()
        pos = torch.as_tensor(np.asarray(positions, dtype=np.int32)).to(self.device)
        B = pos.numel()
        o = self._batch_buffers(B)
        o["idx"] = pos
        _lib.check(_lib.lib().d4pg_replay_gather(self.handle, B, _lib.ptr(pos), _lib.ptr(o["s"]), _lib.ptr(o["a"]),
                                                 _lib.ptr(o["r"]), _lib.ptr(o["s2"]), _lib.ptr(o["d"]), _lib.stream_ptr()),
                   "d4pg_replay_gather")
        return o

    def update_priorities(self, idxes, priorities):
        self.flush()
        idx = torch.as_tensor(np.asarray(idxes, dtype=np.int32)).to(self.device) if not torch.is_tensor(idxes) \
            else idxes.to(device=self.device, dtype=torch.int32)
        pr = torch.as_tensor(np.asarray(priorities, dtype=np.float32)).to(self.device) if not torch.is_tensor(priorities) \
            else priorities.to(device=self.device, dtype=torch.float32)
        assert idx.numel() == pr.numel()                                                   # :328
        if idx.numel():
            # the reference's per-element asserts (:330-331); an unchecked index would be a stray device write into the trees
            assert bool((pr > 0).all()), "priorities must be > 0"
            assert bool(((idx >= 0) & (idx < len(self))).all()), "index out of range"
        _lib.check(_lib.lib().d4pg_replay_update_priorities(self.handle, idx.numel(), _lib.ptr(idx), _lib.ptr(pr),
                                                            _lib.stream_ptr()), "d4pg_replay_update_priorities")

    def reduce(self, start=0, end=None):
        self.flush()
        out = torch.empty(2, dtype=torch.float32, device=self.device)
        e = 0 if end is None else int(end)
        _lib.check(_lib.lib().d4pg_replay_reduce(self.handle, int(start), e, _lib.ptr(out), _lib.stream_ptr()),
                   "d4pg_replay_reduce")
        return out.cpu().numpy()

    @property
    def max_priority(self):
        self.flush()
        return float(self.state[0].item())


class _TreeView(object):
    """`_it_sum` / `_it_min` facade: the reference's SegmentTree read API over the device tree."""

    def __init__(self, store, which):
        self._store, self._which = store, which

    def _tree(self):
        self._store.flush()
        return self._store.sum_tree if self._which == 0 else self._store.min_tree

    def __getitem__(self, idx):
        assert 0 <= idx < self._store.capacity                                              # :111
        return np.float32(self._tree()[self._store.capacity + idx].item())

    def reduce(self, start=0, end=None):
        return np.float32(self._store.reduce(start, end)[self._which])

    def sum(self, start=0, end=None):
        return self.reduce(start, end)

    def min(self, start=0, end=None):
        return self.reduce(start, end)

    def values(self):
        """All 2*capacity node values (host copy)."""
        return self._tree().cpu().numpy()


class SegmentTree(object):
    """Standalone device segment tree with the reference constructor (:34-58).  `operation`
    must be addition or min (the only two the reference instantiates)."""

    def __init__(self, capacity, operation=None, neutral_element=None, _which=0):
        assert capacity > 0 and capacity & (capacity - 1) == 0, "capacity must be positive and a power of 2."
        self._capacity = capacity
        self._which = _which
        self._store = _DeviceReplay(capacity, 1.0, True, obs_dim=1, act_dim=1)
        if self._store.handle is None:
            _lib.require_cuda()

    def __setitem__(self, idx, val):
        st = self._store
        i = torch.tensor([idx], dtype=torch.int32, device=st.device)
        v = torch.tensor([val], dtype=torch.float32, device=st.device)
        _lib.check(_lib.lib().d4pg_replay_set_leaves(st.handle, 1, _lib.ptr(i), _lib.ptr(v), _lib.ptr(v),
                                                     _lib.stream_ptr()), "d4pg_replay_set_leaves")

    def __getitem__(self, idx):
        assert 0 <= idx < self._capacity
        tree = self._store.sum_tree if self._which == 0 else self._store.min_tree
        return np.float32(tree[self._capacity + idx].item())

    def reduce(self, start=0, end=None):
        return np.float32(self._store.reduce(start, end)[self._which])


class SumSegmentTree(SegmentTree):
    def __init__(self, capacity):
        super(SumSegmentTree, self).__init__(capacity, _which=0)

    def sum(self, start=0, end=None):
        return self.reduce(start, end)

    def find_prefixsum_idx(self, prefixsum):
        st = self._store
        assert 0 <= prefixsum <= self.sum() + 1e-5                                          # :141
        m = torch.tensor([float(prefixsum)], dtype=torch.float64, device=st.device)
        out = torch.empty(1, dtype=torch.int32, device=st.device)
        _lib.check(_lib.lib().d4pg_replay_find_prefixsum(st.handle, 1, _lib.ptr(m), _lib.ptr(out), _lib.stream_ptr()),
                   "d4pg_replay_find_prefixsum")
        return int(out.item())


class MinSegmentTree(SegmentTree):
    def __init__(self, capacity):
        super(MinSegmentTree, self).__init__(capacity, _which=1)

    def min(self, start=0, end=None):
        return self.reduce(start, end)


def _to_host_batch(o):
    return (o["s"].cpu().numpy(), o["a"].cpu().numpy(), o["r"].cpu().numpy(), o["s2"].cpu().numpy(),
            o["d"].cpu().numpy().astype(bool))


class ReplayBuffer(object):
    """Uniform-sampling ring buffer (:164-222)."""

    _prioritized = False

    def __init__(self, size, obs_dim=None, act_dim=None, device=None, _alpha=1.0):
        self._maxsize = size
        self._store = _DeviceReplay(size, _alpha, self._prioritized, obs_dim, act_dim, device)

    def __len__(self):
        return len(self._store)

    @property
    def _next_idx(self):
        return (self._store._next_idx + self._store._n_staged) % self._store.size

    def add(self, obs_t, action, reward, obs_tp1, done):
        self._store.add(obs_t, action, reward, obs_tp1, done)

    def add_batch(self, obs_t, action, reward, obs_tp1, done):
        """n transitions at once.  Host arrays of up to 4096 rows take the packed single-copy path."""
        if isinstance(obs_t, np.ndarray) or (torch.is_tensor(obs_t) and not obs_t.is_cuda):
            self._store.add_batch_host(obs_t, action, reward, obs_tp1, done)
        else:
            self._store.add_batch(obs_t, action, reward, obs_tp1, done)

    def add_episode(self, obs, action, reward, obs_next, done, n_steps=1, gamma=0.99):
        """One episode of consecutive steps; the n-step return is accumulated on the device at insert
        (the arithmetic of replay_memory.py:38-45).  Returns the number of transitions inserted."""
        return self._store.add_episode_nstep(obs, action, reward, obs_next, done, n_steps, gamma)

    def add_her_episode(self, obs, obs_next, goal, achieved_goal_next, action, reward, done, **kw):
        """Hindsight relabelling on the device (main.py:154-184); see _DeviceReplay.add_her_episode."""
        return self._store.add_her_episode(obs, obs_next, goal, achieved_goal_next, action, reward, done, **kw)

    def _encode_sample(self, idxes):
        return _to_host_batch(self._store.gather(idxes))

    def sample(self, batch_size):
        idxes = [random.randint(0, len(self) - 1) for _ in range(batch_size)]              # :221
        return self._encode_sample(idxes)


class PrioritizedReplayBuffer(ReplayBuffer):
    """Proportional prioritized replay (:224-335) on GPU-resident fp32 sum/min trees."""

    _prioritized = True

    def __init__(self, size, alpha, obs_dim=None, act_dim=None, device=None):
        assert alpha >= 0                                                                    # :240
        self._alpha = alpha
        super(PrioritizedReplayBuffer, self).__init__(size, obs_dim, act_dim, device, _alpha=alpha)
        self._it_sum = _TreeView(self._store, 0)
        self._it_min = _TreeView(self._store, 1)

    @property
    def _max_priority(self):
        return self._store.max_priority

    def sample(self, batch_size, beta, uniforms=None):
        """-> (obs, act, rew, obs2, done, weights, idxes) as the reference returns them:
        host numpy arrays, `idxes` a list of ints.  `uniforms` (optional) overrides the
        `random.random()` draws."""
        assert beta > 0                                                                      # :299
        o = self._store.sample_proportional(batch_size, beta, uniforms)
        idxes = [int(i) for i in o["idx"].cpu().numpy()]
        return tuple(list(_to_host_batch(o)) + [o["w"].cpu().numpy(), idxes])

    def update_priorities(self, idxes, priorities):
        self._store.update_priorities(idxes, priorities)
