"""CPU restatement of the reference's D4PG learner hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker the CUDA path is
compared against, and the CPU arm of bench.py.  Never imported by the product.

Parity: PINNED against the unmodified reference run in the build container
(tests/test_oracle_vs_reference.py) and against tests/golden/*.npz generated
from it (tests/golden/make_golden.py).  dtype semantics are those of
NumPy 2.x (NEP 50) + torch 2.11 CPU, which is what the reference's Python
evaluates to in this image.

Every function cites the reference lines (relative to /root/reference) whose
arithmetic it restates.  The code is written array-first (numpy fp32 trees,
explicit dtype casts) rather than as a transcription of the reference's
list-of-Python-floats code; equality with the reference is established by the
tests above, not by textual similarity.
"""
from __future__ import annotations

import math
import numpy as np
import torch
import torch.nn.functional as F

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------
# Atom support
# --------------------------------------------------------------------------
def atom_support(v_min, v_max, n_atoms):
    """delta and bin centres exactly as ddpg.py:43-47 (Python-float arithmetic)."""
    delta = (v_max - v_min) / float(n_atoms - 1)
    centers = np.array([v_min + i * delta for i in range(n_atoms)], dtype=F64)
    return delta, centers


# --------------------------------------------------------------------------
# P1: live projection, ddpg.py:142-185 (`reproject2`)
# --------------------------------------------------------------------------
def project_live(target_probs, rewards, dones, v_min, v_max, n_atoms, gamma,
                 return_bins=False):
    """Categorical Bellman projection as the reference's live code computes it.

    target_probs [B,N] f32, rewards [B] f64, dones [B] bool -> m [B,N] f32.

    * atom constants `(v_min + j*delta)*gamma` in Python-float (f64) arithmetic,
      discount is gamma (NOT gamma**n)           ddpg.py:155
    * b_j in f64, l=floor, u=ceil as int64                      ddpg.py:156-158
    * accumulation into an f32 array, each add carried out in f64 then
      rounded to f32, atoms visited j=0..N-1                    ddpg.py:150,160-163
    * terminal rows: zeroed, then Dirac at clip(r) with weights cast to f32
                                                                ddpg.py:165-181
      (per-row; the reference's batched version crashes on mixed batches,
      SURVEY.md H6 -- this is what it computes whenever it does not crash).
    """
    p = np.ascontiguousarray(target_probs, dtype=F32)
    r = np.asarray(rewards, dtype=F64).reshape(-1)
    d = np.asarray(dones).reshape(-1).astype(bool)
    B = r.shape[0]
    delta, _ = atom_support(v_min, v_max, n_atoms)
    m = np.zeros((B, n_atoms), dtype=F32)
    rows = np.arange(B)
    bins_l = np.zeros((B, n_atoms), dtype=np.int64)
    bins_u = np.zeros((B, n_atoms), dtype=np.int64)
    for j in range(n_atoms):
        c_j = (v_min + j * delta) * gamma                       # Python floats
        tz = np.minimum(v_max, np.maximum(v_min, r + c_j))
        b = (tz - v_min) / delta
        l = np.floor(b).astype(np.int64)
        u = np.ceil(b).astype(np.int64)
        bins_l[:, j] = l
        bins_u[:, j] = u
        pj = p[:, j].astype(F64)
        eq = (u == l)
        ne = ~eq
        # f32 + f64 -> f64 add, result stored back as f32
        m[rows[eq], l[eq]] = (m[rows[eq], l[eq]].astype(F64) + pj[eq]).astype(F32)
        m[rows[ne], l[ne]] = (m[rows[ne], l[ne]].astype(F64)
                              + pj[ne] * (u - b)[ne]).astype(F32)
        m[rows[ne], u[ne]] = (m[rows[ne], u[ne]].astype(F64)
                              + pj[ne] * (b - l)[ne]).astype(F32)
    for i in np.nonzero(d)[0]:
        m[i, :] = 0.0
        tz = min(v_max, max(v_min, float(r[i])))
        b = (tz - v_min) / delta
        l = int(math.floor(b))
        u = int(math.ceil(b))
        bins_l[i, :] = l
        bins_u[i, :] = u
        if l == u:
            m[i, l] = 1.0
        else:
            m[i, l] = F32(u - b)
            m[i, u] = F32(b - l)
    if return_bins:
        return m, bins_l, bins_u
    return m


# --------------------------------------------------------------------------
# P2: n-step projection, ddpg.py:122-140 (`reproj_categorical_dist`, dead code
# in the reference but its only gamma**n variant: the config-5 oracle)
# --------------------------------------------------------------------------
def project_nstep(target_probs, rewards, dones, v_min, v_max, n_atoms, gamma,
                  n_steps, return_bins=False):
    """f64 projection with discount gamma**n_steps and a (1-done) mask.

    Integer b_j (l==u) are split as (l-1,u) for u>0 and (0,1) for u==0
    (ddpg.py:133-134); scatter order: all l-contributions row-major, then all
    u-contributions (two np.add.at calls, ddpg.py:137-138).  Returns f64.
    """
    p = np.asarray(target_probs)
    r = np.asarray(rewards, dtype=F64).reshape(-1, 1)
    d = np.asarray(dones, dtype=F64).reshape(-1, 1)
    B = r.shape[0]
    delta, centers = atom_support(v_min, v_max, n_atoms)
    disc = gamma ** n_steps                                      # ddpg.py:24
    tz = r + disc * (1 - d) * centers.reshape(1, -1)
    tz = np.minimum(v_max, np.maximum(v_min, tz))
    b = (tz - v_min) / delta
    l = np.floor(b).astype(np.int64)
    u = np.ceil(b).astype(np.int64)
    l[(u > 0) & (l == u)] -= 1
    u[(l < (n_atoms - 1)) & (l == u)] += 1
    m = np.zeros(B * n_atoms, dtype=F64)
    off = (np.arange(B, dtype=np.int64) * n_atoms).reshape(-1, 1)
    np.add.at(m, (l + off).reshape(-1), (p * (u.astype(F64) - b)).reshape(-1))
    np.add.at(m, (u + off).reshape(-1), (p * (b - l.astype(F64))).reshape(-1))
    m = m.reshape(B, n_atoms)
    if return_bins:
        return m, l, u
    return m


# --------------------------------------------------------------------------
# L1/L2: critic loss, TD proxy and priorities, ddpg.py:217,220-222,253
# --------------------------------------------------------------------------
def critic_loss_terms(m, q):
    """m,q [B,N] f32 (q = softmax output).  Returns dict with
    loss (f32 scalar), loss_rows [B], td [B], prio [B] and dlogits [B,N]
    (gradient of the mean loss w.r.t. the pre-softmax logits)."""
    mt = torch.as_tensor(np.asarray(m, dtype=F32))
    qt = torch.as_tensor(np.asarray(q, dtype=F32))
    rows = -(mt * torch.log(qt + 1e-10)).sum(dim=1)              # ddpg.py:217
    loss = rows.mean()
    td = -(mt * qt).sum(dim=1)                                   # ddpg.py:221-222
    prio = np.abs(td.numpy()) + 1e-6                             # ddpg.py:253 (f32)
    B = mt.shape[0]
    gq = -(mt / (qt + 1e-10)) / B
    dlogits = qt * (gq - (qt * gq).sum(dim=1, keepdim=True))
    return dict(loss=loss.numpy(), loss_rows=rows.numpy(), td=td.numpy(),
                prio=prio.astype(F32), dlogits=dlogits.numpy())


# --------------------------------------------------------------------------
# T: segment trees, prioritized_replay_memory.py:33-162
# --------------------------------------------------------------------------
class SegmentTree32:
    """Array-embedded binary tree, root at 1, leaves at [cap, 2cap).

    Node values are fp32 (what the reference's list of Python/NumPy scalars
    evaluates to under NumPy 2).  `kind` is 'sum' or 'min'.
    """

    def __init__(self, capacity, kind):
        assert capacity > 0 and capacity & (capacity - 1) == 0   # :56
        self.capacity = capacity
        self.kind = kind
        neutral = 0.0 if kind == "sum" else np.inf               # :116-120,152-156
        self.value = np.full(2 * capacity, neutral, dtype=F32)

    def _op(self, a, b):
        if self.kind == "sum":
            return F32(a) + F32(b)
        return min(F32(a), F32(b))

    def set(self, idx, val):
        """`__setitem__`: leaf write + parent recompute to the root (:98-108)."""
        i = idx + self.capacity
        self.value[i] = F32(val)
        i //= 2
        while i >= 1:
            self.value[i] = self._op(self.value[2 * i], self.value[2 * i + 1])
            i //= 2

    def get(self, idx):
        assert 0 <= idx < self.capacity                           # :111
        return self.value[self.capacity + idx]

    def rebuild(self):
        """Bulk parent recompute, level by level.  Equals the state reached by
        any sequence of `set` calls that leaves the same leaves (every node is
        op(left,right) of its final children)."""
        c = self.capacity
        while c > 1:
            lo, hi = c // 2, c
            kids = self.value[c:2 * c].reshape(-1, 2)
            if self.kind == "sum":
                self.value[lo:hi] = kids[:, 0] + kids[:, 1]
            else:
                self.value[lo:hi] = np.minimum(kids[:, 0], kids[:, 1])
            c //= 2

    def reduce_prefix(self, end_inclusive):
        """`reduce(0, end_inclusive+1)` (:61-96): op over leaves [0,end].

        `_reduce_helper` with start=0 returns V[left] op (recursive right part),
        i.e. a RIGHT-nested association over the canonical left-to-right node
        cover; evaluated innermost-first in fp32."""
        if end_inclusive < 0:
            raise AssertionError("empty range")
        terms = []
        node, lo, hi = 1, 0, self.capacity - 1
        e = end_inclusive
        while True:
            if e == hi:
                terms.append(self.value[node])
                break
            mid = (lo + hi) // 2
            if e <= mid:
                node, hi = 2 * node, mid
            else:
                terms.append(self.value[2 * node])
                node, lo = 2 * node + 1, mid + 1
        acc = F32(terms[-1])
        for t in reversed(terms[:-1]):
            acc = self._op(t, acc)
        return acc

    def root(self):
        return self.value[1]


def find_prefixsum_idx(values, capacity, mass):
    """Root-to-leaf descent, prioritized_replay_memory.py:126-149.
    `mass` keeps its dtype (np.float32 or Python float); the subtraction runs in
    that dtype, the comparison is exact."""
    i = 1
    while i < capacity:
        left = values[2 * i]
        if left > mass:                                           # strict, :144
            i = 2 * i
        else:
            # pristine reference tree: Python-float nodes -> the subtraction stays f64
            mass = (mass - float(left)) if isinstance(mass, float) else F32(mass - left)
            i = 2 * i + 1
    return i - capacity


# --------------------------------------------------------------------------
# S/V/W/E: prioritized replay, prioritized_replay_memory.py:164-335
# --------------------------------------------------------------------------
class PrioritizedReplayOracle:
    """State: SoA storage + fp32 sum/min trees + max_priority.

    `pristine` mirrors a reference tree that still holds only Python floats
    (no `update_priorities` call yet): there `mass = u * sum` is an f64 product
    and the descent subtracts in f64 (all node values are integers, so the tree
    contents themselves are identical in f32)."""

    def __init__(self, size, alpha, obs_dim, act_dim):
        assert alpha >= 0                                         # :240
        self.size = int(size)
        self.alpha = alpha
        cap = 1
        while cap < size:                                         # :243-245
            cap *= 2
        self.capacity = cap
        self.sum = SegmentTree32(cap, "sum")
        self.min = SegmentTree32(cap, "min")
        self.max_priority = 1.0                                   # Python float, :249
        self.max_priority_is_f32 = False
        self.pristine = True
        self.obs = np.zeros((self.size, obs_dim), dtype=F32)
        self.act = np.zeros((self.size, act_dim), dtype=F32)
        self.rew = np.zeros((self.size,), dtype=F64)
        self.obs2 = np.zeros((self.size, obs_dim), dtype=F32)
        self.done = np.zeros((self.size,), dtype=bool)
        self.length = 0
        self.next_idx = 0

    def __len__(self):
        return self.length

    def _new_leaf(self):
        """`max_priority ** alpha` (:255-256): Python-float pow while
        max_priority is still the initial Python 1.0, fp32 powf afterwards."""
        if self.max_priority_is_f32:
            return F32(self.max_priority) ** self.alpha
        return F32(float(self.max_priority) ** self.alpha)

    def add(self, s, a, r, s2, done):
        i = self.next_idx                                         # :180-187,251-256
        self.obs[i] = s
        self.act[i] = a
        self.rew[i] = r
        self.obs2[i] = s2
        self.done[i] = done
        self.length = max(self.length, i + 1)
        self.next_idx = (i + 1) % self.size
        leaf = self._new_leaf()
        self.sum.set(i, leaf)
        self.min.set(i, leaf)

    def add_batch(self, s, a, r, s2, done):
        """Same final state as calling add() row by row (bulk tree rebuild)."""
        n = len(r)
        leaf = self._new_leaf()
        idx = (self.next_idx + np.arange(n)) % self.size
        self.obs[idx] = s
        self.act[idx] = a
        self.rew[idx] = r
        self.obs2[idx] = s2
        self.done[idx] = done
        self.sum.value[self.capacity + idx] = leaf
        self.min.value[self.capacity + idx] = leaf
        self.length = min(self.size, max(self.length, self.next_idx + n))
        self.next_idx = (self.next_idx + n) % self.size
        self.sum.rebuild()
        self.min.rebuild()

    def sample_indices(self, uniforms):
        """`_sample_proportional` (:258-265) with caller-supplied U[0,1) f64
        draws standing in for `random.random()`."""
        total = self.sum.reduce_prefix(self.length - 2)          # sum(0, len-1)
        out = []
        for u in uniforms:
            if self.pristine:
                mass = float(u) * float(total)                    # Python floats
            else:
                mass = F32(F32(u) * F32(total))                   # weak float * np.float32
            out.append(find_prefixsum_idx(self.sum.value, self.capacity, mass))
        return np.asarray(out, dtype=np.int64)

    def is_weights(self, idxes, beta):
        """:303-311, fp32 under NumPy 2.  Unused by the loss."""
        assert beta > 0                                           # :299
        if self.pristine:                                         # all-Python-float tree: f64
            total = float(self.sum.root())
            max_w = (float(self.min.root()) / total * self.length) ** (-beta)
            return np.asarray([(float(self.sum.get(int(i))) / total * self.length) ** (-beta) / max_w
                               for i in idxes], dtype=F64)
        total = self.sum.root()
        p_min = F32(self.min.root() / total)
        max_w = F32(F32(p_min * F32(self.length)) ** F32(-beta))
        w = []
        for i in idxes:
            p_s = F32(self.sum.get(int(i)) / total)
            w.append(F32(F32(p_s * F32(self.length)) ** F32(-beta)) / max_w)
        return np.asarray(w, dtype=F32)

    def encode(self, idxes):
        i = np.asarray(idxes, dtype=np.int64)                     # :189-199
        return self.obs[i], self.act[i], self.rew[i], self.obs2[i], self.done[i]

    def sample(self, batch_size, beta, uniforms):
        idx = self.sample_indices(uniforms[:batch_size])
        w = self.is_weights(idx, beta)
        return tuple(list(self.encode(idx)) + [w, idx])

    def update_priorities(self, idxes, priorities):
        """:315-335; priorities arrive as an f32 ndarray so the leaf is
        powf(p, 0.6f) evaluated by NumPy's *scalar* power."""
        assert len(idxes) == len(priorities)
        for i, p in zip(idxes, priorities):
            p = F32(p)
            assert p > 0 and 0 <= i < self.length
            leaf = p ** self.alpha                                 # np.float32 ** float
            self.sum.set(int(i), leaf)
            self.min.set(int(i), leaf)
            if p > self.max_priority:
                self.max_priority = p
                self.max_priority_is_f32 = True
        self.pristine = False


def nstep_transitions(states, actions, rewards, next_states, dones, n_steps, gamma):
    """replay_memory.py:31-45 for ONE episode: at every step t >= n-1 the reference adds
    (s_{t-n+1}, a_{t-n+1}, sum_k gamma^k r_{t-n+1+k}, s'_t, done_t); the return is accumulated left to right in
    Python floats (`cum_reward += exp_gamma * r; exp_gamma *= gamma`).  Returns the list of added tuples."""
    out = []
    T = len(rewards)
    for t in range(n_steps - 1, T):
        cum, eg = 0., 1
        for k in range(t - n_steps + 1, t + 1):
            cum += eg * rewards[k]
            eg *= gamma
        i = t - n_steps + 1
        out.append((np.asarray(states[i]).reshape(-1), actions[i], cum, next_states[t], dones[t]))
    return out


def her_relabel(obs, obs_next, goal, ag_next, act, rew, done, select, future, threshold, last_action):
    """main.py:154-184 for one episode whose rollout did not succeed (`args.her and not done`), "future" strategy.
    PARITY UNPINNED: main.py cannot be imported here (gym / pybullet_envs / tensorboard-pytorch are absent), so this
    is a restatement of the listed lines only; `select[t]` stands for `np.random.uniform() < her_ratio` (:166),
    `future[t]` for `np.random.randint(t, len(episode_buffer))` (:170), the reward for the sparse gym-robotics
    `env.compute_reward` = -(||achieved - goal||_2 > distance_threshold).  The relabelled copy is stored with `action`,
    the LAST action of the rollout loop (:184 -- not the step's own `a`), passed here as `last_action`."""
    rows = []
    T = len(rew)
    for t in range(T):
        s = np.concatenate((obs[t], goal[t]))
        s_n = np.concatenate((obs_next[t], goal[t]))
        rows.append((s, act[t], rew[t], s_n, bool(done[t])))                                 # :160-163
        if select[t]:
            dummy_goal = ag_next[future[t]]                                                 # :170-171
            her_s = np.concatenate((obs[t], dummy_goal))
            her_sn = np.concatenate((obs_next[t], dummy_goal))
            d = np.linalg.norm(np.asarray(ag_next[t], dtype=np.float64) - np.asarray(dummy_goal, dtype=np.float64), axis=-1)
            her_r = -float(d > threshold)                                                   # :177
            rows.append((her_s, last_action, her_r, her_sn, her_r == 0.))                   # :182-184
    return rows


class LinearScheduleOracle:
    """prioritized_replay_memory.py:5-29 (post-incrementing beta schedule)."""

    def __init__(self, schedule_timesteps, final_p, initial_p=1.0):
        self.n, self.final_p, self.initial_p, self.t = schedule_timesteps, final_p, initial_p, 0

    def value(self):
        frac = min(float(self.t) / self.n, 1.0)
        self.t += 1
        return self.initial_p + frac * (self.final_p - self.initial_p)


# --------------------------------------------------------------------------
# U: uniform replay, replay_memory.py:14-19,61-80
# --------------------------------------------------------------------------
def uniform_sample_positions(py_random, n, k):
    """`random.sample(buffer, k)` picks the same positions as
    `random.sample(range(n), k)` for the same generator state."""
    return py_random.sample(range(n), k)


# --------------------------------------------------------------------------
# M1/M2: networks, models.py:32-41,76-88 (functional form over a weight dict)
# --------------------------------------------------------------------------
def actor_forward(w, s):
    h = F.relu(F.linear(s, w["fc1.weight"], w["fc1.bias"]))
    h = F.linear(h, w["fc2.weight"], w["fc2.bias"])               # no ReLU (H9)
    h = F.relu(F.linear(h, w["fc2_2.weight"], w["fc2_2.bias"]))
    return torch.tanh(F.linear(h, w["fc3.weight"], w["fc3.bias"]))


def critic_forward(w, s, a, logits=False):
    h = F.relu(F.linear(s, w["fc1.weight"], w["fc1.bias"]))
    h = F.relu(F.linear(torch.cat([h, a], 1), w["fc2.weight"], w["fc2.bias"]))
    h = F.relu(F.linear(h, w["fc2_2.weight"], w["fc2_2.bias"]))
    z = F.linear(h, w["fc3.weight"], w["fc3.bias"])
    return z if logits else F.softmax(z, dim=1)


PARAM_ORDER = ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias",
               "fc2_2.weight", "fc2_2.bias", "fc3.weight", "fc3.bias"]


def init_actor(obs_dim, act_dim, hidden=256):
    """Same RNG consumption as models.py:16-30 (4 nn.Linear ctors, then
    fan-"in" normals on size[0]=out_features, then fc3 ~ N(0,3e-3))."""
    import torch.nn as nn
    ls = [nn.Linear(obs_dim, hidden), nn.Linear(hidden, hidden),
          nn.Linear(hidden, hidden), nn.Linear(hidden, act_dim)]
    for l in ls[:3]:
        l.weight.data = torch.Tensor(l.weight.size()).normal_(0.0, 1.0 / np.sqrt(l.weight.size(0)))
    ls[3].weight.data.normal_(0, 3e-3)
    return _pack(ls)


def init_critic(obs_dim, act_dim, n_atoms, hidden=256):
    """models.py:52-73."""
    import torch.nn as nn
    ls = [nn.Linear(obs_dim, hidden), nn.Linear(hidden + act_dim, hidden),
          nn.Linear(hidden, hidden), nn.Linear(hidden, n_atoms)]
    for l in ls[:3]:
        l.weight.data = torch.Tensor(l.weight.size()).normal_(0.0, 1.0 / np.sqrt(l.weight.size(0)))
    ls[3].weight.data.normal_(0, 3e-4)
    return _pack(ls)


def _pack(ls):
    names = ["fc1", "fc2", "fc2_2", "fc3"]
    out = {}
    for n, l in zip(names, ls):
        out[n + ".weight"] = l.weight.data.clone()
        out[n + ".bias"] = l.bias.data.clone()
    return out


# --------------------------------------------------------------------------
# O/Z: Adam (torch 2.11 single-tensor CPU form) and Polyak
# --------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.9, eps=1e-8):
    """In-place on torch tensors; arithmetic of torch/optim/adam.py
    `_single_tensor_adam` (non-capturable, wd=0, amsgrad off) as called through
    SharedAdam at ddpg.py:232,244.  `step` is the post-increment step count."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-step_size)


def polyak(target, source, tau):
    """ddpg.py:110-116: t <- (1-tau)*t + tau*s (fp32, two products then add)."""
    target.copy_((1 - tau) * target + tau * source)


# --------------------------------------------------------------------------
# A: one full learner step, ddpg.py:200-255
# --------------------------------------------------------------------------
class LearnerOracle:
    """Single-worker learner: local == global parameters (the state the
    reference is in after every `sync_local_global`, ddpg.py:247)."""

    def __init__(self, obs_dim, act_dim, dist_info, gamma=0.99, tau=0.001,
                 n_steps=1, lr=1e-3, betas=(0.9, 0.9), eps=1e-8,
                 actor_w=None, critic_w=None, projection="live"):
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.v_min = dist_info["v_min"]
        self.v_max = dist_info["v_max"]
        self.n_atoms = dist_info["n_atoms"]
        self.gamma, self.tau, self.n_steps = gamma, tau, n_steps
        self.lr, self.betas, self.eps = lr, betas, eps
        self.projection = projection
        self.delta, centers = atom_support(self.v_min, self.v_max, self.n_atoms)
        self.z = torch.from_numpy(centers.reshape(-1, 1)).float()   # ddpg.py:47,238
        self.actor = actor_w if actor_w is not None else init_actor(obs_dim, act_dim)
        self.actor_target = {k: v.clone() for k, v in self.actor.items()}
        self.critic = critic_w if critic_w is not None else init_critic(obs_dim, act_dim, self.n_atoms)
        self.critic_target = {k: v.clone() for k, v in self.critic.items()}
        self.m_a = {k: torch.zeros_like(v) for k, v in self.actor.items()}
        self.v_a = {k: torch.zeros_like(v) for k, v in self.actor.items()}
        self.m_c = {k: torch.zeros_like(v) for k, v in self.critic.items()}
        self.v_c = {k: torch.zeros_like(v) for k, v in self.critic.items()}
        self.step_a = 0
        self.step_c = 0

    def project(self, target_probs, r, done):
        if self.projection == "live":
            return project_live(target_probs, r, done, self.v_min, self.v_max,
                                self.n_atoms, self.gamma)
        return project_nstep(target_probs, r, done, self.v_min, self.v_max,
                             self.n_atoms, self.gamma, self.n_steps).astype(F32)

    def train_step(self, s, a, r, s2, done, grad_hook=None, is_weights=None, ce_priority=False, post_update_critic=False):
        """One `DDPG.train` body on a given batch.  `grad_hook(flat_grads)` lets a
        data-parallel test average gradients across ranks before Adam.
        `is_weights` / `ce_priority` / `post_update_critic` are the corrected-semantics variants of SURVEY.md section
        8f.4 (DERIVED oracle: the reference implements none of them, ddpg.py:217,221-222,229-247): with
        `post_update_critic` the critic's Adam step runs BEFORE the policy loss is evaluated, so the actor gradient
        flows through the updated critic (what ddpg.py would do if sync_local_global preceded line 236)."""
        s_t = torch.from_numpy(np.asarray(s, dtype=F32))
        a_t = torch.from_numpy(np.asarray(a, dtype=F32))
        s2_t = torch.from_numpy(np.asarray(s2, dtype=F32))
        B = s_t.shape[0]
        out = {}
        with torch.no_grad():
            a2 = actor_forward(self.actor_target, s2_t)            # ddpg.py:205-206
            tz = critic_forward(self.critic_target, s2_t, a2)
        cw = {k: v.clone().requires_grad_(True) for k, v in self.critic.items()}
        q = critic_forward(cw, s_t, a_t)                           # ddpg.py:208
        m = self.project(tz.numpy(), np.asarray(r, dtype=F64), np.asarray(done))
        m_t = torch.from_numpy(m)
        rows_c = -(m_t * torch.log(q + 1e-10)).sum(dim=1)
        if is_weights is not None:
            rows_c = rows_c * torch.from_numpy(np.asarray(is_weights, dtype=F32))
        loss_c = rows_c.mean()                                     # ddpg.py:217 (unweighted there)
        td = -(m_t * q).sum(dim=1)                                 # ddpg.py:221-222
        loss_c.backward()                                          # ddpg.py:230
        g_c = {k: cw[k].grad.detach().clone() for k in PARAM_ORDER}

        if post_update_critic:                                     # derived variant: critic first
            if grad_hook is not None:
                raise ValueError("post_update_critic and grad_hook are not combined")
            self.step_c += 1
            for k in PARAM_ORDER:
                adam_step(self.critic[k], g_c[k], self.m_c[k], self.v_c[k], self.step_c,
                          self.lr, self.betas[0], self.betas[1], self.eps)
        # actor loss through the PRE-update critic   ddpg.py:236-242
        aw = {k: v.clone().requires_grad_(True) for k, v in self.actor.items()}
        qp = critic_forward(self.critic, s_t, actor_forward(aw, s_t))
        loss_a = -qp.matmul(self.z).mean()
        loss_a.backward()
        g_a = {k: aw[k].grad.detach().clone() for k in PARAM_ORDER}

        if grad_hook is not None:
            grad_hook(g_a, g_c)

        if not post_update_critic:
            self.step_c += 1                                       # ddpg.py:232
            for k in PARAM_ORDER:
                adam_step(self.critic[k], g_c[k], self.m_c[k], self.v_c[k], self.step_c,
                          self.lr, self.betas[0], self.betas[1], self.eps)
        self.step_a += 1                                           # ddpg.py:244
        for k in PARAM_ORDER:
            adam_step(self.actor[k], g_a[k], self.m_a[k], self.v_a[k], self.step_a,
                      self.lr, self.betas[0], self.betas[1], self.eps)
        for k in PARAM_ORDER:                                      # ddpg.py:250
            polyak(self.actor_target[k], self.actor[k], self.tau)
            polyak(self.critic_target[k], self.critic[k], self.tau)
        prio = (np.abs(td.detach().numpy()) + 1e-6).astype(F32)    # ddpg.py:253
        if ce_priority:
            prio = ((-(m_t * torch.log(q + 1e-10)).sum(dim=1)).detach().numpy() + F32(1e-6)).astype(F32)
        out.update(target_probs=tz.numpy(), q=q.detach().numpy(), m=m,
                   loss_critic=loss_c.detach().numpy(), loss_actor=loss_a.detach().numpy(),
                   td=td.detach().numpy(), prio=prio, grads_actor=g_a, grads_critic=g_c)
        return out


def flatten(wdict):
    return np.concatenate([wdict[k].detach().numpy().reshape(-1) for k in PARAM_ORDER])
