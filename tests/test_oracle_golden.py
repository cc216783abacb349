"""oracle/ vs the committed golden fixtures (generated from the unmodified
reference by tests/golden/make_golden.py).  CPU only; runs everywhere."""
import numpy as np
import pytest
import torch

from oracle import d4pg_oracle as O
from tests import helpers as H


def test_projection_kats_and_random_batches_bit_exact():
    g = H.load("projection.npz")
    for k in ("kat_nt", "kat_t", "kat_ti"):
        m = O.project_live(g[k + "_probs"], g[k + "_r"], g[k + "_done"], -50.0, 0.0, 51, 0.99)
        assert np.array_equal(m, g[k + "_m"]), k
    # terminal r=-1.6346495489906907 -> bins (48,49)
    m = g["kat_t_m"]
    assert abs(m[0, 48] - 0.6346496) < 1e-6 and abs(m[0, 49] - 0.36535046) < 1e-6
    for c in range(int(g["n_rand"])):
        k = "rand%d" % c
        v_min, v_max, N, gamma = g[k + "_meta"]
        m = O.project_live(g[k + "_probs"], g[k + "_r"], g[k + "_done"], float(v_min), float(v_max),
                           int(N), float(gamma))
        assert m.dtype == np.float32
        assert np.array_equal(m, g[k + "_m"]), k


def test_projection_nstep_matches_reference():
    g = H.load("projection.npz")
    for i in range(2):
        k = "nstep%d" % i
        m = O.project_nstep(g[k + "_probs"], g[k + "_r"], g[k + "_done"], -150.0, 150.0, 101, 0.99, 5)
        assert np.array_equal(m, g[k + "_m"]), k


def test_projection_nstep_config5_size_matches_reference():
    """B=4096, 101 atoms, n_steps=5 (BASELINE.json configs[4]) against ddpg.py:122-140 run at that size."""
    g, p, r, done = H.projection_c5_inputs()
    m = O.project_nstep(p, r, done, -150.0, 150.0, 101, 0.99, 5)
    H.check_compact(g, "m", m, 0.0)


def test_nstep_accumulation_at_insert_matches_reference_initialize():
    """replay_memory.py:21-59 run on a scripted env (tests/golden/make_golden.py:gen_nstep): the restatement of its
    n-step accumulation reproduces the reference buffer, returns bit for bit."""
    g = H.load("nstep_init.npz")
    n_steps, init_length, n_buf, n_eps = [int(x) for x in g["meta"]]
    rows = []
    for i in range(n_eps):
        e = [g["ep%d_%s" % (i, k)] for k in ("s", "a", "r", "s2", "d")]
        rows += O.nstep_transitions(e[0], e[1], [float(x) for x in e[2]], e[3], e[4], n_steps, float(g["gamma"]))
    assert len(rows) == n_buf
    assert np.array_equal(np.stack([r[0] for r in rows]), g["buf_s"]) and np.array_equal(np.stack([r[1] for r in rows]), g["buf_a"])
    assert np.array_equal(np.array([r[2] for r in rows]), g["buf_r"])
    assert np.array_equal(np.stack([r[3] for r in rows]), g["buf_s2"]) and np.array_equal(np.array([bool(r[4]) for r in rows]), g["buf_d"])


def _replay(g, name):
    size, n_fill, B, rounds = [int(x) for x in g[name + "_meta"]]
    buf = O.PrioritizedReplayOracle(size, 0.6, 2, 1)
    for i in range(n_fill):
        buf.add(np.full(2, i, np.float32), np.zeros(1, np.float32), -1.0, np.zeros(2, np.float32), False)
    return buf, size, n_fill, B, rounds


@pytest.mark.parametrize("name", ["full", "part", "wrap"])
def test_tree_states_indices_weights(name):
    g = H.load("tree.npz")
    buf, size, n_fill, B, rounds = _replay(g, name)
    assert np.array_equal(buf.sum.value.astype(np.float64), g[name + "_sum_r0"])
    assert np.array_equal(buf.min.value.astype(np.float64), g[name + "_min_r0"])
    for k in range(rounds):
        idx = buf.sample_indices(g[name + "_u"][k])
        assert np.array_equal(idx, g[name + "_idx"][k]), (name, k)
        w = buf.is_weights(idx, float(g[name + "_beta"][k]))
        np.testing.assert_allclose(w, g[name + "_w"][k], rtol=2e-6, atol=0)
        buf.update_priorities(g[name + "_upd_idx"][k], g[name + "_upd_prio"][k])
        assert np.array_equal(buf.sum.value.astype(np.float64), g["%s_sum_r%d" % (name, k + 1)]), (name, k)
        assert np.array_equal(buf.min.value.astype(np.float64), g["%s_min_r%d" % (name, k + 1)]), (name, k)
        if k == 3:
            for j in range(7):
                buf.add(np.zeros(2, np.float32), np.zeros(1, np.float32), -1.0, np.zeros(2, np.float32), False)
            assert np.array_equal(buf.sum.value.astype(np.float64), g[name + "_sum_after_add"])
            assert np.array_equal(buf.min.value.astype(np.float64), g[name + "_min_after_add"])
    assert float(buf.max_priority) == float(g[name + "_max_priority"])


def test_tree_bulk_rebuild_equals_sequential_sets():
    rng = np.random.RandomState(0)
    a = O.SegmentTree32(64, "sum")
    b = O.SegmentTree32(64, "sum")
    vals = rng.rand(50).astype(np.float32)
    for i, v in enumerate(vals):
        a.set(i, v)
    b.value[64:64 + 50] = vals
    b.rebuild()
    assert np.array_equal(a.value, b.value)


def test_init_rng_parity_and_forward():
    g = H.load("init.npz")
    torch.manual_seed(5)
    a = O.init_actor(17, 6)
    c = O.init_critic(17, 6, 51)
    for k in H.NAMES:
        H.check_compact(g, "actor_" + k, a[k].numpy(), 0.0)
        H.check_compact(g, "critic_" + k, c[k].numpy(), 0.0)
    x, act = torch.from_numpy(g["x"]), torch.from_numpy(g["act"])
    assert np.array_equal(O.actor_forward(a, x).numpy(), g["actor_out"])
    assert np.array_equal(O.critic_forward(c, x, act).numpy(), g["critic_out"])


@pytest.mark.parametrize("tag", ["per_c2", "per_part", "uniform_c1", "per_c2_b256", "per_c3_b1024", "per_c5_b256"])
def test_full_train_steps_bit_exact(tag):
    """3 consecutive DDPG.train() steps: indices, projection, losses, priorities,
    tree, gradients, post-step parameters, targets, Adam moments."""
    g = H.load("train_%s.npz" % tag)
    obs_dim, act_dim, N, B, mem, n_fill, per, steps = [int(x) for x in g["meta"]]
    v_min, v_max = [float(x) for x in g["dist"]]
    a, c = H.regen_init(int(g["seed"]), obs_dim, act_dim, N)
    for k in H.NAMES:
        H.check_compact(g, "init_actor_" + k, a[k].numpy(), 0.0)
        H.check_compact(g, "init_critic_" + k, c[k].numpy(), 0.0)
    info = {"type": "categorical", "v_min": v_min, "v_max": v_max, "n_atoms": N}
    n_steps = int(g["n_steps"]) if "n_steps" in g.files else 1
    lo = O.LearnerOracle(obs_dim, act_dim, info, actor_w=a, critic_w=c, n_steps=n_steps)   # live projection: gamma (H5)
    buf = O.PrioritizedReplayOracle(mem, 0.6, obs_dim, act_dim)
    S, A, R, S2, D = H.train_data(g)
    for i in range(n_fill):
        buf.add(S[i], A[i], float(R[i]), S2[i], bool(D[i]))
    sched = O.LinearScheduleOracle(100000, 1.0, 0.4)
    torch.set_num_threads(1)
    for t in range(steps):
        if per:
            batch = buf.sample(B, sched.value(), g["u_%d" % t])
            idx = batch[6]
            assert np.array_equal(idx, g["idx_%d" % t])
        else:
            idx = g["idx_%d" % t]
        s, a_, r, s2, d = buf.encode(idx)
        out = lo.train_step(s, a_, r, s2, d)
        H.check_compact(g, "target_probs_%d" % t, out["target_probs"], 0.0)
        H.check_compact(g, "m_%d" % t, out["m"], 0.0)
        H.check_compact(g, "q_%d" % t, out["q"], 0.0)
        assert np.array_equal(out["loss_critic"], g["loss_critic_%d" % t])
        assert np.array_equal(out["loss_actor"], g["loss_actor_%d" % t])
        if per:
            assert np.array_equal(out["prio"], g["prio_%d" % t])
            buf.update_priorities(idx, out["prio"])
            assert np.array_equal(buf.sum.value.astype(np.float64), g["tree_sum_%d" % t])
            assert np.array_equal(buf.min.value.astype(np.float64), g["tree_min_%d" % t])
        for k in H.NAMES:
            H.check_compact(g, "g_critic_%s_%d" % (k, t), out["grads_critic"][k].numpy(), 0.0)
            H.check_compact(g, "g_actor_%s_%d" % (k, t), out["grads_actor"][k].numpy(), 0.0)
            H.check_compact(g, "actor_%s_%d" % (k, t), lo.actor[k].numpy(), 0.0)
            H.check_compact(g, "critic_%s_%d" % (k, t), lo.critic[k].numpy(), 0.0)
    t = steps - 1
    for k in H.NAMES:
        H.check_compact(g, "actor_target_%s_%d" % (k, t), lo.actor_target[k].numpy(), 0.0)
        H.check_compact(g, "critic_target_%s_%d" % (k, t), lo.critic_target[k].numpy(), 0.0)
        H.check_compact(g, "adam_m_actor_%s_%d" % (k, t), lo.m_a[k].numpy(), 0.0)
        H.check_compact(g, "adam_v_critic_%s_%d" % (k, t), lo.v_c[k].numpy(), 0.0)


def test_loss_terms_closed_form_gradient():
    rng = np.random.RandomState(3)
    z = torch.tensor(rng.randn(8, 51).astype(np.float32), requires_grad=True)
    q = torch.softmax(z, dim=1)
    m = torch.softmax(torch.tensor(rng.randn(8, 51).astype(np.float32)), dim=1)
    loss = -(m * torch.log(q + 1e-10)).sum(dim=1).mean()
    loss.backward()
    t = O.critic_loss_terms(m.numpy(), q.detach().numpy())
    np.testing.assert_allclose(t["dlogits"], z.grad.numpy(), atol=1e-7)
    np.testing.assert_allclose(t["loss"], loss.detach().numpy(), atol=1e-6)
