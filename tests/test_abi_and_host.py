"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the host-side
mirror of the reference interface behaves, and the DP plumbing works over gloo (world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "d4pg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(d4pg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import d4pg_b200
    from d4pg_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        d4pg_b200.build()
    L = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), "missing export: " + name
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, set(_lib.EXPORTED_SYMBOLS) ^ set(declared)
    assert L.d4pg_version() >= 100


def test_struct_mirrors_have_the_c_sizes():
    """The ctypes mirrors of the structs passed by pointer must have exactly the C sizes (a stale mirror would make
    the library read garbage configuration)."""
    import ctypes as C
    from d4pg_b200 import _lib
    L = _lib.lib()
    assert L.d4pg_struct_size(0) == C.sizeof(_lib.LearnerConfig)
    assert L.d4pg_struct_size(1) == C.sizeof(_lib.LearnerBuffers)
    assert L.d4pg_struct_size(2) == C.sizeof(_lib.NetLayout)
    assert L.d4pg_struct_size(99) == -1


def test_no_compute_without_gpu_fails_loudly():
    import d4pg_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = d4pg_b200.actor(5, 2)
    with pytest.raises(d4pg_b200.D4PGError):
        a(torch.zeros(1, 5))
    with pytest.raises(d4pg_b200.D4PGError):
        d4pg_b200.PrioritizedReplayBuffer(8, 0.6).add(np.zeros(3), np.zeros(1), 0.0, np.zeros(3), False)


def test_layouts_match_between_python_mirror_and_c():
    import d4pg_b200
    from d4pg_b200 import _lib
    for (s, a, n) in ((17, 6, 51), (376, 17, 51), (3, 1, 51), (17, 6, 101)):
        act = d4pg_b200.actor(s, a, device="cpu")
        cri = d4pg_b200.critic(s, a, {"type": "categorical", "v_min": -1., "v_max": 1., "n_atoms": n}, device="cpu")
        assert (act._offsets, act._sizes, act._total, act._pitch) == _lib.actor_layout(s, a)
        assert (cri._offsets, cri._sizes, cri._total, cri._pitch) == _lib.critic_layout(s, a, n)
        assert all(p % 4 == 0 for p in cri._pitch) and cri._pitch[1] >= 256 + a
        assert sum(p.numel() for p in act.parameters()) == s * 256 + 256 + 2 * (256 * 256 + 256) + 256 * a + a


def test_module_state_dict_keys_and_flat_aliasing():
    import d4pg_b200
    torch.manual_seed(0)
    a = d4pg_b200.actor(17, 6, device="cpu")
    assert list(a.state_dict().keys()) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias",
                                           "fc2_2.weight", "fc2_2.bias", "fc3.weight", "fc3.bias"]
    assert a.fc2.weight.shape == (256, 256) and a.fc3.weight.shape == (6, 256)
    a.fc1.bias.data.fill_(7.0)
    off = a._offsets[1]
    assert torch.all(a.flat_params()[off:off + 256] == 7.0)
    b = d4pg_b200.actor(17, 6, device="cpu")
    b.load_state_dict(a.state_dict())
    assert torch.equal(a.flat_params(), b.flat_params())
    b.adopt_flat(a.flat_params())
    a.fc3.weight.data.zero_()
    assert torch.all(b.fc3.weight == 0)


def test_linear_schedule_and_aliases():
    import d4pg_b200
    s = d4pg_b200.LinearSchedule(10, final_p=1.0, initial_p=0.4)
    vals = [s.value() for _ in range(12)]
    assert vals[0] == 0.4 and abs(vals[5] - 0.7) < 1e-12 and vals[-1] == 1.0 and s.t == 12
    assert d4pg_b200.ReplayMemory is d4pg_b200.Replay
    assert d4pg_b200.PrioritizedReplayMemory is d4pg_b200.PrioritizedReplayBuffer
    d4pg_b200.install_reference_aliases()
    import ddpg, shared_adam          # noqa: E401  (the reference's module names now bind here)
    assert ddpg.DDPG is d4pg_b200.DDPG and shared_adam.SharedAdam is d4pg_b200.SharedAdam
    for n in ("ddpg", "models", "shared_adam", "prioritized_replay_memory", "replay_memory", "utils", "random_process"):
        sys.modules.pop(n, None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "d4pg-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f


def test_shard_range_partitions():
    from d4pg_b200 import dist
    for n in (7, 8, 1000003):
        for w in (1, 2, 3, 8):
            spans = [dist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


_GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["D4PG_ROOT"])
from d4pg_b200 import dist as ddist
from oracle import d4pg_oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
# 1. unique-id style byte broadcast over the default group
payload = bytes(range(128)) if rank == 0 else bytes(128)
got = ddist.broadcast_bytes(payload, 128, src=0)
assert got == bytes(range(128))
# 2. synchronous DP == one big batch: each rank trains on its shard with gradient averaging
info = {"type": "categorical", "v_min": -50.0, "v_max": 0.0, "n_atoms": 51}
torch.manual_seed(0)
a0, c0 = O.init_actor(5, 2), O.init_critic(5, 2, 51)
rng = np.random.RandomState(0)
B = 16
S = rng.randn(2 * B, 5).astype(np.float32); A = rng.uniform(-1, 1, (2 * B, 2)).astype(np.float32)
R = -3 * rng.rand(2 * B); S2 = rng.randn(2 * B, 5).astype(np.float32); D = rng.rand(2 * B) < 0.1
lo, hi = ddist.shard_range(2 * B, rank, 2)
mine = O.LearnerOracle(5, 2, info, actor_w={k: v.clone() for k, v in a0.items()}, critic_w={k: v.clone() for k, v in c0.items()})
def hook(ga, gc):
    for gd in (ga, gc):
        for k in gd:
            dist.all_reduce(gd[k]); gd[k] /= 2
mine.train_step(S[lo:hi], A[lo:hi], R[lo:hi], S2[lo:hi], D[lo:hi], grad_hook=hook)
big = O.LearnerOracle(5, 2, info, actor_w={k: v.clone() for k, v in a0.items()}, critic_w={k: v.clone() for k, v in c0.items()})
big.train_step(S, A, R, S2, D)
for k in O.PARAM_ORDER:
    assert (mine.actor[k] - big.actor[k]).abs().max() < 2e-6, k
    assert (mine.critic[k] - big.critic[k]).abs().max() < 2e-6, k
# replicas bit-identical across ranks
flat = torch.from_numpy(O.flatten(mine.critic)).clone()
other = flat.clone(); dist.broadcast(other, src=0)
assert torch.equal(flat, other)
dist.destroy_process_group()
print("OK", rank)
'''


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port), D4PG_ROOT=ROOT, OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0 and "OK" in out, out
