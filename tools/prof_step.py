"""A few eager (non-graph) config-2 learner steps, for ncu:  ncu --set full --import-source on -k regex:<kernel> ... python tools/prof_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import d4pg_b200 as d4pg
B, S, A, N = 256, 17, 6, 51
info = {"type": "categorical", "v_min": -50.0, "v_max": 0.0, "n_atoms": N}
n = 1 << 20
chain = int(os.environ.get("CHAIN", "1"))
dd = d4pg.DDPG(S, A, memory_size=n, batch_size=B, critic_dist_info=info, sampling="device", use_graph=False,
               chain={0: "levels", 1: "cluster"}[chain], precision=os.environ.get("PRECISION", "fp32"))
dd.assign_global_optimizer(d4pg.SharedAdam(dd.actor.parameters(), lr=1e-3), d4pg.SharedAdam(dd.critic.parameters(), lr=1e-3))
rng = np.random.RandomState(0)
dd.replayBuffer.add_batch(rng.randn(n, S).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
                          (-3 * rng.rand(n)).astype(np.float32).astype(np.float64), rng.randn(n, S).astype(np.float32), np.zeros(n, bool))
for _ in range(int(os.environ.get("STEPS", "4"))):
    dd.train()
torch.cuda.synchronize()
print("ok", dd.last_losses(), dd.kernels_per_step())
