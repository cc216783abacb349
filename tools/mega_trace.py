"""Phase timeline (ns) of CTA 0 of the persistent step kernel (D4PG_TC_TRACE=1)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["D4PG_TC_TRACE"] = "1"
import numpy as np, torch
import d4pg_b200 as d4pg
from d4pg_b200 import _lib
info = {"type": "categorical", "v_min": -50.0, "v_max": 0.0, "n_atoms": 51}
dd = d4pg.DDPG(17, 6, memory_size=1 << 20, batch_size=256, critic_dist_info=info, sampling="device", persistent=True)
dd.assign_global_optimizer(d4pg.SharedAdam(dd.actor.parameters()), d4pg.SharedAdam(dd.critic.parameters()))
rng = np.random.RandomState(0)
n = 1 << 20
dd.replayBuffer._store.add_batch(rng.randn(n, 17).astype(np.float32), rng.uniform(-1, 1, (n, 6)).astype(np.float32), -rng.rand(n), rng.randn(n, 17).astype(np.float32), np.zeros(n, bool))
dd.train_n(50); torch.cuda.synchronize()
out = (C.c_ulonglong * 32)()
_lib.check(_lib.lib().d4pg_debug_tc_trace(out), "trace")
names = ["start", "sample_done(cta0)", "bar0"] + ["fwd%d+bar" % i for i in range(1, 8)] + ["heads_done(cta0)", "bar_heads"] + ["bwd%d+bar" % i for i in range(1, 8)] + ["adam_done(cta0)"]
prev = out[0]
for i, nme in enumerate(names):
    if out[i]:
        print("%-20s t=%7d ns  (+%6d)" % (nme, out[i] - out[0], out[i] - prev)); prev = out[i]
