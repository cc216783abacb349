"""python tools/rows_trace.py : %globaltimer timeline (ns) of CTA 0 of the two mlp_rows launches of one step
(2 stamps per layer: before its first weight stage, after its last)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["D4PG_TC_TRACE"] = "1"
import numpy as np
import torch
import d4pg_b200 as d4pg
from d4pg_b200 import _lib
B, S, A, N = 256, 17, 6, 51
info = {"type": "categorical", "v_min": -50.0, "v_max": 0.0, "n_atoms": N}
n = 1 << 16
dd = d4pg.DDPG(S, A, memory_size=n, batch_size=B, critic_dist_info=info, sampling="device", chain="rows")
dd.assign_global_optimizer(d4pg.SharedAdam(dd.actor.parameters(), lr=1e-3), d4pg.SharedAdam(dd.critic.parameters(), lr=1e-3))
rng = np.random.RandomState(0)
dd.replayBuffer.add_batch(rng.randn(n, S).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
                          -rng.rand(n), rng.randn(n, S).astype(np.float32), np.zeros(n, bool))
dd.train_n(20)
torch.cuda.synchronize()
out = (C.c_ulonglong * 64)()
_lib.check(_lib.lib().d4pg_debug_trace_read(out, 64), "trace")
for base, nm, nl in ((0, "forward chain T (8 layers)", 8), (24, "backward chain C (3 layers)", 3)):
    t0 = out[base]
    print(nm)
    for l in range(nl):
        print("  layer %d: start @%6d ns  stages %5d ns" % (l, out[base + 2 * l] - t0, out[base + 2 * l + 1] - out[base + 2 * l]))
    print("  total %d ns" % (out[base + 2 * nl - 1] - t0))
    print("  SM cycles of thread 0: wait-for-weights %d  compute %d  epilogue+sync %d | producer lane 0: wait-for-empty %d  issue %d" % tuple(out[base + 16 + i] for i in range(5)))
