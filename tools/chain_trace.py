"""python tools/chain_trace.py : %globaltimer phase timeline (ns) of CTA 0 of the two mlp_chain launches of one step.
Stamps per layer slot: 0 slot start, 1 cluster wait over, 2 operands landed (cp.async wait + bar), 3 FMA loop done,
4 reduce + epilogue done, 5 cluster arrive issued."""
import ctypes as C, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["D4PG_TC_TRACE"] = "1"
import numpy as np
import torch
import d4pg_b200 as d4pg
from d4pg_b200 import _lib
B, S, A, N = 256, 17, 6, 51
info = {"type": "categorical", "v_min": -50.0, "v_max": 0.0, "n_atoms": N}
n = 1 << 16
dd = d4pg.DDPG(S, A, memory_size=n, batch_size=B, critic_dist_info=info, sampling="device")
dd.assign_global_optimizer(d4pg.SharedAdam(dd.actor.parameters(), lr=1e-3), d4pg.SharedAdam(dd.critic.parameters(), lr=1e-3))
rng = np.random.RandomState(0)
dd.replayBuffer.add_batch(rng.randn(n, S).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
                          -rng.rand(n), rng.randn(n, S).astype(np.float32), np.zeros(n, bool))
dd.train_n(20)
torch.cuda.synchronize()
out = (C.c_ulonglong * 96)()
_lib.check(_lib.lib().d4pg_debug_trace_read(out, 96), "trace")
names = ["start", "wait_over", "landed", "fma_done", "epi_done", "arrived"]
for base, nm, ns in ((0, "forward chain T (8 slots)", 8), (48, "backward chain C (3 slots)", 3)):
    t0 = out[base]
    print(nm)
    for l in range(ns):
        st = [out[base + 6 * l + i] for i in range(6)]
        print("  slot %d @%6d ns: " % (l, st[0] - t0) + "  ".join("%s +%d" % (names[i], st[i] - st[i - 1]) for i in range(1, 6)))
    tot = out[base + 6 * (ns - 1) + 5] - t0
    print("  total %d ns, %d SM cycles -> %.0f MHz" % (tot, out[base + 47], out[base + 47] / max(tot, 1) * 1e3))
