"""python tools/tcc_trace.py : %globaltimer phase timeline (ns) of one CTA (D4PG_TRACE_CTA, default 0) of the two
mlp_tc_chain launches of one step.  Stamps per layer slot: 0 loader: slot start (cluster wait over), 1 loader: A loads
issued, 2 MMA: weights landed, 3 MMA: first A chunk landed, 4 MMA: all MMAs issued, 5 epilogue: accumulator complete,
6 epilogue: stores + fence done, 7 cluster barrier passed."""
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
dd = d4pg.DDPG(S, A, memory_size=n, batch_size=B, critic_dist_info=info, sampling="device", precision=os.environ.get("PRECISION", "tf32x3"))
dd.assign_global_optimizer(d4pg.SharedAdam(dd.actor.parameters(), lr=1e-3), d4pg.SharedAdam(dd.critic.parameters(), lr=1e-3))
rng = np.random.RandomState(0)
dd.replayBuffer.add_batch(rng.randn(n, S).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
                          -rng.rand(n), rng.randn(n, S).astype(np.float32), np.zeros(n, bool))
dd.train_n(20)
torch.cuda.synchronize()
out = (C.c_ulonglong * 512)()
_lib.check(_lib.lib().d4pg_debug_trace_read(out, 512), "trace")
names = ["start", "a_issued", "w_landed", "a0_landed", "mma_issued", "acc_done", "tmem_ld", "math", "img_st", "epi_end", "barrier", "-"]
for base, nm, ns in ((256, "forward launch, traced CTA's chain", 8), (384, "backward launch, traced CTA's chain", 8)):
    t0 = out[base]
    print(nm)
    for l in range(ns):
        st = [out[base + 12 * l + i] for i in range(12)]
        if st[0] == 0:
            break
        print("  slot %d @%6d ns: " % (l, st[0] - t0) + "  ".join("%s %+d" % (names[i], st[i] - st[0]) for i in range(1, 11) if st[i]))
st = [out[96 + i] for i in range(32)]
print("step timeline (entry, exit) ns relative to the forward launch:")
for k, nm in ((0, "sample"), (1, "fwd chains"), (2, "heads"), (3, "tree update"), (4, "sample t+1"), (5, "dX chains"), (6, "dW"), (7, "adam")):
    if st[k] and st[16 + k]:
        print("  %-12s start %8.2f us  runs %7.2f us" % (nm, (st[k] - st[1]) / 1e3, (st[16 + k] - st[k]) / 1e3))
