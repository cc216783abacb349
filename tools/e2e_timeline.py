"""python tools/e2e_timeline.py : %globaltimer timeline (us) of the LAST step of bench.py's e2e loop (add_batch -> train ->
last_losses(lag=1)) with the host pipeline: ingest-stream kernels (gate, ring write, tree add, sample) and the step graph."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["D4PG_TC_TRACE"] = "1"
import numpy as np
import torch
import d4pg_b200 as d4pg
from d4pg_b200 import _lib
import bench
cfg = bench.CFG["c2"]
info = {"type": "categorical", "v_min": cfg["v_min"], "v_max": cfg["v_max"], "n_atoms": cfg["atoms"]}
B, cap = cfg["batch"], cfg["cap"]
dd = d4pg.DDPG(cfg["obs"], cfg["act"], memory_size=cap, batch_size=B, critic_dist_info=info, precision=os.environ.get("PRECISION", "tf32x3"))
dd.assign_global_optimizer(d4pg.SharedAdam(dd.actor.parameters(), lr=1e-3), d4pg.SharedAdam(dd.critic.parameters(), lr=1e-3))
dd.replayBuffer.add_batch(*bench.synth(cfg, cap, seed=0))
S, A_, R, S2, D = bench.synth(cfg, B * 8, seed=100)
pin = [torch.from_numpy(x).pin_memory() for x in (S, A_, R, S2, D)]
N = 400
for i in range(N):
    lo = (i % 8) * B
    dd.replayBuffer.add_batch(*[p[lo:lo + B] for p in pin])
    dd.train()
    if i: dd.last_losses(lag=1)
    if i == N - 101: torch.cuda.synchronize(); t0 = time.perf_counter()
torch.cuda.synchronize()
period = (time.perf_counter() - t0) / 100 * 1e6
out = (C.c_ulonglong * 256)()
_lib.check(_lib.lib().d4pg_debug_trace_read(out, 256), "trace")
names = {10: "gate wait (in tree add) [ing]", 11: "ring write [ing]", 12: "tree add [ing]", 0: "sample [ing]", 1: "fwd chains", 2: "heads",
         3: "tree update [side]", 5: "dX chains", 6: "dW", 7: "adam"}
ev = [(out[96 + k], out[112 + k], names[k]) for k in names if out[96 + k]]
t1 = [e[0] for e in ev if e[2] == "fwd chains"][0]
print("e2e period %.1f us/step (wall, last 100 steps)" % period)
for a, b, nm in sorted(ev):
    if b < a:
        print("%-30s start %8.2f us" % (nm, (a - t1) / 1e3))
    else:
        print("%-30s start %8.2f us   end %8.2f us   (CTA0 runs %6.2f us)" % (nm, (a - t1) / 1e3, (b - t1) / 1e3, (b - a) / 1e3))
