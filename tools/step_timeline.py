"""python tools/step_timeline.py : %globaltimer timeline (us) of one steady-state config-2 learner step:
entry / exit of CTA 0 of every kernel, main branch and prefetch side branch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["D4PG_TC_TRACE"] = "1"
import numpy as np
import torch
import d4pg_b200 as d4pg
from d4pg_b200 import _lib
B, S, A, N = 256, 17, 6, 51
info = {"type": "categorical", "v_min": -50.0, "v_max": 0.0, "n_atoms": N}
n = 1 << 20
dd = d4pg.DDPG(S, A, memory_size=n, batch_size=B, critic_dist_info=info, sampling="device", precision=os.environ.get("PRECISION", "fp32"))
dd.assign_global_optimizer(d4pg.SharedAdam(dd.actor.parameters(), lr=1e-3), d4pg.SharedAdam(dd.critic.parameters(), lr=1e-3))
rng = np.random.RandomState(0)
dd.replayBuffer.add_batch(rng.randn(n, S).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
                          -rng.rand(n), rng.randn(n, S).astype(np.float32), np.zeros(n, bool))
dd.train_n(50)
torch.cuda.synchronize()
out = (C.c_ulonglong * 256)()
_lib.check(_lib.lib().d4pg_debug_trace_read(out, 256), "trace")
names = {0: "sample (cold)", 1: "fwd chains", 2: "heads", 3: "tree update [side]", 4: "sample t+1 [side]", 5: "dX chains", 6: "dW", 7: "adam"}
ev = [(out[96 + k], out[112 + k], names[k]) for k in names if out[96 + k]]
t0 = min(e[0] for e in ev if e[2] != "sample (cold)")
for a, b, nm in sorted(ev):
    if nm == "sample (cold)": continue
    print("%-22s start %8.2f us   CTA0 runs %7.2f us" % (nm, (a - t0) / 1e3, (b - a) / 1e3))

if out[96 + 24]:
    print("tree update phases: prefetch+barrier %.2f us, pow/max %.2f us, level walk %.2f us" % (
        (out[96 + 24] - out[96 + 3]) / 1e3, (out[96 + 25] - out[96 + 24]) / 1e3, (out[112 + 3] - out[96 + 25]) / 1e3))
print("per-slot stamps of CTA %s of each chain launch (ns): start / wait_over / landed / fma_done / epi_done / arrived" % os.environ.get("D4PG_TRACE_CTA", "0"))
for base, nm in ((0, "fwd launch"), (48, "dX launch")):
    t0 = out[base]
    for l in range(8):
        st = [out[base + 6 * l + i] for i in range(6)]
        if l and st[0] <= out[base + 6 * (l - 1)]: break
        if st[0] == 0: break
        print("  %s slot %d @%6d: " % (nm, l, st[0] - t0) + " ".join("+%d" % (st[i] - st[i - 1]) for i in range(1, 6) if st[i] > st[i-1] and st[i] - st[i-1] < 10**8))

import collections
raw = bytes(C.cast(out, C.POINTER(C.c_ubyte * (256 * 8))).contents)[128 * 8:128 * 8 + 192]
sm = collections.defaultdict(list)
for cta, s_ in enumerate(raw):
    sm[s_].append("TPQ"[cta // 64])
pat = collections.Counter("".join(sorted(v)) for v in sm.values())
print("forward launch: CTAs per SM by chain (T = target, Q = critic, P = policy):", dict(pat), " SMs used:", len(sm))
