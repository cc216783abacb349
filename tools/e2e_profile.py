"""python tools/e2e_profile.py : wall-clock split of the e2e loop of bench.py (add_batch / train / last_losses) + cProfile."""
import cProfile, os, pstats, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import d4pg_b200 as d4pg
import bench
cfg = bench.CFG["c2"]
info = {"type": "categorical", "v_min": cfg["v_min"], "v_max": cfg["v_max"], "n_atoms": cfg["atoms"]}
B, cap = cfg["batch"], cfg["cap"]
sampling = os.environ.get("SAMPLING", "reference")
LAG = int(os.environ.get("LAG", "1"))
dd = d4pg.DDPG(cfg["obs"], cfg["act"], memory_size=cap, batch_size=B, critic_dist_info=info, sampling=sampling)
dd.assign_global_optimizer(d4pg.SharedAdam(dd.actor.parameters(), lr=1e-3), d4pg.SharedAdam(dd.critic.parameters(), lr=1e-3))
dd.replayBuffer.add_batch(*bench.synth(cfg, cap, seed=0))
S, A_, R, S2, D = bench.synth(cfg, B * 8, seed=100)
pin = [torch.from_numpy(x).pin_memory() for x in (S, A_, R, S2, D)]
t = {"add": 0.0, "train": 0.0, "loss": 0.0}
def step(i, acc=True):
    lo = (i % 8) * B
    a = time.perf_counter()
    dd.replayBuffer.add_batch(*[p[lo:lo + B] for p in pin])
    b = time.perf_counter()
    dd.train()
    c = time.perf_counter()
    if LAG and i > 0: dd.last_losses(lag=1)
    elif not LAG: dd.last_losses()
    d = time.perf_counter()
    if acc:
        t["add"] += b - a; t["train"] += c - b; t["loss"] += d - c
for i in range(50): step(i, False)
dd.last_losses()
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for i in range(N): step(i)
dd.last_losses()
tot = time.perf_counter() - t0
print("sampling=%s  %.1f us/step: add_batch %.1f  train %.1f  last_losses(sync) %.1f" % (sampling, tot / N * 1e6, t["add"] / N * 1e6, t["train"] / N * 1e6, t["loss"] / N * 1e6))
pr = cProfile.Profile(); pr.enable()
for i in range(500): step(i, False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
