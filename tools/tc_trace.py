"""D4PG_TC_TRACE=1 python tools/tc_trace.py : phase timeline (ns) of CTA 0 of one gemm_tc2 launch."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["D4PG_TC_TRACE"] = "1"
import torch
import d4pg_b200 as d4pg
from d4pg_b200 import _lib
names = ["start", "setup_done", "tma_c0_issued", "tma_all_issued", "conv_full0", "conv_done0", "mma_c0_committed",
         "mma_all_issued", "conv_loop_done", "done_wait_over", "epilogue_done", "dealloc_done"]
for (S, A, B) in ((16, 8, 256), (256, 8, 256)):
    a = d4pg.actor(S, A); a.precision = 1
    x = torch.randn(B, S, device="cuda")
    for rep in range(3):
        y = a(x); torch.cuda.synchronize()
    out = (C.c_ulonglong * 16)()
    _lib.check(_lib.lib().d4pg_debug_tc_trace(out), "trace")
    t0 = out[0]
    print("actor(%d,%d) B=%d last launch (fc3: K=256, N=%d):" % (S, A, B, A))
    for i, n in enumerate(names):
        print("  %-18s %8d ns" % (n, out[i] - t0 if out[i] else -1))
