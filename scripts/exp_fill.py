"""Fill-kernel timings at the headline size for the current library: with and without R (experiments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
from bench import make_positions, CUTOFF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
X, C, L = make_positions(n, 10)
Xd = torch.from_numpy(X).cuda()
tag = os.environ.get("EXP_TAG", "")
for with_R in (True, False):
    best = [1e9, 1e9, 1e9]
    for it in range(5):
        timers = {}
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        cl = nl.build_cell_list(Xd, CUTOFF, C, (True, True, True))
        pl = nl.materialize_pairlist(cl, with_R=with_R, timers=timers)
        e1.record(); torch.cuda.synchronize()
        ev = timers["events"][-1]
        t = [e0.elapsed_time(e1), ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])]
        if it >= 1:
            best = [min(a, b) for a, b in zip(best, t)]
        chk = int(pl.j.sum().item()) ^ int(pl.S.sum().item())
        del pl, cl
    print(f"{tag} with_R={with_R}: step {best[0]:.2f} ms  count {best[1]:.2f}  fill {best[2]:.2f}  chk {chk}", flush=True)
