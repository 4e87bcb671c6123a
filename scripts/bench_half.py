"""Half list (NL_FLAG_HALF) vs full list at the headline size; min of 5, device-resident."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
from bench import make_positions, CUTOFF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
X, C, L = make_positions(n, 10)
Xd = torch.from_numpy(X).cuda()
out = {}
for half in (False, True):
    for with_R in (True, False):
        best = [1e9, 1e9, 1e9]
        for it in range(6):
            timers = {}
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            cl = nl.build_cell_list(Xd, CUTOFF, C, (True, True, True))
            pl = nl.materialize_pairlist(cl, with_R=with_R, timers=timers, half=half)
            e1.record(); torch.cuda.synchronize()
            ev = timers["events"][-1]
            t = [e0.elapsed_time(e1), ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])]
            if it >= 1:
                best = [min(a, b) for a, b in zip(best, t)]
            P = nl.npairs(pl)
            del pl, cl
        out[f"half={half},R={with_R}"] = dict(pairs=P, step_ms=round(best[0], 3), count_ms=round(best[1], 3), fill_ms=round(best[2], 3))
        print(f"half={half} with_R={with_R}: pairs {P} step {best[0]:.2f} ms count {best[1]:.2f} fill {best[2]:.2f}", flush=True)
print(json.dumps(out))
