"""Diagnostic: single-GPU cost of one rank's local problem when it is a z-slab of a G-times larger global box
(same local atom count), vs. its own box.  No communication involved."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
n = 10_000_000
for G in (1, 2, 8):
    L = (n * G / 0.05) ** (1 / 3)
    rng = np.random.Generator(np.random.PCG64(10))
    X = rng.random((n, 3))
    X[:, 2] = X[:, 2] / G          # rank 0's slab (halo omitted: only the cost structure matters here)
    X *= L
    C = np.eye(3) * L
    Xd = torch.from_numpy(X).cuda()
    best = None
    for it in range(4):
        tm = {}
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); cl = nl.build_cell_list(Xd, 5.0, C, (True, True, True)); e1.record()
        pl = nl.materialize_pairlist(cl, with_R=True, timers=tm)
        torch.cuda.synchronize()
        ev = tm["events"][0]
        t = (e0.elapsed_time(e1), ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]))
        best = t if best is None else tuple(min(a, b) for a, b in zip(best, t))
        P = nl.npairs(pl); del pl, cl
    print(f"G={G} ncells={int(L//5)}^3 pairs={P} build/count/fill ms: {[round(v,2) for v in best]}", flush=True)
