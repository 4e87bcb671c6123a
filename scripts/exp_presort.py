"""Diagnostic: stage times with atoms in random vs spatially pre-sorted input order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
from bench import make_positions, CUTOFF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
X, C, L = make_positions(n, 10)
for mode in ("random", "presorted"):
    Xm = X
    if mode == "presorted":
        nc = int(L // CUTOFF)
        c = np.floor(X / L * nc).astype(np.int64)
        key = c[:, 0] + nc * (c[:, 1] + nc * c[:, 2])
        Xm = X[np.argsort(key, kind="stable")]
    Xd = torch.from_numpy(Xm).cuda()
    for with_R in (True, False):
        ts = []
        for it in range(4):
            tm = {}
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            cl = nl.build_cell_list(Xd, CUTOFF, C, (True, True, True))
            e1.record()
            pl = nl.materialize_pairlist(cl, with_R=with_R, timers=tm)
            torch.cuda.synchronize()
            ev = tm["events"][0]
            ts.append((e0.elapsed_time(e1), ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])))
            del pl, cl
        print(mode, "with_R" if with_R else "no_R", "build/count/fill ms:", [round(v, 3) for v in np.min(np.array(ts[1:]), axis=0)])
