"""What a slab shard costs on ONE GPU: 10 M atoms confined to 1/G of a box that holds G x 10 M atoms at the same density
(the local problem of rank r in a G-rank run, without halo): stage times vs G."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
n = 10_000_000
import sys as _s
WINDOW = len(_s.argv) > 1 and _s.argv[1] == "window"  # pass the active z planes (what sharded.py does for z slabs)
for G in (1, 2, 8):
    L = (n * G / 0.05) ** (1 / 3)
    rng = np.random.Generator(np.random.PCG64(10))
    X = rng.random((n, 3)); X[:, 2] /= G; X *= L
    C = np.eye(3) * L
    Xd = torch.from_numpy(X).cuda()
    best = [1e9] * 4
    for it in range(5):
        tm = {}
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); cl = nl.build_cell_list(Xd, 5.0, C, (True, True, False)); e1.record()
        pa = None
        if WINDOW:
            nz = int(cl.ncells[2]); pa = np.zeros(nz, np.uint8); pa[:min(nz, int(np.ceil(nz / G)) + 1)] = 1
        pl = nl.materialize_pairlist(cl, with_R=True, timers=tm, plane_active=pa); e2.record(); torch.cuda.synchronize()
        ev = tm["events"][0]
        t = [e0.elapsed_time(e1), ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]), e0.elapsed_time(e2)]
        if it: best = [min(a, b) for a, b in zip(best, t)]
        P = nl.npairs(pl); del pl, cl
    print(f"{'window ' if WINDOW else ''}G={G}: ncells {int(L // 5)}^3 pairs {P} build {best[0]:.2f} count {best[1]:.2f} fill {best[2]:.2f} step {best[3]:.2f} ms", flush=True)
