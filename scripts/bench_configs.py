"""Timings of the BASELINE configs other than the headline (device-resident, min of 5 after 2 warm-ups)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
from tests import util as U

def timeit(X, cutoff, cell, pbc, int_type=np.int32, with_R=False, reps=5):
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    best = 1e9; P = 0
    for it in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pl = nl.neighbour_list(Xd, cutoff, cell, pbc, int_type=int_type, with_R=with_R)
        e1.record(); torch.cuda.synchronize()
        if it >= 2: best = min(best, e0.elapsed_time(e1))
        P = nl.npairs(pl); del pl
    return best, P

rows = []
X, C, L = U.rand_config(10000, seed=1)
rows.append(("C1 10k atoms cubic pbc TTT F64 (reference GPU on A4500: 2.4 ms)", *timeit(X, 5.0, C, (True,)*3)))
X, C = U.fcc(3.61)
rows.append(("C2 fcc Cu 4x4x4 rc=5 F64", *timeit(X, 5.0, C, (True,)*3)))
s = (1e6 / 0.05 / 720.0) ** (1 / 3); cell = s * U.TRICLINIC
rows.append(("C3 1M atoms triclinic pbc TTF F64", *timeit(U.rand_in_cell(1_000_000, cell, seed=3), 5.0, cell, (True, True, False))))
rng = np.random.Generator(np.random.PCG64(10)); N = 10_000_000; L = (N / 0.05) ** (1 / 3); Xh = rng.random((N, 3)) * L; C = np.eye(3) * L
rows.append(("headline 10M F64/I32 no R (reference PairList layout, 20 B/pair)", *timeit(Xh, 5.0, C, (True,)*3)))
rows.append(("headline 10M F64/I32 with R (44 B/pair)", *timeit(Xh, 5.0, C, (True,)*3, with_R=True)))
rows.append(("headline 10M F64/I64 with R (64 B/pair)", *timeit(Xh, 5.0, C, (True,)*3, int_type=np.int64, with_R=True)))
rows.append(("headline 10M F32/I32 with R (32 B/pair)", *timeit(Xh.astype(np.float32), 5.0, C.astype(np.float32), (True,)*3, with_R=True)))
for name, ms, P in rows:
    print(f"| {name} | {P} | {ms:.3f} ms | {P / ms / 1e6:.2f} G pairs/s |")
