"""One or two full steps of the headline workload, for ncu captures (never a bench number)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
from bench import make_positions, CUTOFF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dt = np.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else np.float64
X, C, L = make_positions(n, 10)
Xd = torch.from_numpy(X.astype(dt)).cuda()
for _ in range(steps):
    cl = nl.build_cell_list(Xd, CUTOFF, C, (True, True, True))
    pl = nl.materialize_pairlist(cl, with_R=True)
    torch.cuda.synchronize()
    print(nl.npairs(pl))
    del pl, cl
