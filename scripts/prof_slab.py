import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
n, G = 10_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 8
L = (n * G / 0.05) ** (1 / 3)
rng = np.random.Generator(np.random.PCG64(10))
X = rng.random((n, 3)); X[:, 2] /= G; X *= L
C = np.eye(3) * L
Xd = torch.from_numpy(X).cuda()
for _ in range(2):
    cl = nl.build_cell_list(Xd, 5.0, C, (True, True, True)); pl = nl.materialize_pairlist(cl, with_R=True); torch.cuda.synchronize(); del pl, cl
