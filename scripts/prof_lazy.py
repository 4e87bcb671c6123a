import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
N = 10_000_000
rng = np.random.Generator(np.random.PCG64(10))
L = (N / 0.05) ** (1 / 3)
X = (rng.random((N, 3)) * L).astype(np.float32)
C = (np.eye(3) * L).astype(np.float32)
clist = nl.neighbour_list(torch.from_numpy(X).cuda(), 6.0, C, (True, True, True), lazy=True)
for _ in range(2):
    e = nl.lj_energy(clist, 1.0, 3.4); torch.cuda.synchronize()
print(float(e.item()))
