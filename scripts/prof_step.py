"""One build_cell_list + materialize_pairlist step at the headline workload, repeated; for ncu captures.
usage: prof_step.py [n_atoms] [reps] [with_R 0/1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neighbourlists_jl_b200 as nl
from bench import make_positions, CUTOFF
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with_R = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
X, C, L = make_positions(n, 10)
Xd = torch.from_numpy(X).cuda()
for _ in range(reps):
    cl = nl.build_cell_list(Xd, CUTOFF, C, (True, True, True))
    pl = nl.materialize_pairlist(cl, with_R=with_R)
    torch.cuda.synchronize()
    print("pairs", nl.npairs(pl), flush=True)
    del pl, cl
