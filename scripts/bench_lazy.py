"""BASELINE config 5: lazy SortedCellList at 10 M atoms, Float32, rc = 6, fused LJ energy + fused count (no pair materialisation)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.Generator(np.random.PCG64(10))
L = (N / 0.05) ** (1 / 3)
X = (rng.random((N, 3)) * L).astype(np.float32)
C = (np.eye(3) * L).astype(np.float32)
Xd = torch.from_numpy(X).cuda()
res = {}
for name, fn in (("build_cell_list", lambda: nl.neighbour_list(Xd, 6.0, C, (True, True, True), lazy=True)),):
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); clist = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[name + "_ms"] = min(ts[2:])
for name, fn in (("lj_energy", lambda: nl.lj_energy(clist, 1.0, 3.4)), ("lj_forces", lambda: nl.lj_forces(clist, 1.0, 3.4)), ("count_neighbours", lambda: (setattr(clist, "_counts", None), nl.count_neighbours(clist))[1])):
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[name + "_ms"] = min(ts[2:])
pairs = int(nl.count_neighbours(clist).sum().item())
res.update(atoms=N, virtual_pairs=pairs, energy=float(nl.lj_energy(clist, 1.0, 3.4).item()),
           ljf_pairs_per_s=pairs / (res["lj_forces_ms"] * 1e-3), energy_from_forces=float(nl.lj_forces(clist, 1.0, 3.4)[1].double().sum().item()),
           lj_pairs_per_s=pairs / (res["lj_energy_ms"] * 1e-3), count_pairs_per_s=pairs / (res["count_neighbours_ms"] * 1e-3))
print(json.dumps(res))
