"""BASELINE config 4 on ONE GPU: 100 M atoms, rho = 0.05, rc = 5 (2.6e9 pairs > 2^31 - 1): Int32 must be refused with
NL_ERR_OVERFLOW, Int64 must work (reference PairList layout, no R: 40 B/pair = 105 GB).  Prints timings and checks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
rng = np.random.Generator(np.random.PCG64(100))
L = (N / 0.05) ** (1 / 3)
t0 = time.time()
X = rng.random((N, 3)); X *= L
C = np.eye(3) * L
Xd = torch.from_numpy(X).cuda(); del X
print(f"generated {N} atoms in {time.time()-t0:.1f} s, L = {L:.2f}", flush=True)
cl = nl.build_cell_list(Xd, 5.0, C, (True, True, True), int_type=np.int32)
try:
    nl.materialize_pairlist(cl)
    print("ERROR: Int32 materialisation did not overflow")
except nl.NlError as e:
    print("Int32 refused as expected:", e.code, str(e)[:60], flush=True)
del cl
torch.cuda.empty_cache()
for it in range(2):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    cl = nl.build_cell_list(Xd, 5.0, C, (True, True, True), int_type=np.int64)
    e1.record()
    pl = nl.materialize_pairlist(cl)
    e2.record(); torch.cuda.synchronize()
    P = nl.npairs(pl)
    print(f"run {it}: ncells {cl.ncells.tolist()} pairs {P} build {e0.elapsed_time(e1):.2f} ms materialise {e1.elapsed_time(e2):.2f} ms "
          f"-> {P / (e0.elapsed_time(e2) * 1e-3):.3e} pairs/s, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    if it == 0:
        first = pl.first
        assert int(first[0]) == 1 and int(first[-1]) == P + 1 and P > 2**31
        assert bool((first[1:] >= first[:-1]).all())
        assert abs(P / N - 4 / 3 * np.pi * 125 * 0.05) < 0.05
        cnt = nl.count_neighbours(cl)
        assert torch.equal(cnt, first[1:] - first[:-1])
        sel = torch.randint(0, P, (1_000_000,), device="cuda")
        i, j, S = pl.i[sel], pl.j[sel], pl.S[sel].double()
        R = Xd[j - 1] - Xd[i - 1] + S * L
        assert float((R * R).sum(1).max()) < 25.0 and int(pl.j.max()) <= N and int(pl.j.min()) >= 1
        print("checks ok", flush=True)
    del pl, cl
    torch.cuda.empty_cache()
