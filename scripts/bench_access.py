"""Timings of the SURVEY 8f kernels at the headline size (10 M atoms, 2.6e8 pairs): min of 5 after warm-up,
CUDA events on the current stream; algorithmic bytes -> GB/s next to the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import neighbourlists_jl_b200 as nl
from bench import make_positions, CUTOFF

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
X, C, L = make_positions(n, 10)
Xd = torch.from_numpy(X).cuda()
pl = nl.neighbour_list(Xd, CUTOFF, C, (True, True, True))
P = nl.npairs(pl)
peak = 6448.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        del r
    return best


rows = []
def report(name, ms, nbytes, note=""):
    gbs = nbytes / ms / 1e6
    rows.append(dict(kernel=name, ms=round(ms, 3), algorithmic_GB=round(nbytes / 1e9, 3), GBps=round(gbs, 1), frac_of_hbm_peak=round(gbs / peak, 3), note=note))
    print(rows[-1], flush=True)

R = torch.empty((P, 3), dtype=Xd.dtype, device="cuda")
L_ = nl._lib.lib()
from neighbourlists_jl_b200 import api
prm = api._list_params(pl)
def f_R():
    nl._lib.check(L_.nl_pairs_R(prm, Xd.data_ptr(), n, pl.i.data_ptr(), pl.j.data_ptr(), pl.S.data_ptr(), 0, P, R.data_ptr(),
                                torch.cuda.current_stream().cuda_stream))
report("nl_pairs_R (all pairs)", timed(f_R), P * (4 + 4 + 12 + 24 + 24) + 24 * n, "i,j,S read + X[j] gather + R write (+X[i] once per row)")
del R
report("nl_max_neighbours", timed(lambda: nl.maxneigs(pl)), 4 * (n + 1), "includes the host read of the result")
w = nl.maxneigs(pl)
m = min(n, 2_000_000)
sel = torch.arange(1, m + 1, dtype=torch.int32, device="cuda")
t = timed(lambda: nl.sites_padded(pl, sel, w))
Pm = int(pl.first[m].item()) - 1
report(f"nl_rows_padded ({m} rows, width {w})", t, m * w * (4 + 12 + 24) + Pm * (4 + 12 + 24), "padded j,S,R blocks written; j,S read + X[j] gather; includes torch.empty of the outputs")
report("nl_bounding_box", timed(lambda: nl.bounding_box(Xd)), 24 * n)
Y = Xd + 0.01
report("nl_max_displacement2", timed(lambda: nl.max_displacement2(Y, Xd)), 48 * n)
sl = nl.SkinList(Xd, CUTOFF, 0.5, C, (True, True, True))
Ps = nl.npairs(sl.nlist)
t_upd = timed(lambda: sl.update(Y))
t_full = timed(lambda: nl.neighbour_list(Y, CUTOFF + 0.5, C, (True, True, True), with_R=True))
print(json.dumps(dict(skin_list=dict(pairs=Ps, update_ms=round(t_upd, 3), rebuild_ms=round(t_full, 3), builds=sl.builds), kernels=rows, hbm_peak_GBps=peak)))
