"""How exposed is parity to the ONE thing the oracle cannot pin: the association order inside StaticArrays' unrolled 3x3
mat-vec and 3-vector dot (the reference evaluates `cell' * S` and `dot(R, R)` through StaticArrays, which is not
vendored and cannot run here).  The contract restates the left fold StaticArrays 1.x generates
(src/matrix_multiply.jl `mul_unrolled`: reduce(+, a[k,j] * b[j]) over j = 1..3, i.e. (m1 v1 + m2 v2) + m3 v3, and
src/linalg.jl `_vecdot`: ret = a1 b1; ret += a2 b2; ret += a3 b3).  This test takes EVERY candidate pair whose contract r^2 lies
within +-4 ulp of rc^2 -- hits and near misses alike, from a list built with a slightly larger cutoff -- and re-decides
`r^2 < rc^2` under the alternative associations a + (b + c) of the mat-vec, of the dot, and of both.  The number of decisions
that flip is the number of pairs a wrong inference could move in or out of the list; it is written to
gpurun_out/association_exposure.json and bounded here."""
import json
import os

import numpy as np
import pytest

from tests import util as U

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _variants(X, i, j, S, C, rc, T):
    """r^2 of the given pairs under the four association choices, in T, separate IEEE operations (numpy never fuses)."""
    Xi, Xj = X[i - 1].astype(T), X[j - 1].astype(T)
    Sf = S.astype(T)
    C = C.astype(T)
    out = {}
    for mv in ("left", "right"):
        cs = []
        for k in range(3):
            a, b, c = C[0, k] * Sf[:, 0], C[1, k] * Sf[:, 1], C[2, k] * Sf[:, 2]
            cs.append((a + b) + c if mv == "left" else a + (b + c))
        R = [(Xj[:, k] - Xi[:, k]) + cs[k] for k in range(3)]
        p = [R[k] * R[k] for k in range(3)]
        out[(mv, "left")] = (p[0] + p[1]) + p[2]
        out[(mv, "right")] = p[0] + (p[1] + p[2])
    rc2 = T(rc) * T(rc)
    base = out[("left", "left")] < rc2
    return {f"matvec_{m}__dot_{d}": int(np.count_nonzero((v < rc2) != base)) for (m, d), v in out.items() if (m, d) != ("left", "left")}, base


def _exposure(nl, X, C, pbc, rc, T, widen):
    import torch
    Xd = torch.from_numpy(X).cuda()
    pl = nl.neighbour_list(Xd, float(T(rc) * T(1 + widen)), C.astype(T), pbc, with_R=True)
    R = pl.R
    r2 = (R[:, 0] * R[:, 0] + R[:, 1] * R[:, 1]) + R[:, 2] * R[:, 2]   # separate IEEE operations in T: the contract's r^2
    rc2 = T(rc) * T(rc)
    ulp = np.spacing(rc2)
    band = ((r2 >= float(rc2 - 4 * ulp)) & (r2 <= float(rc2 + 4 * ulp))).nonzero().flatten()
    n_hits = int((r2 < float(rc2)).sum())
    i, j, S = pl.i[band].cpu().numpy().astype(np.int64), pl.j[band].cpu().numpy().astype(np.int64), pl.S[band].cpu().numpy()
    flips, base = _variants(X, i, j, S, C, rc, T)
    # the contract variant recomputed on the host must reproduce the device decisions of the band
    assert np.array_equal(base, (r2[band] < float(rc2)).cpu().numpy())
    rcw = T(rc) * T(1 + widen)
    assert rcw * rcw > rc2 + T(64) * ulp, "the widened cutoff must reach well beyond the band"
    return dict(pairs_within_cutoff=n_hits, candidates_within_4ulp=int(band.numel()), decision_flips=flips, ulp_of_rc2=float(ulp))


def test_association_exposure():
    import torch
    assert torch.cuda.is_available()
    import neighbourlists_jl_b200 as nl
    report = {}
    # headline: 10 M atoms, Float64, cubic, rc = 5
    N = 10_000_000
    rng = np.random.Generator(np.random.PCG64(10))
    L = (N / 0.05) ** (1 / 3)
    X = rng.random((N, 3)) * L
    report["headline_f64_10m"] = _exposure(nl, X, np.eye(3) * L, (True, True, True), 5.0, np.float64, 1e-9)
    # BASELINE config 5: 10 M atoms, Float32, cubic, rc = 6
    X32 = X.astype(np.float32)
    report["c5_f32_10m"] = _exposure(nl, X32, (np.eye(3) * L).astype(np.float32), (True, True, True), 6.0, np.float32, 3e-5)
    del X, X32
    # BASELINE config 3 geometry (triclinic, pbc T T F: the mat-vec association matters here), 1 M atoms, both precisions
    N = 1_000_000
    cell = (N / 0.05 / 720.0) ** (1.0 / 3.0) * U.TRICLINIC
    for T, widen, key in ((np.float64, 1e-9, "c3_triclinic_f64_1m"), (np.float32, 3e-5, "c3_triclinic_f32_1m")):
        Xt = U.rand_in_cell(N, cell, seed=3, dtype=T)
        report[key] = _exposure(nl, Xt, cell, (True, True, False), 5.0, T, widen)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "association_exposure.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))
    # Float64: nothing within 4 ulp at these sizes (expected 4e-7 pairs), so no inference about the association can matter
    assert report["headline_f64_10m"]["candidates_within_4ulp"] == 0
    assert report["c3_triclinic_f64_1m"]["candidates_within_4ulp"] == 0
    # Float32: a few hundred candidates in the band out of 4.5e8 pairs; a wrong inference could move at most this many
    f32 = report["c5_f32_10m"]
    assert max(f32["decision_flips"].values()) <= f32["candidates_within_4ulp"] <= 1e-5 * f32["pairs_within_cutoff"]
