"""Shared test helpers: seeded generators for the BASELINE configs, the reference's
canonical comparison (test/test_utils.jl:68-95) and the oracle <-> engine comparison."""
from __future__ import annotations

import numpy as np

from oracle import nl_oracle as O

ALL_PBC = [(True, True, True), (False, False, False), (True, False, False), (False, True, False), (False, False, True),
           (True, True, False), (True, False, True), (False, True, True)]  # test/test_utils.jl:41-52

TRICLINIC = np.array([[10.0, 2.0, 1.0], [0.0, 9.0, 1.5], [0.0, 0.0, 8.0]])  # test/test_utils.jl:297-299


def rand_config(N, seed, density=0.05, dtype=np.float64):
    """rand_config (test/test_utils.jl:28-34) with a fixed seed: cubic box, x = L * U[0,1)^3."""
    rng = np.random.Generator(np.random.PCG64(seed))
    L = (N / density) ** (1.0 / 3.0)
    C = np.eye(3) * L
    X = (rng.random((N, 3)) @ C).astype(dtype)
    return X, C.astype(dtype), L


def rand_in_cell(N, cell, seed, dtype=np.float64):
    """x = cell' * f, f ~ U[0,1)^3 in float64, cast to T last."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = rng.random((N, 3))
    return (f @ np.asarray(cell, np.float64)).astype(dtype)


def displace_by_lattice(X, cell, pbc, seed=None):
    """Issue #6 pattern (test/test_sortbased.jl:185-195): per-atom shifts in {-1,0,+1} lattice vectors on periodic axes."""
    X = np.asarray(X)
    N = X.shape[0]
    idx = np.arange(1, N + 1)
    sh = np.stack([(idx % 3) - 1, ((idx + 1) % 3) - 1, ((idx + 2) % 3) - 1], axis=1).astype(np.float64)
    sh = sh * np.asarray(pbc, dtype=np.float64)[None, :]
    return (X.astype(np.float64) + sh @ np.asarray(cell, np.float64)).astype(X.dtype)


def fcc(a=3.61, reps=(4, 4, 4), dtype=np.float64):
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]]) * a
    pts = [basis + np.array([i, j, k]) * a for i in range(reps[0]) for j in range(reps[1]) for k in range(reps[2])]
    return np.concatenate(pts).astype(dtype), (np.diag(reps) * a).astype(dtype)


def pair_tuples(d):
    """sorted (i, j, S1, S2, S3) rows as an (P,5) int64 array."""
    i, j, S = O.canonical(d["i"], d["j"], d["S"])
    return np.concatenate([i[:, None], j[:, None], S], axis=1)


def assert_same_pairs(a, b, msg=""):
    ta, tb = pair_tuples(a), pair_tuples(b)
    assert ta.shape == tb.shape, f"{msg}: {ta.shape[0]} vs {tb.shape[0]} pairs"
    assert np.array_equal(ta, tb), f"{msg}: pair sets differ"


def assert_engine_matches_oracle(eng, orc, rtol, check_R=True, msg=""):
    """eng / orc: dicts with first, i, j, S (and R).  CSR offsets and the (i,j,S) set bit-exact after the
    canonical per-row ordering; R within rtol relative (entrywise, scaled by |R|)."""
    assert np.array_equal(np.asarray(eng["first"]).astype(np.int64), np.asarray(orc["first"]).astype(np.int64)), f"{msg}: first differs"
    ei, ej, eS, *eR = O.canonical(eng["i"], eng["j"], eng["S"], eng.get("R") if check_R else None)
    oi, oj, oS, *oR = O.canonical(orc["i"], orc["j"], orc["S"], orc.get("R") if check_R else None)
    assert ei.shape == oi.shape, f"{msg}: {ei.shape[0]} vs {oi.shape[0]} pairs"
    assert np.array_equal(ei, oi) and np.array_equal(ej, oj) and np.array_equal(eS, oS), f"{msg}: (i,j,S) differ"
    # rows must be grouped by i in CSR order
    f = np.asarray(eng["first"]).astype(np.int64)
    ii = np.asarray(eng["i"]).astype(np.int64)
    assert np.array_equal(ii, np.repeat(np.arange(1, len(f)), np.diff(f))), f"{msg}: i is not the CSR row index"
    if check_R and eR and eR[0] is not None and oR[0] is not None and len(ei):
        er, orr = eR[0].astype(np.float64), oR[0].astype(np.float64)
        scale = np.maximum(np.linalg.norm(orr, axis=1, keepdims=True), 1e-300)
        err = np.abs(er - orr) / scale
        assert err.max() <= rtol, f"{msg}: R rel err {err.max()} > {rtol}"
