"""Pins the CPU oracle (oracle/nl_oracle.cpp) against every analytic known answer the reference's
own tests hold for this path and against its two sibling oracles (legacy
linked-list restatement, brute-force image enumeration).  CPU only."""
import numpy as np
import pytest

from oracle import nl_oracle as O
from tests import util as U

FULL, NONE = (True, True, True), (False, False, False)


def same(a, b):
    return np.array_equal(U.pair_tuples(a), U.pair_tuples(b))


def test_edge_cases():
    # test/test_utils.jl:480-494, test/test_unified_api.jl:67-83
    C = np.eye(3) * 10.0
    assert O.sortbased([[5.0, 5.0, 5.0]], 3.0, C, FULL)["npairs"] == 0
    assert O.sortbased([[5.0, 5.0, 5.0], [5.0, 5.0, 6.0]], 3.0, C, FULL)["npairs"] == 2
    assert O.sortbased([[1.0, 1.0, 1.0], [8.0, 8.0, 8.0]], 3.0, C, NONE)["npairs"] == 0
    r = O.sortbased(np.zeros((0, 3)), 3.0, C, FULL)
    assert r["first"].tolist() == [1] and r["cell_offsets"].tolist() == [1] * 28  # src/gpu_kernels.jl:268-271,303-312


@pytest.mark.parametrize("reps", [(4, 2, 3), (3, 3, 3), (2, 2, 2)])
def test_fcc_twelve_neighbours(reps):
    # test/test_atoms_base.jl:13-69: fcc Cu, rc = 3.5 -> 12 neighbours for EVERY atom
    X, C = U.fcc(3.61, reps)
    r = O.sortbased(X, 3.5, C, FULL)
    assert np.all(np.diff(r["first"]) == 12)
    assert same(r, O.brute(X, 3.5, C, FULL)) and same(r, O.legacy(X, 3.5, C, FULL))


def test_fcc_cutoffs():
    # test/test_atoms_base.jl:106-120: rc 2.6 -> 12, 3.7 -> > 12, 1.0 -> 0 pairs
    X, C = U.fcc(3.61, (3, 3, 3))
    assert np.all(np.diff(O.sortbased(X, 2.6, C, FULL)["first"]) == 12)
    assert np.all(np.diff(O.sortbased(X, 3.7, C, FULL)["first"]) > 12)
    assert O.sortbased(X, 1.0, C, FULL)["npairs"] == 0


@pytest.mark.parametrize("a", [3.61, 3.615])
def test_config2_fcc_4x4x4(a):
    # BASELINE config 2: 10 752 pairs, 42 per atom, 3 720 with S != 0
    X, C = U.fcc(a)
    r = O.sortbased(X, 5.0, C, FULL)
    assert r["npairs"] == 10752 and np.all(np.diff(r["first"]) == 42)
    assert int((np.abs(r["S"]).sum(1) > 0).sum()) == 3720
    assert same(r, O.brute(X, 5.0, C, FULL)) and same(r, O.legacy(X, 5.0, C, FULL))


def test_isolated_h3():
    # test/test_atoms_base.jl:71-104: H3 at z=0, z=2 and y=10 in the open bbox+1 cell diag(1,11,3), rc = 5
    X = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 2.0], [0.0, 10.0, 0.0]])
    r = O.sortbased(X, 5.0, np.diag([1.0, 11.0, 3.0]), NONE)
    assert r["first"].tolist() == [1, 2, 3, 3] and r["j"].tolist() == [2, 1]
    assert np.array_equal(r["R"], [[0, 0, 2.0], [0, 0, -2.0]]) and not r["S"].any()


def _multiset(r):
    i, j = r["i"].astype(int), r["j"].astype(int)
    d = np.round(np.linalg.norm(r["R"], axis=1), 10)
    return sorted(zip(np.minimum(i, j).tolist(), np.maximum(i, j).tolist(), d.tolist()))


def test_issue6_outside_box():
    # test/test_sortbased.jl:143-169
    C = np.eye(3) * 8.0
    ref = _multiset(O.sortbased([[0.5, 0.5, 0.5], [0.5, 1.5, 0.5]], 1.5, C, FULL))
    assert ref == [(1, 2, 1.0), (1, 2, 1.0)]
    assert _multiset(O.sortbased([[0.5, 0.5, 0.5], [0.5, -0.5, 0.5]], 1.5, C, FULL)) == ref
    assert _multiset(O.sortbased([[0.5, 0.5, 0.5], [0.5, -6.5, 0.5]], 1.5, C, FULL)) == ref


@pytest.mark.parametrize("pbc", U.ALL_PBC)
def test_issue6_lattice_shifts_are_noop(pbc):
    # test/test_sortbased.jl:171-200
    X, C, L = U.rand_config(60, seed=40)
    ref = _multiset(O.sortbased(X, L * 0.25, C, pbc))
    assert _multiset(O.sortbased(U.displace_by_lattice(X, C, pbc), L * 0.25, C, pbc)) == ref


def test_issue6_tiny_negative_frac():
    # test/test_sortbased.jl:202-228
    L = 8.0
    tiny = -5e-17 * L
    assert (tiny / L) - np.floor(tiny / L) == 1.0
    X = np.array([[0.5, L - 0.3, 0.5], [0.5, tiny, 0.5]])
    r = O.sortbased(X, 1.5, np.eye(3) * L, FULL)
    row = slice(r["first"][0] - 1, r["first"][1] - 1)
    assert 2 in r["j"][row].tolist()
    k = r["j"][row].tolist().index(2)
    assert abs(np.linalg.norm(r["R"][row][k]) - 0.3) < 1e-9
    assert np.allclose(X[1] - X[0] + r["S"][row][k] @ (np.eye(3) * L), r["R"][row][k], atol=1e-12)


@pytest.mark.parametrize("seed", range(5))
def test_sortbased_vs_legacy_random(seed):
    # test/test_sortbased.jl:17-21
    rng = np.random.default_rng(seed)
    X, C, L = U.rand_config(int(rng.integers(50, 201)), seed=seed)
    r = O.sortbased(X, L * 0.25, C, FULL)
    assert same(r, O.legacy(X, L * 0.25, C, FULL)) and same(r, O.brute(X, L * 0.25, C, FULL))


@pytest.mark.parametrize("pbc", U.ALL_PBC)
def test_all_pbc_vs_legacy_and_brute(pbc):
    # test/test_sortbased.jl:23-41 (cubic + triclinic), plus displaced atoms against brute force
    X, C, L = U.rand_config(100, seed=50)
    r = O.sortbased(X, L * 0.25, C, pbc)
    assert same(r, O.legacy(X, L * 0.25, C, pbc)) and same(r, O.brute(X, L * 0.25, C, pbc))
    Xt = U.rand_in_cell(80, U.TRICLINIC, seed=51)
    rt = O.sortbased(Xt, 3.0, U.TRICLINIC, pbc)
    assert same(rt, O.legacy(Xt, 3.0, U.TRICLINIC, pbc)) and same(rt, O.brute(Xt, 3.0, U.TRICLINIC, pbc))
    Xd = U.displace_by_lattice(Xt, U.TRICLINIC, pbc)
    assert same(O.sortbased(Xd, 3.0, U.TRICLINIC, pbc), O.brute(Xd, 3.0, U.TRICLINIC, pbc))


def test_elongated_large_cutoff_int_types():
    # test/test_sortbased.jl:36-63
    C2 = np.diag([5.0, 5.0, 20.0])
    X2 = U.rand_in_cell(80, C2, seed=11)
    assert same(O.sortbased(X2, 3.0, C2, FULL), O.legacy(X2, 3.0, C2, FULL))
    X, C, L = U.rand_config(30, seed=12)
    r = O.sortbased(X, L * 0.6, C, FULL)
    assert same(r, O.legacy(X, L * 0.6, C, FULL)) and same(r, O.brute(X, L * 0.6, C, FULL))
    r64 = O.sortbased(X, L * 0.6, C, FULL, int_type=np.int64)
    assert r64["i"].dtype == np.int64 and same(r, r64)


@pytest.mark.parametrize("N", [500, 1000, 2000])
def test_sizes_counts(N):
    # test/test_sortbased.jl:65-71
    X, C, L = U.rand_config(N, seed=N)
    assert O.sortbased(X, L * 0.25, C, FULL)["npairs"] == O.legacy(X, L * 0.25, C, FULL)["npairs"]


def test_float32_vs_brute():
    # test/test_gpu.jl runs every case for Float32 too
    X, C, L = U.rand_config(150, seed=70, dtype=np.float32)
    assert same(O.sortbased(X, L * 0.25, C, FULL, dtype=np.float32), O.brute(X, L * 0.25, C, FULL, dtype=np.float32))
    rng = np.random.Generator(np.random.PCG64(5))
    Xd = (rng.random((200, 3)) * 5.0).astype(np.float32)  # high density, test/test_gpu.jl:78-86
    C5 = (np.eye(3) * 5.0).astype(np.float32)
    assert same(O.sortbased(Xd, 2.0, C5, FULL, dtype=np.float32), O.brute(Xd, 2.0, C5, FULL, dtype=np.float32))


def test_lazy_fields_and_stable_perm():
    X, C, L = U.rand_config(500, seed=80)
    r = O.sortbased(X, L * 0.25, C, FULL)
    ids = np.empty(500, np.int64)
    ids[r["perm"] - 1] = r["cell_id"]
    assert np.array_equal(r["perm"] - 1, np.argsort(ids, kind="stable"))       # == sortperm (stable)
    assert np.array_equal(r["Xs"], X[r["perm"] - 1])
    counts = np.bincount(ids, minlength=r["ncells_total"] + 1)[1:]
    assert np.array_equal(np.diff(r["cell_offsets"]), counts) and r["cell_offsets"][0] == 1


def test_fuzz_sortbased_vs_brute_force():
    """The same seeded generator as the GPU fuzz (tests/test_parity_gpu.py::test_randomised_geometries): random triclinic /
    left-handed cells, random pbc, cutoffs up to 1.3 box widths, atoms several boxes away or outside the box.  The sort-based
    restatement must equal the brute-force image sum as (i, j, S) sets -- the set-level invariant of SURVEY 8a."""
    rng = np.random.default_rng(20260101)
    checked = 0
    for case in range(40):
        dtype = np.float64 if case % 2 == 0 else np.float32
        N = int(rng.integers(1, 600))
        A = np.diag(rng.uniform(4.0, 14.0, size=3)) + rng.uniform(-1.5, 1.5, size=(3, 3)) * (rng.random() < 0.7)
        if rng.random() < 0.3:
            A[2] = -A[2]
        if abs(np.linalg.det(A)) < 20.0:
            A = np.diag(np.diag(A))
        pbc = tuple(bool(b) for b in rng.integers(0, 2, size=3))
        f = rng.random((N, 3))
        f += rng.integers(-2, 3, size=(N, 3)) * (rng.random() < 0.5)
        f += rng.normal(scale=0.2, size=(N, 3)) * (rng.random() < 0.5)
        X = f @ A
        lens = np.abs(np.linalg.det(A)) / np.array([np.linalg.norm(np.cross(A[(k + 1) % 3], A[(k + 2) % 3])) for k in range(3)])
        cutoff = float(rng.choice([0.25, 0.45, 0.9, 1.3]) * lens.min())
        if dtype != np.float64 or N > 320:
            continue  # brute force in Float64 only (rounding-sensitive pairs differ between formulations in Float32)
        d = O.sortbased(X, cutoff, A, pbc)
        b = O.brute(X, cutoff, A, pbc)
        U.assert_same_pairs(d, b, f"fuzz case {case}")
        assert np.array_equal(np.diff(d["first"]), np.bincount(d["i"] - 1, minlength=N))
        checked += 1
    assert checked >= 8
