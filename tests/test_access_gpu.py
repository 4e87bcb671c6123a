"""SURVEY 8f rows on the GPU: device-side PairList accessors (f2), the AtomsBase-style adapter with the device
bounding box (f1) and the skin list (f3), each against the numpy restatement of the reference (oracle/nl_oracle.py).
Integer outputs and R must be BIT-exact: the kernels evaluate the same expression in the same association."""
import numpy as np
import pytest

from oracle import nl_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nl():
    import torch
    assert torch.cuda.is_available()
    import neighbourlists_jl_b200 as nl
    nl._lib.lib()
    return nl


@pytest.mark.parametrize("dtype,int_type", [(np.float64, np.int32), (np.float32, np.int32), (np.float64, np.int64), (np.float32, np.int64)])
def test_pairs_R_bit_exact(nl, dtype, int_type):
    # _getR (src/cell_list.jl:525-531) for every pair, triclinic cell, atoms displaced by lattice vectors
    import torch
    pbc = (True, True, False)
    C = (U.TRICLINIC * 2.0).astype(dtype)
    X = U.displace_by_lattice(U.rand_in_cell(3000, C, seed=71, dtype=dtype), C, pbc)
    Xd = torch.from_numpy(X).cuda()
    pl = nl.neighbour_list(Xd, 3.0, C, pbc, int_type=int_type, with_R=True)
    h = pl.cpu()
    want = O.pairs_R(X, h["i"], h["j"], h["S"], C, dtype)
    got = nl.pairs_R(pl).cpu().numpy()
    assert got.dtype == np.dtype(dtype) and np.array_equal(got, want)
    assert np.array_equal(got, h["R"]), "accessor R == fill-pass R (same expression, SURVEY 8a13)"
    lo, hi = 1234, 1234 + 777
    assert np.array_equal(nl.pairs_R(pl, lo, hi).cpu().numpy(), want[lo:hi])
    assert nl.pairs_R(pl, 5, 5).shape == (0, 3)
    with pytest.raises(IndexError):
        nl.pairs_R(pl, 0, nl.npairs(pl) + 1)


def test_maxneigs_and_rows_padded(nl):
    import torch
    X, C, L = U.rand_config(5000, seed=72)
    rng = np.random.default_rng(5)
    X[:40] = X[0] + rng.normal(scale=0.7, size=(40, 3))  # a cluster: one long row
    Xd = torch.from_numpy(X).cuda()
    pl = nl.neighbour_list(Xd, 5.0, C, (True, True, True))
    h = pl.cpu()
    assert nl.maxneigs(pl) == O.maxneigs(h["first"]) >= 39
    rows = np.array([1, 2, 40, 41, 5000, 7, 7, 2500], dtype=np.int64)
    for width in (O.maxneigs(h["first"]), 16, 0):
        n, j, R, S = nl.sites_padded(pl, rows, width)
        wn, wj, wS, wR = O.rows_padded(X, h["first"], h["j"], h["S"], C, rows, width)
        assert np.array_equal(n.cpu().numpy(), wn)
        assert np.array_equal(j.cpu().numpy(), wj) and np.array_equal(S.cpu().numpy(), wS) and np.array_equal(R.cpu().numpy(), wR)
    # all atoms, default width; without R / S
    n, j, R, S = nl.sites_padded(pl, with_R=False, with_S=False)
    assert R is None and S is None and j.shape == (5000, nl.maxneigs(pl))
    assert np.array_equal(n.cpu().numpy(), np.diff(h["first"]))
    assert int((j != 0).sum().item()) == nl.npairs(pl)
    with pytest.raises(IndexError):
        nl.sites_padded(pl, [0])
    with pytest.raises(IndexError):
        nl.sites_padded(pl, [5001])


def test_iterators(nl):
    # src/iterators.jl: pairs -> (i, j, R) over all pairs; sites -> (i, j, R) per atom
    import torch
    X, C, L = U.rand_config(120, seed=73)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), 4.5, C, (True, True, False))
    h = pl.cpu()
    wantR = O.pairs_R(X, h["i"], h["j"], h["S"], C)
    got = list(nl.pairs(pl, chunk=1000))
    assert len(got) == nl.npairs(pl)
    assert [g[0] for g in got] == h["i"].tolist() and [g[1] for g in got] == h["j"].tolist()
    assert np.array_equal(np.array([g[2] for g in got]), wantR)
    seen = 0
    for i, j, R in nl.sites(pl, chunk=50):
        seen += 1
        lo, hi = h["first"][i - 1] - 1, h["first"][i] - 1
        assert i == seen and np.array_equal(j, h["j"][lo:hi]) and np.array_equal(R, wantR[lo:hi])
    assert seen == 120


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_bounding_box_and_isolated_system(nl, dtype):
    import torch
    rng = np.random.default_rng(9)
    X = (rng.normal(size=(100_003, 3)) * np.array([30.0, 2.0, 11.0]) + np.array([5.0, -40.0, 0.25])).astype(dtype)
    Xd = torch.from_numpy(X).cuda()
    mm = nl.bounding_box(Xd).cpu().numpy()
    assert mm.dtype == np.dtype(dtype)
    assert np.array_equal(mm[:3], X.min(axis=0)) and np.array_equal(mm[3:], X.max(axis=0))
    assert np.array_equal(nl.bounding_cell(Xd), O.bounding_cell(X, dtype))
    one = torch.tensor([[1.5, -2.0, 3.0]], dtype=Xd.dtype).cuda()
    assert np.array_equal(nl.bounding_cell(one), np.eye(3, dtype=dtype))
    with pytest.raises(ValueError):
        nl.bounding_box(Xd[:0])
    # an isolated system end to end == the explicit call with the bounding cell and open boundaries
    Xs = X[:4000] * np.asarray(0.2, dtype=dtype)
    sysm = nl.isolated_system(torch.from_numpy(Xs).cuda())
    pl = nl.neighbour_list(sysm, 1.5)
    orc = O.sortbased(Xs, 1.5, O.bounding_cell(Xs, dtype), (False, False, False), dtype=dtype)
    U.assert_engine_matches_oracle(pl.cpu(), orc, rtol=0, check_R=False, msg="isolated system")


def test_atoms_adapter_known_answers(nl):
    # test/test_atoms_base.jl:13-69 (fcc Cu: 12 neighbours), :71-104 (H3), :106-120 (cutoffs), :122-133 (units), :135-144 (2-D)
    import torch
    for reps in ((4, 2, 3), (3, 3, 3), (2, 2, 2)):
        X, C = U.fcc(3.61, reps)
        cu = nl.periodic_system(X, C)  # host positions are uploaded
        pl = nl.neighbour_list(cu, 3.5)
        assert isinstance(pl, nl.PairList) and nl.nsites(pl) == len(cu) == X.shape[0] and nl.npairs(pl) == 12 * X.shape[0]
        assert all(nl.num_neighbours(pl, i) == 12 for i in range(1, X.shape[0] + 1))
        cl = nl.neighbour_list(cu, 3.5, lazy=True)
        assert isinstance(cl, nl.SortedCellList) and nl.count_neighbours(cl, 1) == 12
        assert int(nl.count_neighbours(cl).sum().item()) == nl.npairs(pl)
        assert len(nl.neighbours(nl.build_cell_list(cu, 3.5), 1)[0]) == 12
    X, C = U.fcc(3.61, (3, 3, 3))
    cu = nl.periodic_system(X, C)
    assert nl.num_neighbours(nl.neighbour_list(cu, 2.6), 1) == 12
    assert nl.num_neighbours(nl.neighbour_list(cu, 3.7), 1) > 12
    assert nl.npairs(nl.neighbour_list(cu, 1.0)) == 0
    nA, nnm, npm = (nl.npairs(nl.neighbour_list(cu, c)) for c in (3.5, (0.35, "nm"), (350.0, "pm")))
    assert nA == nnm == npm == 12 * X.shape[0]
    # isolated H3
    h3 = nl.isolated_system(np.array([[0., 0., 0.], [0., 0., 2.], [0., 10., 0.]]))
    pl = nl.neighbour_list(h3, 5.0)
    assert np.array_equal(pl.C, np.diag([1.0, 11.0, 3.0]))
    j, R, S = nl.neighbours(pl, 1)
    assert j.tolist() == [2] and R.tolist() == [[0.0, 0.0, 2.0]]
    j, R, S = nl.neighbours(pl, 2)
    assert j.tolist() == [1] and R.tolist() == [[0.0, 0.0, -2.0]]
    assert [nl.num_neighbours(pl, i) for i in (1, 2, 3)] == [1, 1, 0]
    assert len(nl.neighbours(nl.neighbour_list(h3, 5.0, lazy=True), 1)[0]) == 1
    # 2-D systems are an error in both entry points
    h2d = nl.isolated_system(np.array([[0., 0.], [0., 2.], [10., 0.]]))
    with pytest.raises(nl.NlError):
        nl.neighbour_list(h2d, 5.0)
    with pytest.raises(nl.NlError):
        nl.build_cell_list(h2d, 5.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_skin_list(nl, dtype):
    import torch
    rc, skin = 4.0, 1.0
    X, C, L = U.rand_config(4000, seed=74, dtype=dtype)
    pbc = (True, True, False)
    Xd = torch.from_numpy(X).cuda()
    sl = nl.SkinList(Xd, rc, skin, C, pbc)
    assert sl.builds == 1
    rng = np.random.default_rng(3)
    X1 = (X + rng.uniform(-0.28, 0.28, size=X.shape)).astype(dtype)  # |d| <= 0.485 < skin / 2
    d2 = nl.max_displacement2(torch.from_numpy(X1).cuda(), Xd).cpu().numpy()[0]
    assert d2 == O.max_displacement2(X1, X, dtype)
    assert sl.update(torch.from_numpy(X1).cuda()) is False and sl.builds == 1
    h = sl.nlist.cpu()
    assert np.array_equal(h["R"], O.pairs_R(X1, h["i"], h["j"], h["S"], C, dtype)), "R refreshed for the moved atoms"
    # completeness: the pairs within rc at the NEW positions are exactly the list's pairs with r^2 < rc^2 (contract arithmetic)
    T = np.dtype(dtype).type
    R = h["R"]
    r2 = (R[:, 0] * R[:, 0] + R[:, 1] * R[:, 1]) + R[:, 2] * R[:, 2]
    keep = r2 < T(rc) * T(rc)
    orc = O.sortbased(X1, rc, C, pbc, dtype=dtype)
    sub = dict(i=h["i"][keep], j=h["j"][keep], S=h["S"][keep])
    U.assert_same_pairs(sub, orc, "skin list filtered to rc at the new positions")
    # a large move forces a rebuild, after which the list equals a fresh one
    X2 = X1.copy()
    X2[17] += np.asarray([0.9, 0.0, 0.0], dtype=dtype)
    assert sl.update(torch.from_numpy(X2).cuda()) is True and sl.builds == 2
    fresh = O.sortbased(X2, rc + skin, C, pbc, dtype=dtype)
    U.assert_engine_matches_oracle(sl.nlist.cpu(), fresh, rtol=1e-12 if dtype == np.float64 else 1e-5, msg="rebuilt skin list")


def _check_lj_forces(nl, X, C, pbc, rc, dtype, eps=0.8, sig=2.5, tol=None):
    import torch
    N = X.shape[0]
    cl = nl.neighbour_list(torch.from_numpy(X).cuda(), rc, C, pbc, lazy=True)
    F, e = nl.lj_forces(cl, eps, sig)
    assert F.shape == (N, 3) and e.shape == (N,) and F.dtype == cl.X.dtype
    d = O.sortbased(X, rc, C, pbc, dtype=dtype)
    wF, we = O.lj_forces(d, eps, sig, N)
    tol = tol or (1e-10 if dtype == np.float64 else 2e-5)
    F, e = F.cpu().numpy().astype(np.float64), e.cpu().numpy().astype(np.float64)
    # per-atom scale: the sum of the magnitudes of the terms (forces cancel)
    R = d["R"].astype(np.float64)
    r2 = (R * R).sum(axis=1)
    s6 = (sig * sig / r2) ** 3
    mag = np.zeros(N)
    np.add.at(mag, d["j"].astype(np.int64) - 1, np.abs(24 * eps * (2 * s6 * s6 - s6) / r2) * np.sqrt(r2))
    emag = np.zeros(N)
    np.add.at(emag, d["j"].astype(np.int64) - 1, np.abs(4 * eps * (s6 * s6 - s6)))
    assert (np.abs(F - wF).max(axis=1) <= tol * np.maximum(mag, 1e-30) + 1e-300).all(), np.abs(F - wF).max()
    assert (np.abs(e - we) <= tol * np.maximum(emag, 1e-30) + 1e-300).all()
    etot = float(nl.lj_energy(cl, eps, sig).item())
    assert abs(e.sum() - etot) <= 10 * tol * emag.sum()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lj_forces_match_oracle(nl, dtype):
    # fast packed path (Float32) / exact tiled path (Float64): cubic periodic, triclinic mixed pbc with displaced atoms
    X, C, L = U.rand_config(20000, seed=81, dtype=dtype)
    _check_lj_forces(nl, X, C, (True, True, True), 5.0, dtype)
    Ct = (U.TRICLINIC * 3.0).astype(dtype)
    pbc = (True, False, True)
    Xt = U.displace_by_lattice(U.rand_in_cell(8000, Ct, seed=82, dtype=dtype), Ct, pbc)
    _check_lj_forces(nl, Xt, Ct, pbc, 3.0, dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lj_forces_fallback_routes(nl, dtype):
    # generic route: box narrower than two cutoffs (self images, nxyz-wide stencils); dense cluster: cells beyond the tables
    X, C = U.fcc(3.61, (2, 2, 2), dtype=dtype)
    _check_lj_forces(nl, X, C, (True, True, True), 5.0, dtype)
    X, C, L = U.rand_config(6000, seed=83, dtype=dtype)
    rng = np.random.default_rng(4)
    X[:700] = (X[0] + rng.uniform(0, 4.0, size=(700, 3))).astype(dtype)
    _check_lj_forces(nl, X, C, (True, True, False), 4.0, dtype, sig=0.6)


def _check_half(nl, X, C, pbc, rc, dtype, int_type=np.int32):
    import torch
    Xd = torch.from_numpy(X).cuda()
    full = O.sortbased(X, rc, C, pbc, dtype=dtype)
    want = O.half_list(full)
    pl = nl.neighbour_list(Xd, rc, C, pbc, half=True, with_R=True, int_type=int_type)
    h = pl.cpu()
    assert pl.half and nl.npairs(pl) * 2 == len(full["i"])
    got = O.mirror_canonical(h["i"], h["j"], h["S"])
    assert np.array_equal(got, want), "half list == one representative of every mirror couple of the reference's list"
    # rows are still CSR rows of i in original order, and R follows the contract for the stored orientation
    f = h["first"].astype(np.int64)
    assert f[0] == 1 and f[-1] - 1 == len(h["i"]) and np.array_equal(h["i"].astype(np.int64), np.repeat(np.arange(1, len(f)), np.diff(f)))
    assert np.array_equal(h["R"], O.pairs_R(X, h["i"], h["j"], h["S"], C, dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_half_list_fast_route(nl, dtype):
    X, C, L = U.rand_config(30000, seed=91, dtype=dtype)
    _check_half(nl, X, C, (True, True, True), 5.0, dtype)
    Ct = (U.TRICLINIC * 3.0).astype(dtype)
    pbc = (True, True, False)
    Xt = U.displace_by_lattice(U.rand_in_cell(9000, Ct, seed=92, dtype=dtype), Ct, pbc)
    _check_half(nl, Xt, Ct, pbc, 3.0, dtype, int_type=np.int64)
    # cells with more than 32 atoms: several home groups per cell
    X, C, L = U.rand_config(20000, seed=93, dtype=dtype, density=0.25)
    _check_half(nl, X, C, (True, False, True), 5.0, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_half_list_other_routes(nl, dtype):
    # self images and repeated cells (2 cells per axis: the same real cell under several shifts), 1-cell box (generic route),
    # dense cluster (cells beyond the mask capacity -> generic per-cell route mixed with the fast one)
    X, C = U.fcc(3.61, (4, 4, 4), dtype=dtype)
    _check_half(nl, X, C, (True, True, True), 5.0, dtype)
    X, C = U.fcc(3.61, (2, 2, 2), dtype=dtype)
    _check_half(nl, X, C, (True, True, True), 5.0, dtype)
    X, C = U.fcc(3.61, (1, 1, 1), dtype=dtype)
    _check_half(nl, X, C, (True, True, True), 4.0, dtype)
    X, C, L = U.rand_config(6000, seed=94, dtype=dtype)
    rng = np.random.default_rng(4)
    X[:700] = (X[0] + rng.uniform(0, 4.0, size=(700, 3))).astype(dtype)
    _check_half(nl, X, C, (True, True, False), 4.0, dtype)
    # density beyond the mask path altogether (27 * dens > 200): exact tiled route
    X, C, L = U.rand_config(40000, seed=95, dtype=dtype, density=0.08)
    _check_half(nl, X, C, (True, True, True), 5.0, dtype)


@pytest.mark.parametrize("dtype,int_type", [(np.float64, np.int32), (np.float32, np.int64)])
def test_lazy_neighbours_rows_in_reference_order(nl, dtype, int_type):
    # neighbours(clist, i) (src/cell_list.jl:821-833): rows straight from the cell list, in the reference's traversal
    # order (the oracle emits its rows in that order too), bit-exact j / S / R
    import torch
    cases = []
    X, C, L = U.rand_config(4000, seed=101, dtype=dtype)
    cases.append((X, C, (True, True, True), 5.0))
    Ct = (U.TRICLINIC * 2.0).astype(dtype)
    cases.append((U.displace_by_lattice(U.rand_in_cell(3000, Ct, seed=102, dtype=dtype), Ct, (True, False, True)), Ct, (True, False, True), 3.0))
    Xf, Cf = U.fcc(3.61, (2, 2, 2), dtype=dtype)       # 2 cells per axis: repeated cells, self images
    cases.append((Xf, Cf, (True, True, True), 5.0))
    Xf1, Cf1 = U.fcc(3.61, (1, 1, 1), dtype=dtype)     # 1 cell, stencil wider than the box
    cases.append((Xf1, Cf1, (True, True, True), 4.0))
    for X, C, pbc, rc in cases:
        N = X.shape[0]
        cl = nl.neighbour_list(torch.from_numpy(X).cuda(), rc, C, pbc, lazy=True, int_type=int_type)
        orc = O.sortbased(X, rc, C, pbc, dtype=dtype, int_type=int_type)
        f = orc["first"].astype(np.int64)
        atoms = np.unique(np.concatenate([[1, N], np.random.default_rng(1).integers(1, N + 1, size=200)]))
        width = int(np.diff(f).max())
        n, j, R, S = nl.neighbours_padded(cl, atoms, width)
        n, j, R, S = n.cpu().numpy(), j.cpu().numpy(), R.cpu().numpy(), S.cpu().numpy()
        for s, a in enumerate(atoms):
            lo, hi = f[a - 1] - 1, f[a] - 1
            assert n[s] == hi - lo
            assert np.array_equal(j[s, :n[s]], orc["j"][lo:hi]) and np.array_equal(S[s, :n[s]], orc["S"][lo:hi])
            assert np.array_equal(R[s, :n[s]], orc["R"][lo:hi])
            assert not j[s, n[s]:].any() and not S[s, n[s]:].any() and not R[s, n[s]:].any()
        # truncation keeps the full count; the single-atom accessor grows its block as needed
        k5 = min(5, len(atoms))
        n2, j2, _, _ = nl.neighbours_padded(cl, atoms[:k5], 3, with_R=False, with_S=False)
        assert np.array_equal(n2.cpu().numpy(), n[:k5]) and j2.shape == (k5, 3)
        jj, RR, SS = nl.neighbours(cl, int(atoms[0]))
        lo, hi = f[atoms[0] - 1] - 1, f[atoms[0]] - 1
        assert np.array_equal(jj.cpu().numpy(), orc["j"][lo:hi]) and np.array_equal(RR.cpu().numpy(), orc["R"][lo:hi])
        assert cl._pl is None, "no pair list was materialised"
    with pytest.raises(IndexError):
        nl.neighbours_padded(cl, [0], 4)


def test_reductions_propagate_nan(nl):
    """maximum / minimum in the reference propagate NaN (ext/NeighbourListsAtomsBaseExt.jl:17-31); so do the device
    reductions: a NaN position must neither pass for a valid bounding box nor for "nothing moved" (SkinList then rebuilds)."""
    import torch
    X, C, L = U.rand_config(3000, seed=5)
    Xd = torch.from_numpy(X).cuda()
    Y = Xd.clone()
    Y[1234, 1] = float("nan")
    assert np.isnan(nl.max_displacement2(Y, Xd).cpu().numpy()[0])
    assert not np.isnan(nl.max_displacement2(Xd, Xd).cpu().numpy()[0])
    bb = nl.bounding_box(Y).cpu().numpy()   # (min x, min y, min z, max x, max y, max z)
    assert np.isnan(bb[1]) and np.isnan(bb[4]) and not np.isnan(bb[0]) and not np.isnan(bb[5])
