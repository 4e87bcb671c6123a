"""BASELINE-size checks on the GPU.  Where the oracle is too slow, parity is checked through
size-independent properties: sortedness, permutation, CSR consistency, pair symmetry
((i,j,S) present <=> (j,i,-S) present, via order-independent checksums), R against its definition,
lazy == materialised, and the expected pair density."""
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


def _mix(a, b, s0, s1, s2):
    """order-independent 64-bit fingerprint of a pair record (wrapping int64 arithmetic)"""
    h = a * 0x9E3779B97F4A7C15 + b * 0x4F1BBCDCBFA53E0B + (s0 + 7) * 0x1000193 + (s1 + 7) * 0x27D4EB2F + (s2 + 7) * 0x165667B1
    h = h ^ (h >> 29)
    return h * 0x2545F4914F6CDD1D


def test_headline_10m_properties(nl):
    import torch
    N, rc = 10_000_000, 5.0
    rng = np.random.Generator(np.random.PCG64(10))
    L = (N / 0.05) ** (1 / 3)
    X = rng.random((N, 3)) * L
    C = np.eye(3) * L
    Xd = torch.from_numpy(X).cuda()
    clist = nl.build_cell_list(Xd, rc, C, (True, True, True))
    assert clist.ncells.tolist() == [116, 116, 116]
    # stage 1-3: sorted keys, permutation, offsets
    cid = clist.cell_id.long()
    assert bool((cid[1:] >= cid[:-1]).all()) and int(cid.min()) >= 1 and int(cid.max()) <= 116 ** 3
    perm = clist.perm.long()
    assert int(torch.bincount(perm - 1, minlength=N).max()) == 1 and int(perm.min()) == 1 and int(perm.max()) == N
    same_cell = cid[1:] == cid[:-1]
    assert bool((perm[1:][same_cell] > perm[:-1][same_cell]).all()), "stable: ascending original index inside a cell"
    assert torch.equal(clist.X, Xd[perm - 1])
    co = clist.cell_offsets.long()
    assert int(co[0]) == 1 and int(co[-1]) == N + 1
    assert torch.equal(co[1:] - co[:-1], torch.bincount(cid - 1, minlength=116 ** 3))
    del same_cell
    # stage 4-6
    pl = nl.materialize_pairlist(clist, with_R=True)
    P = nl.npairs(pl)
    assert abs(P / N - 4 / 3 * np.pi * rc ** 3 * 0.05) < 0.05, P / N
    first = pl.first.long()
    assert int(first[0]) == 1 and int(first[-1]) == P + 1 and bool((first[1:] >= first[:-1]).all())
    counts = first[1:] - first[:-1]
    assert torch.equal(pl.i.long(), torch.repeat_interleave(torch.arange(1, N + 1, device="cuda"), counts))
    assert torch.equal(nl.count_neighbours(clist).long(), counts), "lazy count == materialised row sizes"
    i, j, S = pl.i.long(), pl.j.long(), pl.S.long()
    assert int(j.min()) >= 1 and int(j.max()) <= N and int(S.abs().max()) <= 1
    fwd = _mix(i, j, S[:, 0], S[:, 1], S[:, 2]).sum()
    bwd = _mix(j, i, -S[:, 0], -S[:, 1], -S[:, 2]).sum()
    assert int(fwd) == int(bwd), "pair set is symmetric under (i,j,S) -> (j,i,-S)"
    del fwd, bwd
    # R against its definition, bit for bit, on a 2 M-pair sample; and inside the cutoff everywhere
    r2 = (pl.R * pl.R).sum(1)
    assert float(r2.max()) < rc * rc
    sel = torch.randint(0, P, (2_000_000,), device="cuda")
    Cm = torch.as_tensor(C, device="cuda")
    Sf = S[sel].double()
    cs = torch.stack([(Cm[0, k] * Sf[:, 0] + Cm[1, k] * Sf[:, 1]) + Cm[2, k] * Sf[:, 2] for k in range(3)], 1)
    Rdef = (Xd[j[sel] - 1] - Xd[i[sel] - 1]) + cs
    assert torch.equal(Rdef, pl.R[sel])
    # no duplicate (j, S) inside a row: check on the first 200k rows
    m = int(first[200_000]) - 1
    key = (i[:m] * (N + 1) + j[:m]) * 27 + (S[:m, 0] + 1) * 9 + (S[:m, 1] + 1) * 3 + (S[:m, 2] + 1)
    assert int(torch.unique(key).numel()) == m


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_config3_1m_triclinic_vs_oracle(nl, dtype):
    # BASELINE config 3: 1 M atoms, triclinic cell, pbc (T,T,F), rc = 5 -> full comparison with the oracle
    import torch
    N = 1_000_000
    s = (N / 0.05 / 720.0) ** (1.0 / 3.0)
    cell = s * U.TRICLINIC
    X = U.rand_in_cell(N, cell, seed=3, dtype=dtype)
    clist = nl.build_cell_list(torch.from_numpy(X).cuda(), 5.0, cell.astype(dtype), (True, True, False))
    if dtype == np.float64:
        assert clist.ncells.tolist() == [60, 53, 47]
    pl = nl.materialize_pairlist(clist, with_R=True)
    orc = O.sortbased(X, 5.0, cell.astype(dtype), (True, True, False), dtype=dtype)
    assert np.array_equal(clist.perm.cpu().numpy(), orc["perm"]) and np.array_equal(clist.cell_offsets.cpu().numpy(), orc["cell_offsets"])
    U.assert_engine_matches_oracle(pl.cpu(), orc, 1e-12 if dtype == np.float64 else 1e-5, msg="C3")


def test_config5_lazy_lj_f32(nl):
    # BASELINE config 5 at 1 M atoms against the oracle, and at 10 M atoms against the materialised list
    import torch
    for N, full in ((1_000_000, True), (10_000_000, False)):
        rng = np.random.Generator(np.random.PCG64(10))
        L = (N / 0.05) ** (1 / 3)
        X = (rng.random((N, 3)) * L).astype(np.float32)
        C = (np.eye(3) * L).astype(np.float32)
        clist = nl.neighbour_list(torch.from_numpy(X).cuda(), 6.0, C, (True, True, True), lazy=True)
        e = float(nl.lj_energy(clist, 1.0, 3.4).item())
        counts = nl.count_neighbours(clist)
        if full:
            lazy = O.sortbased(X, 6.0, C, (True, True, True), dtype=np.float32, lazy=True)
            e_ref = O.lj_energy(lazy, 1.0, 3.4)
            assert abs(e - e_ref) <= 1e-9 * abs(e_ref), (e, e_ref)
        else:
            pl = nl.materialize_pairlist(clist, with_R=True)
            assert torch.equal(counts.long(), (pl.first[1:] - pl.first[:-1]).long())
            r2 = (pl.R.double() ** 2).sum(1)
            s6 = (3.4 * 3.4 / r2) ** 3
            e_mat = float((4.0 * (s6 * s6 - s6)).sum().item())
            assert abs(e - e_mat) <= 1e-5 * abs(e_mat), (e, e_mat)
            assert abs(nl.npairs(pl) / N - 4 / 3 * np.pi * 216 * 0.05) < 0.1


def test_headline_10m_half_list_accessors_and_forces(nl):
    """The 8f rows at the headline size, through size-independent properties: the half list holds exactly one pair of every
    mirror couple of the full list (order-independent fingerprints); pairs_R reproduces the fill pass's R bit for bit;
    maxneigs / padded rows agree with the CSR; the fused force sink obeys Newton's third law and its energies sum to
    the fused energy sink."""
    import torch
    N, rc = 10_000_000, 5.0
    rng = np.random.Generator(np.random.PCG64(10))
    L = (N / 0.05) ** (1 / 3)
    X = rng.random((N, 3)) * L
    C = np.eye(3) * L
    Xd = torch.from_numpy(X).cuda()
    clist = nl.build_cell_list(Xd, rc, C, (True, True, True))
    full = nl.materialize_pairlist(clist, with_R=True)
    P = nl.npairs(full)
    fi, fj, fS = full.i.long(), full.j.long(), full.S.long()
    # canonical fingerprint of a mirror couple: min/max of the two orientations' fingerprints, combined symmetrically
    a = _mix(fi, fj, fS[:, 0], fS[:, 1], fS[:, 2])
    b = _mix(fj, fi, -fS[:, 0], -fS[:, 1], -fS[:, 2])
    couple_full = (torch.minimum(a, b) * 31 + torch.maximum(a, b)).sum()
    del a, b
    assert torch.equal(nl.pairs_R(full), full.R), "accessor R == fill-pass R"
    counts = (full.first[1:] - full.first[:-1]).long()
    w = nl.maxneigs(full)
    assert w == int(counts.max())
    rows = torch.randint(1, N + 1, (100_000,), device="cuda", dtype=torch.int32)
    n, j, R, S = nl.sites_padded(full, rows, w)
    assert torch.equal(n.long(), counts[rows.long() - 1])
    k = torch.arange(w, device="cuda")[None, :]
    valid = k < n.long()[:, None]
    src = (full.first.long()[rows.long() - 1] - 1)[:, None] + k
    assert torch.equal(j[valid], full.j[src[valid]]) and torch.equal(R[valid], full.R[src[valid]]) and not bool(j[~valid].any())
    del n, j, R, S, valid, src, fi, fj, fS
    Pf = P
    del full
    half = nl.materialize_pairlist(clist, half=True)
    assert nl.npairs(half) * 2 == Pf
    hi_, hj, hS = half.i.long(), half.j.long(), half.S.long()
    a = _mix(hi_, hj, hS[:, 0], hS[:, 1], hS[:, 2])
    b = _mix(hj, hi_, -hS[:, 0], -hS[:, 1], -hS[:, 2])
    couple_half = (torch.minimum(a, b) * 31 + torch.maximum(a, b)).sum()
    assert int(couple_half * 2 - couple_full) == 0, "half list = one pair per mirror couple (wrapping int64 fingerprints)"
    del a, b, half, hi_, hj, hS
    # fused force sink in Float64 (exact tiled route) at full size
    F, e = nl.lj_forces(clist, 1.0, 3.4)
    etot = float(nl.lj_energy(clist, 1.0, 3.4).item())
    assert abs(float(e.sum().item()) - etot) <= 1e-9 * abs(etot)
    assert float(F.sum(0).abs().max()) <= 1e-9 * float(F.abs().sum())


def _row_fingerprints_np(first, j, S):
    """per-row sums of the pair fingerprints (wrapping uint64), numpy side: rows are contiguous in the CSR"""
    with np.errstate(over="ignore"):
        row = np.repeat(np.arange(1, first.shape[0], dtype=np.uint64), np.diff(first.astype(np.int64)))
        h = (row * np.uint64(0x9E3779B97F4A7C15) + j.astype(np.int64).astype(np.uint64) * np.uint64(0x4F1BBCDCBFA53E0B)
             + (S[:, 0].astype(np.int64) + 7).astype(np.uint64) * np.uint64(0x1000193)
             + (S[:, 1].astype(np.int64) + 7).astype(np.uint64) * np.uint64(0x27D4EB2F)
             + (S[:, 2].astype(np.int64) + 7).astype(np.uint64) * np.uint64(0x165667B1))
        h ^= (h.view(np.int64) >> np.int64(29)).view(np.uint64)   # arithmetic shift, as torch's int64 >> in _mix
        h *= np.uint64(0x2545F4914F6CDD1D)
        c = np.concatenate([[np.uint64(0)], np.cumsum(h, dtype=np.uint64)])
        f0 = first.astype(np.int64) - 1
        return c[f0[1:]] - c[f0[:-1]]


def test_headline_10m_list_against_oracle(nl):
    """The reference's bar (test/test_gpu.jl:41-68: GPU list == CPU list) at the BASELINE size: `first` exactly, and an
    order-independent fingerprint of the (j, S) multiset of EVERY row, against the CPU oracle on the same 10 M atoms; R
    against the oracle's own R on the rows of a sample of atoms."""
    import torch
    N, rc = 10_000_000, 5.0
    rng = np.random.Generator(np.random.PCG64(10))
    L = (N / 0.05) ** (1 / 3)
    X = rng.random((N, 3)) * L
    C = np.eye(3) * L
    pbc = (True, True, True)
    orc = O.sortbased(X, rc, C, pbc, want_R=False)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), rc, C, pbc, with_R=True)
    first = pl.first.cpu().numpy()
    assert np.array_equal(first, orc["first"]), "CSR offsets"
    P = int(first[-1]) - 1
    assert P == orc["npairs"] == nl.npairs(pl)
    fo = _row_fingerprints_np(orc["first"], orc["j"], orc["S"])
    # engine side on the device: the same wrapping arithmetic in int64
    i, j, S = pl.i.long(), pl.j.long(), pl.S.long()
    assert torch.equal(i, torch.repeat_interleave(torch.arange(1, N + 1, device="cuda"), (pl.first[1:] - pl.first[:-1]).long()))
    h = _mix(i, j, S[:, 0], S[:, 1], S[:, 2])          # _mix's logical shift: emulate >> on the unsigned value
    del i, j, S
    c = torch.cumsum(h, 0)
    f0 = pl.first.long() - 1
    cz = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), c])
    fe = (cz[f0[1:]] - cz[f0[:-1]]).cpu().numpy().view(np.uint64)
    assert np.array_equal(fe, fo), "per-row (j, S) multisets differ from the oracle"
    # R on the rows of 50 000 sampled atoms against the oracle's contract arithmetic
    sel = np.sort(np.random.default_rng(1).choice(N, 50_000, replace=False))
    lo, hi = first[sel] - 1, first[sel + 1] - 1
    idx = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)])
    it = torch.from_numpy(idx).cuda()
    sub = dict(i=pl.i[it].cpu().numpy(), j=pl.j[it].cpu().numpy(), S=pl.S[it].cpu().numpy())
    Rref = O.pairs_R(X, sub["i"], sub["j"], sub["S"], C, np.float64)
    assert np.array_equal(pl.R[it].cpu().numpy(), Rref), "R == (X[j] - X[i]) + C' S in the contract's association"


def test_config5_10m_counts_against_oracle(nl):
    """BASELINE config 5 (10 M atoms, Float32, rc = 6) per-atom neighbour counts of the lazy sink against the oracle's CSR."""
    import torch
    N = 10_000_000
    rng = np.random.Generator(np.random.PCG64(10))
    L = (N / 0.05) ** (1 / 3)
    X = (rng.random((N, 3)) * L).astype(np.float32)
    C = (np.eye(3) * L).astype(np.float32)
    clist = nl.neighbour_list(torch.from_numpy(X).cuda(), 6.0, C, (True, True, True), lazy=True)
    counts = nl.count_neighbours(clist).cpu().numpy()
    orc = O.sortbased(X, 6.0, C, (True, True, True), dtype=np.float32, want_R=False)
    assert np.array_equal(counts, np.diff(orc["first"]))
    assert np.array_equal(clist.perm.cpu().numpy(), orc["perm"])
