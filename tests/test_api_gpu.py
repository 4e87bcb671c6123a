"""API surface of the host mirror, following the reference's test/test_unified_api.jl: return types, lazy == materialised,
int_type, neighbours == neigs == neigss, num_neighbours, max_neighbours, host inputs, error behaviour."""
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


def test_return_types_and_lazy(nl):
    # test_unified_api.jl:26-51
    import torch
    X, C, L = U.rand_config(300, seed=31)
    Xd = torch.from_numpy(X).cuda()
    pl = nl.neighbour_list(Xd, L / 3, C, (True, True, True))
    cl = nl.neighbour_list(Xd, L / 3, C, (True, True, True), lazy=True)
    assert isinstance(pl, nl.PairList) and isinstance(cl, nl.SortedCellList)
    assert pl.i.is_cuda and pl.S.shape == (nl.npairs(pl), 3) and pl.first.shape[0] == 301 and pl.R is None
    assert cl.X_orig.data_ptr() == Xd.data_ptr(), "X_orig aliases the caller's array (src/cell_list.jl:676)"
    assert int(nl.count_neighbours(cl).sum().item()) == nl.npairs(pl)
    assert nl.nsites(pl) == nl.nsites(cl) == 300 and nl.cutoff(pl) == nl.cutoff(cl) == L / 3
    for it, tt in ((np.int32, torch.int32), (np.int64, torch.int64)):
        p2 = nl.neighbour_list(Xd, L / 3, C, (True, True, True), int_type=it)
        assert p2.i.dtype == p2.j.dtype == p2.S.dtype == p2.first.dtype == tt
        assert torch.equal(p2.first.long(), pl.first.long())


def test_accessors(nl):
    # test_unified_api.jl:88-176
    import torch
    X, C, L = U.rand_config(200, seed=32)
    Xd = torch.from_numpy(X).cuda()
    pl = nl.neighbour_list(Xd, L * 0.3, C, (True, False, True))
    cl = nl.neighbour_list(Xd, L * 0.3, C, (True, False, True), lazy=True)
    orc = O.sortbased(X, L * 0.3, C, (True, False, True))
    counts = np.diff(orc["first"])
    assert nl.maxneigs(pl) == nl.max_neighbours(pl) == int(counts.max())
    for i in (1, 7, 200):
        j1, R1 = nl.neigs(pl, i)
        j2, R2, S2 = nl.neigss(pl, i)
        j3, R3, S3 = nl.neighbours(pl, i)
        jc, Rc, Sc = nl.neighbours(cl, i)
        assert torch.equal(j1, j2) and torch.equal(j2, j3) and torch.equal(R1, R2) and torch.equal(S2, S3)
        assert nl.nneigs(pl, i) == nl.num_neighbours(pl, i) == nl.num_neighbours(cl, i) == nl.count_neighbours(cl, i) == counts[i - 1]
        assert sorted(jc.cpu().tolist()) == sorted(j1.cpu().tolist())
        # R from the accessor follows _getR: X[j] - X[i] + C' S, and equals the oracle's R for that row
        row = slice(orc["first"][i - 1] - 1, orc["first"][i] - 1)
        key_e = np.lexsort((S2.cpu().numpy()[:, 2], S2.cpu().numpy()[:, 1], S2.cpu().numpy()[:, 0], j2.cpu().numpy()))
        key_o = np.lexsort((orc["S"][row][:, 2], orc["S"][row][:, 1], orc["S"][row][:, 0], orc["j"][row]))
        assert np.array_equal(R2.cpu().numpy()[key_e], orc["R"][row][key_o])
        seen = []
        nl.for_each_neighbour(lambda j, R, S: seen.append(j), cl, i)
        assert sorted(seen) == sorted(j1.cpu().tolist())
    # single atom: max_neighbours == 0 (test_unified_api.jl:164-175)
    p1 = nl.neighbour_list(torch.tensor([[5.0, 5.0, 5.0]], dtype=torch.float64).cuda(), 3.0, np.eye(3) * 10, (True,) * 3)
    assert nl.maxneigs(p1) == 0 and nl.nneigs(p1, 1) == 0


def test_host_inputs_and_errors(nl):
    import torch
    X, C, L = U.rand_config(500, seed=33)
    pl = nl.neighbour_list(X, L * 0.25, C, (True, True, True))          # numpy input: uploaded, list stays on the device
    orc = O.sortbased(X, L * 0.25, C, (True, True, True))
    assert pl.i.is_cuda
    U.assert_engine_matches_oracle(pl.cpu(), orc, 1e-12, check_R=False)
    with pytest.raises(ValueError):
        nl.neighbour_list(np.zeros((5, 2)), 1.0, C, (True, True, True))   # 2-D systems are rejected (test_atoms_base.jl:135-144)
    with pytest.raises(TypeError):
        nl.neighbour_list(torch.zeros((5, 3), dtype=torch.float16).cuda(), 1.0, C, (True, True, True))
    with pytest.raises(nl.NlError):
        nl.build_cell_list(torch.from_numpy(X).cuda(), 1e-4, C * 1000, (True, True, True))  # too many cells for Int32


def test_fill_rejects_a_workspace_count_did_not_stamp():
    """include/nlcuda.h: nl_fill_pairs* return NL_ERR_WORKSPACE when the workspace is not the one nl_count_pairs filled for
    this very problem (the hit masks, counts and records of the count pass live in it)."""
    import ctypes as C
    import torch
    import neighbourlists_jl_b200 as nl
    from neighbourlists_jl_b200 import _lib, api
    X, cell, _ = U.rand_config(5000, seed=11)
    clist = nl.build_cell_list(torch.from_numpy(X).cuda(), 5.0, cell, (True, True, True))
    pl = nl.materialize_pairlist(clist)          # a good list, for the sizes
    L = _lib.lib()
    N, P = clist.X.shape[0], nl.npairs(pl)
    dev = clist.X.device
    need = L.nl_workspace_bytes(clist.params, N, _lib.NL_STAGE_PAIRS)
    i = torch.empty(P, dtype=torch.int32, device=dev)
    j = torch.empty(P, dtype=torch.int32, device=dev)
    S = torch.empty((P, 3), dtype=torch.int32, device=dev)

    def fill(ws, params=clist.params, n=N):
        return L.nl_fill_pairs(params, api._ptr(clist.X), n, api._ptr(clist.perm), api._ptr(clist.cell_offsets), api._ptr(pl.first),
                               api._ptr(i), api._ptr(j), api._ptr(S), None, api._ptr(ws), ws.numel(), api._stream(dev))

    fresh = torch.zeros(need, dtype=torch.uint8, device=dev)          # never seen by nl_count_pairs
    assert fill(fresh) == _lib.NL_ERR_WORKSPACE
    assert fill(clist._ws) == _lib.NL_OK                                # the stamped one works (and again: fill is repeatable)
    torch.cuda.synchronize()
    assert torch.equal(i, pl.i) and torch.equal(j, pl.j) and torch.equal(S, pl.S)
    half = type(clist.params).from_buffer_copy(clist.params)           # same workspace, different problem (half-list flag)
    half.reserved[0] = _lib.NL_FLAG_HALF
    assert fill(clist._ws, params=half) == _lib.NL_ERR_WORKSPACE


def test_window_promise_is_checked():
    """nl_count_pairs_window: atoms in a z plane of cells the caller declared empty -> NL_ERR_BAD_ARG, not a silently
    wrong list."""
    import torch
    import neighbourlists_jl_b200 as nl
    X, cell, _ = U.rand_config(20000, seed=12)
    clist = nl.build_cell_list(torch.from_numpy(X).cuda(), 5.0, cell, (True, True, True))
    nz = int(clist.ncells[2])
    ok = np.ones(nz, dtype=np.uint8)
    pl = nl.materialize_pairlist(clist, plane_active=ok)
    ref = nl.materialize_pairlist(clist)
    assert torch.equal(pl.first, ref.first)
    bad = ok.copy()
    bad[nz // 2] = 0                                                    # that plane is full of atoms
    with pytest.raises(nl.NlError):
        nl.materialize_pairlist(clist, plane_active=bad)
