"""nl_pairs_to_host (include/nlcuda.h): the list in host memory must equal the plain field-by-field copy of the device list,
which is how the reference compares its GPU list (Array(...) of every field, test/test_utils.jl:127-131), bit for bit."""
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


def _check(nl, pl, **kw):
    ref = pl.cpu()
    h = nl.to_host(pl, **kw)
    for k in ("first", "i", "j", "S"):
        a = getattr(h, k)
        assert a.dtype == ref[k].dtype and a.shape == ref[k].shape, k
        assert np.array_equal(a, ref[k]), k
    return h


@pytest.mark.parametrize("int_type", [np.int32, np.int64])
@pytest.mark.parametrize("pbc", [(True, True, True), (True, False, True), (False, False, False)])
def test_to_host_equals_field_copies(nl, pbc, int_type):
    import torch
    X, C, L = U.rand_config(20000, seed=61)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), 5.0, C, pbc, int_type=int_type)
    h = _check(nl, pl)
    orc = O.sortbased(X, 5.0, C, pbc)
    assert np.array_equal(h.first, orc["first"])
    U.assert_same_pairs(dict(i=h.i, j=h.j, S=h.S), orc)
    for nt in (1, 3, 64):   # any number of host threads gives the same arrays
        _check(nl, pl, nthreads=nt)
    for f in (0.0, 0.5, 1.0):   # any split of i between the host threads and the bus as well
        _check(nl, pl, i_copy_fraction=f)


def test_to_host_triclinic_many_chunks(nl):
    import torch
    cell = U.TRICLINIC * 12
    X = U.rand_in_cell(300000, cell, seed=62)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), 3.0, cell, (True, True, False))
    assert nl.npairs(pl) > 1 << 20
    _check(nl, pl)


def test_to_host_wide_shifts_take_the_plain_copy(nl):
    # atoms many lattice vectors outside the cell: shift components beyond {-1, 0, 1} cannot be coded in a byte
    import torch
    X, C, L = U.rand_config(3000, seed=63)
    X = U.displace_by_lattice(X, C, (True, True, True), seed=3)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), 5.0, C, (True, True, True))
    assert int(pl.S.abs().max().item()) > 1
    _check(nl, pl)


def test_to_host_small_box_self_images(nl):
    import torch
    X, C, L = U.rand_config(40, seed=64)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), L * 0.9, C, (True, True, True))   # 1 cell per axis, nxyz = 1
    _check(nl, pl)


def test_to_host_empty_and_tiny(nl):
    import torch
    X = np.array([[0.0, 0.0, 0.0], [50.0, 50.0, 50.0]])
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), 1.0, np.eye(3) * 100.0, (False, False, False))
    assert nl.npairs(pl) == 0
    h = _check(nl, pl)
    assert h.first.tolist() == [1, 1, 1]
    X = np.array([[0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [50.0, 50.0, 50.0]])
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), 1.0, np.eye(3) * 100.0, (False, False, False))
    h = _check(nl, pl)
    assert h.i.tolist() == [1, 2] and h.j.tolist() == [2, 1]


def test_to_host_copied_i_and_buffer_reuse(nl):
    import torch
    X, C, L = U.rand_config(5000, seed=65)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), 5.0, C, (True, True, True))
    buf = nl.HostPairBuffers(nl.npairs(pl) + 10, 5000)
    _check(nl, pl, out=buf, rebuild_i=False)       # i copied, as for shard lists whose i carries global indices
    _check(nl, pl, out=buf, rebuild_i=True)
    small = nl.HostPairBuffers(10, 5000)
    with pytest.raises(ValueError):
        nl.to_host(pl, out=small)
    assert nl.to_host_bytes(pl, i_copy_fraction=0.0) == 4 * 5001 + 5 * nl.npairs(pl) + 4


def test_to_host_rejects_first_that_does_not_match(nl):
    import torch
    X, C, L = U.rand_config(2000, seed=66)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), 5.0, C, (True, True, True))
    bad = nl.PairList(X=pl.X, C=pl.C, cutoff=pl.cutoff, i=pl.i[:-3], j=pl.j[:-3], S=pl.S[:-3], first=pl.first, params=pl.params)
    with pytest.raises(nl.NlError) as e:
        nl.to_host(bad)
    assert e.value.code == nl._lib.NL_ERR_BAD_ARG


@pytest.mark.parametrize("int_type", [np.int32, np.int64])
def test_host_out_overlapped_transfer(nl, int_type):
    """neighbour_list(..., host_out=...) = nl_pairs_to_host_begin right after the counting pass + _finish after the fill."""
    import torch
    X, C, L = U.rand_config(30000, seed=67)
    Xd = torch.from_numpy(X).cuda()
    ref = nl.neighbour_list(Xd, 5.0, C, (True, True, False), int_type=int_type).cpu()
    for host_out in (True, nl.HostPairBuffers(len(ref["i"]) + 7, 30000, int_type)):
        for nt in (0, 1, 5):
            h = nl.neighbour_list(Xd, 5.0, C, (True, True, False), int_type=int_type, host_out=host_out, host_threads=nt)
            assert isinstance(h, nl.HostPairList)
            for k in ("first", "i", "j", "S"):
                assert np.array_equal(getattr(h, k), ref[k]), k
    # host positions in, host list out: the whole end-to-end call
    h = nl.neighbour_list(X, 5.0, C, (True, True, False), int_type=int_type, host_out=True)
    assert np.array_equal(h.j, ref["j"]) and np.array_equal(h.i, ref["i"])
    # wide shifts (S copied as it is), an empty list, and a half list
    Xw = U.displace_by_lattice(X[:3000], C, (True, True, True), seed=3)
    pw = nl.neighbour_list(torch.from_numpy(Xw).cuda(), 5.0, C, (True, True, True), int_type=int_type)
    hw = nl.neighbour_list(torch.from_numpy(Xw).cuda(), 5.0, C, (True, True, True), int_type=int_type, host_out=True)
    assert np.array_equal(hw.S, pw.S.cpu().numpy()) and np.array_equal(hw.i, pw.i.cpu().numpy())
    he = nl.neighbour_list(np.array([[0.0, 0, 0], [50.0, 50, 50]]), 1.0, np.eye(3) * 100.0, (False, False, False), int_type=int_type, host_out=True)
    assert he.first.tolist() == [1, 1, 1] and he.i.shape[0] == 0
    ph = nl.neighbour_list(Xd, 5.0, C, (True, True, False), int_type=int_type, half=True)
    hh = nl.neighbour_list(Xd, 5.0, C, (True, True, False), int_type=int_type, half=True, host_out=True)
    assert np.array_equal(hh.j, ph.j.cpu().numpy()) and np.array_equal(hh.first, ph.first.cpu().numpy())
    with pytest.raises(ValueError):
        nl.neighbour_list(Xd, 5.0, C, (True, True, False), with_R=True, host_out=True)
    small = nl.HostPairBuffers(10, 30000, int_type)
    with pytest.raises(ValueError):
        nl.neighbour_list(Xd, 5.0, C, (True, True, False), int_type=int_type, host_out=small)
