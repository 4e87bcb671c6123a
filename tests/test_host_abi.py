"""CPU-only checks of the host logic and the C-ABI boundary: the shared library loads, exports every
symbol include/nlcuda.h declares, agrees on the nl_params layout, and validates arguments without
touching a GPU.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import neighbourlists_jl_b200 as nl
from oracle import nl_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nlcuda.h")).read()
    return sorted(set(re.findall(r"NL_API[^;(]*?\b(nl_[A-Za-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = nl._lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 9 and set(syms) == set(nl._lib.EXPORTS)
    for s in syms:
        assert hasattr(L, s), s
    assert L.nl_version() == 200
    assert b"workspace" in L.nl_strerror(nl._lib.NL_ERR_WORKSPACE)
    assert L.nl_strerror(0) == b"ok"


def test_params_layout_and_validation():
    assert C.sizeof(nl._lib.NlParams) == 192
    assert C.sizeof(nl._lib.NlShardInfo) == 1632 + 8 + 8 * 64 + 16     # static_assert in csrc/nlcuda.cu
    assert C.sizeof(nl._lib.NlShardPeers) == 32 + 16 * 64
    geo = nl.cellmath.geometry(np.eye(3) * 20.0, 5.0, (True, True, False), np.float64)
    p = nl._lib.make_params(geo, np.float64, np.int32)
    L = nl._lib.lib()
    assert list(p.ncells) == [4, 4, 4] and list(p.nxyz) == [1, 1, 1] and list(p.pbc) == [1, 1, 0]
    assert L.nl_workspace_bytes(p, 1000, nl._lib.NL_STAGE_BUILD) > 16000
    assert L.nl_workspace_bytes(p, 1000, nl._lib.NL_STAGE_PAIRS) > 1000 * 36
    assert L.nl_workspace_bytes(p, 0, nl._lib.NL_STAGE_PAIRS) > 0
    assert L.nl_workspace_bytes(p, -1, nl._lib.NL_STAGE_BUILD) == 0
    # argument validation happens before any CUDA call
    assert L.nl_build_cells(None, None, 0, None, None, None, None, None, 0, None) == nl._lib.NL_ERR_BAD_ARG
    assert L.nl_build_cells(p, None, 10, None, None, None, None, None, 0, None) == nl._lib.NL_ERR_BAD_ARG
    dummy = (C.c_char * 64)()
    ptr = C.cast(dummy, C.c_void_p)
    assert L.nl_build_cells(p, ptr, 10, ptr, ptr, ptr, ptr, None, 0, None) == nl._lib.NL_ERR_WORKSPACE
    total = C.c_int64(0)
    assert L.nl_count_pairs(p, ptr, 10, ptr, ptr, ptr, C.byref(total), None, 0, None) == nl._lib.NL_ERR_WORKSPACE
    bad = nl._lib.make_params(geo, np.float64, np.int32)
    bad.nxyz[1] = 0
    assert L.nl_workspace_bytes(bad, 10, 0) == 0
    big = nl._lib.make_params(geo, np.float64, np.int64)
    big.ncells[0] = big.ncells[1] = big.ncells[2] = 2000
    assert L.nl_build_cells(big, ptr, 10, ptr, ptr, ptr, ptr, ptr, 64, None) == nl._lib.NL_ERR_UNSUPPORTED
    with pytest.raises(nl.NlError):
        nl._lib.check(nl._lib.NL_ERR_OVERFLOW)
    # reserved bytes: only the documented flag bits are accepted
    flg = nl._lib.make_params(geo, np.float64, np.int32)
    flg.reserved[0] = nl._lib.NL_FLAG_HALF
    assert L.nl_workspace_bytes(flg, 10, nl._lib.NL_STAGE_PAIRS) > 0
    flg.reserved[0] = 2
    assert L.nl_workspace_bytes(flg, 10, nl._lib.NL_STAGE_PAIRS) == 0
    flg.reserved[0] = 0
    flg.reserved[3] = 1
    assert L.nl_build_cells(flg, ptr, 10, ptr, ptr, ptr, ptr, ptr, 64, None) == nl._lib.NL_ERR_BAD_ARG


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_host_cell_analysis_matches_oracle(dtype):
    # cellmath.analyze_cell restates src/cell_list.jl:152-170; must equal the oracle's bit for bit
    rng = np.random.default_rng(1)
    for t in range(300):
        Cm = rng.normal(size=(3, 3)) * 5 + np.eye(3) * 10
        if t % 5 == 0:
            Cm = np.diag(rng.random(3) * 50 + 3)
        if t % 7 == 0:
            Cm[2] *= -1  # left-handed
        rc = rng.random() * 6 + 0.5
        inv, nc, lens, nxyz = nl.cellmath.analyze_cell(Cm, rc, dtype)
        o = O.analyze_cell(Cm, rc, dtype)
        assert np.array_equal(inv, o["inv_mat"]) and np.array_equal(nc, o["ncells"])
        assert np.array_equal(lens, o["lens"]) and np.array_equal(nxyz, o["nxyz"])


def test_headline_geometry():
    # 10 M atoms -> L = 584.80, 116^3 cells; C3 triclinic -> (60, 53, 47)
    L = (1e7 / 0.05) ** (1 / 3)
    g = nl.cellmath.geometry(np.eye(3) * L, 5.0, (True, True, True), np.float64)
    assert g.ncells.tolist() == [116, 116, 116] and g.nxyz.tolist() == [1, 1, 1]
    s = (1e6 / 0.05 / 720) ** (1 / 3)
    g = nl.cellmath.geometry(s * np.array([[10, 2, 1], [0, 9, 1.5], [0, 0, 8.0]]), 5.0, (True, True, False), np.float64)
    assert g.ncells.tolist() == [60, 53, 47]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nl.NlError):
        nl.neighbour_list(np.zeros((4, 3)), 1.0, np.eye(3) * 5, (True, True, True))


def test_shard_plan_matches_host_logic():
    """nl_shard_plan (C, for the Julia driver) == sharded.plan_slabs (Python driver) on random histograms."""
    from importlib import import_module
    sh = import_module("neighbourlists_jl_b200.sharded")
    L = nl._lib.lib()
    rng = np.random.default_rng(4)
    for t in range(300):
        nplanes = int(rng.integers(1, 200))
        world = int(rng.integers(1, 9))
        halo = int(rng.integers(0, 3))
        hist = rng.integers(0, 1000, nplanes).astype(np.int64)
        if t % 7 == 0:
            hist[rng.integers(0, nplanes)] += 10 ** 6  # one very dense plane
        out = np.zeros(world + 1, np.int64)
        rc = L.nl_shard_plan(hist.ctypes.data_as(C.c_void_p), nplanes, world, halo, out.ctypes.data_as(C.c_void_p))
        minw = 2 * halo + 1 if world > 1 else 1
        if world * minw > nplanes:
            assert rc == nl._lib.NL_ERR_BAD_ARG
            with pytest.raises(ValueError):
                sh.plan_slabs(hist, world, halo, True, 0)
            continue
        assert rc == 0
        ref = sh.plan_slabs(hist, world, halo, True, 0).bounds
        assert np.array_equal(out, ref), (hist, world, halo, out, ref)
        assert out[0] == 0 and out[-1] == nplanes and np.all(np.diff(out) >= minw)


def test_accessor_entry_points_validate_arguments():
    # SURVEY 8f entry points: argument errors come back before any CUDA call
    geo = nl.cellmath.geometry(np.eye(3) * 20.0, 5.0, (True, True, False), np.float64)
    p = nl._lib.make_params(geo, np.float64, np.int32)
    L = nl._lib.lib()
    E = nl._lib
    dummy = (C.c_char * 64)()
    ptr = C.cast(dummy, C.c_void_p)
    assert L.nl_pairs_R(None, ptr, 1, ptr, ptr, ptr, 0, 1, ptr, None) == E.NL_ERR_BAD_ARG
    assert L.nl_pairs_R(p, ptr, 1, ptr, ptr, ptr, 3, 2, ptr, None) == E.NL_ERR_BAD_ARG
    assert L.nl_pairs_R(p, None, 1, ptr, ptr, ptr, 0, 1, ptr, None) == E.NL_ERR_BAD_ARG
    assert L.nl_pairs_R(p, None, 1, None, None, None, 4, 4, None, None) == E.NL_OK  # empty range: nothing to do
    assert L.nl_max_neighbours(p, ptr, 0, ptr, None) == E.NL_ERR_BAD_ARG  # maximum over an empty collection throws
    assert L.nl_max_neighbours(p, None, 5, ptr, None) == E.NL_ERR_BAD_ARG
    assert L.nl_rows_padded(p, ptr, 5, ptr, ptr, ptr, ptr, -1, 4, ptr, ptr, ptr, ptr, None) == E.NL_ERR_BAD_ARG
    assert L.nl_rows_padded(p, ptr, 5, ptr, ptr, ptr, None, 2, 4, ptr, ptr, ptr, ptr, None) == E.NL_ERR_BAD_ARG
    assert L.nl_rows_padded(p, ptr, 5, ptr, ptr, ptr, ptr, 0, 4, ptr, ptr, ptr, ptr, None) == E.NL_OK
    assert L.nl_lazy_neighbours(p, ptr, ptr, 5, ptr, ptr, None, 2, 4, ptr, ptr, ptr, ptr, None) == E.NL_ERR_BAD_ARG
    assert L.nl_lazy_neighbours(p, ptr, ptr, 5, ptr, ptr, ptr, 0, 4, ptr, ptr, ptr, ptr, None) == E.NL_OK
    assert L.nl_lazy_neighbours(p, ptr, ptr, 5, ptr, ptr, ptr, 2, -1, ptr, ptr, ptr, ptr, None) == E.NL_ERR_BAD_ARG
    total = C.c_int64(0)
    assert L.nl_count_pairs_window(p, ptr, 10, ptr, ptr, ptr, C.byref(total), None, None, 0, None) == E.NL_ERR_WORKSPACE
    assert L.nl_count_pairs_window(p, None, 10, ptr, ptr, ptr, C.byref(total), None, ptr, 1 << 30, None) == E.NL_ERR_BAD_ARG
    assert L.nl_fill_pairs_window(p, ptr, 10, ptr, ptr, ptr, 11, None, None, ptr, ptr, ptr, None, ptr, 1 << 30, None) == E.NL_ERR_BAD_ARG
    assert L.nl_fill_pairs_window(p, ptr, 10, ptr, ptr, ptr, 0, None, None, ptr, ptr, ptr, None, ptr, 1 << 30, None) == E.NL_OK
    assert L.nl_bounding_box(7, ptr, 5, ptr, ptr, 1 << 20, None) == E.NL_ERR_BAD_ARG
    assert L.nl_bounding_box(E.NL_F64, ptr, 0, ptr, ptr, 1 << 20, None) == E.NL_ERR_BAD_ARG
    assert L.nl_bounding_box(E.NL_F64, ptr, 5, ptr, None, 0, None) == E.NL_ERR_WORKSPACE
    assert L.nl_max_displacement2(E.NL_F32, None, ptr, 5, ptr, ptr, 1 << 20, None) == E.NL_ERR_BAD_ARG
    assert L.nl_max_displacement2(E.NL_F32, ptr, ptr, 5, ptr, ptr, 16, None) == E.NL_ERR_WORKSPACE
    assert E.NL_REDUCE_WS_BYTES == int(re.search(r"#define NL_REDUCE_WS_BYTES (\d+)", open(os.path.join(ROOT, "include", "nlcuda.h")).read()).group(1))


def test_accessor_restatements_known_answers():
    # the numpy restatements used as the checker for the 8f rows, pinned on the reference's own known answers
    h3 = np.array([[0., 0., 0.], [0., 0., 2.], [0., 10., 0.]])
    Cb = O.bounding_cell(h3)
    assert np.array_equal(Cb, np.diag([1.0, 11.0, 3.0]))  # ext/NeighbourListsAtomsBaseExt.jl:27-31 (bbox + 1)
    d = O.sortbased(h3, 5.0, Cb, (False, False, False))
    assert d["i"].tolist() == [1, 2] and d["j"].tolist() == [2, 1]  # test_atoms_base.jl:71-104
    R = O.pairs_R(h3, d["i"], d["j"], d["S"], Cb)
    assert R.tolist() == [[0.0, 0.0, 2.0], [0.0, 0.0, -2.0]] and np.array_equal(R, d["R"])
    assert O.maxneigs(d["first"]) == 1
    with pytest.raises(ValueError):
        O.maxneigs(np.array([1]))
    # _getR == the traversal's R for a periodic triclinic case with shifts (same expression)
    from tests import util as U
    Ct = U.TRICLINIC.copy()
    X = U.displace_by_lattice(U.rand_in_cell(200, Ct, seed=5), Ct, (True, True, True))
    d = O.sortbased(X, 3.0, Ct, (True, True, True))
    assert (d["S"] != 0).any() and np.array_equal(O.pairs_R(X, d["i"], d["j"], d["S"], Ct), d["R"])
    n, j, S, Rp = O.rows_padded(X, d["first"], d["j"], d["S"], Ct, [1, 200], 4)
    assert n.tolist() == [int(d["first"][1] - d["first"][0]), int(d["first"][200] - d["first"][199])] and j.shape == (2, 4)
    Z = np.round(X)
    assert O.max_displacement2(Z + 0.5, Z) == 0.75


def test_units_and_system_dispatch_host_logic():
    from neighbourlists_jl_b200 import atoms
    assert atoms._unit_scale("Å", 3.5) == (3.5, 1.0)
    v, s = atoms._unit_scale("Å", (0.35, "nm"))
    assert v == 0.35 and abs(s - 0.1) < 1e-15
    v, s = atoms._unit_scale("Å", (350.0, "pm"))
    assert abs(s - 100.0) < 1e-9
    with pytest.raises(ValueError):
        atoms._unit_scale("Å", (1.0, "furlong"))
    assert atoms.is_system(atoms.isolated_system(np.zeros((2, 3)))) and not atoms.is_system(np.zeros((2, 3)))


def test_lj_force_restatement_is_minus_gradient():
    # pins the checker of nl_lazy_lj_forces: F = -dE/dX (central differences) and sum(e) = the oracle's LJ energy
    from tests import util as U
    X, C = U.fcc(3.61, (3, 3, 3))
    rng = np.random.default_rng(11)
    X = X + rng.uniform(-0.05, 0.05, size=X.shape)
    pbc, rc, eps, sig = (True, True, False), 4.6, 0.7, 2.3  # cutoff between the 3rd (4.42) and 4th (5.11) fcc shells
    N = X.shape[0]

    def energy(Y):
        d = O.sortbased(Y, rc, C, pbc)
        return O.lj_forces(d, eps, sig, N)[1].sum() / 2.0, d

    E0, d0 = energy(X)
    F, e = O.lj_forces(d0, eps, sig, N)
    assert abs(e.sum() - O.lj_energy(O.sortbased(X, rc, C, pbc, lazy=True), eps, sig)) <= 1e-9 * abs(e.sum())
    h = 1e-5
    for n, k in ((0, 0), (17, 1), (55, 2), (107, 0)):
        Xp, Xm = X.copy(), X.copy()
        Xp[n, k] += h
        Xm[n, k] -= h
        fd = -(energy(Xp)[0] - energy(Xm)[0]) / (2 * h)
        assert abs(fd - F[n, k]) <= 1e-6 * max(1.0, abs(F[n, k])), (n, k, fd, F[n, k])
    assert np.abs(F.sum(axis=0)).max() <= 1e-9 * np.abs(F).max() * N  # Newton's third law


def test_round2_entry_points_validate_arguments():
    """Shard peer path and host transfer: every argument check happens before the first CUDA / NCCL call."""
    L = nl._lib.lib()
    E = nl._lib
    geo = nl.cellmath.geometry(np.eye(3) * 40.0, 5.0, (True, True, True), np.float64)
    p = E.make_params(geo, np.float64, np.int32)
    dummy = (C.c_char * 256)()
    ptr = C.cast(dummy, C.c_void_p)
    peers = E.NlShardPeers()
    info = E.NlShardInfo()
    # connect: null outputs, bad ranks, capacity < 1, workspace missing / too small
    assert L.nl_shard_connect(p, 1000, None, 0, 2, ptr, 256, None, None) == E.NL_ERR_BAD_ARG
    assert L.nl_shard_connect(p, 1000, None, 2, 2, ptr, 256, C.byref(peers), None) == E.NL_ERR_BAD_ARG
    assert L.nl_shard_connect(p, 0, None, 0, 2, ptr, 256, C.byref(peers), None) == E.NL_ERR_BAD_ARG
    assert L.nl_shard_connect(p, 1000, None, 0, 2, None, 0, C.byref(peers), None) == E.NL_ERR_WORKSPACE
    assert L.nl_shard_workspace_bytes(p, 1000, 2) > 1000 * 28 * 2          # send + halo buffers
    assert L.nl_shard_workspace_bytes(p, 1000, 65) == 0                    # more ranks than NL_MAX_RANKS
    # exchange_peer: peers must belong to the info and to this workspace, and the capacity must suffice
    info.nranks, info.rank, info.n_local, info.n_max_all = 2, 0, 10, 5000
    peers.nranks, peers.rank, peers.cap, peers.ws_bytes, peers.ws = 2, 0, 1000, 256, ptr.value
    assert L.nl_shard_exchange_peer(p, C.byref(info), ptr, ptr, 10, None, None, ptr, ptr, None, ptr, 256, None) == E.NL_ERR_BAD_ARG
    assert L.nl_shard_exchange_peer(p, C.byref(info), ptr, ptr, 10, None, C.byref(peers), ptr, ptr, None, ptr, 256, None) == E.NL_ERR_WORKSPACE
    info.n_max_all = 500
    assert L.nl_shard_exchange_peer(p, C.byref(info), ptr, ptr, 10, None, C.byref(peers), ptr, ptr, None, ptr, 128, None) == E.NL_ERR_WORKSPACE
    peers.rank = 1
    assert L.nl_shard_exchange_peer(p, C.byref(info), ptr, ptr, 10, None, C.byref(peers), ptr, ptr, None, ptr, 256, None) == E.NL_ERR_BAD_ARG
    assert L.nl_shard_disconnect(None) == E.NL_ERR_BAD_ARG
    fresh = E.NlShardPeers()
    assert L.nl_shard_disconnect(C.byref(fresh)) == 0                       # nothing mapped: nothing to do
    # host transfer
    job = C.c_void_p()
    assert L.nl_pairs_to_host_begin(p, None, 3, 5, ptr, ptr, 0, None, C.byref(job)) == E.NL_ERR_BAD_ARG        # first missing
    assert L.nl_pairs_to_host_begin(p, ptr, 3, 5, ptr, None, 0, None, C.byref(job)) == E.NL_ERR_BAD_ARG         # i_host missing, P > 0
    assert L.nl_pairs_to_host_begin(p, ptr, -1, 5, ptr, ptr, 0, None, C.byref(job)) == E.NL_ERR_BAD_ARG
    assert L.nl_pairs_to_host_begin(p, ptr, 3, 5, ptr, ptr, 0, None, None) == E.NL_ERR_BAD_ARG
    assert L.nl_pairs_to_host_finish(None, ptr, ptr, ptr, ptr, ptr, ptr, 256, None) == E.NL_ERR_BAD_ARG
    # whole-list call: i_copy_from out of range, row_index without a host buffer, scratch too small
    assert L.nl_pairs_to_host(p, ptr, 3, None, 6, None, None, ptr, ptr, 5, ptr, ptr, ptr, ptr, ptr, ptr, 4096, 0, None) == E.NL_ERR_BAD_ARG
    assert L.nl_pairs_to_host(p, ptr, 3, None, 2, None, None, ptr, ptr, 5, ptr, ptr, ptr, ptr, ptr, ptr, 4096, 0, None) == E.NL_ERR_BAD_ARG
    assert L.nl_pairs_to_host(p, ptr, 3, None, 5, ptr, None, ptr, ptr, 5, ptr, ptr, ptr, ptr, ptr, ptr, 4096, 0, None) == E.NL_ERR_BAD_ARG
    assert L.nl_pairs_to_host(p, ptr, 3, None, 5, None, None, ptr, ptr, 5, ptr, ptr, ptr, ptr, ptr, ptr, 16, 0, None) == E.NL_ERR_WORKSPACE


def test_header_is_plain_c_and_the_c_example_builds():
    """include/nlcuda.h must be usable from C99 without CUDA or C++ headers (the drop-in boundary); examples/host_list.c is the
    whole path -- host positions in, host list out -- written against it (ran on a B200: 26.2 pairs per atom)."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    probe = '#include "nlcuda.h"\nint main(void) { nl_params p; nl_shard_info s; nl_shard_peers q; (void)p; (void)s; (void)q; return NL_VERSION != 200; }\n'
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"), "-x", "c", "-"], input=probe,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cuda_inc = "/usr/local/cuda/include"
    if os.path.isdir(cuda_inc):
        r = subprocess.run([gcc, "-std=c99", "-Wall", "-fsyntax-only", "-I", os.path.join(root, "include"), "-I", cuda_inc,
                            os.path.join(root, "examples", "host_list.c")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
