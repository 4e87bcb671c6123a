"""Hardware parity of the COMPOSED multi-GPU path: one process per GPU, NCCL inside libnlcuda.so
(nl_shard_prepare / nl_shard_exchange), the unchanged local CUDA stages, i / j as global indices.  Concatenating the ranks'
rows by global i must equal the single-process CPU oracle CSR bit for bit -- the reference's own CPU-vs-GPU comparison
(test/test_utils.jl:127-131, compare_cpu_gpu_full) across ranks.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket
import tempfile

import numpy as np
import pytest

from oracle import nl_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, outdir):
    import torch
    import torch.distributed as dist
    X, cell, pbc, cutoff, mode, driver = case
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}",
                            device_id=torch.device("cuda", rank))
    import neighbourlists_jl_b200  # noqa: F401
    from importlib import import_module
    sh = import_module("neighbourlists_jl_b200.sharded")
    N = X.shape[0]
    if mode == "by_index":       # block distribution by index: rank r starts with a contiguous block of the global index range
        mine = np.arange(rank * N // world, (rank + 1) * N // world)
    else:                        # pre-slabbed: equal-width slabs along the axis the planner will pick (almost nothing moves)
        nc = O.analyze_cell(cell, cutoff, X.dtype)["ncells"]
        axis = 2 - int(np.argmax(nc[::-1]))
        frac = X @ np.linalg.inv(np.asarray(cell, np.float64))
        slab = np.clip(np.floor((frac[:, axis] % 1.0) * world).astype(int), 0, world - 1)
        mine = np.flatnonzero(slab == rank)
    Xl = torch.from_numpy(X[mine]).cuda()
    gl = torch.from_numpy(mine + 1).cuda()
    if driver == "native":
        comm = sh.make_nccl_comm()
        res = sh.neighbour_list_sharded_native(Xl, gl, cutoff, cell, pbc, comm, rank, world, with_R=True)
        # a second list on the same communicator (steady state of an MD loop) must give the same rows
        res2 = sh.neighbour_list_sharded_native(Xl, gl, cutoff, cell, pbc, comm, rank, world, with_R=True)
        assert torch.equal(res.first, res2.first) and torch.equal(res.owned_index, res2.owned_index)
        # the shard's rows in host memory through the compressed transfer (global i is copied, not rebuilt)
        import neighbourlists_jl_b200 as nl
        h = nl.to_host(res)
        for k in ("first", "i", "j", "S"):
            assert np.array_equal(getattr(h, k), getattr(res, k).cpu().numpy()), k
    else:
        res = sh.neighbour_list_sharded(Xl, gl, cutoff, cell, pbc, with_R=True)
    torch.cuda.synchronize()
    np.savez(os.path.join(outdir, f"r{rank}.npz"), owned=res.owned_index.cpu().numpy(), first=res.first.cpu().numpy(), i=res.i.cpu().numpy(),
             j=res.j.cpu().numpy(), S=res.S.cpu().numpy(), R=res.R.cpu().numpy(), bounds=res.plan.bounds, axis=res.plan.axis, n_halo=res.n_halo)
    if driver == "native":
        from neighbourlists_jl_b200 import _lib
        sh.shard_disconnect(comm)
        _lib.check(_lib.lib().nl_nccl_comm_destroy(comm))
    dist.destroy_process_group()


def _run_and_check(world, X, cell, pbc, cutoff, mode, driver="native"):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), (X, cell, pbc, cutoff, mode, driver), d), nprocs=world, join=True)
        parts = [dict(np.load(os.path.join(d, f"r{r}.npz"))) for r in range(world)]
    orc = O.sortbased(X, cutoff, cell, pbc)
    N = X.shape[0]
    owned = np.concatenate([p["owned"] for p in parts]).astype(np.int64)
    assert np.array_equal(np.sort(owned), np.arange(1, N + 1)), "every atom owned exactly once"
    counts = np.zeros(N + 1, np.int64)
    for p in parts:
        counts[p["owned"]] = np.diff(p["first"])
        assert np.array_equal(p["i"], np.repeat(p["owned"], np.diff(p["first"]))), "rows are grouped by owned atom, global i"
    assert np.array_equal(counts[1:], np.diff(orc["first"])), "CSR row sizes"
    mi, mj, mS, mR = O.canonical(np.concatenate([p["i"] for p in parts]), np.concatenate([p["j"] for p in parts]),
                                 np.concatenate([p["S"] for p in parts]), np.concatenate([p["R"] for p in parts]))
    oi, oj, oS, oR = O.canonical(orc["i"], orc["j"], orc["S"], orc["R"])
    assert np.array_equal(mi, oi) and np.array_equal(mj, oj) and np.array_equal(mS, oS), "(i, j, S) differ from the oracle"
    assert np.array_equal(mR, oR), "R differs (same arithmetic on the same positions: must be bit-equal)"
    return parts


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["by_index", "slabbed"])
def test_two_gpus_headline_density_1m_atoms(mode):
    """1 M atoms at the headline density / cutoff, full PBC: z slabs, both halos from the same peer (ring of 2)."""
    X, C, L = U.rand_config(1_000_000, seed=31)
    parts = _run_and_check(2, X, C, (True, True, True), 5.0, mode)
    assert all(int(p["axis"]) == 2 and int(p["n_halo"]) > 0 for p in parts)


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpus_open_slab_axis_x_and_displaced_atoms():
    """Slab axis x (not the key-slowest axis: no plane window), open along x, atoms displaced by lattice vectors on the
    periodic axes (issue #6 pattern), triclinic tilt."""
    cell = np.array([[120.0, 0.0, 0.0], [6.0, 40.0, 0.0], [3.0, 4.0, 36.0]])
    pbc = (False, True, True)
    X = U.displace_by_lattice(U.rand_in_cell(150_000, cell, seed=32), cell, pbc)
    parts = _run_and_check(2, X, cell, pbc, 4.0, "by_index")
    assert all(int(p["axis"]) == 0 for p in parts)


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpus_python_driver_matches_too():
    """The torch.distributed driver (sharded.neighbour_list_sharded with CudaEngine) on the same hardware path."""
    X, C, L = U.rand_config(200_000, seed=33)
    _run_and_check(2, X, C, (True, True, False), 5.0, "by_index", driver="python")


@pytest.mark.skipif(_ngpu() < 4, reason="needs 4 GPUs")
def test_four_gpus_periodic_and_open():
    X, C, L = U.rand_config(600_000, seed=34)
    _run_and_check(4, X, C, (True, True, True), 5.0, "by_index")
    _run_and_check(4, X, C, (True, False, False), 5.0, "slabbed")


def test_single_rank_shard_entry_points():
    """nranks = 1 needs no NCCL: prepare + exchange reduce to the plan of one slab and a copy; the composed list equals the
    plain single-GPU list."""
    import ctypes as C
    import torch
    import neighbourlists_jl_b200 as nl
    from importlib import import_module
    sh = import_module("neighbourlists_jl_b200.sharded")
    X, cell, L = U.rand_config(30_000, seed=35)
    Xd = torch.from_numpy(X).cuda()
    g = torch.arange(1, X.shape[0] + 1, dtype=torch.int32).cuda()
    res = sh.neighbour_list_sharded_native(Xd, g, 5.0, cell, (True, True, True), C.c_void_p(), 0, 1, with_R=True)
    orc = O.sortbased(X, 5.0, cell, (True, True, True))
    assert np.array_equal(res.owned_index.cpu().numpy(), np.arange(1, X.shape[0] + 1))
    U.assert_engine_matches_oracle(dict(first=res.first.cpu().numpy(), i=res.i.cpu().numpy(), j=res.j.cpu().numpy(), S=res.S.cpu().numpy(),
                                        R=res.R.cpu().numpy()), orc, 1e-12, msg="one-rank shard path")
    h = nl.to_host(res)
    for k in ("first", "i", "j", "S"):
        assert np.array_equal(getattr(h, k), getattr(res, k).cpu().numpy()), k
