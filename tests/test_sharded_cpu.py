"""world_size-2 (and 3) gloo tests of the slab sharding host logic on CPU: slab plan, all-to-all-v
redistribution and halo exchange.  The local stages are played by the CPU oracle (test infrastructure);
concatenating the ranks' rows by global i must reproduce the single-process oracle CSR bit for bit."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nl_oracle as O
from tests import util as U


class OracleEngine:
    """Stands in for CudaEngine: the same two local stages, computed by the oracle on CPU tensors."""
    device = torch.device("cpu")

    def cell_ids(self, X, cutoff, cell, pbc):
        Xn = X.numpy()
        r = O.sortbased(Xn, cutoff, cell, pbc, dtype=Xn.dtype, int_type=np.int64, lazy=True)
        ids = np.empty(Xn.shape[0], np.int64)
        ids[r["perm"] - 1] = r["cell_id"]
        return torch.from_numpy(ids)

    def build(self, X_all, n_owned, gmap, cutoff, cell, pbc, int_type, with_R, plane_active=None):
        if plane_active is not None:  # the promise made to the windowed entry points: no local atom outside the active planes
            ids = self.cell_ids(X_all, cutoff, cell, pbc).numpy() - 1
            g = O.analyze_cell(cell, cutoff, X_all.numpy().dtype)
            nc = [int(v) for v in g["ncells"]]
            assert plane_active.shape == (nc[2],) and plane_active[ids // (nc[0] * nc[1])].all()
        Xn = X_all.numpy()
        r = O.sortbased(Xn, cutoff, cell, pbc, dtype=Xn.dtype, int_type=int_type)
        P = int(r["first"][n_owned]) - 1
        g = gmap.numpy().astype(int_type)
        return dict(first=torch.from_numpy(r["first"][:n_owned + 1].copy()), i=torch.from_numpy(g[r["i"][:P] - 1]),
                    j=torch.from_numpy(g[r["j"][:P] - 1]), S=torch.from_numpy(r["S"][:P].copy()), R=torch.from_numpy(r["R"][:P].copy()))


def _worker(rank, world, port, case, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import neighbourlists_jl_b200  # noqa: F401
    from importlib import import_module
    sh = import_module("neighbourlists_jl_b200.sharded")
    X, cell, pbc, cutoff = case[:4]
    mode = case[4] if len(case) > 4 else "by_index"
    N = X.shape[0]
    if mode == "by_index":
        # block distribution BY INDEX (not by space): rank r starts with atoms r::world
        mine = np.arange(rank, N, world)
        kw = {}
    else:
        # spatially pre-distributed: equal-width z slabs of the box (nothing has to move if the plan agrees);
        # "placed" additionally promises it (redistribute=False), "misplaced" breaks the promise on rank 0
        z = X[:, 2] / cell[2, 2]
        slab = np.minimum((z * world).astype(int), world - 1)
        if mode == "misplaced":
            slab = (slab + 1) % world
        mine = np.flatnonzero(slab == rank)
        kw = {} if mode == "prespatial" else dict(redistribute=False)
    try:
        res = sh.neighbour_list_sharded(torch.from_numpy(X[mine]), torch.from_numpy(mine + 1), cutoff, cell, pbc, engine=OracleEngine(),
                                        with_R=True, **kw)
    except ValueError as e:
        open(os.path.join(outdir, f"r{rank}.err"), "w").write(str(e))
        dist.destroy_process_group()
        return
    np.savez(os.path.join(outdir, f"r{rank}.npz"), owned=res.owned_index.numpy(), first=res.first.numpy(), i=res.i.numpy(),
             j=res.j.numpy(), S=res.S.numpy(), R=res.R.numpy(), bounds=res.plan.bounds, axis=res.plan.axis, n_halo=res.n_halo)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, case):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), case, d), nprocs=world, join=True)
        return [dict(np.load(os.path.join(d, f"r{r}.npz"))) for r in range(world)]


def _check(world, X, cell, pbc, cutoff):
    parts = _run(world, (X, cell, pbc, cutoff))
    orc = O.sortbased(X, cutoff, cell, pbc)
    N = X.shape[0]
    owned = np.concatenate([p["owned"] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(1, N + 1)), "every atom owned exactly once"
    counts = np.zeros(N + 1, np.int64)
    for p in parts:
        counts[p["owned"]] = np.diff(p["first"])
        assert np.array_equal(p["i"], np.repeat(p["owned"], np.diff(p["first"]))), "rows are grouped by owned atom, global i"
    assert np.array_equal(counts[1:], np.diff(orc["first"])), "CSR row sizes"
    merged = dict(i=np.concatenate([p["i"] for p in parts]), j=np.concatenate([p["j"] for p in parts]),
                  S=np.concatenate([p["S"] for p in parts]), R=np.concatenate([p["R"] for p in parts]))
    mi, mj, mS, mR = O.canonical(merged["i"], merged["j"], merged["S"], merged["R"])
    oi, oj, oS, oR = O.canonical(orc["i"], orc["j"], orc["S"], orc["R"])
    assert np.array_equal(mi, oi) and np.array_equal(mj, oj) and np.array_equal(mS, oS)
    assert np.array_equal(mR, oR)  # same arithmetic on the same positions
    return parts


def test_two_ranks_periodic():
    # elongated periodic box: slabs along z (the axis with the most cells); with G = 2 both halos come from the same peer
    cell = np.diag([12.0, 12.0, 48.0])
    X = U.rand_in_cell(1500, cell, seed=5)
    parts = _check(2, X, cell, (True, True, True), 3.0)
    assert all(int(p["axis"]) == 2 and int(p["n_halo"]) > 0 for p in parts)


def test_two_ranks_open_axis_and_displaced():
    cell = np.diag([40.0, 10.0, 10.0])
    X = U.displace_by_lattice(U.rand_in_cell(1200, cell, seed=6), cell, (False, True, True))
    _check(2, X, cell, (False, True, True), 2.5)


def test_three_ranks_triclinic_unbalanced():
    cell = 2.0 * U.TRICLINIC * np.array([[3.0], [1.0], [1.0]])
    rng = np.random.Generator(np.random.PCG64(7))
    f = rng.random((2000, 3))
    f[:, 0] = f[:, 0] ** 2  # denser at one end: unequal slab widths
    X = f @ cell
    parts = _check(3, X, cell, (True, True, False), 3.0)
    widths = np.diff(parts[0]["bounds"])
    assert widths.min() >= 3 and len(set(widths.tolist())) > 1


def test_plan_rejects_too_many_ranks():
    from importlib import import_module
    import neighbourlists_jl_b200  # noqa: F401
    sh = import_module("neighbourlists_jl_b200.sharded")
    with pytest.raises(ValueError):
        sh.plan_slabs(np.ones(5, np.int64), 2, 1, True, 0)
    p = sh.plan_slabs(np.ones(6, np.int64), 2, 1, True, 0)
    assert p.bounds.tolist() == [0, 3, 6]


def test_pre_distributed_inputs_skip_the_redistribution():
    # a uniform grid of atoms: 24 z planes of 3 A, 2 equal slabs = the balanced plan, so no atom has to move
    cell = np.diag([9.0, 9.0, 72.0])
    g = np.stack(np.meshgrid(np.arange(6) * 1.5 + 0.3, np.arange(6) * 1.5 + 0.4, np.arange(48) * 1.5 + 0.2, indexing="ij"), -1).reshape(-1, 3)
    for mode in ("prespatial", "placed"):
        parts = _run(2, (g, cell, (True, True, True), 3.0, mode))
        orc = O.sortbased(g, 3.0, cell, (True, True, True))
        assert parts[0]["bounds"].tolist() == [0, 12, 24]
        merged = dict(i=np.concatenate([p["i"] for p in parts]), j=np.concatenate([p["j"] for p in parts]), S=np.concatenate([p["S"] for p in parts]))
        U.assert_same_pairs(merged, orc, mode)


def test_misplaced_atoms_without_redistribution_are_an_error():
    cell = np.diag([9.0, 9.0, 72.0])
    X = U.rand_in_cell(800, cell, seed=9)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), (X, cell, (True, True, True), 3.0, "misplaced"), d), nprocs=2, join=True)
        errs = [f for f in os.listdir(d) if f.endswith(".err")]
        assert len(errs) == 2 and "not in this rank's slab" in open(os.path.join(d, errs[0])).read()


def test_three_ranks_z_slabs_open_axis_plane_window():
    # slabs along an OPEN z axis: the plane window of the end ranks is clipped (no wrap); OracleEngine.build asserts the promise
    cell = np.diag([10.0, 10.0, 60.0])
    X = U.rand_in_cell(1500, cell, seed=11)
    parts = _check(3, X, cell, (True, True, False), 2.5)
    assert all(int(p["axis"]) == 2 for p in parts)
