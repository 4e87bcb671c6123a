"""GPU parity: libnlcuda.so (through the C ABI, via the host mirror) against the CPU oracle on the
same seeded inputs.  Mirrors the structure of the reference's test/test_gpu.jl and
test/test_sortbased.jl.  Bars: SortedCellList fields, CSR offsets and the (i,j,S) set bit-exact;
R within 1e-12 relative (Float64) / 1e-5 (Float32)."""
import numpy as np
import pytest

from oracle import nl_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu

RTOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


@pytest.fixture(scope="module")
def nl():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import neighbourlists_jl_b200 as nl
    nl._lib.lib()  # fail loudly if the extension is missing
    return nl


def run_engine(nl, X, cutoff, cell, pbc, int_type=np.int32, with_R=True):
    import torch
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    clist = nl.build_cell_list(Xd, cutoff, cell, pbc, int_type=int_type)
    pl = nl.materialize_pairlist(clist, with_R=with_R)
    torch.cuda.synchronize()
    return clist, pl


def check_case(nl, X, cutoff, cell, pbc, dtype=np.float64, int_type=np.int32, msg=""):
    X = np.asarray(X, dtype=dtype)
    clist, pl = run_engine(nl, X, cutoff, cell, pbc, int_type)
    geo = dict(inv=np.asarray(clist.inv_cell).ravel(order="F"), ncells=clist.ncells, nxyz=clist.geo.nxyz)
    orc = O.sortbased(X, cutoff, cell, pbc, dtype=dtype, int_type=int_type, geo=geo)
    # host geometry == oracle's own analyze_cell
    own = O.analyze_cell(cell, cutoff, dtype)
    assert np.array_equal(own["ncells"], clist.ncells) and np.array_equal(own["nxyz"], clist.geo.nxyz), msg
    assert np.array_equal(own["inv_mat"], clist.inv_cell), msg
    # SortedCellList fields, bit-exact (stable sort => same perm as the CPU sortperm)
    assert clist.perm.dtype == pl.i.dtype
    assert np.array_equal(clist.perm.cpu().numpy(), orc["perm"]), f"{msg}: perm"
    assert np.array_equal(clist.cell_id.cpu().numpy(), orc["cell_id"]), f"{msg}: cell_id"
    assert np.array_equal(clist.cell_offsets.cpu().numpy(), orc["cell_offsets"]), f"{msg}: cell_offsets"
    assert np.array_equal(clist.X.cpu().numpy(), orc["Xs"]), f"{msg}: X sorted"
    eng = pl.cpu()
    assert eng["i"].dtype == np.dtype(int_type) and eng["S"].dtype == np.dtype(int_type) and eng["first"].dtype == np.dtype(int_type)
    U.assert_engine_matches_oracle(eng, orc, RTOL[np.dtype(dtype)], msg=msg)
    # lazy sinks
    counts = nl.count_neighbours(clist).cpu().numpy()
    assert np.array_equal(counts.astype(np.int64), np.diff(orc["first"].astype(np.int64))), f"{msg}: lazy counts"
    return clist, pl, orc


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("int_type", [np.int32, np.int64])
def test_random_configs(nl, dtype, int_type):
    # test/test_gpu.jl:43-47 (5 random configs, N in 50:200, cutoff L/4 via test_cpu_vs_gpu)
    rng = np.random.default_rng(7)
    for k in range(5):
        N = int(rng.integers(50, 201))
        X, C, L = U.rand_config(N, seed=100 + k, dtype=dtype)
        check_case(nl, X, L * 0.25, C, (True, True, True), dtype, int_type, msg=f"rand{k}")
        check_case(nl, X, L / 3, C, (True, True, True), dtype, int_type, msg=f"rand{k} L/3")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_all_pbc(nl, dtype):
    for k, pbc in enumerate(U.ALL_PBC):
        X, C, L = U.rand_config(120, seed=200 + k, dtype=dtype)
        check_case(nl, X, L * 0.25, C, pbc, dtype, msg=f"pbc{pbc}")
        Xd = U.displace_by_lattice(X, C, pbc)
        check_case(nl, Xd, L * 0.25, C, pbc, dtype, msg=f"pbc{pbc} displaced")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_triclinic_all_pbc(nl, dtype):
    # test/test_gpu.jl:53-60 plus every pbc combination and displaced atoms
    for k, pbc in enumerate(U.ALL_PBC):
        X = U.rand_in_cell(80, U.TRICLINIC, seed=300 + k, dtype=dtype)
        check_case(nl, X, 3.0, U.TRICLINIC.astype(dtype), pbc, dtype, msg=f"tri{pbc}")
        Xd = U.displace_by_lattice(X, U.TRICLINIC, pbc)
        check_case(nl, Xd, 3.0, U.TRICLINIC.astype(dtype), pbc, dtype, msg=f"tri{pbc} displaced")


def test_edge_cases(nl):
    # test/test_utils.jl:480-494
    C = np.eye(3) * 10.0
    _, pl, _ = check_case(nl, [[5.0, 5.0, 5.0]], 3.0, C, (True, True, True), msg="single")
    assert nl.npairs(pl) == 0 and pl.first.cpu().tolist() == [1, 1]
    _, pl, _ = check_case(nl, [[5.0, 5.0, 5.0], [5.0, 5.0, 6.0]], 3.0, C, (True, True, True), msg="two")
    assert nl.npairs(pl) == 2
    _, pl, _ = check_case(nl, [[1.0, 1.0, 1.0], [8.0, 8.0, 8.0]], 3.0, C, (False, False, False), msg="far")
    assert nl.npairs(pl) == 0 and pl.first.cpu().tolist() == [1, 1, 1]


def test_empty(nl):
    # nat == 0: first = [1], cell_offsets all ones (src/gpu_kernels.jl:268-271, 303-312)
    import torch
    C = np.eye(3) * 10.0
    clist = nl.build_cell_list(torch.zeros((0, 3), dtype=torch.float64, device="cuda"), 3.0, C, (True, True, True))
    assert clist.cell_offsets.cpu().tolist() == [1] * 28
    pl = nl.materialize_pairlist(clist)
    assert pl.first.cpu().tolist() == [1] and nl.npairs(pl) == 0 and nl.nsites(pl) == 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_large_systems(nl, dtype):
    # test/test_utils.jl:540-549 (N = 500, 1000, 2000; cutoff 0.25 L) -- full comparison, not just counts
    for N in (500, 1000, 2000):
        X, C, L = U.rand_config(N, seed=N, dtype=dtype)
        check_case(nl, X, L * 0.25, C, (True, True, True), dtype, msg=f"N={N}")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_high_density(nl, dtype):
    # test/test_gpu.jl:78-86: 200 atoms, L = 5, rc = 2
    rng = np.random.Generator(np.random.PCG64(5))
    X = (rng.random((200, 3)) * 5.0).astype(dtype)
    check_case(nl, X, 2.0, (np.eye(3) * 5.0).astype(dtype), (True, True, True), dtype, msg="dense")


def test_elongated_and_large_cutoff(nl):
    # test/test_sortbased.jl:36-41, 49-55
    C2 = np.diag([5.0, 5.0, 20.0])
    X2 = U.rand_in_cell(80, C2, seed=11)
    check_case(nl, X2, 3.0, C2, (True, True, True), msg="elongated")
    X, C, L = U.rand_config(30, seed=12)
    check_case(nl, X, L * 0.6, C, (True, True, True), msg="cutoff 0.6 L")
    check_case(nl, X, L * 1.7, C, (True, False, True), msg="cutoff 1.7 L mixed")


def test_fcc_cu(nl):
    # BASELINE config 2: fcc Cu 4x4x4, rc = 5 -> 10 752 pairs, 42 per atom, 3 720 with S != 0
    for a in (3.61, 3.615):
        X, C = U.fcc(a)
        _, pl, _ = check_case(nl, X, 5.0, C, (True, True, True), msg=f"fcc a={a}")
        e = pl.cpu()
        assert nl.npairs(pl) == 10752 and np.all(np.diff(e["first"]) == 42) and int((np.abs(e["S"]).sum(1) > 0).sum()) == 3720
    # test/test_atoms_base.jl:13-69: rc = 3.5 -> 12 neighbours for every atom
    for reps in ((4, 2, 3), (3, 3, 3), (2, 2, 2)):
        X, C = U.fcc(3.61, reps)
        _, pl, _ = check_case(nl, X, 3.5, C, (True, True, True), msg=f"fcc {reps}")
        assert np.all(np.diff(pl.cpu()["first"]) == 12)


def test_issue6_outside_box(nl):
    # test/test_sortbased.jl:143-228
    C = np.eye(3) * 8.0
    for X in ([[0.5, 0.5, 0.5], [0.5, 1.5, 0.5]], [[0.5, 0.5, 0.5], [0.5, -0.5, 0.5]], [[0.5, 0.5, 0.5], [0.5, -6.5, 0.5]]):
        _, pl, _ = check_case(nl, X, 1.5, C, (True, True, True), msg="issue6")
        e = pl.cpu()
        assert nl.npairs(pl) == 2
        assert np.allclose(np.linalg.norm(e["R"], axis=1), 1.0, atol=1e-12)
    tiny = -5e-17 * 8.0
    X = np.array([[0.5, 8.0 - 0.3, 0.5], [0.5, tiny, 0.5]])
    clist, pl, _ = check_case(nl, X, 1.5, C, (True, True, True), msg="tiny negative frac")
    j, R, S = nl.neighbours(pl, 1)
    assert 2 in j.cpu().tolist()
    k = j.cpu().tolist().index(2)
    Rk = R[k].cpu().numpy()
    assert abs(np.linalg.norm(Rk) - 0.3) < 1e-9
    assert np.allclose(X[1] - X[0] + S[k].cpu().numpy() @ C, Rk, atol=1e-12)
    # far excursions: windings beyond the packed 10-bit range take the recompute path
    X, C, L = U.rand_config(100, seed=66)
    Xf = X.copy()
    Xf[::3, 0] += 1000 * L
    Xf[1::3, 2] -= 777 * L
    check_case(nl, Xf, L * 0.25, C, (True, True, True), msg="far windings")


def test_left_handed_and_hcp_fixtures(nl):
    # harvested from the reference's dead test/test_julip.jl:80-86, :126-133
    X = np.array([[0.0, 0.0, 0.0], [1.92333044, 6.63816518e-17, -1.36], [1.92333044, 1.92333044, -2.72], [3.84666089, 1.92333044, -4.08]])
    C = np.diag([3.84666089, 3.84666089, -5.44])
    check_case(nl, X, 2.3 * 2.35, C, (True, True, True), msg="left-handed Si")
    C1 = np.array([[5.71757, -1.81834e-15, 9.74255e-41], [-2.85879, 4.95156, 4.93924e-25], [4.56368e-40, 9.05692e-25, 9.05629]])
    X1 = np.array([[0.00533847, 2.85879, -1.42939, 1.42939, 0.0, 2.85879, -1.42939, 1.42939, -1.43e-6, 2.85878, -1.42939, 1.42939, -1.43e-6, 2.85878, -1.42939, 1.42939],
                   [-0.0, -0.0, 2.47578, 2.47578, 0.0, -0.0, 2.47578, 2.47578, 1.65052, 1.65052, 4.1263, 4.1263, 1.65052, 1.65052, 4.1263, 4.1263],
                   [0.00845581, 0.0, 0.0, 0.0, 4.52815, 4.52815, 4.52815, 4.52815, 2.26407, 2.26407, 2.26407, 2.26407, 6.79222, 6.79222, 6.79222, 6.79222]]).T
    check_case(nl, X1, 2.5 * 2.95, C1, (True, True, True), msg="hcp Ti 1")


def test_lazy_equals_materialised_and_lj(nl):
    # test/test_sortbased.jl:74-115 and BASELINE config 5 (Float32, rc = 6, LJ sink) at a small size
    for dtype, tol in ((np.float32, 1e-6), (np.float64, 1e-12)):
        X, C, L = U.rand_config(3000, seed=21, dtype=dtype)
        clist, pl, orc = check_case(nl, X, 6.0, C, (True, True, True), dtype, msg="lazy")
        assert int(nl.count_neighbours(clist).sum().item()) == nl.npairs(pl)
        assert nl.count_neighbours(clist, 17) == nl.nneigs(pl, 17) == nl.num_neighbours(pl, 17)
        e = float(nl.lj_energy(clist, 1.0, 3.4).item())
        lazy = O.sortbased(X, 6.0, C, (True, True, True), dtype=dtype, lazy=True)
        e_ref = O.lj_energy(lazy, 1.0, 3.4)
        assert abs(e - e_ref) <= tol * abs(e_ref), (e, e_ref)


def test_config1_10k(nl):
    # BASELINE config 1: 10k atoms, rho = 0.05, rc = 5, full PBC, Float64 (~262k pairs)
    X, C, L = U.rand_config(10000, seed=1)
    _, pl, orc = check_case(nl, X, 5.0, C, (True, True, True), msg="C1")
    assert 250000 < nl.npairs(pl) < 275000


def test_overflow_and_errors(nl):
    import torch
    X, C, L = U.rand_config(10, seed=3)
    with pytest.raises(nl.NlError):
        nl.build_cell_list(torch.from_numpy(X).cuda(), 1e-4, C * 1000, (True, True, True))  # prod(ncells) > typemax(Int32)
    with pytest.raises(ValueError):
        nl.build_cell_list(torch.zeros((4, 2), dtype=torch.float64, device="cuda"), 1.0, C, (True, True, True))


def test_nonuniform_density(nl):
    # a dense blob inside a dilute box: tiles over the blob exceed the shared-memory staging capacity and
    # must take the generic per-atom route inside the tiled kernel
    rng = np.random.Generator(np.random.PCG64(91))
    L = 60.0
    C = np.eye(3) * L
    bg = rng.random((3000, 3)) * L
    blob = 30.0 + rng.normal(size=(6000, 3)) * 1.5
    X = np.concatenate([bg, blob])
    for dtype in (np.float64, np.float32):
        check_case(nl, X.astype(dtype), 5.0, C.astype(dtype), (True, True, False), dtype, msg="blob")


@pytest.mark.parametrize("dtype,int_type", [(np.float64, np.int32), (np.float32, np.int64)])
def test_100k_mixed_pbc_triclinic(nl, dtype, int_type):
    # BASELINE config 3 scaled down: triclinic cell, pbc (T,T,F), rc = 5, 100k atoms
    s = (1e5 / 0.05 / 720.0) ** (1.0 / 3.0)
    cell = s * U.TRICLINIC
    X = U.rand_in_cell(100000, cell, seed=3, dtype=dtype)
    clist, pl, orc = check_case(nl, X, 5.0, cell.astype(dtype), (True, True, False), dtype, int_type, msg="C3/10")
    Xd = U.displace_by_lattice(X, cell, (True, True, False))
    check_case(nl, Xd, 5.0, cell.astype(dtype), (True, True, False), dtype, int_type, msg="C3/10 displaced")


def test_shard_mode_rows_and_global_indices(nl):
    """nl_fill_pairs_rows on the GPU: emulate the slab sharding of sharded.py for G = 3 ranks one after the
    other (the exchange itself is covered by the gloo tests) and merge the ranks' rows."""
    import torch
    from importlib import import_module
    sh = import_module("neighbourlists_jl_b200.sharded")
    for dtype, pbc, cutoff in ((np.float64, (True, True, True), 5.0), (np.float32, (True, False, True), 5.0)):
        cell = np.diag([25.0, 25.0, 120.0])
        N = 4000
        X = U.rand_in_cell(N, cell, seed=8, dtype=dtype)
        orc = O.sortbased(X, cutoff, cell, pbc, dtype=dtype)
        eng = sh.CudaEngine()
        cid = eng.cell_ids(torch.from_numpy(X).cuda(), cutoff, cell, pbc).cpu().numpy() - 1
        nc = nl.cellmath.geometry(cell, cutoff, pbc, dtype).ncells
        planes = (cid // (nc[0] * nc[1])) % nc[2]
        plan = sh.plan_slabs(np.bincount(planes, minlength=nc[2]), 3, 1, bool(pbc[2]), 2)
        merged = dict(i=[], j=[], S=[], R=[])
        counts = np.zeros(N, np.int64)
        for r in range(3):
            lo, hi = plan.bounds[r], plan.bounds[r + 1]
            owned = np.nonzero((planes >= lo) & (planes < hi))[0]
            below = np.nonzero(planes == (lo - 1) % nc[2])[0] if (pbc[2] or lo > 0) else np.zeros(0, np.int64)
            above = np.nonzero(planes == hi % nc[2])[0] if (pbc[2] or hi < nc[2]) else np.zeros(0, np.int64)
            local = np.concatenate([owned, below, above])
            res = eng.build(torch.from_numpy(X[local]).cuda(), len(owned), torch.from_numpy(local + 1).cuda(), cutoff, cell, pbc,
                            np.int32, True)
            f = res["first"].cpu().numpy()
            assert f.shape[0] == len(owned) + 1 and res["i"].shape[0] == f[-1] - 1
            counts[owned] = np.diff(f)
            assert np.array_equal(res["i"].cpu().numpy(), np.repeat(owned + 1, np.diff(f)))
            for k in merged:
                merged[k].append(res[k].cpu().numpy())
        merged = {k: np.concatenate(v) for k, v in merged.items()}
        merged["first"] = np.concatenate([[1], 1 + np.cumsum(counts)])
        order = np.argsort(merged["i"], kind="stable")
        merged = dict(first=merged["first"], i=merged["i"][order], j=merged["j"][order], S=merged["S"][order], R=merged["R"][order])
        U.assert_engine_matches_oracle(merged, orc, RTOL[np.dtype(dtype)], msg="shards")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_near_threshold_decisions(nl, dtype):
    """Rounding-sensitive decisions: 60 000 isolated dimers whose separation is within a few ulps of the
    cutoff (both sides), in random orientations, some straddling a periodic face.  Any fused multiply-add,
    reassociation or pre-filter slip in the kernels flips pairs here."""
    rng = np.random.Generator(np.random.PCG64(77))
    n_d, rc, L = 60000, 5.0, 4000.0
    g = int(np.ceil(n_d ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(g)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n_d]
    centres = (grid + 0.5) * (L / g)                                    # dimers ~100 A apart
    centres[: n_d // 8, 0] = rng.random(n_d // 8) * 2.0                 # some straddle the x = 0 face
    u = rng.normal(size=(n_d, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    eps = np.finfo(dtype).eps
    r = rc * (1.0 + rng.integers(-6, 7, n_d) * eps)
    X = np.concatenate([centres - 0.5 * r[:, None] * u, centres + 0.5 * r[:, None] * u]).astype(dtype)
    C = (np.eye(3) * L).astype(dtype)
    _, pl, orc = check_case(nl, X, rc, C, (True, True, True), dtype, msg="near threshold")
    # the designed dimer pairs (atom k with atom k + n_d) really straddle the threshold
    e = pl.cpu()
    dimer = int((np.abs(e["i"].astype(np.int64) - e["j"].astype(np.int64)) == n_d).sum()) / (2 * n_d)
    assert 0.2 < dimer < 0.8, dimer


@pytest.mark.parametrize("switch", ["NL_FILL_ROWS=1", "NL_FILL=park", "NL_FILL=2", "NL_FILL=legacy", "NL_BUILD=radix", "NL_COUNT=legacy",
                                    "NL_FILL_SZERO=0", "NL_FILL_PREFETCH=0"])
def test_kernel_variants_behind_environment_switches(switch):
    """The alternative kernels kept for A/B runs (environment switches: original-order fill k_fill_rows, the
    boundary-parking fill, the lean in-place k_fill_park, round 1's k_fill_mask and k_count_mask, the radix-sort build, ...) are
    selected by variables read once per process; each must produce the same lists.  Runs in a subprocess."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np, torch, sys
sys.path.insert(0, %r)
import neighbourlists_jl_b200 as nl
from oracle import nl_oracle as O
from tests import util as U
for dtype, pbc, cell, N, rc in ((np.float64, (True, True, True), None, 20000, 5.0), (np.float32, (True, False, True), U.TRICLINIC * 3, 4000, 3.0)):
    if cell is None:
        X, cell, _ = U.rand_config(N, seed=5, dtype=dtype)
    else:
        X = U.displace_by_lattice(U.rand_in_cell(N, cell, seed=6, dtype=dtype), cell, pbc)
    pl = nl.neighbour_list(torch.from_numpy(X).cuda(), rc, cell.astype(dtype), pbc, with_R=True)
    orc = O.sortbased(X, rc, cell.astype(dtype), pbc, dtype=dtype)
    U.assert_engine_matches_oracle(pl.cpu(), orc, 1e-12 if dtype == np.float64 else 1e-5, msg="variant")
    cl = nl.build_cell_list(torch.from_numpy(X).cuda(), rc, cell.astype(dtype), pbc)
    assert np.array_equal(cl.perm.cpu().numpy(), orc["perm"]) and np.array_equal(cl.cell_offsets.cpu().numpy(), orc["cell_offsets"])
print("OK")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    name, value = switch.split("=")
    env = dict(os.environ, **{name: value})
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_randomised_geometries(nl):
    """Fuzz: 40 seeded random problems -- random triclinic (sometimes left-handed) cells, random pbc, cutoffs from a
    fraction of the box to more than the box (wide stencils, self images), atoms scattered several boxes away on periodic
    axes and outside the box on open ones, both element types -- each compared field by field with the oracle."""
    rng = np.random.default_rng(20260101)
    for case in range(40):
        dtype = np.float64 if case % 2 == 0 else np.float32
        int_type = np.int32 if case % 3 else np.int64
        N = int(rng.integers(1, 600))
        A = np.diag(rng.uniform(4.0, 14.0, size=3)) + rng.uniform(-1.5, 1.5, size=(3, 3)) * (rng.random() < 0.7)
        if rng.random() < 0.3:
            A[2] = -A[2]  # left-handed
        if abs(np.linalg.det(A)) < 20.0:
            A = np.diag(np.diag(A))
        cell = A.astype(dtype)
        pbc = tuple(bool(b) for b in rng.integers(0, 2, size=3))
        f = rng.random((N, 3))
        f += rng.integers(-2, 3, size=(N, 3)) * (rng.random() < 0.5)          # several boxes away
        f += rng.normal(scale=0.2, size=(N, 3)) * (rng.random() < 0.5)        # slightly outside
        X = (f @ A).astype(dtype)
        lens = np.abs(np.linalg.det(A)) / np.array([np.linalg.norm(np.cross(A[(k + 1) % 3], A[(k + 2) % 3])) for k in range(3)])
        cutoff = float(rng.choice([0.25, 0.45, 0.9, 1.3]) * lens.min())
        check_case(nl, X, cutoff, cell, pbc, dtype=dtype, int_type=int_type, msg=f"fuzz case {case}: N={N} pbc={pbc} rc={cutoff:.3f}")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_plane_window_equals_full_launch(nl, dtype):
    """nl_count_pairs_window / nl_fill_pairs_window: a slab shard promises which z planes of cells can hold atoms and only
    those tile layers are launched.  The result must equal the full launch and the oracle (contiguous and wrapped windows)."""
    import torch
    rng = np.random.default_rng(17)
    L = 60.0
    C = (np.eye(3) * L).astype(dtype)
    pbc = (True, True, True)
    for zr in ([(0.30, 0.55)], [(0.0, 0.12), (0.9, 1.0)]):
        parts = []
        for lo, hi in zr:
            f = rng.random((6000, 3))
            f[:, 2] = lo + f[:, 2] * (hi - lo)
            parts.append(f)
        X = (np.concatenate(parts) * L).astype(dtype)
        Xd = torch.from_numpy(X).cuda()
        cl = nl.build_cell_list(Xd, 5.0, C, pbc)
        nz = int(cl.ncells[2])
        zc = ((cl.cell_id.long() - 1) // int(cl.ncells[0] * cl.ncells[1])).cpu().numpy()
        pa = np.zeros(nz, np.uint8)
        pa[np.unique(zc)] = 1
        assert 0 < pa.sum() < nz
        full = nl.materialize_pairlist(cl, with_R=True).cpu()
        win = nl.materialize_pairlist(cl, with_R=True, plane_active=pa).cpu()
        for k in ("first", "i", "j", "S", "R"):
            assert np.array_equal(full[k], win[k]), k
        U.assert_engine_matches_oracle(win, O.sortbased(X, 5.0, C, pbc, dtype=dtype), RTOL[np.dtype(dtype)], msg="window")
        # shard-style call: rows for the first half of the atoms only, global indices
        gmap = torch.arange(100, 100 + X.shape[0], device="cuda", dtype=torch.int32)
        a = nl.materialize_pairlist(cl, n_rows=3000, index_map=gmap).cpu()
        b = nl.materialize_pairlist(cl, n_rows=3000, index_map=gmap, plane_active=pa).cpu()
        for k in ("first", "i", "j", "S"):
            assert np.array_equal(a[k], b[k]), k
    with pytest.raises(ValueError):
        nl.materialize_pairlist(cl, plane_active=np.ones(3, np.uint8))
