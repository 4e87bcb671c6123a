"""ctypes loader for the CPU oracle (oracle/nl_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Matrices cross this boundary as numpy (3,3) arrays C[r, c] whose ROWS are lattice vectors (the
reference's convention, /root/reference/src/cell_list.jl:627); they are flattened in Julia
column-major order for the C side.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_F = {np.dtype(np.float32): ("f32", C.c_float), np.dtype(np.float64): ("f64", C.c_double)}
_I = {np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}


def build(force: bool = False) -> str:
    """Compile libnl_oracle.so with the committed Makefile (building the checker is not using it)."""
    so = os.path.join(_HERE, "libnl_oracle.so")
    src = os.path.join(_HERE, "nl_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.nlo_lj_energy_f32_i32.restype = C.c_double
        _LIB.nlo_lj_energy_f32_i64.restype = C.c_double
        _LIB.nlo_lj_energy_f64_i32.restype = C.c_double
        _LIB.nlo_lj_energy_f64_i64.restype = C.c_double
        for n in ("nlo_legacy_f64", "nlo_legacy_f32", "nlo_brute_f64", "nlo_brute_f32"):
            getattr(_LIB, n).restype = C.c_void_p
        _LIB.nlo_pairset_npairs.restype = C.c_int64
        _LIB.nlo_pairset_nsites.restype = C.c_int64
    return _LIB


def max_threads() -> int:
    return int(lib().nlo_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _colmajor(M, dt):
    return np.ascontiguousarray(np.asarray(M, dtype=dt).reshape(3, 3).ravel(order="F"))


def analyze_cell(cell, cutoff, dtype=np.float64):
    """analyze_cell (src/cell_list.jl:152-170) + nxyz (src/gpu_kernels.jl:315-316), all in `dtype`."""
    dt = np.dtype(dtype)
    suf, cty = _F[dt]
    c = _colmajor(cell, dt)
    inv = np.zeros(9, dt)
    lens = np.zeros(3, dt)
    nc = np.zeros(3, np.int64)
    nxyz = np.zeros(3, np.int64)
    warn = getattr(lib(), f"nlo_analyze_cell_{suf}")(_p(c), cty(float(dt.type(cutoff))), _p(inv), _p(lens), _p(nc), _p(nxyz))
    return dict(cell=c, inv=inv, inv_mat=inv.reshape(3, 3, order="F"), lens=lens, ncells=nc, nxyz=nxyz, warn=bool(warn))


def sortbased(X, cutoff, cell, pbc, dtype=np.float64, int_type=np.int32, nthreads=None, want_R=True,
              lazy=False, geo=None):
    """build_cell_list + materialize_pairlist on the reference's CPU path.

    Returns a dict with the SortedCellList fields (Xs, perm, cell_id, cell_offsets, ncells, nxyz,
    inv) and, unless lazy, the PairList fields (first, i, j, S) plus R.  `geo` may carry
    host-computed (inv, ncells, nxyz) to use instead of the oracle's own analyze_cell.
    """
    dt, it = np.dtype(dtype), np.dtype(int_type)
    fs, cty = _F[dt]
    suf = f"{fs}_{_I[it]}"
    L = lib()
    X = np.ascontiguousarray(np.asarray(X, dtype=dt).reshape(-1, 3))
    N = X.shape[0]
    cut = dt.type(cutoff)
    g = analyze_cell(cell, cut, dt)
    if geo is not None:
        g = dict(g)
        g["inv"] = np.ascontiguousarray(np.asarray(geo["inv"], dt).ravel())
        g["ncells"] = np.asarray(geo["ncells"], np.int64).copy()
        g["nxyz"] = np.asarray(geo["nxyz"], np.int64).copy()
    nct = int(np.prod(g["ncells"].astype(object)))
    pb = np.ascontiguousarray(np.asarray(pbc, dtype=bool).astype(np.uint8))
    nthreads = nthreads or max_threads()
    Xs = np.zeros((N, 3), dt)
    perm = np.zeros(N, it)
    cell_id = np.zeros(N, it)
    cell_offsets = np.zeros(nct + 1, it)
    getattr(L, f"nlo_build_cells_{suf}")(_p(X), C.c_int64(N), _p(g["cell"]), _p(g["inv"]), cty(float(cut)),
                                         _p(g["ncells"]), _p(pb), _p(Xs), _p(perm), _p(cell_id), _p(cell_offsets))
    out = dict(X=X, Xs=Xs, perm=perm, cell_id=cell_id, cell_offsets=cell_offsets, ncells=g["ncells"],
               nxyz=g["nxyz"], inv=g["inv"], cell=g["cell"], ncells_total=nct, cutoff=cut, pbc=pb, N=N)
    if lazy:
        return out
    geo_args = (_p(g["cell"]), _p(g["inv"]), cty(float(cut)), _p(g["ncells"]), _p(g["nxyz"]), _p(pb), C.c_int(nthreads))
    counts = np.zeros(N, it)
    getattr(L, f"nlo_count_pairs_{suf}")(_p(X), C.c_int64(N), _p(perm), _p(cell_offsets), *geo_args, _p(counts))
    first = np.zeros(N + 1, it)
    getattr(L, f"nlo_pair_offsets_{suf}")(_p(counts), C.c_int64(N), _p(first))
    P = int(first[-1]) - 1
    i = np.zeros(P, it)
    j = np.zeros(P, it)
    S = np.zeros((P, 3), it)
    R = np.zeros((P, 3), dt) if want_R else None
    if P > 0:
        getattr(L, f"nlo_fill_pairs_{suf}")(_p(X), C.c_int64(N), _p(perm), _p(cell_offsets), *geo_args,
                                            _p(first), _p(i), _p(j), _p(S), _p(R))
    out.update(first=first, i=i, j=j, S=S, R=R, counts=counts, npairs=P)
    return out


def lj_energy(cl, eps, sigma, nthreads=None):
    """LJ sink over for_each_neighbour on a lazy sortbased() result (ordered-pair sum, f64 accumulate)."""
    dt, it = cl["Xs"].dtype, cl["perm"].dtype
    fs, cty = _F[dt]
    fn = getattr(lib(), f"nlo_lj_energy_{fs}_{_I[it]}")
    return float(fn(_p(cl["X"]), C.c_int64(cl["N"]), _p(cl["perm"]), _p(cl["cell_offsets"]), _p(cl["cell"]),
                    _p(cl["inv"]), cty(float(cl["cutoff"])), _p(cl["ncells"]), _p(cl["nxyz"]), _p(cl["pbc"]),
                    C.c_int(nthreads or max_threads()), C.c_double(eps), C.c_double(sigma)))


def _pairset(h, want_R):
    L = lib()
    h = C.c_void_p(h)
    P, N = int(L.nlo_pairset_npairs(h)), int(L.nlo_pairset_nsites(h))
    i = np.zeros(P, np.int64)
    j = np.zeros(P, np.int64)
    S = np.zeros((P, 3), np.int64)
    first = np.zeros(N + 1, np.int64)
    R = np.zeros((P, 3), np.float64) if want_R else None
    X = np.zeros((N, 3), np.float64) if not want_R else None
    Cm = np.zeros(9, np.float64) if not want_R else None
    L.nlo_pairset_copy(h, _p(i), _p(j), _p(S), _p(first), _p(R), _p(X), _p(Cm))
    L.nlo_pairset_free(h)
    return dict(i=i, j=j, S=S, first=first, R=R, X=X, C=None if Cm is None else Cm.reshape(3, 3, order="F"), npairs=P)


def legacy(X, cutoff, cell, pbc, dtype=np.float64, fixcell=True):
    """Legacy linked-list PairList (src/cell_list.jl:11-20, 369-384): second oracle; rows sorted by j."""
    dt = np.dtype(dtype)
    fs, cty = _F[dt]
    X = np.ascontiguousarray(np.asarray(X, dtype=dt).reshape(-1, 3))
    pb = np.ascontiguousarray(np.asarray(pbc, dtype=bool).astype(np.uint8))
    h = getattr(lib(), f"nlo_legacy_{fs}")(_p(X), C.c_int64(X.shape[0]), _p(_colmajor(cell, dt)), _p(pb),
                                          cty(float(dt.type(cutoff))), C.c_int(int(fixcell)))
    return _pairset(h, want_R=False)


def brute(X, cutoff, cell, pbc, dtype=np.float64):
    """O(N^2 * images) enumeration of {(i,j,S): |X[j]-X[i]+C'S| < rc} \\ {(i,i,0)}: third oracle."""
    dt = np.dtype(dtype)
    fs, cty = _F[dt]
    X = np.ascontiguousarray(np.asarray(X, dtype=dt).reshape(-1, 3))
    pb = np.ascontiguousarray(np.asarray(pbc, dtype=bool).astype(np.uint8))
    h = getattr(lib(), f"nlo_brute_{fs}")(_p(X), C.c_int64(X.shape[0]), _p(_colmajor(cell, dt)), _p(pb),
                                         cty(float(dt.type(cutoff))))
    return _pairset(h, want_R=True)


def canonical(i, j, S, R=None):
    """The reference's canonical ordering: sort by (i, j, S1, S2, S3) (test/test_utils.jl:68-70)."""
    i = np.asarray(i).astype(np.int64)
    j = np.asarray(j).astype(np.int64)
    S = np.asarray(S).astype(np.int64).reshape(-1, 3)
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))
    out = [i[order], j[order], S[order]]
    if R is not None:
        out.append(np.asarray(R).reshape(-1, 3)[order])
    return tuple(out)


# ---------------------------------------------------------------------------------------------------------
# numpy restatements of the reference's accessors / adapter pieces.  numpy evaluates every
# elementwise operation separately in the array dtype, i.e. unfused and in the written association.

def pairs_R(X, i, j, S, cell, dtype=np.float64):
    """_getR for all pairs: R = (X[j] - X[i]) + C' * S (/root/reference/src/cell_list.jl:525-531)."""
    T = np.dtype(dtype).type
    X = np.asarray(X, dtype=T)
    C_ = np.asarray(cell, dtype=T)
    i0, j0 = np.asarray(i, dtype=np.int64) - 1, np.asarray(j, dtype=np.int64) - 1
    Sf = np.asarray(S).astype(T)
    d = X[j0] - X[i0]
    cs = np.stack([(C_[0, k] * Sf[:, 0] + C_[1, k] * Sf[:, 1]) + C_[2, k] * Sf[:, 2] for k in range(3)], axis=1)
    return (d + cs).astype(T)


def maxneigs(first):
    """maximum(nneigs(nlist, n) for n = 1:nsites) (/root/reference/src/cell_list.jl:513,523)."""
    f = np.asarray(first, dtype=np.int64)
    if f.shape[0] < 2:
        raise ValueError("maximum over an empty collection")
    return int((f[1:] - f[:-1]).max())


def rows_padded(X, first, j, S, cell, rows, width, dtype=np.float64):
    """neigss(nlist, i) (/root/reference/src/cell_list.jl:583-597) for every i in rows, padded to `width`."""
    T = np.dtype(dtype).type
    f = np.asarray(first, dtype=np.int64)
    rows = np.asarray(rows, dtype=np.int64)
    n = np.zeros(len(rows), dtype=np.int64)
    jo = np.zeros((len(rows), width), dtype=np.asarray(j).dtype)
    So = np.zeros((len(rows), width, 3), dtype=np.asarray(S).dtype)
    Ro = np.zeros((len(rows), width, 3), dtype=T)
    for s, r in enumerate(rows):
        lo, hi = f[r - 1] - 1, f[r] - 1
        n[s] = hi - lo
        m = min(hi - lo, width)
        jo[s, :m] = j[lo:lo + m]
        So[s, :m] = S[lo:lo + m]
        Ro[s, :m] = pairs_R(X, np.full(m, r), j[lo:lo + m], S[lo:lo + m], cell, dtype)
    return n, jo, So, Ro


def bounding_cell(X, dtype=np.float64):
    """IsolatedCell branch of _get_cell_matrix (/root/reference/ext/NeighbourListsAtomsBaseExt.jl:17-31):
    diag(max - min + 1) per axis, in the positions' element type."""
    T = np.dtype(dtype).type
    X = np.asarray(X, dtype=T)
    return np.diag((X.max(axis=0) - X.min(axis=0)) + T(1)).astype(T)


def max_displacement2(X, X_ref, dtype=np.float64):
    T = np.dtype(dtype).type
    d = np.asarray(X, dtype=T) - np.asarray(X_ref, dtype=T)
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return T(d2.max()) if d2.size else T(0)


def lj_forces(d, eps, sigma, N):
    """Per-atom Lennard-Jones forces and energies from an oracle pair list d (with R), in float64: a visit of the
    ordered pair (i, j), R = x_j - x_i + C'S, adds 24 eps (2 s^12 - s^6) / r^2 * R to the force on j and
    phi = 4 eps (s^12 - s^6) to j's energy (s = sigma / r).  Returns (F (N,3), e (N,)); F = -dE/dX for E = e.sum() / 2."""
    R = np.asarray(d["R"], dtype=np.float64)
    j0 = np.asarray(d["j"], dtype=np.int64) - 1
    r2 = (R * R).sum(axis=1)
    s2 = sigma * sigma / r2
    s6 = s2 * s2 * s2
    phi = 4.0 * eps * (s6 * s6 - s6)
    gg = 24.0 * eps * (2.0 * s6 * s6 - s6) / r2
    F = np.zeros((N, 3))
    e = np.zeros(N)
    np.add.at(F, j0, gg[:, None] * R)
    np.add.at(e, j0, phi)
    return F, e


def mirror_canonical(i, j, S):
    """Canonical representative of each pair's mirror couple {(i, j, S), (j, i, -S)}: the lexicographically smaller
    (i, j, S1, S2, S3) tuple.  Returns the sorted (P, 5) int64 array (duplicates kept)."""
    a = np.concatenate([np.asarray(i, np.int64)[:, None], np.asarray(j, np.int64)[:, None], np.asarray(S, np.int64).reshape(-1, 3)], axis=1)
    b = np.concatenate([a[:, 1:2], a[:, 0:1], -a[:, 2:]], axis=1)
    swap = np.zeros(len(a), dtype=bool)
    undecided = np.ones(len(a), dtype=bool)
    for k in range(5):
        lt, gt = undecided & (b[:, k] < a[:, k]), undecided & (b[:, k] > a[:, k])
        swap |= lt
        undecided &= ~(lt | gt)
    c = np.where(swap[:, None], b, a)
    return c[np.lexsort(c.T[::-1])]


def half_list(d):
    """The half list as a SET: one canonical representative per mirror couple of the full oracle list d."""
    c = mirror_canonical(d["i"], d["j"], d["S"])
    assert len(c) % 2 == 0 and np.array_equal(c[0::2], c[1::2]), "the full list consists of mirror couples"
    return c[0::2]
