"""Host-side cell analysis in the position element type T.

Mirrors what the reference does on the host before any device work
(/root/reference/src/cell_list.jl:152-170 `analyze_cell`, :94-95 `lengths`,
/root/reference/src/gpu_kernels.jl:315-317 `nxyz`, `cutoff_sq`) with StaticArrays' unrolled 3x3
formulas (inv via the adjugate, det = x0 . (x1 x x2), left-associated sums, no FMA).  Every scalar
operation below is carried out on numpy scalars of dtype T, so Float32 inputs give Float32-rounded
results exactly as `SMat{Float32}` does.  The results are handed to libnlcuda.so through
`nl_params`; the library never re-derives them.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass

import numpy as np


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


@dataclass
class CellGeometry:
    dtype: np.dtype
    cell: np.ndarray       # (3,3) T, rows = lattice vectors
    inv_cell: np.ndarray   # (3,3) T
    lens: np.ndarray       # (3,) T, |perpendicular widths|
    ncells: np.ndarray     # (3,) int64
    nxyz: np.ndarray       # (3,) int64
    cutoff: float          # rounded to T
    pbc: tuple

    @property
    def ncells_total(self) -> int:
        return int(self.ncells[0]) * int(self.ncells[1]) * int(self.ncells[2])


def analyze_cell(cell, cutoff, dtype) -> tuple:
    """Returns (inv_cell, ncells, lens) like the reference's analyze_cell, plus nxyz."""
    T = np.dtype(dtype).type
    with np.errstate(all="ignore"):
        C = np.asarray(cell, dtype=T).reshape(3, 3)
        rc = T(cutoff)
        col = [tuple(T(C[r, c]) for r in range(3)) for c in range(3)]   # columns of C
        row = [tuple(T(C[r, c]) for c in range(3)) for r in range(3)]   # rows of C
        det = _dot(col[0], _cross(col[1], col[2]))
        if abs(float(det)) < 1e-12:
            warnings.warn("zero volume detected - proceed at your own risk")
        # inv(::SMatrix{3,3})
        x0, x1, x2 = col
        y0 = _cross(x1, x2)
        d = _dot(x0, y0)
        x0 = tuple(v / d for v in x0)
        y0 = tuple(v / d for v in y0)
        y1 = _cross(x2, x0)
        y2 = _cross(x0, x1)
        inv = np.empty((3, 3), dtype=T)
        inv[0, :] = y0
        inv[1, :] = y1
        inv[2, :] = y2
        # lengths(C) = det(C) ./ (|r2 x r3|, |r3 x r1|, |r1 x r2|)
        def nrm(v):
            return np.sqrt(_dot(v, v))
        lens_signed = (det / nrm(_cross(row[1], row[2])), det / nrm(_cross(row[2], row[0])), det / nrm(_cross(row[0], row[1])))
        lens = np.array([abs(v) for v in lens_signed], dtype=T)
        ncells = np.array([max(int(math.floor(float(T(lens[k]) / rc))), 1) for k in range(3)], dtype=np.int64)
        nxyz = np.array([int(math.ceil(float(rc * (T(int(ncells[k])) / T(lens[k]))))) for k in range(3)], dtype=np.int64)
    return inv, ncells, lens, nxyz


_GEO_CACHE: dict = {}


def geometry(cell, cutoff, pbc, dtype) -> CellGeometry:
    """analyze_cell + nxyz for (cell, cutoff, pbc) in `dtype`; memoised (the scalar numpy arithmetic costs ~0.1 ms,
    which matters for the 10k-atom case where the whole list takes 0.4 ms)."""
    dt = np.dtype(dtype)
    key = (np.asarray(cell, dtype=dt).tobytes(), float(dt.type(cutoff)), tuple(bool(b) for b in pbc), dt.str)
    hit = _GEO_CACHE.get(key)
    if hit is not None:
        return hit
    if len(_GEO_CACHE) > 256:
        _GEO_CACHE.clear()
    geo = _geometry_uncached(cell, cutoff, pbc, dt)
    _GEO_CACHE[key] = geo
    return geo


def _geometry_uncached(cell, cutoff, pbc, dtype) -> CellGeometry:
    dt = np.dtype(dtype)
    inv, ncells, lens, nxyz = analyze_cell(cell, cutoff, dt)
    return CellGeometry(dtype=dt, cell=np.asarray(cell, dtype=dt).reshape(3, 3).copy(), inv_cell=inv, lens=lens, ncells=ncells,
                        nxyz=nxyz, cutoff=float(dt.type(cutoff)), pbc=tuple(bool(b) for b in pbc))
