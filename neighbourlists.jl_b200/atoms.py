"""AtomsBase-style adapter on the device (reference: ext/NeighbourListsAtomsBaseExt.jl).

The reference adapter always builds a host Vector of positions first (:54-56, :76-78) and derives the cell of an
isolated system from a host min/max over the atoms (:17-31).  Here the positions stay on the device and the
bounding box is one device reduction (nl_bounding_box); the rest is the unchanged hot path.

A "system" is anything with `positions` ((N,3) array-like or torch tensor), `cell` ((3,3), rows = cell vectors,
or None for an isolated system: AtomsBase's IsolatedCell) and `pbc` (3 bools); `unit` names the length unit of
positions and cell.  Cutoffs are floats in the system's unit or (value, unit) pairs; like the reference
(length_unit = unit(cutoff)) everything is converted to the CUTOFF's unit before the list is built.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib, api

# length units in Angstrom (Unitful's definitions)
_UNITS = {"Å": 1.0, "A": 1.0, "angstrom": 1.0, "nm": 10.0, "pm": 0.01, "bohr": 0.529177210903, "m": 1e10}


@dataclass
class System:
    """Minimal AtomsBase.AbstractSystem stand-in: positions, cell vectors (None = IsolatedCell), periodicity."""
    positions: object
    cell: Optional[object] = None
    pbc: tuple = (False, False, False)
    unit: str = "Å"

    def __len__(self):
        return int(self.positions.shape[0])


def isolated_system(positions, unit: str = "Å") -> System:
    """AtomsBase.isolated_system: open boundaries, no cell."""
    return System(positions=positions, cell=None, pbc=(False, False, False), unit=unit)


def periodic_system(positions, cell, pbc=(True, True, True), unit: str = "Å") -> System:
    return System(positions=positions, cell=cell, pbc=tuple(bool(b) for b in pbc), unit=unit)


def is_system(obj) -> bool:
    return hasattr(obj, "positions") and hasattr(obj, "pbc") and hasattr(obj, "cell")


def bounding_box(X: torch.Tensor) -> torch.Tensor:
    """(min x, min y, min z, max x, max y, max z) of device positions, as a device tensor (nl_bounding_box)."""
    X = api._as_device_positions(X)
    N = X.shape[0]
    if N == 0:
        raise ValueError("bounding box of an empty system")  # reference: maximum over an empty collection throws
    dev = X.device
    with torch.cuda.device(dev):
        out = torch.empty(6, dtype=X.dtype, device=dev)
        ws = torch.empty(_lib.NL_REDUCE_WS_BYTES, dtype=torch.uint8, device=dev)
        ft = _lib.NL_F64 if X.dtype == torch.float64 else _lib.NL_F32
        _lib.check(_lib.lib().nl_bounding_box(ft, api._ptr(X), N, api._ptr(out), api._ptr(ws), ws.numel(), api._stream(dev)))
    return out


def bounding_cell(X: torch.Tensor) -> np.ndarray:
    """Cell of an isolated 3-D system: diag(max - min + 1) in the positions' element type
    (_get_cell_matrix, ext/NeighbourListsAtomsBaseExt.jl:17-31)."""
    mm = bounding_box(X).cpu().numpy()
    T = mm.dtype.type
    d = (mm[3:] - mm[:3]) + T(1)
    return np.diag(d).astype(mm.dtype)


def _unit_scale(sys_unit: str, cutoff):
    """(cutoff value, factor that converts the system's lengths into the cutoff's unit)."""
    if isinstance(cutoff, tuple):
        value, cu = cutoff
    else:
        value, cu = cutoff, sys_unit
    try:
        return float(value), _UNITS[sys_unit] / _UNITS[cu]
    except KeyError as e:
        raise ValueError(f"unknown length unit {e}") from None


def _system_inputs(system, cutoff, device=None):
    pos = system.positions
    shape = tuple(pos.shape) if hasattr(pos, "shape") else np.asarray(pos).shape
    if len(shape) != 2 or shape[1] != 3:
        # the reference throws for 2-D isolated systems (ext/NeighbourListsAtomsBaseExt.jl:32-34, test_atoms_base.jl:135-144)
        D = shape[1] if len(shape) == 2 else "?"
        raise _lib.NlError(_lib.NL_ERR_BAD_ARG, f"NeighbourLists does not support {D}-dimensional AtomsBase systems yet.")
    rc, scale = _unit_scale(getattr(system, "unit", "Å"), cutoff)
    X = api._as_device_positions(pos, device)
    if scale != 1.0:
        X = X * scale
    if system.cell is None:
        cell = bounding_cell(X)
    else:
        cell = np.asarray(system.cell, dtype=np.float64) * scale
    return X, rc, cell, tuple(bool(b) for b in system.pbc)


def build_cell_list(system, cutoff, *, int_type=np.int32, device=None):
    """build_cell_list(system, cutoff) (ext/NeighbourListsAtomsBaseExt.jl:72-84)."""
    X, rc, cell, pbc = _system_inputs(system, cutoff, device)
    return api.build_cell_list(X, rc, cell, pbc, int_type=int_type)


def neighbour_list(system, cutoff, *, lazy: bool = False, int_type=np.int32, with_R: bool = False, device=None, half: bool = False):
    """neighbour_list(system, cutoff; lazy, int_type) (ext/NeighbourListsAtomsBaseExt.jl:116-138)."""
    clist = build_cell_list(system, cutoff, int_type=int_type, device=device)
    return clist if lazy else api.materialize_pairlist(clist, with_R=with_R, half=half)


def pair_list(system, cutoff, *, int_type=np.int32, device=None):
    """PairList(system, cutoff) (ext/NeighbourListsAtomsBaseExt.jl:50-63)."""
    return neighbour_list(system, cutoff, lazy=False, int_type=int_type, device=device)
