"""Skin (Verlet) list: reuse of a neighbour list across MD steps.

The reference rebuilds on every call (src/cell_list.jl:906-916).  A list built with cutoff + skin contains every
pair within `cutoff` for as long as no atom has moved further than skin / 2 from where it was at build time; until
then only R has to be refreshed for the unchanged (i, j, S) topology -- S stays valid because the engine, like the
reference, never wraps positions (src/cell_list.jl:661-664).  Both steps are single device passes:
nl_max_displacement2 (one read of the positions) and nl_pairs_R (_getR for all pairs).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, api


def max_displacement2(X: torch.Tensor, X_ref: torch.Tensor) -> torch.Tensor:
    """max_n |X[n] - X_ref[n]|^2 as a device scalar tensor (nl_max_displacement2)."""
    if X.shape != X_ref.shape or X.dtype != X_ref.dtype or X.device != X_ref.device:
        raise ValueError("X and X_ref must match in shape, dtype and device")
    X = api._as_device_positions(X)
    X_ref = X_ref.contiguous()
    dev = X.device
    with torch.cuda.device(dev):
        out = torch.empty(1, dtype=X.dtype, device=dev)
        ws = torch.empty(_lib.NL_REDUCE_WS_BYTES, dtype=torch.uint8, device=dev)
        ft = _lib.NL_F64 if X.dtype == torch.float64 else _lib.NL_F32
        _lib.check(_lib.lib().nl_max_displacement2(ft, api._ptr(X), api._ptr(X_ref), X.shape[0], api._ptr(out), api._ptr(ws),
                                                   ws.numel(), api._stream(dev)))
    return out


class SkinList:
    """PairList for `cutoff + skin`, kept across position updates.

    update(X) returns True when the list had to be rebuilt.  `nlist.R` always holds
    R = (X[j] - X[i]) + C' S for the CURRENT positions; pairs with |R| >= cutoff are part of the list (that is
    the point of the skin) and are filtered by the consumer."""

    def __init__(self, X, cutoff: float, skin: float, cell, pbc, *, int_type=np.int32, device=None):
        if not skin > 0:
            raise ValueError("skin must be positive")
        self.cutoff, self.skin = float(cutoff), float(skin)
        self.cell, self.pbc, self.int_type = cell, pbc, int_type
        self.builds = 0
        self._build(api._as_device_positions(X, device))

    def _build(self, X: torch.Tensor):
        self.X_ref = X.clone()  # the caller may update its tensor in place
        self.nlist = api.neighbour_list(self.X_ref, self.cutoff + self.skin, self.cell, self.pbc, int_type=self.int_type, with_R=True)
        self.builds += 1

    def update(self, X: torch.Tensor) -> bool:
        X = api._as_device_positions(X, self.X_ref.device)
        if X.shape != self.X_ref.shape or X.dtype != self.X_ref.dtype:
            raise ValueError("positions changed shape or dtype")
        d2 = float(max_displacement2(X, self.X_ref).item())
        half = 0.5 * self.skin
        if not d2 < half * half:  # also rebuilds on NaN
            self._build(X)
            return True
        self.nlist.R = api.pairs_R(self.nlist, X=X)
        return False
