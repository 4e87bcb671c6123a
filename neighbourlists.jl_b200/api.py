"""Host-side mirror of the reference's operator interface for the sort-based neighbour-list path.

Same names, argument meaning and error behaviour as NeighbourLists.jl
(/root/reference/src/cell_list.jl:632-642 build_cell_list, :897-916 neighbour_list,
/root/reference/src/gpu_kernels.jl:299-364 materialize_pairlist, /root/reference/src/types.jl:34-82
PairList / SortedCellList), with torch CUDA tensors standing in for CuArrays.  Every stage runs in
libnlcuda.so through the C ABI (include/nlcuda.h); there is no CPU fallback.

Like the reference's arrays, every integer array is 1-BASED, and atom arguments `i` are 1-based.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib
from .cellmath import CellGeometry, geometry

_T2N = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}
_N2T = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
        np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


def _int_dtype(int_type) -> torch.dtype:
    if isinstance(int_type, torch.dtype):
        if int_type not in (torch.int32, torch.int64):
            raise TypeError("int_type must be int32 or int64")
        return int_type
    return _N2T[np.dtype(int_type)]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _as_device_positions(X, device=None) -> torch.Tensor:
    """Accepts an (N,3) torch tensor or array-like.  Host inputs are uploaded (pinned, async): the
    reference would run its CPU path for them, this engine has none."""
    if not torch.cuda.is_available():
        raise _lib.NlError(-101, "no CUDA device: this engine has no CPU path")
    if isinstance(X, torch.Tensor):
        t = X
    else:
        a = np.asarray(X)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype not in (torch.float32, torch.float64):
        raise TypeError("positions must be float32 or float64")
    if t.dim() != 2 or t.shape[1] != 3:
        # the reference only accepts 3-vectors (AbstractVector{<:SVec}); 2-D systems raise (test_atoms_base.jl:135-144)
        raise ValueError("positions must be an (N, 3) array")
    if not t.is_cuda:
        dev = torch.device(device if device is not None else "cuda")
        if not t.is_pinned():
            t = t.pin_memory()
        t = t.to(dev, non_blocking=True)
    return t.contiguous()


@dataclass
class SortedCellList:
    """Field-for-field twin of the reference's SortedCellList (src/types.jl:70-82)."""
    X: torch.Tensor             # positions sorted by cell, (N,3) T
    X_orig: torch.Tensor        # the caller's positions (aliased, never modified)
    perm: torch.Tensor          # sorted slot -> original index (1-based), (N,) TI
    cell_id: torch.Tensor       # linear cell of each sorted slot (1-based), (N,) TI
    cell_offsets: torch.Tensor  # (ncells_total+1,) TI, 1-based
    cell: np.ndarray            # (3,3) T, rows = lattice vectors
    inv_cell: np.ndarray
    pbc: tuple
    cutoff: float
    ncells: np.ndarray
    ncells_total: int
    # host-side extras (not part of the reference struct)
    geo: CellGeometry = field(repr=False, default=None)
    params: object = field(repr=False, default=None)
    _ws: Optional[torch.Tensor] = field(repr=False, default=None)
    _pl: object = field(repr=False, default=None)
    _counts: Optional[torch.Tensor] = field(repr=False, default=None)


@dataclass
class PairList:
    """Twin of the reference's PairList (src/types.jl:34-43); R is an optional extra (not a reference field)."""
    X: torch.Tensor
    C: np.ndarray
    cutoff: float
    i: torch.Tensor
    j: torch.Tensor
    S: torch.Tensor       # (P,3) TI
    first: torch.Tensor   # (N+1,) TI, 1-based
    R: Optional[torch.Tensor] = None
    params: object = field(repr=False, default=None)  # host-side extra: nl_params of the list (types, cell)
    half: bool = False    # half list: one pair per mirror couple (i, j, S) / (j, i, -S)  (NL_FLAG_HALF)

    def cpu(self):
        """Device -> host copies of every array (numpy), for inspection and tests."""
        torch.cuda.current_stream(self.i.device).synchronize()
        return dict(i=self.i.cpu().numpy(), j=self.j.cpu().numpy(), S=self.S.cpu().numpy(), first=self.first.cpu().numpy(),
                    R=None if self.R is None else self.R.cpu().numpy(), X=self.X.cpu().numpy())


# ------------------------------------------------------------------ construction
def _pairs_workspace(clist: SortedCellList) -> torch.Tensor:
    N = clist.X.shape[0]
    need = _lib.lib().nl_workspace_bytes(clist.params, N, _lib.NL_STAGE_PAIRS)
    if clist._ws is None or clist._ws.numel() < need:
        clist._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=clist.X.device)
    return clist._ws


def build_cell_list(X, cutoff, cell=None, pbc=None, *, int_type=np.int32, device=None) -> SortedCellList:
    """build_cell_list(X, cutoff, cell, pbc; int_type) -> SortedCellList  (src/cell_list.jl:632-679);
    build_cell_list(system, cutoff) for AtomsBase-style systems (atoms.py)."""
    if cell is None and pbc is None and hasattr(X, "positions"):
        from . import atoms
        return atoms.build_cell_list(X, cutoff, int_type=int_type, device=device)
    Xd = _as_device_positions(X, device)
    it = _int_dtype(int_type)
    fdt = np.dtype(_T2N[Xd.dtype])
    geo = geometry(cell, cutoff, pbc, fdt)
    nct = geo.ncells_total
    tmax = 2**31 - 1 if it == torch.int32 else 2**63 - 1
    if nct > tmax:
        # src/cell_list.jl:655-659
        raise _lib.NlError(_lib.NL_ERR_UNSUPPORTED,
                           "Ratio of simulation cell size to cutoff is very large. Use a larger integer type (e.g. Int64), "
                           "larger cutoff, or smaller simulation cell.")
    params = _lib.make_params(geo, fdt, _T2N[it])
    N = Xd.shape[0]
    dev = Xd.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        Xs = torch.empty_like(Xd)
        perm = torch.empty(N, dtype=it, device=dev)
        cell_id = torch.empty(N, dtype=it, device=dev)
        cell_offsets = torch.empty(nct + 1, dtype=it, device=dev)
        need = L.nl_workspace_bytes(params, N, _lib.NL_STAGE_BUILD)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        _lib.check(L.nl_build_cells(params, _ptr(Xd), N, _ptr(Xs), _ptr(perm), _ptr(cell_id), _ptr(cell_offsets), _ptr(ws),
                                    ws.numel(), _stream(dev)))
    return SortedCellList(X=Xs, X_orig=Xd, perm=perm, cell_id=cell_id, cell_offsets=cell_offsets, cell=geo.cell,
                          inv_cell=geo.inv_cell, pbc=geo.pbc, cutoff=geo.cutoff, ncells=geo.ncells, ncells_total=nct,
                          geo=geo, params=params)


def materialize_pairlist(clist: SortedCellList, *, with_R: bool = False, timers: Optional[dict] = None,
                         n_rows: Optional[int] = None, index_map: Optional[torch.Tensor] = None, half: bool = False,
                         plane_active: Optional[np.ndarray] = None, host_out=None, host_threads: int = 0):
    """materialize_pairlist(clist) -> PairList  (src/gpu_kernels.jl:299-364).  with_R additionally
    stores R = X[j] - X[i] + C' S per pair (what the reference recomputes in _getR).

    Shard mode (sharded.py): with n_rows, only the first n_rows atoms get rows (`first` has n_rows+1
    entries) and i/j are written through index_map (global 1-based indices, one per local atom).

    plane_active (slab shards cut along z): uint8 per z plane of cells, nonzero where the local set may have atoms; the
    caller promises all other planes are empty and only those tile layers are launched (nl_*_window).

    half=True stores one pair of every mirror couple (i, j, S) / (j, i, -S) (NL_FLAG_HALF, include/nlcuda.h): half the
    pairs, half the output traffic; which of the two is kept is unspecified.

    host_out (HostPairBuffers, or True for the module-level set): the list goes straight to HOST memory and a HostPairList is
    returned (nl_pairs_to_host_begin / _finish: `first` is copied and i rebuilt by host threads while the fill pass runs)."""
    import ctypes as C
    params = clist.params
    if half:
        params = type(clist.params).from_buffer_copy(clist.params)
        params.reserved[0] = _lib.NL_FLAG_HALF
    N = clist.X.shape[0]
    dev = clist.X.device
    it = clist.perm.dtype
    L = _lib.lib()
    with torch.cuda.device(dev):
        first = torch.empty(N + 1, dtype=it, device=dev)
        ws = _pairs_workspace(clist)
        total = C.c_int64(0)
        if timers is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        pa = None
        if plane_active is not None:
            pa = np.ascontiguousarray(plane_active, dtype=np.uint8)
            if pa.shape != (int(clist.ncells[2]),):
                raise ValueError("plane_active must have one entry per z plane of cells")
        pa_ptr = None if pa is None else pa.ctypes.data
        if pa is None:
            _lib.check(L.nl_count_pairs(params, _ptr(clist.X), N, _ptr(clist.perm), _ptr(clist.cell_offsets), _ptr(first),
                                        C.byref(total), _ptr(ws), ws.numel(), _stream(dev)))
        else:
            _lib.check(L.nl_count_pairs_window(params, _ptr(clist.X), N, _ptr(clist.perm), _ptr(clist.cell_offsets), _ptr(first),
                                               C.byref(total), pa_ptr, _ptr(ws), ws.numel(), _stream(dev)))
        P = int(total.value)
        if n_rows is not None and not 0 <= n_rows <= N:
            raise ValueError("n_rows out of range")
        job = None
        if host_out is not None and host_out is not False:
            if n_rows is not None or index_map is not None or with_R:
                raise ValueError("host_out is for whole lists without R")
            host_out = _host_buffers_for(P, N, it, dev, None if host_out is True else host_out)
            job = C.c_void_p()
            _lib.check(L.nl_pairs_to_host_begin(params, _ptr(first), N, P, host_out.first.data_ptr(), host_out.i.data_ptr(), int(host_threads),
                                                _stream(dev), C.byref(job)))
        # shard mode: the owned rows hold first[n_rows] - 1 <= total pairs.  The arrays are allocated for `total` and trimmed
        # AFTER the fill has been enqueued, so that this second host read does not leave the GPU idle between the two passes.
        if timers is not None:
            ev[1].record()
        i = torch.empty(P, dtype=it, device=dev)
        j = torch.empty(P, dtype=it, device=dev)
        S = torch.empty((P, 3), dtype=it, device=dev)
        R = torch.empty((P, 3), dtype=clist.X.dtype, device=dev) if with_R else None
        if timers is not None:
            ev[2].record()
        try:
            if P > 0 and pa is not None:
                if index_map is not None:
                    index_map = index_map.to(device=dev, dtype=it).contiguous()
                    assert index_map.shape[0] == N
                _lib.check(L.nl_fill_pairs_window(params, _ptr(clist.X), N, _ptr(clist.perm), _ptr(clist.cell_offsets), _ptr(first),
                                                  N if n_rows is None else n_rows, _ptr(index_map), pa_ptr, _ptr(i), _ptr(j), _ptr(S), _ptr(R),
                                                  _ptr(ws), ws.numel(), _stream(dev)))
            elif P > 0 and n_rows is None and index_map is None:
                _lib.check(L.nl_fill_pairs(params, _ptr(clist.X), N, _ptr(clist.perm), _ptr(clist.cell_offsets), _ptr(first),
                                           _ptr(i), _ptr(j), _ptr(S), _ptr(R), _ptr(ws), ws.numel(), _stream(dev)))
            elif P > 0:
                if index_map is not None:
                    index_map = index_map.to(device=dev, dtype=it).contiguous()
                    assert index_map.shape[0] == N
                _lib.check(L.nl_fill_pairs_rows(params, _ptr(clist.X), N, _ptr(clist.perm), _ptr(clist.cell_offsets), _ptr(first),
                                                N if n_rows is None else n_rows, _ptr(index_map), _ptr(i), _ptr(j), _ptr(S), _ptr(R),
                                                _ptr(ws), ws.numel(), _stream(dev)))
        except BaseException:
            if job is not None:   # a job must be finished exactly once: this releases its host threads
                L.nl_pairs_to_host_finish(job, None, None, None, None, None, None, 0, _stream(dev))
            raise
        if timers is not None:
            ev[3].record()
            timers.setdefault("events", []).append(ev)
        if job is not None:
            _lib.check(L.nl_pairs_to_host_finish(job, _ptr(j), _ptr(S), host_out.j.data_ptr(), host_out.S.data_ptr(), host_out.dev_scratch.data_ptr(),
                                                 host_out.host_scratch.data_ptr(), host_out.dev_scratch.numel(), _stream(dev)))
            return HostPairList(X=clist.X_orig, C=clist.cell, cutoff=clist.cutoff, i=host_out.i[:P].numpy(), j=host_out.j[:P].numpy(),
                                S=host_out.S[:P].numpy(), first=host_out.first[:N + 1].numpy())
        if n_rows is not None:
            Pr = int(first[n_rows].item()) - 1
            first = first[:n_rows + 1]
            i, j, S = i[:Pr], j[:Pr], S[:Pr]
            R = None if R is None else R[:Pr]
    return PairList(X=clist.X_orig, C=clist.cell, cutoff=clist.cutoff, i=i, j=j, S=S, first=first, R=R, params=clist.params, half=half)


def cell_ids(X, cutoff, cell, pbc, *, int_type=np.int32, device=None) -> torch.Tensor:
    """1-based linear cell id of every atom in the caller's order (_compute_cell_ids, src/gpu_kernels.jl:244-255)."""
    Xd = _as_device_positions(X, device)
    it = _int_dtype(int_type)
    fdt = np.dtype(_T2N[Xd.dtype])
    geo = geometry(cell, cutoff, pbc, fdt)
    params = _lib.make_params(geo, fdt, _T2N[it])
    out = torch.empty(Xd.shape[0], dtype=it, device=Xd.device)
    with torch.cuda.device(Xd.device):
        _lib.check(_lib.lib().nl_cell_ids(params, _ptr(Xd), Xd.shape[0], _ptr(out), _stream(Xd.device)))
    return out


def neighbour_list(X, cutoff, cell=None, pbc=None, *, lazy: bool = False, int_type=np.int32, with_R: bool = False, device=None,
                   half: bool = False, host_out=None, host_threads: int = 0):
    """neighbour_list(X, cutoff, cell, pbc; lazy, int_type)  (src/cell_list.jl:897-916);
    neighbour_list(system, cutoff; lazy, int_type) for AtomsBase-style systems (atoms.py)."""
    if cell is None and pbc is None and hasattr(X, "positions"):
        from . import atoms
        return atoms.neighbour_list(X, cutoff, lazy=lazy, int_type=int_type, with_R=with_R, device=device, half=half)
    clist = build_cell_list(X, cutoff, cell, pbc, int_type=int_type, device=device)
    if lazy:
        return clist
    return materialize_pairlist(clist, with_R=with_R, half=half, host_out=host_out, host_threads=host_threads)



# ------------------------------------------------------------------ Array(PairList): the list in host memory
@dataclass
class HostPairList:
    """PairList with every array in HOST memory (numpy views of the pinned buffers they were copied into): what
    `Array(...)` of each field gives in the reference (test/test_utils.jl:127-131).  1-based like the device list."""
    X: object
    C: np.ndarray
    cutoff: float
    i: np.ndarray
    j: np.ndarray
    S: np.ndarray
    first: np.ndarray


class HostPairBuffers:
    """Pinned host buffers (and the device / host scratch of the transfer) for `to_host`, reusable across lists:
    allocating 5 GB of pinned memory costs more than building the list."""

    def __init__(self, pair_capacity: int, n_rows: int, int_type=np.int32, device=None):
        it = _int_dtype(int_type)
        self.int_dtype = it
        self.pair_capacity = int(pair_capacity)
        self.row_capacity = int(n_rows)
        self.device = torch.device(device if device is not None else "cuda")
        nb = _lib.lib().nl_to_host_scratch_bytes(self.pair_capacity)
        self.first = torch.empty(self.row_capacity + 1, dtype=it).pin_memory()
        self.rows = torch.empty(max(self.row_capacity, 1), dtype=it).pin_memory()   # a shard's row -> global index map
        self.i = torch.empty(max(self.pair_capacity, 1), dtype=it).pin_memory()
        self.j = torch.empty(max(self.pair_capacity, 1), dtype=it).pin_memory()
        self.S = torch.empty((max(self.pair_capacity, 1), 3), dtype=it).pin_memory()
        self.host_scratch = torch.empty(nb, dtype=torch.uint8).pin_memory()
        self.dev_scratch = torch.empty(nb, dtype=torch.uint8, device=self.device)

    def fits(self, P: int, n_rows: int, it) -> bool:
        return P <= self.pair_capacity and n_rows <= self.row_capacity and it == self.int_dtype


_host_buffers: Optional[HostPairBuffers] = None


def _host_buffers_for(P: int, n_rows: int, it, dev, out: Optional[HostPairBuffers]) -> HostPairBuffers:
    global _host_buffers
    if out is None:
        if _host_buffers is None or not _host_buffers.fits(P, n_rows, it) or _host_buffers.device != dev:
            _host_buffers = None
            _host_buffers = HostPairBuffers(int(P * 1.05) + 1024, n_rows, it, dev)
        return _host_buffers
    if not out.fits(P, n_rows, it):
        raise ValueError("host buffers too small for this list")
    return out


def to_host(nlist, out: Optional[HostPairBuffers] = None, nthreads: int = 0, rebuild_i: Optional[bool] = None,
            i_copy_fraction: float = 0.0) -> HostPairList:
    """The whole list into host memory through nl_pairs_to_host (include/nlcuda.h): `first`, `j` and one byte per pair for S
    cross the bus; i and S are rebuilt by host threads of the library while the copies run.  Returns when every array is
    complete.  `out`: buffers to reuse (default: a module-level set grown on demand).  i is rebuilt from `first` -- for a shard
    list (sharded.ShardedPairList) through its row -> global index map, which is copied instead of i; rebuild_i=False copies i;
    otherwise the last i_copy_fraction of i is copied and the rest rebuilt (0 is fastest where measured: the host's memory
    bandwidth, which DMA writes and host stores share, is the bound, not the host threads' instruction rate)."""
    P = int(nlist.i.shape[0])
    n_rows = int(nlist.first.shape[0]) - 1
    it = nlist.i.dtype
    dev = nlist.i.device
    out = _host_buffers_for(P, n_rows, it, dev, out)
    whole = isinstance(nlist, PairList) and n_rows == int(nlist.X.shape[0])   # a whole list: i[p] is the row of p
    row_index = getattr(nlist, "owned_index", None)                          # sharded.ShardedPairList: i[p] = owned_index[row of p]
    if rebuild_i is None:
        rebuild_i = whole or row_index is not None
    elif rebuild_i and not (whole or row_index is not None):
        raise ValueError("rebuild_i needs a whole PairList or a shard list with its row -> global index map")
    if whole or not rebuild_i:
        row_index = None
    params = getattr(nlist, "params", None)
    if params is None:   # sharded.ShardedPairList: only the integer type is read
        params = _lib.NlParams()
        params.int_type = _lib.NL_I64 if it == torch.int64 else _lib.NL_I32
    S = nlist.S if nlist.S.is_contiguous() else nlist.S.contiguous()
    with torch.cuda.device(dev):
        i_from = _i_copy_from(P, rebuild_i, i_copy_fraction)
        _lib.check(_lib.lib().nl_pairs_to_host(params, _ptr(nlist.first), n_rows, _ptr(nlist.i), i_from, _ptr(row_index),
                                               None if row_index is None else out.rows.data_ptr(), _ptr(nlist.j),
                                               _ptr(S), P, out.first.data_ptr(), out.i.data_ptr(), out.j.data_ptr(), out.S.data_ptr(),
                                               out.dev_scratch.data_ptr(), out.host_scratch.data_ptr(), out.dev_scratch.numel(),
                                               int(nthreads), _stream(dev)))
    return HostPairList(X=getattr(nlist, "X", getattr(nlist, "X_owned", None)), C=getattr(nlist, "C", None), cutoff=getattr(nlist, "cutoff", None),
                        i=out.i[:P].numpy(), j=out.j[:P].numpy(), S=out.S[:P].numpy(),
                        first=out.first[:n_rows + 1].numpy())


def _i_copy_from(P: int, rebuild_i: bool, i_copy_fraction: float) -> int:
    if not rebuild_i:
        return 0
    if i_copy_fraction <= 0.0:
        return P
    return min(P, max(0, int(P * (1.0 - min(1.0, max(0.0, i_copy_fraction))))) & ~63)


def to_host_bytes(nlist, rebuild_i: bool = True, i_copy_fraction: float = 0.0) -> int:
    """Bytes nl_pairs_to_host moves over the bus for this list (when no shift component escapes the one-byte code)."""
    P = int(nlist.i.shape[0])
    w = nlist.i.element_size()
    rows = int(nlist.first.shape[0]) - 1
    rowmap = rows * w if (rebuild_i and getattr(nlist, "owned_index", None) is not None) else 0
    return (rows + 1) * w + rowmap + P * w + P + 4 + (P - _i_copy_from(P, rebuild_i, i_copy_fraction)) * w

# ------------------------------------------------------------------ accessors (src/cell_list.jl:25-27, 507-611, 753-833, 919-927)
def npairs(nlist: PairList) -> int:
    return int(nlist.i.shape[0])


def nsites(nl) -> int:
    if isinstance(nl, PairList):
        return int(nl.first.shape[0]) - 1
    return int(nl.X.shape[0])


def cutoff(nl) -> float:
    return nl.cutoff


def nneigs(nlist: PairList, i0: int) -> int:
    f = nlist.first[i0 - 1:i0 + 1].tolist()
    return int(f[1] - f[0])


def maxneigs(nlist: PairList) -> int:
    """maxneigs(nlist) (src/cell_list.jl:513) as one device reduction (nl_max_neighbours)."""
    N = nsites(nlist)
    if N == 0:
        raise ValueError("maxneigs of an empty list")  # reference: maximum over an empty collection throws
    dev = nlist.first.device
    with torch.cuda.device(dev):
        out = torch.empty(1, dtype=torch.int64, device=dev)
        _lib.check(_lib.lib().nl_max_neighbours(_list_params(nlist), _ptr(nlist.first), N, _ptr(out), _stream(dev)))
    return int(out.item())


max_neighbours = maxneigs
max_neigs = maxneigs


def _list_params(nlist: PairList):
    """nl_params of a PairList (built by materialize_pairlist; rebuilt from C/cutoff for hand-made lists)."""
    if nlist.params is None:
        fdt = np.dtype(_T2N[nlist.X.dtype])
        geo = geometry(nlist.C, nlist.cutoff, (False, False, False), fdt)
        nlist.params = _lib.make_params(geo, fdt, _T2N[nlist.first.dtype])
    return nlist.params


def pairs_R(nlist: PairList, lo: int = 0, hi: Optional[int] = None, X: Optional[torch.Tensor] = None) -> torch.Tensor:
    """R[n] = (X[j] - X[i]) + C' * S[n] for the pairs [lo, hi) (0-based) with the reference's association
    (_getR, src/cell_list.jl:525-531), in one kernel (nl_pairs_R) instead of scalar indexing.  `X` overrides
    the stored positions (same atoms, moved: the skin-list refresh)."""
    P = npairs(nlist)
    hi = P if hi is None else hi
    if not 0 <= lo <= hi <= P:
        raise IndexError("pair range out of bounds")
    Xp = nlist.X if X is None else X
    if Xp.dtype != nlist.X.dtype or Xp.shape != nlist.X.shape or Xp.device != nlist.X.device:
        raise ValueError("X must match the list's positions in dtype, shape and device")
    Xp = Xp.contiguous()
    dev = Xp.device
    with torch.cuda.device(dev):
        R = torch.empty((hi - lo, 3), dtype=Xp.dtype, device=dev)
        _lib.check(_lib.lib().nl_pairs_R(_list_params(nlist), _ptr(Xp), Xp.shape[0], _ptr(nlist.i), _ptr(nlist.j), _ptr(nlist.S),
                                         lo, hi, _ptr(R), _stream(dev)))
    return R


def _getR(nlist: PairList, lo: int, hi: int) -> torch.Tensor:
    return pairs_R(nlist, lo, hi)


def neigss(nlist: PairList, i0: int):
    """(j, R, S) of atom i0 (1-based)."""
    f = nlist.first[i0 - 1:i0 + 1].tolist()
    lo, hi = int(f[0]) - 1, int(f[1]) - 1
    return nlist.j[lo:hi], _getR(nlist, lo, hi), nlist.S[lo:hi]


def sites_padded(nlist: PairList, rows=None, width: Optional[int] = None, with_R: bool = True, with_S: bool = True):
    """Neighbourhoods of many atoms at once: the sites() loop of src/iterators.jl:27-40 over neigss!
    (src/cell_list.jl:583-592) as fixed-width device blocks (nl_rows_padded).

    rows: 1-based atom indices (default: all atoms); width: block width (default: maxneigs of the list).
    Returns (n, j, R, S): n[s] = nneigs(rows[s]); j (n_sel, width) with 0 padding; R (n_sel, width, 3) or None;
    S (n_sel, width, 3) or None.  Rows longer than `width` are truncated (n still holds the full count)."""
    N = nsites(nlist)
    dev = nlist.first.device
    it = nlist.first.dtype
    if rows is None:
        rows_t = torch.arange(1, N + 1, dtype=it, device=dev)
    else:
        rows_t = torch.as_tensor(rows, device=dev).to(it).contiguous().reshape(-1)
        if rows_t.numel() and (int(rows_t.min()) < 1 or int(rows_t.max()) > N):
            raise IndexError("atom index out of range")  # BoundsError in the reference
    n_sel = int(rows_t.numel())
    if width is None:
        width = maxneigs(nlist) if N > 0 else 0
    with torch.cuda.device(dev):
        n_out = torch.empty(n_sel, dtype=it, device=dev)
        j_out = torch.empty((n_sel, width), dtype=it, device=dev)
        S_out = torch.empty((n_sel, width, 3), dtype=it, device=dev) if with_S else None
        R_out = torch.empty((n_sel, width, 3), dtype=nlist.X.dtype, device=dev) if with_R else None
        if n_sel > 0:
            _lib.check(_lib.lib().nl_rows_padded(_list_params(nlist), _ptr(nlist.X), nlist.X.shape[0], _ptr(nlist.first), _ptr(nlist.j),
                                                 _ptr(nlist.S), _ptr(rows_t), n_sel, width, _ptr(n_out), _ptr(j_out), _ptr(S_out),
                                                 _ptr(R_out), _stream(dev)))
    return n_out, j_out, R_out, S_out


def pairs(nlist: PairList, chunk: int = 1 << 20):
    """pairs(nlist) (src/iterators.jl:12-25): iterates (i, j, R) over all pairs; R comes from nl_pairs_R in
    chunks, so the host loop never indexes device memory element by element."""
    P = npairs(nlist)
    for lo in range(0, P, chunk):
        hi = min(P, lo + chunk)
        ii, jj = nlist.i[lo:hi].tolist(), nlist.j[lo:hi].tolist()
        RR = pairs_R(nlist, lo, hi).tolist()
        for n in range(hi - lo):
            yield ii[n], jj[n], RR[n]


def sites(nlist: PairList, chunk: int = 1 << 14):
    """sites(nlist) (src/iterators.jl:27-40): iterates (i, j, R) with j, R the neighbourhood of atom i, gathered
    chunk rows at a time on the device (nl_rows_padded)."""
    N = nsites(nlist)
    if N == 0:
        return
    width = maxneigs(nlist)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        n, j, R, _ = sites_padded(nlist, torch.arange(lo + 1, hi + 1, device=nlist.first.device), width, with_R=True, with_S=False)
        n, j, R = n.tolist(), j.cpu().numpy(), R.cpu().numpy()
        for s in range(hi - lo):
            yield lo + s + 1, j[s, :n[s]], R[s, :n[s]]


def neigs(nlist: PairList, i0: int):
    j, R, _ = neigss(nlist, i0)
    return j, R


def count_neighbours(clist: SortedCellList, i: Optional[int] = None):
    """count_neighbours(clist, i) (src/cell_list.jl:808-814); with i=None, the counts of ALL atoms
    as a device tensor in original order (one fused traversal, nl_lazy_count)."""
    if clist._counts is None:
        N = clist.X.shape[0]
        dev = clist.X.device
        with torch.cuda.device(dev):
            out = torch.zeros(N, dtype=clist.perm.dtype, device=dev)
            ws = _pairs_workspace(clist)
            _lib.check(_lib.lib().nl_lazy_count(clist.params, _ptr(clist.X), N, _ptr(clist.perm), _ptr(clist.cell_offsets),
                                                _ptr(out), _ptr(ws), ws.numel(), _stream(dev)))
        clist._counts = out
    if i is None:
        return clist._counts
    return int(clist._counts[i - 1].item())


def lj_energy(clist: SortedCellList, eps: float, sigma: float) -> torch.Tensor:
    """Fused for_each_neighbour traversal with a Lennard-Jones sink: sum over ORDERED pairs of
    4 eps ((sigma/r)^12 - (sigma/r)^6); returns a device float64 scalar (nl_lazy_lj_energy)."""
    N = clist.X.shape[0]
    dev = clist.X.device
    with torch.cuda.device(dev):
        e = torch.zeros(1, dtype=torch.float64, device=dev)
        ws = _pairs_workspace(clist)
        _lib.check(_lib.lib().nl_lazy_lj_energy(clist.params, _ptr(clist.X), N, _ptr(clist.perm), _ptr(clist.cell_offsets),
                                                float(eps), float(sigma), _ptr(e), _ptr(ws), ws.numel(), _stream(dev)))
    return e


def lj_forces(clist: SortedCellList, eps: float, sigma: float):
    """Fused traversal with a Lennard-Jones FORCE sink (nl_lazy_lj_forces): returns (F, e), views of one (N,4) device
    tensor in original atom order: F[n] = -dE/dx_n for E = 1/2 sum(e), e[n] = sum over n's neighbours of
    4 eps ((sigma/r)^12 - (sigma/r)^6); e.sum() equals lj_energy(clist) (ordered-pair convention)."""
    N = clist.X.shape[0]
    dev = clist.X.device
    with torch.cuda.device(dev):
        fe = torch.zeros((N, 4), dtype=clist.X.dtype, device=dev)
        ws = _pairs_workspace(clist)
        _lib.check(_lib.lib().nl_lazy_lj_forces(clist.params, _ptr(clist.X), N, _ptr(clist.perm), _ptr(clist.cell_offsets),
                                                float(eps), float(sigma), _ptr(fe), _ptr(ws), ws.numel(), _stream(dev)))
    return fe[:, :3], fe[:, 3]


def neighbours_padded(clist: SortedCellList, atoms, width: int, with_R: bool = True, with_S: bool = True):
    """neighbours(clist, i) (src/cell_list.jl:821-833) for MANY atoms at once, straight from the cell list -- nothing is
    materialised (nl_lazy_neighbours).  atoms: 1-based indices.  Returns (n, j, R, S) as sites_padded does; rows are in
    the reference's traversal order (dz, dy, dx, then sorted slot) and truncated to `width` (n holds the full count)."""
    N = clist.X.shape[0]
    dev = clist.X.device
    it = clist.perm.dtype
    at = torch.as_tensor(atoms, device=dev).to(it).contiguous().reshape(-1)
    if at.numel() and (int(at.min()) < 1 or int(at.max()) > N):
        raise IndexError("atom index out of range")  # BoundsError in the reference
    n_sel = int(at.numel())
    with torch.cuda.device(dev):
        n_out = torch.empty(n_sel, dtype=it, device=dev)
        j_out = torch.empty((n_sel, width), dtype=it, device=dev)
        S_out = torch.empty((n_sel, width, 3), dtype=it, device=dev) if with_S else None
        R_out = torch.empty((n_sel, width, 3), dtype=clist.X.dtype, device=dev) if with_R else None
        if n_sel > 0:
            _lib.check(_lib.lib().nl_lazy_neighbours(clist.params, _ptr(clist.X_orig), _ptr(clist.X), N, _ptr(clist.perm),
                                                     _ptr(clist.cell_offsets), _ptr(at), n_sel, width, _ptr(n_out), _ptr(j_out),
                                                     _ptr(S_out), _ptr(R_out), _stream(dev)))
    return n_out, j_out, R_out, S_out


def neighbours(nl, i: int):
    """neighbours(nlist_or_clist, i) -> (j, R, S)  (src/cell_list.jl:606, :821-833).  For a SortedCellList this is one
    warp's traversal of atom i's stencil on the device; no pair list is built."""
    if isinstance(nl, PairList):
        return neigss(nl, i)
    width = 64
    while True:
        n, j, R, S = neighbours_padded(nl, [i], width)
        k = int(n.item())
        if k <= width:
            return j[0, :k], R[0, :k], S[0, :k]
        width = k


def for_each_neighbour(f, clist: SortedCellList, i: int):
    """for_each_neighbour(f, clist, i): calls f(j, R, S) per neighbour of atom i (host-side
    convenience over the device lists; device-side traversal sinks are count_neighbours / lj_energy)."""
    js, Rs, Ss = neighbours(clist, i)
    js, Rs, Ss = js.tolist(), Rs.tolist(), Ss.tolist()
    for n in range(len(js)):
        f(js[n], Rs[n], Ss[n])


def num_neighbours(nl, i: int) -> int:
    return nneigs(nl, i) if isinstance(nl, PairList) else count_neighbours(nl, i)
