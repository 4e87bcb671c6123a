"""Multi-GPU neighbour lists: 1-D spatial slabs with a cutoff-wide halo.

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU in the tests).  The box is cut
into slabs of whole cell PLANES along the axis with the most cells (ties: z, the slowest key axis); rank r owns planes
[bounds[r], bounds[r+1]) balanced by atom count.  Three exchanges, all plain point-to-point data
movement (no reduction on the data path):

  0. all-to-all-v: atoms (position + global index) move to the rank that owns their plane;
  1. halo: every rank sends the atoms of its `halo` top planes to rank r+1 and of its `halo` bottom
     planes to rank r-1 (ring when the slab axis is periodic; the same peer twice when G = 2);
  2. none afterwards: the local atom set (owned first, then halo) goes through the UNCHANGED
     single-GPU pipeline with the GLOBAL cell, cutoff and pbc.  Every neighbour of an owned atom is
     present locally and planes owned by nobody are empty, so the rows of the owned atoms -- and
     their periodic shifts S -- are exactly the global ones; the kernel writes i/j as global indices
     and skips the (incomplete) rows of the halo atoms (nl_fill_pairs_rows).

The reference has no multi-GPU code; concatenating the ranks' rows by global i reproduces its
single-device CSR (tests/test_sharded_*).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from .cellmath import CellGeometry, geometry

_PROFILE = bool(os.environ.get("NL_SHARD_PROFILE"))
_PEER = os.environ.get("NL_SHARD_PEER", "1") != "0"   # 0: ncclSend / ncclRecv instead of NVLink peer copies (A/B runs)
_peer_state = {}   # (device index, communicator) -> persistent workspace + nl_shard_peers of the peer path


def shard_disconnect(comm=None):
    """Unmaps the peer workspaces of the peer path (all communicators, or one).  Call before destroying the communicator."""
    from . import _lib
    import ctypes as C
    for key in list(_peer_state):
        if comm is None or key[1] == int(comm.value or 0):
            st = _peer_state.pop(key)
            _lib.check(_lib.lib().nl_shard_disconnect(C.byref(st["peers"])))


class _Phases:
    """Optional per-phase device timing (NL_SHARD_PROFILE=1): CUDA events between the phases of one call."""

    def __init__(self, on):
        self.on, self.marks = on, []

    def mark(self, name):
        if self.on and torch.cuda.is_available():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def report(self, rank):
        if self.on and self.marks and rank == 0:
            torch.cuda.synchronize()
            print("[shard phases ms] " + " ".join(f"{n}={a.elapsed_time(b):.2f}" for (_, a), (n, b) in zip(self.marks[:-1], self.marks[1:])),
                  flush=True)


@dataclass
class SlabPlan:
    axis: int            # slab axis (0, 1, 2)
    bounds: np.ndarray   # (G+1,) plane boundaries, bounds[0] = 0, bounds[G] = ncells[axis]
    halo: int            # halo width in planes = nxyz[axis]
    periodic: bool


def plan_slabs(plane_hist: np.ndarray, world: int, halo: int, periodic: bool, axis: int) -> SlabPlan:
    """Plane boundaries balanced by atom count; every slab at least 2*halo+1 planes wide so that a
    halo only ever comes from the two adjacent ranks and the two halos of a rank never overlap."""
    n = int(plane_hist.shape[0])
    minw = 2 * halo + 1 if world > 1 else 1
    if world * minw > n:
        raise ValueError(f"cannot cut {n} cell planes into {world} slabs of >= {minw} planes: use fewer ranks (replicas)")
    cum = np.concatenate([[0], np.cumsum(plane_hist.astype(np.int64))])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        # smallest b with cum[b] >= total * r / world, compared exactly (same rule as nl_shard_plan in the library)
        b = int(np.searchsorted(cum * world, total * r, side="left"))
        b = max(b, bounds[-1] + minw)            # this slab wide enough
        b = min(b, n - (world - r) * minw)       # room for the remaining slabs
        bounds.append(b)
    bounds.append(n)
    return SlabPlan(axis=axis, bounds=np.asarray(bounds, dtype=np.int64), halo=halo, periodic=periodic)


class CudaEngine:
    """Local stages on this rank's GPU through libnlcuda.so."""

    def __init__(self, device=None, timers=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.timers = timers  # optional dict: CUDA events around the count and fill stages (bench.py)

    def cell_ids(self, X, cutoff, cell, pbc):
        from . import api
        return api.cell_ids(X, cutoff, cell, pbc, int_type=np.int64).long()

    def build(self, X_all, n_owned, gmap, cutoff, cell, pbc, int_type, with_R, plane_active=None):
        from . import api
        clist = api.build_cell_list(X_all, cutoff, cell, pbc, int_type=int_type)
        pl = api.materialize_pairlist(clist, with_R=with_R, n_rows=n_owned, index_map=gmap, timers=self.timers, plane_active=plane_active)
        return dict(first=pl.first, i=pl.i, j=pl.j, S=pl.S, R=pl.R)


@dataclass
class ShardedPairList:
    """Rows of this rank's owned atoms: row m belongs to global atom owned_index[m]; i, j global (1-based)."""
    owned_index: torch.Tensor
    X_owned: torch.Tensor
    first: torch.Tensor
    i: torch.Tensor
    j: torch.Tensor
    S: torch.Tensor
    R: Optional[torch.Tensor]
    n_halo: int
    plan: SlabPlan


def _a2a(t: torch.Tensor, send_counts, recv_counts, group):
    out = t.new_empty((int(sum(recv_counts)),) + tuple(t.shape[1:]))
    dist.all_to_all_single(out, t.contiguous(), output_split_sizes=[int(c) for c in recv_counts],
                           input_split_sizes=[int(c) for c in send_counts], group=group)
    return out


def neighbour_list_sharded(X_local, gidx_local, cutoff, cell, pbc, *, group=None, int_type=np.int32, with_R=False,
                           engine=None, redistribute=True) -> ShardedPairList:
    """Neighbour list of a system distributed over the ranks of `group`.

    X_local (n,3) and gidx_local (n,) hold ANY subset of the atoms per rank (global 1-based indices);
    with redistribute=False the caller guarantees they already lie in this rank's slab of `plan_slabs`.
    """
    engine = engine or CudaEngine()
    ph = _Phases(_PROFILE)
    ph.mark("start")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = engine.device
    X = torch.as_tensor(X_local).to(dev)
    gidx = torch.as_tensor(gidx_local).to(dev).long()
    fdt = np.float64 if X.dtype == torch.float64 else np.float32
    geo: CellGeometry = geometry(cell, cutoff, pbc, fdt)
    nc = [int(v) for v in geo.ncells]
    axis = 2 - int(np.argmax(nc[::-1]))              # most planes; ties go to the SLOWEST key axis (z), so that a slab is one
                                                     # contiguous range of cell keys / sorted atoms
    halo = int(geo.nxyz[axis])
    stride = [1, nc[0], nc[0] * nc[1]][axis]

    def planes_of(Xt):
        cid = engine.cell_ids(Xt, cutoff, cell, pbc) - 1
        return (cid // stride) % nc[axis]

    # ---- slab plan from the global plane histogram
    planes = planes_of(X) if X.shape[0] else torch.zeros(0, dtype=torch.long, device=dev)
    hist = torch.bincount(planes, minlength=nc[axis]).long()
    if world > 1:
        dist.all_reduce(hist, group=group)
    ph.mark("bin+hist")
    hist_host = hist.cpu().numpy()
    plan = plan_slabs(hist_host, world, halo, bool(geo.pbc[axis]), axis)
    ph.mark("plan")
    bounds_t = torch.as_tensor(plan.bounds, device=dev)

    # ---- step 0: all-to-all-v to the owners
    if world > 1 and redistribute:
        owner = torch.bucketize(planes, bounds_t[1:], right=True)
        stay = owner == rank
        # one small all-reduce decides whether ANY rank has atoms to move (spatially pre-distributed inputs: none)
        n_leave = (~stay).sum().reshape(1)
        dist.all_reduce(n_leave, group=group)
        if int(n_leave.item()) > 0:
            # atoms already on their owner stay put; only the leavers are sorted by destination and exchanged
            leave = (~stay).nonzero().flatten()
            lo_owner = owner[leave]
            order = leave[torch.argsort(lo_owner, stable=True)]
            send_counts = torch.bincount(lo_owner, minlength=world).long()
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts, group=group)
            sc, rc = send_counts.cpu().tolist(), recv_counts.cpu().tolist()
            # the exchanges are collective: every rank takes part even with nothing to send or receive
            rX, rg, rp = _a2a(X[order], sc, rc, group), _a2a(gidx[order], sc, rc, group), _a2a(planes[order], sc, rc, group)
            if leave.numel() > 0:                          # compact only if something left; one index for the three arrays
                keep = stay.nonzero().flatten()
                X, gidx, planes = X.index_select(0, keep), gidx.index_select(0, keep), planes.index_select(0, keep)
            if rX.shape[0] > 0:
                X, gidx, planes = torch.cat([X, rX]), torch.cat([gidx, rg]), torch.cat([planes, rp])
    n_owned = int(X.shape[0])
    ph.mark("redistribute")

    # ---- step 1: halo exchange with ranks r-1 / r+1
    halos_X, halos_g = [], []
    if world > 1:
        lo, hi = int(plan.bounds[rank]), int(plan.bounds[rank + 1])
        up_peer, dn_peer = rank + 1, rank - 1
        if plan.periodic:
            up_peer %= world
            dn_peer %= world
        has_up, has_dn = 0 <= up_peer < world, 0 <= dn_peer < world
        sel_up = (planes >= hi - halo).nonzero().flatten() if has_up else planes.new_zeros(0)
        sel_dn = (planes < lo + halo).nonzero().flatten() if has_dn else planes.new_zeros(0)
        # Message sizes need no exchange: after step 0 every rank holds exactly the atoms of its planes, so the number of
        # atoms in any rank's boundary planes follows from the all-reduced plane histogram every rank already has.
        hist_h = hist_host
        def n_top(r):
            return int(hist_h[int(plan.bounds[r + 1]) - halo:int(plan.bounds[r + 1])].sum())
        def n_bottom(r):
            return int(hist_h[int(plan.bounds[r]):int(plan.bounds[r]) + halo].sum())
        sizes = [[n_top(r), n_bottom(r)] for r in range(world)]
        if sel_up.numel() != (sizes[rank][0] if has_up else 0) or sel_dn.numel() != (sizes[rank][1] if has_dn else 0):
            raise ValueError("atoms are not in this rank's slab (redistribute=False with misplaced atoms?)")
        # Message order matters when both neighbours are the same rank (G = 2, periodic): every rank sends
        # [up, down] and receives [from below, from above], so the k-th send to a peer meets its k-th receive.
        sends, recvs, keep = [], [], []
        rX_dn = rg_dn = rX_up = rg_up = None
        if has_up:   # my top planes are rank up_peer's lower halo
            sX, sg = X[sel_up].contiguous(), gidx[sel_up].contiguous()
            sends += [dist.P2POp(dist.isend, sX, up_peer, group, tag=1), dist.P2POp(dist.isend, sg, up_peer, group, tag=2)]
            keep += [sX, sg]
        if has_dn:   # my bottom planes are rank dn_peer's upper halo
            sX, sg = X[sel_dn].contiguous(), gidx[sel_dn].contiguous()
            sends += [dist.P2POp(dist.isend, sX, dn_peer, group, tag=3), dist.P2POp(dist.isend, sg, dn_peer, group, tag=4)]
            keep += [sX, sg]
        if has_dn:   # lower halo: the top planes of the rank below
            rX_dn = X.new_empty((sizes[dn_peer][0], 3))
            rg_dn = gidx.new_empty((sizes[dn_peer][0],))
            recvs += [dist.P2POp(dist.irecv, rX_dn, dn_peer, group, tag=1), dist.P2POp(dist.irecv, rg_dn, dn_peer, group, tag=2)]
            halos_X.append(rX_dn); halos_g.append(rg_dn)
        if has_up:   # upper halo: the bottom planes of the rank above
            rX_up = X.new_empty((sizes[up_peer][1], 3))
            rg_up = gidx.new_empty((sizes[up_peer][1],))
            recvs += [dist.P2POp(dist.irecv, rX_up, up_peer, group, tag=3), dist.P2POp(dist.irecv, rg_up, up_peer, group, tag=4)]
            halos_X.append(rX_up); halos_g.append(rg_up)
        ops = sends + recvs
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    X_all = torch.cat([X] + halos_X) if halos_X else X
    g_all = torch.cat([gidx] + halos_g) if halos_g else gidx
    n_halo = int(X_all.shape[0]) - n_owned
    ph.mark("halo")

    # ---- step 2: the unchanged single-GPU pipeline on owned + halo atoms, global geometry
    # z slabs: tell the local stage which cell planes can hold atoms (owned planes + halo planes), so that it launches only
    # those tile layers of the global grid
    plane_active = None
    if world > 1 and axis == 2:
        plane_active = np.zeros(nc[2], dtype=np.uint8)
        lo, hi = int(plan.bounds[rank]), int(plan.bounds[rank + 1])
        for z in range(lo - halo, hi + halo):
            if 0 <= z < nc[2]:
                plane_active[z] = 1
            elif plan.periodic:
                plane_active[z % nc[2]] = 1
    res = engine.build(X_all, n_owned, g_all, cutoff, cell, pbc, int_type, with_R, plane_active=plane_active)
    ph.mark("local build")
    ph.report(rank)
    return ShardedPairList(owned_index=gidx, X_owned=X, first=res["first"], i=res["i"], j=res["j"], S=res["S"], R=res["R"],
                           n_halo=n_halo, plan=plan)


# ------------------------------------------------------------------ native driver: nl_shard_prepare / nl_shard_exchange (NCCL inside the library)
def make_nccl_comm(group=None):
    """ncclComm_t (ctypes.c_void_p) over the ranks of `group`, created by the LIBRARY (nl_nccl_comm_init) in the NCCL instance
    this process has loaded; the 128-byte unique id travels over torch.distributed (any backend).  The current CUDA device
    must be this rank's GPU.  Collective."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    buf = (C.c_char * 128)()
    if rank == 0:
        _lib.check(L.nl_nccl_unique_id(C.cast(buf, C.c_void_p)))
    t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
    on_gpu = dist.get_backend(group) == "nccl"
    if on_gpu:
        t = t.cuda()
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    idb = (C.c_char * 128).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
    comm = C.c_void_p()
    _lib.check(L.nl_nccl_comm_init(C.byref(comm), world, C.cast(idb, C.c_void_p), rank))
    return comm


def neighbour_list_sharded_native(X_local, gidx_local, cutoff, cell, pbc, comm, rank: int, world: int, *, int_type=np.int32,
                                  with_R=False, timers=None) -> ShardedPairList:
    """The same list as neighbour_list_sharded, driven entirely through the C ABI: nl_shard_prepare (bin, all-gathered plane
    histograms, slab plan: ONE host read), nl_shard_exchange (owner partition, all-to-all-v and halo exchange with ncclSend /
    ncclRecv inside the library) and the unchanged local stages.  This is the call sequence a Julia driver makes."""
    import ctypes as C
    from . import _lib, api
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    X = api._as_device_positions(X_local, dev)
    it = api._int_dtype(int_type)
    gidx = torch.as_tensor(gidx_local).to(device=dev, dtype=it).contiguous()
    fdt = np.dtype(api._T2N[X.dtype])
    geo = geometry(cell, cutoff, pbc, fdt)
    params = _lib.make_params(geo, fdt, api._T2N[it])
    n = int(X.shape[0])
    st = api._stream(dev)
    ph = _Phases(_PROFILE)
    ph.mark("start")
    with torch.cuda.device(dev):
        use_peer = world > 1 and _PEER
        state = _peer_state.get((dev.index, int(comm.value or 0))) if use_peer else None
        need_n = L.nl_shard_workspace_bytes(params, n, world)
        if state is not None and state["ws"].numel() >= need_n:
            ws = state["ws"]
        else:
            ws = torch.empty(max(need_n, 256), dtype=torch.uint8, device=dev)
        info = _lib.NlShardInfo()
        _lib.check(L.nl_shard_prepare(params, api._ptr(X), n, comm, rank, world, C.byref(info), api._ptr(ws), ws.numel(), st))
        ph.mark("prepare")
        n_owned, n_all = int(info.n_owned), int(info.n_owned + info.n_halo_dn + info.n_halo_up)
        X_all = torch.empty((n_all, 3), dtype=X.dtype, device=dev)
        g_all = torch.empty(n_all, dtype=it, device=dev)
        plane_active = np.ones(int(geo.ncells[2]), dtype=np.uint8)
        if use_peer:
            # peer path: one persistent workspace per (device, communicator), mapped into every other rank (nl_shard_connect).
            # info.n_max_all is the same number on every rank, so every rank takes the same decision here (connect is collective).
            key = (dev.index, int(comm.value or 0))
            if state is None or int(info.n_max_all) > state["cap"] or state["params_key"] != (params.float_type, params.int_type, int(info.nplanes)):
                if state is not None:
                    _lib.check(L.nl_shard_disconnect(C.byref(state["peers"])))
                cap = int(int(info.n_max_all) * 1.2) + 4096
                wsp = torch.empty(L.nl_shard_workspace_bytes(params, cap, world), dtype=torch.uint8, device=dev)
                peers = _lib.NlShardPeers()
                _lib.check(L.nl_shard_connect(params, cap, comm, rank, world, api._ptr(wsp), wsp.numel(), C.byref(peers), st))
                state = _peer_state[key] = dict(ws=wsp, peers=peers, cap=cap, params_key=(params.float_type, params.int_type, int(info.nplanes)))
            ws = state["ws"]
            _lib.check(L.nl_shard_exchange_peer(params, C.byref(info), api._ptr(X), api._ptr(gidx), n, comm, C.byref(state["peers"]), api._ptr(X_all),
                                                api._ptr(g_all), plane_active.ctypes.data, api._ptr(ws), ws.numel(), st))
        else:
            need = L.nl_shard_workspace_bytes(params, max(n, n_owned), world)
            if need > ws.numel():
                ws = torch.empty(need, dtype=torch.uint8, device=dev)
            _lib.check(L.nl_shard_exchange(params, C.byref(info), api._ptr(X), api._ptr(gidx), n, comm, api._ptr(X_all), api._ptr(g_all),
                                           plane_active.ctypes.data, api._ptr(ws), ws.numel(), st))
        ph.mark("exchange")
        clist = api.build_cell_list(X_all, cutoff, cell, pbc, int_type=int_type)
        ph.mark("build")
        pl = api.materialize_pairlist(clist, with_R=with_R, n_rows=n_owned, index_map=g_all, timers=timers,
                                      plane_active=plane_active if (world > 1 and info.axis == 2) else None)
        ph.mark("count+fill")
    ph.report(rank)
    plan = SlabPlan(axis=int(info.axis), bounds=np.asarray(list(info.bounds[:world + 1]), dtype=np.int64), halo=int(info.halo),
                    periodic=bool(info.periodic))
    return ShardedPairList(owned_index=g_all[:n_owned], X_owned=X_all[:n_owned], first=pl.first, i=pl.i, j=pl.j, S=pl.S, R=pl.R,
                           n_halo=n_all - n_owned, plan=plan)
