"""ctypes binding of libnlcuda.so (include/nlcuda.h).  One Python function per C entry point, same
argument order; this file is the tested twin of the Julia `ccall` stubs in INTEGRATION.md.

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

NL_F32, NL_F64 = 0, 1
NL_I32, NL_I64 = 0, 1
NL_STAGE_BUILD, NL_STAGE_PAIRS = 0, 1
NL_FLAG_HALF = 1
NL_OK, NL_ERR_BAD_ARG, NL_ERR_WORKSPACE, NL_ERR_CUDA, NL_ERR_OVERFLOW, NL_ERR_UNSUPPORTED, NL_ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
NL_MAX_RANKS = 64

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NLCUDA_LIB") or os.path.join(_HERE, "libnlcuda.so")  # NLCUDA_LIB: A/B testing of kernel builds


class NlParams(C.Structure):
    """struct nl_params (include/nlcuda.h)."""
    _fields_ = [
        ("float_type", C.c_int32),
        ("int_type", C.c_int32),
        ("cell", C.c_double * 9),
        ("inv_cell", C.c_double * 9),
        ("cutoff", C.c_double),
        ("ncells", C.c_int32 * 3),
        ("nxyz", C.c_int32 * 3),
        ("pbc", C.c_uint8 * 3),
        ("reserved", C.c_uint8 * 5),
    ]


class NlShardInfo(C.Structure):
    """struct nl_shard_info (include/nlcuda.h): what nl_shard_prepare tells the host about this rank's slab."""
    _fields_ = [
        ("axis", C.c_int32), ("halo", C.c_int32), ("periodic", C.c_int32), ("nranks", C.c_int32), ("rank", C.c_int32), ("nplanes", C.c_int32),
        ("has_dn", C.c_int32), ("has_up", C.c_int32), ("dn_peer", C.c_int32), ("up_peer", C.c_int32),
        ("n_local", C.c_int64), ("n_owned", C.c_int64), ("n_halo_dn", C.c_int64), ("n_halo_up", C.c_int64),
        ("n_send_dn", C.c_int64), ("n_send_up", C.c_int64),
        ("bounds", C.c_int64 * (NL_MAX_RANKS + 1)), ("send_count", C.c_int64 * NL_MAX_RANKS), ("recv_count", C.c_int64 * NL_MAX_RANKS),
        ("n_max_all", C.c_int64), ("src_offset", C.c_int64 * NL_MAX_RANKS), ("halo_src_offset_dn", C.c_int64), ("halo_src_offset_up", C.c_int64),
    ]


class NlShardPeers(C.Structure):
    """struct nl_shard_peers (include/nlcuda.h): the other ranks' workspaces mapped into this process (peer path)."""
    _fields_ = [
        ("nranks", C.c_int32), ("rank", C.c_int32), ("cap", C.c_int64), ("ws_bytes", C.c_uint64), ("ws", C.c_void_p),
        ("peer_ws", C.c_void_p * NL_MAX_RANKS), ("peer_base", C.c_void_p * NL_MAX_RANKS),
    ]


class NlError(RuntimeError):
    """Mirrors the reference's ErrorException convention (src/cell_list.jl:656-658)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_lib = None

EXPORTS = ("nl_version", "nl_strerror", "nl_last_cuda_error", "nl_launch_count", "nl_workspace_bytes", "nl_build_cells", "nl_count_pairs",
           "nl_fill_pairs", "nl_fill_pairs_rows", "nl_count_pairs_window", "nl_fill_pairs_window", "nl_cell_ids", "nl_shard_plan", "nl_lazy_count", "nl_lazy_lj_energy", "nl_lazy_lj_forces",
           "nl_pairs_R", "nl_max_neighbours", "nl_rows_padded", "nl_lazy_neighbours", "nl_bounding_box", "nl_max_displacement2",
           "nl_shard_workspace_bytes", "nl_shard_prepare", "nl_shard_exchange", "nl_nccl_unique_id", "nl_nccl_comm_init", "nl_nccl_comm_destroy",
           "nl_to_host_scratch_bytes", "nl_pairs_to_host", "nl_host_expand_rows", "nl_host_unpack_shifts",
           "nl_shard_connect", "nl_shard_exchange_peer", "nl_shard_disconnect", "nl_pairs_to_host_begin", "nl_pairs_to_host_finish")
NL_REDUCE_WS_BYTES = 32768


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NlError(-100, f"libnlcuda.so not found at {LIB_PATH}: build it with __graft_entry__.build() "
                                "(neighbourlists.jl_b200/csrc/build.sh); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i64, sz = C.c_void_p, C.c_int64, C.c_size_t
        pp = C.POINTER(NlParams)
        L.nl_version.restype = C.c_int
        L.nl_strerror.restype = C.c_char_p
        L.nl_strerror.argtypes = [C.c_int]
        L.nl_last_cuda_error.restype = C.c_int
        L.nl_launch_count.restype = C.c_longlong
        L.nl_workspace_bytes.restype = sz
        L.nl_workspace_bytes.argtypes = [pp, i64, C.c_int]
        L.nl_build_cells.argtypes = [pp, vp, i64, vp, vp, vp, vp, vp, sz, vp]
        L.nl_count_pairs.argtypes = [pp, vp, i64, vp, vp, vp, C.POINTER(C.c_int64), vp, sz, vp]
        L.nl_fill_pairs.argtypes = [pp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
        L.nl_fill_pairs_rows.argtypes = [pp, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, sz, vp]
        L.nl_count_pairs_window.argtypes = [pp, vp, i64, vp, vp, vp, C.POINTER(C.c_int64), vp, vp, sz, vp]
        L.nl_fill_pairs_window.argtypes = [pp, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, sz, vp]
        L.nl_count_pairs_window.restype = C.c_int
        L.nl_fill_pairs_window.restype = C.c_int
        L.nl_cell_ids.argtypes = [pp, vp, i64, vp, vp]
        L.nl_shard_plan.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp]
        L.nl_shard_plan.restype = C.c_int
        L.nl_lazy_count.argtypes = [pp, vp, i64, vp, vp, vp, vp, sz, vp]
        L.nl_lazy_lj_energy.argtypes = [pp, vp, i64, vp, vp, C.c_double, C.c_double, vp, vp, sz, vp]
        L.nl_lazy_lj_forces.argtypes = [pp, vp, i64, vp, vp, C.c_double, C.c_double, vp, vp, sz, vp]
        L.nl_lazy_lj_forces.restype = C.c_int
        L.nl_pairs_R.argtypes = [pp, vp, i64, vp, vp, vp, i64, i64, vp, vp]
        L.nl_max_neighbours.argtypes = [pp, vp, i64, vp, vp]
        L.nl_rows_padded.argtypes = [pp, vp, i64, vp, vp, vp, vp, i64, C.c_int32, vp, vp, vp, vp, vp]
        L.nl_lazy_neighbours.argtypes = [pp, vp, vp, i64, vp, vp, vp, i64, C.c_int32, vp, vp, vp, vp, vp]
        L.nl_lazy_neighbours.restype = C.c_int
        L.nl_bounding_box.argtypes = [C.c_int32, vp, i64, vp, vp, sz, vp]
        L.nl_max_displacement2.argtypes = [C.c_int32, vp, vp, i64, vp, vp, sz, vp]
        ps = C.POINTER(NlShardInfo)
        L.nl_shard_workspace_bytes.restype = sz
        L.nl_shard_workspace_bytes.argtypes = [pp, i64, C.c_int32]
        L.nl_shard_prepare.argtypes = [pp, vp, i64, vp, C.c_int32, C.c_int32, ps, vp, sz, vp]
        L.nl_shard_exchange.argtypes = [pp, ps, vp, vp, i64, vp, vp, vp, vp, vp, sz, vp]
        pq = C.POINTER(NlShardPeers)
        L.nl_shard_connect.argtypes = [pp, i64, vp, C.c_int32, C.c_int32, vp, sz, pq, vp]
        L.nl_shard_exchange_peer.argtypes = [pp, ps, vp, vp, i64, vp, pq, vp, vp, vp, vp, sz, vp]
        L.nl_shard_disconnect.argtypes = [pq]
        for n in ("nl_shard_connect", "nl_shard_exchange_peer", "nl_shard_disconnect", "nl_pairs_to_host_begin", "nl_pairs_to_host_finish"):
            getattr(L, n).restype = C.c_int
        L.nl_nccl_unique_id.argtypes = [vp]
        L.nl_nccl_comm_init.argtypes = [C.POINTER(vp), C.c_int32, vp, C.c_int32]
        L.nl_nccl_comm_destroy.argtypes = [vp]
        for n in ("nl_shard_prepare", "nl_shard_exchange", "nl_nccl_unique_id", "nl_nccl_comm_init", "nl_nccl_comm_destroy"):
            getattr(L, n).restype = C.c_int
        for n in ("nl_pairs_R", "nl_max_neighbours", "nl_rows_padded", "nl_lazy_neighbours", "nl_bounding_box", "nl_max_displacement2",
           "nl_shard_workspace_bytes", "nl_shard_prepare", "nl_shard_exchange", "nl_nccl_unique_id", "nl_nccl_comm_init", "nl_nccl_comm_destroy"):
            getattr(L, n).restype = C.c_int
        for n in ("nl_build_cells", "nl_count_pairs", "nl_fill_pairs", "nl_fill_pairs_rows", "nl_count_pairs_window", "nl_fill_pairs_window", "nl_cell_ids", "nl_shard_plan", "nl_lazy_count",
                  "nl_lazy_lj_energy"):
            getattr(L, n).restype = C.c_int
        L.nl_to_host_scratch_bytes.restype = sz
        L.nl_to_host_scratch_bytes.argtypes = [i64]
        L.nl_pairs_to_host.argtypes = [pp, vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, sz, C.c_int32, vp]
        L.nl_pairs_to_host.restype = C.c_int
        L.nl_pairs_to_host_begin.argtypes = [pp, vp, i64, i64, vp, vp, C.c_int32, vp, C.POINTER(vp)]
        L.nl_pairs_to_host_begin.restype = C.c_int
        L.nl_pairs_to_host_finish.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz, vp]
        L.nl_pairs_to_host_finish.restype = C.c_int
        L.nl_host_expand_rows.argtypes = [C.c_int32, vp, vp, i64, i64, i64, vp]
        L.nl_host_expand_rows.restype = C.c_int
        L.nl_host_unpack_shifts.argtypes = [C.c_int32, vp, i64, i64, vp]
        L.nl_host_unpack_shifts.restype = C.c_int
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        L = lib()
        msg = L.nl_strerror(rc).decode()
        if rc == NL_ERR_CUDA:
            msg += f" [cudaError_t {L.nl_last_cuda_error()}]"
        raise NlError(rc, msg)


def make_params(geo, float_dtype, int_dtype) -> NlParams:
    """Fill nl_params from a cellmath.CellGeometry (matrices flattened in Julia column-major order)."""
    p = NlParams()
    p.float_type = NL_F64 if np.dtype(float_dtype) == np.float64 else NL_F32
    p.int_type = NL_I64 if np.dtype(int_dtype) == np.int64 else NL_I32
    c = np.asarray(geo.cell, dtype=np.float64).ravel(order="F")
    ic = np.asarray(geo.inv_cell, dtype=np.float64).ravel(order="F")
    for k in range(9):
        p.cell[k] = float(c[k])
        p.inv_cell[k] = float(ic[k])
    p.cutoff = float(geo.cutoff)
    for k in range(3):
        if int(geo.ncells[k]) > 2**31 - 1:
            raise NlError(NL_ERR_UNSUPPORTED, lib().nl_strerror(NL_ERR_UNSUPPORTED).decode())
        p.ncells[k] = int(geo.ncells[k])
        p.nxyz[k] = int(geo.nxyz[k])
        p.pbc[k] = 1 if geo.pbc[k] else 0
    return p
