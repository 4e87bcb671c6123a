"""neighbourlists.jl_b200 -- B200-native drop-in for the sort-based neighbour-list path of
JuliaMolSim/NeighbourLists.jl.  csrc/ holds the sm_100a CUDA kernels and the C ABI
(libnlcuda.so, include/nlcuda.h); api.py mirrors the reference's operator interface on top of it.

The directory name contains a dot, so import it through the repo-root shim:
    import neighbourlists_jl_b200 as nl
"""
from . import _lib, atoms, cellmath, skin
from ._lib import NlError
from .api import (HostPairBuffers, HostPairList, to_host, to_host_bytes, PairList, SortedCellList, build_cell_list, count_neighbours, cutoff, for_each_neighbour, lj_energy, lj_forces,
                  materialize_pairlist, max_neigs, max_neighbours, maxneigs, neighbour_list, neighbours, neighbours_padded, neigs, neigss, nneigs,
                  npairs, nsites, num_neighbours, pairs, pairs_R, sites, sites_padded)
from .atoms import System, bounding_box, bounding_cell, isolated_system, periodic_system
from .skin import SkinList, max_displacement2

__all__ = ["HostPairBuffers", "HostPairList", "to_host", "to_host_bytes", "PairList", "SortedCellList", "build_cell_list", "materialize_pairlist", "neighbour_list", "for_each_neighbour",
           "count_neighbours", "neighbours", "num_neighbours", "npairs", "nsites", "cutoff", "nneigs", "maxneigs",
           "max_neighbours", "max_neigs", "neigs", "neigss", "lj_energy", "lj_forces", "NlError", "cellmath", "pairs", "sites", "pairs_R", "neighbours_padded",
           "sites_padded", "atoms", "skin", "System", "isolated_system", "periodic_system", "bounding_box", "bounding_cell",
           "SkinList", "max_displacement2"]
