# NeighbourListsCUDAExt.jl -- replacement for ext/NeighbourListsCUDAExt.jl (today an empty shell,
# /root/reference/ext/NeighbourListsCUDAExt.jl:1-13): CuArray dispatch of the sort-based path onto
# libnlcuda.so.  Thin ccall wrappers only; PairList / SortedCellList (src/types.jl) are unchanged.
#
# NOT EXECUTED in this repository (no Julia in the build image).  Its tested twin, call for call, is
# neighbourlists.jl_b200/api.py + _lib.py (ctypes); keep the two in sync.
module NeighbourListsCUDAExt

using NeighbourLists
using NeighbourLists: SVec, SMat, SortedCellList, PairList, analyze_cell, lengths
using CUDA
using StaticArrays

const libnlcuda = get(ENV, "NLCUDA_LIB", "libnlcuda.so")

# struct nl_params (include/nlcuda.h) -- 192 bytes
struct NlParams
    float_type::Int32
    int_type::Int32
    cell::NTuple{9,Float64}
    inv_cell::NTuple{9,Float64}
    cutoff::Float64
    ncells::NTuple{3,Int32}
    nxyz::NTuple{3,Int32}
    pbc::NTuple{3,UInt8}
    reserved::NTuple{5,UInt8}
end

_ftag(::Type{Float32}) = Int32(0); _ftag(::Type{Float64}) = Int32(1)
_itag(::Type{Int32}) = Int32(0);   _itag(::Type{Int64}) = Int32(1)

function _check(rc::Cint)
    rc == 0 && return
    error(unsafe_string(ccall((:nl_strerror, libnlcuda), Cstring, (Cint,), rc)))   # ErrorException, as cell_list.jl:656
end

function _params(cell::SMat{T}, inv_cell::SMat{T}, pbc::SVec{Bool}, cutoff::T, ncells::SVec{TI}) where {T,TI}
    lens = abs.(lengths(cell))
    nxyz = ceil.(Int32, cutoff * (ncells ./ lens))                               # gpu_kernels.jl:315-316
    NlParams(_ftag(T), _itag(TI), Tuple(Float64.(cell)), Tuple(Float64.(inv_cell)), Float64(cutoff),
             Tuple(Int32.(ncells)), Tuple(nxyz), Tuple(UInt8.(pbc)), ntuple(_ -> 0x00, 5))
end

_ws(p, N, stage) = CUDA.zeros(UInt8, max(256, ccall((:nl_workspace_bytes, libnlcuda), Csize_t,
                                                    (Ref{NlParams}, Int64, Cint), p, N, stage)))
_stream() = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)

# ---- stage override 1: _build_sorted_celllist (src/cell_list.jl:647-679)
function NeighbourLists._build_sorted_celllist(X::CuVector{SVec{T}}, cell::SMat{T}, pbc::SVec{Bool}, cutoff::T,
                                               ::Type{TI}, backend) where {T,TI}
    nat = length(X)
    inv_cell, ncells, lens = analyze_cell(cell, cutoff, TI)                      # host, unchanged
    prod(BigInt.(ncells)) > typemax(TI) && error("Ratio of simulation cell size to cutoff is very large. ...")
    ncells_total = prod(ncells)
    p = _params(cell, inv_cell, pbc, cutoff, ncells)
    Xs = similar(X); perm = CuVector{TI}(undef, nat); cid = CuVector{TI}(undef, nat)
    offs = CuVector{TI}(undef, ncells_total + 1)
    ws = _ws(p, nat, 0)
    _check(ccall((:nl_build_cells, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, X, nat, Xs, perm, cid, offs, ws, length(ws), _stream()))
    return SortedCellList{T,TI,typeof(Xs),typeof(perm)}(Xs, X, perm, cid, offs, cell, inv_cell, pbc, cutoff, ncells, ncells_total)
end

# ---- stage override 2: materialize_pairlist (src/gpu_kernels.jl:299-364)
function NeighbourLists.materialize_pairlist(clist::SortedCellList{T,TI,<:CuVector}; backend = nothing) where {T,TI}
    nat = length(clist.X)
    p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    first = CuVector{TI}(undef, nat + 1)
    ws = _ws(p, nat, 1)
    total = Ref{Int64}(0)
    _check(ccall((:nl_count_pairs, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ref{Int64}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, clist.X, nat, clist.perm, clist.cell_offsets, first, total, ws, length(ws), _stream()))
    P = total[]
    i = CuVector{TI}(undef, P); j = CuVector{TI}(undef, P); S = CuVector{SVec{TI}}(undef, P)
    if P > 0
        _check(ccall((:nl_fill_pairs, libnlcuda), Cint,
                     (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid},
                      CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                     p, clist.X, nat, clist.perm, clist.cell_offsets, first, i, j, S, CU_NULL, ws, length(ws), _stream()))
    end
    return PairList{T,TI,typeof(clist.X_orig),typeof(i),typeof(S)}(clist.X_orig, clist.cell, clist.cutoff, i, j, S, first)
end

# ---- fused lazy sinks (for_each_neighbour with fixed bodies, src/cell_list.jl:779-814)
function count_neighbours_all(clist::SortedCellList{T,TI,<:CuVector}) where {T,TI}
    nat = length(clist.X); p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    out = CUDA.zeros(TI, nat); ws = _ws(p, nat, 1)
    _check(ccall((:nl_lazy_count, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, clist.X, nat, clist.perm, clist.cell_offsets, out, ws, length(ws), _stream()))
    return out
end

function lj_energy(clist::SortedCellList{T,TI,<:CuVector}, eps, sigma) where {T,TI}
    nat = length(clist.X); p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    e = CUDA.zeros(Float64, 1); ws = _ws(p, nat, 1)
    _check(ccall((:nl_lazy_lj_energy, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Float64, Float64, CuPtr{Float64}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, clist.X, nat, clist.perm, clist.cell_offsets, Float64(eps), Float64(sigma), e, ws, length(ws), _stream()))
    return e
end

end # module
