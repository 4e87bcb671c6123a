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
# `half = true` (not a reference feature) stores one pair of every mirror couple (i, j, S) / (j, i, -S): NL_FLAG_HALF
_with_flags(p::NlParams, flags::UInt8) = NlParams(p.float_type, p.int_type, p.cell, p.inv_cell, p.cutoff, p.ncells, p.nxyz, p.pbc,
                                                  (flags, 0x00, 0x00, 0x00, 0x00))

function NeighbourLists.materialize_pairlist(clist::SortedCellList{T,TI,<:CuVector}; backend = nothing, half::Bool = false) where {T,TI}
    nat = length(clist.X)
    p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    half && (p = _with_flags(p, 0x01))
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

"(F, e): per-atom LJ forces and energies from one fused traversal; fe is N x 4 (F_x, F_y, F_z, e)"
function lj_forces(clist::SortedCellList{T,TI,<:CuVector}, eps, sigma) where {T,TI}
    nat = length(clist.X); p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    fe = CUDA.zeros(T, 4, nat); ws = _ws(p, nat, 1)
    _check(ccall((:nl_lazy_lj_forces, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Float64, Float64, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, clist.X, nat, clist.perm, clist.cell_offsets, Float64(eps), Float64(sigma), fe, ws, length(ws), _stream()))
    return (@view fe[1:3, :]), (@view fe[4, :])
end

# ---- PairList accessors without scalar indexing (src/cell_list.jl:513-606, src/iterators.jl) ------------------
const DevPairList{T,TI} = PairList{T,TI,<:CuVector}

# nl_params of a PairList: only the element types and the cell are read by the accessor kernels
_params(nl::PairList{T,TI}) where {T,TI} =
    NlParams(_ftag(T), _itag(TI), Tuple(Float64.(nl.C)), Tuple(Float64.(inv(nl.C))), Float64(nl.cutoff),
             (Int32(1), Int32(1), Int32(1)), (Int32(1), Int32(1), Int32(1)), (0x00, 0x00, 0x00), ntuple(_ -> 0x00, 5))

"R for the pairs lo:hi (1-based, inclusive) -- the _getR loop of neigss!, on the device; X defaults to nl.X"
function pairs_R(nl::DevPairList{T,TI}, lo::Integer = 1, hi::Integer = length(nl.i); X = nl.X) where {T,TI}
    R = CuVector{SVec{T}}(undef, hi - lo + 1)
    _check(ccall((:nl_pairs_R, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int64, CuPtr{Cvoid}, Ptr{Cvoid}),
                 _params(nl), X, length(X), nl.i, nl.j, nl.S, lo - 1, hi, R, _stream()))
    return R
end

function NeighbourLists.neigss(nl::DevPairList, i0::Integer)
    n1, n2 = CUDA.@allowscalar(nl.first[i0]), CUDA.@allowscalar(nl.first[i0+1]) - 1
    return (@view nl.j[n1:n2]), pairs_R(nl, n1, n2), (@view nl.S[n1:n2])
end

function NeighbourLists.maxneigs(nl::DevPairList)
    out = CUDA.zeros(Int64, 1)
    _check(ccall((:nl_max_neighbours, libnlcuda), Cint, (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Int64}, Ptr{Cvoid}),
                 _params(nl), nl.first, length(nl.first) - 1, out, _stream()))
    return Array(out)[1]
end

"neighbourhoods of the atoms `rows` as padded blocks: (n, j, R, S) with j :: width x n_sel etc."
function sites_padded(nl::DevPairList{T,TI}, rows::CuVector{TI}, width::Integer = maxneigs(nl)) where {T,TI}
    ns = length(rows)
    n = CuVector{TI}(undef, ns); j = CuMatrix{TI}(undef, width, ns)
    S = CuMatrix{SVec{TI}}(undef, width, ns); R = CuMatrix{SVec{T}}(undef, width, ns)
    _check(ccall((:nl_rows_padded, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int32,
                  CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                 _params(nl), nl.X, length(nl.X), nl.first, nl.j, nl.S, rows, ns, width, n, j, S, R, _stream()))
    return n, j, R, S
end

"neighbours(clist, i) for many atoms at once, straight from the cell list (nothing is materialised)"
function neighbours_padded(clist::SortedCellList{T,TI,<:CuVector}, atoms::CuVector{TI}, width::Integer) where {T,TI}
    ns = length(atoms); p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    n = CuVector{TI}(undef, ns); j = CuMatrix{TI}(undef, width, ns)
    S = CuMatrix{SVec{TI}}(undef, width, ns); R = CuMatrix{SVec{T}}(undef, width, ns)
    _check(ccall((:nl_lazy_neighbours, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int32,
                  CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                 p, clist.X_orig, clist.X, length(clist.X), clist.perm, clist.cell_offsets, atoms, ns, width, n, j, S, R, _stream()))
    return n, j, R, S
end

# ---- slab shards (multi-GPU driver): rows of the first n_rows local atoms only, i / j through index_map (global indices);
# plane_active :: Vector{UInt8} (host, one byte per z plane of cells) promises where the local atoms are, so that only
# those tile layers of the global grid are launched
function shard_pairlist(clist::SortedCellList{T,TI,<:CuVector}, n_rows::Integer, index_map::CuVector{TI},
                        plane_active::Union{Nothing,Vector{UInt8}} = nothing) where {T,TI}
    nat = length(clist.X); p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    first = CuVector{TI}(undef, nat + 1); ws = _ws(p, nat, 1); total = Ref{Int64}(0)
    pa = plane_active === nothing ? C_NULL : pointer(plane_active)
    GC.@preserve plane_active begin
        _check(ccall((:nl_count_pairs_window, libnlcuda), Cint,
                     (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ref{Int64}, Ptr{UInt8}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                     p, clist.X, nat, clist.perm, clist.cell_offsets, first, total, pa, ws, length(ws), _stream()))
        P = Int(Array(first[n_rows+1:n_rows+1])[1]) - 1
        i = CuVector{TI}(undef, P); j = CuVector{TI}(undef, P); S = CuVector{SVec{TI}}(undef, P)
        P > 0 && _check(ccall((:nl_fill_pairs_window, libnlcuda), Cint,
                     (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Ptr{UInt8},
                      CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                     p, clist.X, nat, clist.perm, clist.cell_offsets, first, n_rows, index_map, pa, i, j, S, CU_NULL, ws, length(ws), _stream()))
    end
    return first[1:n_rows+1], i, j, S
end

# ---- AtomsBase extension with device positions: the IsolatedCell bounding box (ext/NeighbourListsAtomsBaseExt.jl:17-31)
function bounding_cell(X::CuVector{SVec{T}}) where {T}
    mm = CuVector{T}(undef, 6); ws = CUDA.zeros(UInt8, 32768)
    _check(ccall((:nl_bounding_box, libnlcuda), Cint, (Int32, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 _ftag(T), X, length(X), mm, ws, length(ws), _stream()))
    m = Array(mm)
    return SMat{T}(m[4] - m[1] + 1, 0, 0, 0, m[5] - m[2] + 1, 0, 0, 0, m[6] - m[3] + 1)
end

# ---- skin list: max squared displacement since the list was built
function max_displacement2(X::CuVector{SVec{T}}, Xref::CuVector{SVec{T}}) where {T}
    d2 = CuVector{T}(undef, 1); ws = CUDA.zeros(UInt8, 32768)
    _check(ccall((:nl_max_displacement2, libnlcuda), Cint,
                 (Int32, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 _ftag(T), X, Xref, length(X), d2, ws, length(ws), _stream()))
    return Array(d2)[1]
end

end # module