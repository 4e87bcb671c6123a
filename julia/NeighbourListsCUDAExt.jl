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
import AtomsBase, Unitful      # only the AtomsBase overloads at the end of this file need them

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
@assert sizeof(NlParams) == 192 "NlParams must match struct nl_params (include/nlcuda.h)"

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

# Scratch is uninitialised device memory (the library never reads what it has not written) and is CACHED per cell list:
# materialize_pairlist / the lazy sinks of one SortedCellList reuse one pair workspace instead of allocating 1.2 GB per call.
_ws_bytes(p, N, stage) = max(256, Int(ccall((:nl_workspace_bytes, libnlcuda), Csize_t, (Ref{NlParams}, Int64, Cint), p, N, stage)))
_ws(p, N, stage) = CuVector{UInt8}(undef, _ws_bytes(p, N, stage))
const _PAIR_WS = WeakKeyDict{Any,CuVector{UInt8}}()            # keyed by clist.perm (unique per SortedCellList)
function _pair_ws(clist, p)
    need = _ws_bytes(p, length(clist.X), 1)
    ws = get(_PAIR_WS, clist.perm, nothing)
    if ws === nothing || length(ws) < need
        ws = CuVector{UInt8}(undef, need)
        _PAIR_WS[clist.perm] = ws
    end
    return ws
end
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
    ws = _pair_ws(clist, p)
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
    out = CUDA.zeros(TI, nat); ws = _pair_ws(clist, p)
    _check(ccall((:nl_lazy_count, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, clist.X, nat, clist.perm, clist.cell_offsets, out, ws, length(ws), _stream()))
    return out
end

function lj_energy(clist::SortedCellList{T,TI,<:CuVector}, eps, sigma) where {T,TI}
    nat = length(clist.X); p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    e = CUDA.zeros(Float64, 1); ws = _pair_ws(clist, p)
    _check(ccall((:nl_lazy_lj_energy, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Float64, Float64, CuPtr{Float64}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, clist.X, nat, clist.perm, clist.cell_offsets, Float64(eps), Float64(sigma), e, ws, length(ws), _stream()))
    return e
end

"(F, e): per-atom LJ forces and energies from one fused traversal; fe is N x 4 (F_x, F_y, F_z, e)"
function lj_forces(clist::SortedCellList{T,TI,<:CuVector}, eps, sigma) where {T,TI}
    nat = length(clist.X); p = _params(clist.cell, clist.inv_cell, clist.pbc, clist.cutoff, clist.ncells)
    fe = CUDA.zeros(T, 4, nat); ws = _pair_ws(clist, p)
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
    f = Array(@view nl.first[i0:i0+1])                     # one 2-element copy instead of two scalar reads
    n1, n2 = Int(f[1]), Int(f[2]) - 1
    return (@view nl.j[n1:n2]), pairs_R(nl, n1, n2), (@view nl.S[n1:n2])
end

function NeighbourLists.maxneigs(nl::DevPairList)
    out = CUDA.zeros(Int64, 1)
    _check(ccall((:nl_max_neighbours, libnlcuda), Cint, (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Int64}, Ptr{Cvoid}),
                 _params(nl), nl.first, length(nl.first) - 1, out, _stream()))
    return Array(out)[1]
end

# ---- Array(PairList): the whole list into host memory (include/nlcuda.h: nl_pairs_to_host) --------------------
# 5 B/pair cross the bus (first, j, one byte per pair for S); i and S are rebuilt by host threads of the library while the copies
# run.  Pinned buffers and the transfer scratch are cached per (TI, capacity): pinning 5 GB costs more than building the list.
mutable struct HostPairBuffers{TI}
    cap::Int
    rows::Int
    first::Vector{TI}; i::Vector{TI}; j::Vector{TI}; S::Vector{SVec{TI}}
    hscratch::Vector{UInt8}; dscratch::CuVector{UInt8}
end
function HostPairBuffers{TI}(cap::Integer, rows::Integer) where {TI}
    nb = Int(ccall((:nl_to_host_scratch_bytes, libnlcuda), Csize_t, (Int64,), cap))
    pin(v) = (CUDA.Mem.pin(v); v)
    HostPairBuffers{TI}(cap, rows, pin(Vector{TI}(undef, rows + 1)), pin(Vector{TI}(undef, max(cap, 1))), pin(Vector{TI}(undef, max(cap, 1))),
                        pin(Vector{SVec{TI}}(undef, max(cap, 1))), pin(Vector{UInt8}(undef, nb)), CuVector{UInt8}(undef, nb))
end
const _HOST_BUF = Ref{Any}(nothing)

"PairList with every array in host memory (views of pinned buffers): `Array` of each field, in one compressed transfer"
function to_host(nl::DevPairList{T,TI}; nthreads::Integer = Threads.nthreads(), buffers = nothing) where {T,TI}
    P, rows = length(nl.i), length(nl.first) - 1
    b = buffers
    if b === nothing
        b = _HOST_BUF[]
        if !(b isa HostPairBuffers{TI}) || b.cap < P || b.rows < rows
            b = _HOST_BUF[] = HostPairBuffers{TI}(P + P ÷ 20 + 1024, rows)
        end
    end
    whole = rows == length(nl.X)                       # a whole list: i[p] is the row of p; shard lists copy their global i
    GC.@preserve b _check(ccall((:nl_pairs_to_host, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ptr{Cvoid},
                  Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}, Csize_t, Int32, Ptr{Cvoid}),
                 _params(nl), nl.first, rows, nl.i, whole ? P : 0, CU_NULL, C_NULL, nl.j, nl.S, P, b.first, b.i, b.j, b.S, b.dscratch, b.hscratch,
                 length(b.hscratch), nthreads, _stream()))     # shard lists: pass the row -> global index map instead (row_index), i_copy_from = P
    return PairList(Array(nl.X), nl.C, nl.cutoff, view(b.i, 1:P), view(b.j, 1:P), view(b.S, 1:P), view(b.first, 1:rows+1))
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
    first = CuVector{TI}(undef, nat + 1); ws = _pair_ws(clist, p); total = Ref{Int64}(0)
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
    mm = CuVector{T}(undef, 6); ws = CuVector{UInt8}(undef, 32768)
    _check(ccall((:nl_bounding_box, libnlcuda), Cint, (Int32, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 _ftag(T), X, length(X), mm, ws, length(ws), _stream()))
    m = Array(mm)
    return SMat{T}(m[4] - m[1] + 1, 0, 0, 0, m[5] - m[2] + 1, 0, 0, 0, m[6] - m[3] + 1)
end

# ---- skin list: max squared displacement since the list was built
function max_displacement2(X::CuVector{SVec{T}}, Xref::CuVector{SVec{T}}) where {T}
    d2 = CuVector{T}(undef, 1); ws = CuVector{UInt8}(undef, 32768)
    _check(ccall((:nl_max_displacement2, libnlcuda), Cint,
                 (Int32, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 _ftag(T), X, Xref, length(X), d2, ws, length(ws), _stream()))
    return Array(d2)[1]
end

# ---- AtomsBase systems with DEVICE positions (ext/NeighbourListsAtomsBaseExt.jl:50-140 always builds a CPU Vector, :54-56).
# These overloads take the positions as a CuVector (already stripped to `length_unit`), so nothing touches the host except the
# six numbers of the IsolatedCell bounding box; they mirror PairList / build_cell_list / neighbour_list of the extension.
# (Loaded only when AtomsBase and Unitful are: guard with Requires / a package extension in the real package.)
function _device_cell_matrix(ab, X::CuVector{SVec{T}}, length_unit) where {T}
    c = AtomsBase.cell(ab)
    if c isa AtomsBase.IsolatedCell{3}
        return bounding_cell(X)                                                   # :17-31 on the device (nl_bounding_box)
    elseif c isa AtomsBase.IsolatedCell
        error("NeighbourLists.jl does not support $(AtomsBase.n_dimensions(ab))-dimensional isolated AtomsBase systems yet.")
    end
    return SMat{T}(Unitful.ustrip.(length_unit, hcat(AtomsBase.cell_vectors(ab)...)'))  # :36
end

function NeighbourLists.build_cell_list(ab, X::CuVector{SVec{T}}, cutoff; length_unit = Unitful.unit(cutoff),
                                        int_type::Type = Int32) where {T}
    C = _device_cell_matrix(ab, X, length_unit)
    return NeighbourLists.build_cell_list(X, T(Unitful.ustrip(length_unit, cutoff)), C, AtomsBase.periodicity(ab); int_type = int_type)
end

function NeighbourLists.neighbour_list(ab, X::CuVector{SVec{T}}, cutoff; lazy::Bool = false, kwargs...) where {T}
    clist = NeighbourLists.build_cell_list(ab, X, cutoff; kwargs...)
    return lazy ? clist : NeighbourLists.materialize_pairlist(clist)
end

NeighbourLists.PairList(ab, X::CuVector{SVec{T}}, cutoff; kwargs...) where {T} = NeighbourLists.neighbour_list(ab, X, cutoff; kwargs...)

# ---- multi-GPU slabs through the library (include/nlcuda.h: nl_shard_prepare / nl_shard_exchange; NCCL inside libnlcuda.so).
# One Julia process per GPU.  `comm` is an ncclComm_t: NCCL.jl's `comm.handle`, or nccl_comm(...) below with the unique id
# shipped by MPI.jl / Distributed.
struct NlShardInfo
    axis::Int32; halo::Int32; periodic::Int32; nranks::Int32; rank::Int32; nplanes::Int32
    has_dn::Int32; has_up::Int32; dn_peer::Int32; up_peer::Int32
    n_local::Int64; n_owned::Int64; n_halo_dn::Int64; n_halo_up::Int64; n_send_dn::Int64; n_send_up::Int64
    bounds::NTuple{65,Int64}; send_count::NTuple{64,Int64}; recv_count::NTuple{64,Int64}
    n_max_all::Int64; src_offset::NTuple{64,Int64}; halo_src_offset_dn::Int64; halo_src_offset_up::Int64
end
@assert sizeof(NlShardInfo) == 2168 "NlShardInfo must match struct nl_shard_info (include/nlcuda.h)"

# Peer path (nl_shard_connect / nl_shard_exchange_peer): the ranks' workspaces mapped into each other (CUDA IPC), the all-to-all-v
# and the halos as NVLink copies.  One persistent workspace per communicator, laid out for the same `cap` on every rank.
mutable struct NlShardPeers
    nranks::Int32; rank::Int32; cap::Int64; ws_bytes::UInt64; ws::Ptr{Cvoid}
    peer_ws::NTuple{64,Ptr{Cvoid}}; peer_base::NTuple{64,Ptr{Cvoid}}
    NlShardPeers() = new(0, 0, 0, 0, C_NULL, ntuple(_ -> C_NULL, 64), ntuple(_ -> C_NULL, 64))
end
const _PEERS = Dict{Ptr{Cvoid},Tuple{NlShardPeers,CuVector{UInt8}}}()      # per communicator: (peers, workspace)
function _connect(p::NlParams, cap::Integer, comm::Ptr{Cvoid}, rank::Integer, nranks::Integer)
    nb = Int(ccall((:nl_shard_workspace_bytes, libnlcuda), Csize_t, (Ref{NlParams}, Int64, Int32), p, cap, nranks))
    ws = CuVector{UInt8}(undef, nb); peers = NlShardPeers()
    _check(ccall((:nl_shard_connect, libnlcuda), Cint,
                 (Ref{NlParams}, Int64, Ptr{Cvoid}, Int32, Int32, CuPtr{Cvoid}, Csize_t, Ref{NlShardPeers}, Ptr{Cvoid}),
                 p, cap, comm, rank, nranks, ws, nb, peers, _stream()))
    return _PEERS[comm] = (peers, ws)
end
function shard_disconnect(comm::Ptr{Cvoid})
    haskey(_PEERS, comm) || return
    peers, _ = pop!(_PEERS, comm)
    _check(ccall((:nl_shard_disconnect, libnlcuda), Cint, (Ref{NlShardPeers},), peers))
end

nccl_unique_id() = (id = zeros(UInt8, 128); _check(ccall((:nl_nccl_unique_id, libnlcuda), Cint, (Ptr{UInt8},), id)); id)
function nccl_comm(id::Vector{UInt8}, rank::Integer, nranks::Integer)
    c = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:nl_nccl_comm_init, libnlcuda), Cint, (Ref{Ptr{Cvoid}}, Int32, Ptr{UInt8}, Int32), c, nranks, id, rank))
    return c[]
end

"""
    sharded_pairlist(X, gidx, cutoff, cell, pbc, comm, rank, nranks) -> (owned, first, i, j, S)

Rows of the atoms this rank owns after the slab redistribution, with GLOBAL i / j.  X / gidx: any subset of the system per rank
(positions and global 1-based indices); rank is 0-based.
"""
function sharded_pairlist(X::CuVector{SVec{T}}, gidx::CuVector{TI}, cutoff::T, cell::SMat{T}, pbc::SVec{Bool},
                          comm::Ptr{Cvoid}, rank::Integer, nranks::Integer) where {T,TI}
    inv_cell, ncells, lens = analyze_cell(cell, cutoff, TI)
    p = _params(cell, inv_cell, pbc, cutoff, ncells)
    n = length(X)
    wsb(nmax) = max(256, Int(ccall((:nl_shard_workspace_bytes, libnlcuda), Csize_t, (Ref{NlParams}, Int64, Int32), p, nmax, nranks)))
    ws = CuVector{UInt8}(undef, wsb(n))
    info = Ref{NlShardInfo}()
    _check(ccall((:nl_shard_prepare, libnlcuda), Cint,
                 (Ref{NlParams}, CuPtr{Cvoid}, Int64, Ptr{Cvoid}, Int32, Int32, Ref{NlShardInfo}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, X, n, comm, rank, nranks, info, ws, length(ws), _stream()))
    f = info[]
    nall = f.n_owned + f.n_halo_dn + f.n_halo_up
    Xall = CuVector{SVec{T}}(undef, nall); gall = CuVector{TI}(undef, nall)
    plane_active = ones(UInt8, ncells[3])
    # f.n_max_all is the same number on every rank, so every rank (re)connects in the same call
    if !haskey(_PEERS, comm) || _PEERS[comm][1].cap < f.n_max_all
        shard_disconnect(comm)
        _connect(p, f.n_max_all + f.n_max_all ÷ 5 + 4096, comm, rank, nranks)
    end
    peers, pws = _PEERS[comm]
    _check(ccall((:nl_shard_exchange_peer, libnlcuda), Cint,
                 (Ref{NlParams}, Ref{NlShardInfo}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ptr{Cvoid}, Ref{NlShardPeers}, CuPtr{Cvoid}, CuPtr{Cvoid},
                  Ptr{UInt8}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                 p, info, X, gidx, n, comm, peers, Xall, gall, plane_active, pws, length(pws), _stream()))
    clist = NeighbourLists.build_cell_list(Xall, cutoff, cell, pbc; int_type = TI)
    first, i, j, S = shard_pairlist(clist, f.n_owned, gall, (nranks > 1 && f.axis == 2) ? plane_active : nothing)
    return (@view gall[1:f.n_owned]), first, i, j, S
end

end # module