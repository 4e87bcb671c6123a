/* nlcuda.h -- C ABI of libnlcuda.so, a B200 (sm_100a) neighbour-list engine.
 *
 * Drop-in boundary for the sort-based path of JuliaMolSim/NeighbourLists.jl behind
 *   neighbour_list(X, cutoff, cell, pbc; lazy)            (src/cell_list.jl:897-916)
 * for device-resident inputs.  Each entry point replaces one STAGE of that path; the Julia side
 * keeps PairList / SortedCellList (src/types.jl:34-82) and binds these with ccall.
 *
 * Conventions
 *  - All pointers except `params`, `total_pairs_host` are DEVICE pointers owned by the caller
 *    (CuArray / torch tensor).  The library never allocates or frees device memory and keeps no
 *    global state; scratch comes from the caller through `ws` (size from nl_workspace_bytes).
 *  - Work is enqueued on the caller's `stream` (a cudaStream_t / CUstream passed as void*).  Only
 *    nl_count_pairs synchronises (it must return the data-dependent pair count to the host),
 *    mirroring the reference's single D2H read (src/gpu_kernels.jl:333,367-371).
 *  - Element types are selected by params->float_type (positions, R) and params->int_type
 *    (perm, cell_id, cell_offsets, first, i, j, S).  Arrays are packed AoS exactly like Julia's
 *    Vector{SVector{3,T}}: X, X_sorted, R are N x 3 T; S is P x 3 TI.
 *  - All integer outputs are 1-BASED, as in the reference.
 *  - Matrices are 9 doubles in Julia column-major order, m[r + 3*c] = M[r+1, c+1]; ROWS of `cell`
 *    are the lattice vectors.  They hold values already rounded to T by the caller, which computes
 *    them with the reference's own analyze_cell (src/cell_list.jl:152-170) and nxyz formula
 *    (src/gpu_kernels.jl:315-316): the library never re-derives them.
 *  - Return value: NL_OK (0) or a negative NL_ERR_* code; nl_strerror() names it.  No exceptions,
 *    no abort.  N == 0 is legal everywhere.
 */
#ifndef NLCUDA_H
#define NLCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NL_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define NL_API __attribute__((visibility("default")))
#else
#define NL_API
#endif

enum { NL_F32 = 0, NL_F64 = 1 };
enum { NL_I32 = 0, NL_I64 = 1 };

enum {
  NL_OK = 0,
  NL_ERR_BAD_ARG = -1,     /* null pointer, negative N, bad enum, ncells < 1, nxyz < 1 ...            */
  NL_ERR_WORKSPACE = -2,   /* ws too small / misaligned, or fill called with a ws count did not stamp */
  NL_ERR_CUDA = -3,        /* a CUDA call failed; nl_last_cuda_error() holds the cudaError_t          */
  NL_ERR_OVERFLOW = -4,    /* pair count does not fit int_type (reference: silent Int32 wrap)         */
  NL_ERR_UNSUPPORTED = -5, /* N or prod(ncells) >= 2^31 - 1 (keys are 32-bit internally)              */
  NL_ERR_NCCL = -6         /* NCCL is not loadable in this process, or an NCCL call failed            */
};

/* Geometry of one neighbour-list problem, filled by the caller.
 * Mirrors the scalar fields of SortedCellList (src/types.jl:70-82) plus nxyz. */
typedef struct nl_params {
  int32_t float_type;  /* NL_F32 | NL_F64: element type T of X, X_sorted, R                         */
  int32_t int_type;    /* NL_I32 | NL_I64: element type TI of every integer array                   */
  double cell[9];      /* clist.cell                                                                */
  double inv_cell[9];  /* clist.inv_cell = inv(cell)                                                */
  double cutoff;       /* clist.cutoff (already converted to T, src/cell_list.jl:639)               */
  int32_t ncells[3];   /* clist.ncells = max(floor(lens / cutoff), 1)                               */
  int32_t nxyz[3];     /* ceil(cutoff * ncells / |lens|): stencil half-widths, >= 1                 */
  uint8_t pbc[3];      /* clist.pbc                                                                 */
  uint8_t reserved[5]; /* reserved[0]: NL_FLAG_* bits; the rest must be zero                        */
} nl_params;

/* params->reserved[0] flags.
 * NL_FLAG_HALF (nl_count_pairs / nl_fill_pairs only): HALF list -- of every mirror couple (i, j, S) / (j, i, -S) exactly
 * one pair is stored (which one is unspecified: the engine keeps the pair whose second atom comes later in cell-sorted
 * order; self images keep the lexicographically positive shift).  `first` then holds the half-list row sizes; P is exactly
 * half the full count.  Absent in the reference: halves the output traffic for consumers that use Newton's
 * third law.  The lazy sinks ignore the flag.                                                                          */
#define NL_FLAG_HALF 1

/* Workspace stages for nl_workspace_bytes. */
enum { NL_STAGE_BUILD = 0, NL_STAGE_PAIRS = 1 };

NL_API int nl_version(void);
NL_API const char* nl_strerror(int code);
NL_API int nl_last_cuda_error(void); /* cudaError_t captured by the last NL_ERR_CUDA on this thread */
NL_API long long nl_launch_count(void); /* kernels launched by this library in this process so far */

/* Bytes of scratch the given stage needs for N atoms (ws must be 256-byte aligned).
 * NL_STAGE_BUILD: nl_build_cells.  NL_STAGE_PAIRS: nl_count_pairs / nl_fill_pairs / nl_lazy_*. */
NL_API size_t nl_workspace_bytes(const nl_params* params, int64_t N, int stage);

/* Stage "build_cell_list": replaces _build_sorted_celllist's device stages
 * (src/cell_list.jl:647-679: _compute_cell_ids gpu_kernels.jl:244-255, _get_sortperm
 * cell_list.jl:711-718, the two gathers :669-670, _compute_cell_offsets gpu_kernels.jl:262-285).
 *   X            in   N x 3 T   caller's positions (never written; becomes clist.X_orig)
 *   X_sorted     out  N x 3 T   clist.X       = X[perm]
 *   perm         out  N TI      clist.perm    : sorted slot -> original index; the unique STABLE
 *                               permutation (equals the CPU path's sortperm, cell_list.jl:706-708)
 *   cell_id      out  N TI      clist.cell_id : linear cell of each sorted slot
 *   cell_offsets out  prod(ncells)+1 TI  clist.cell_offsets; all ones when N == 0               */
NL_API int nl_build_cells(const nl_params* params, const void* X, int64_t N, void* X_sorted, void* perm,
                   void* cell_id, void* cell_offsets, void* ws, size_t ws_bytes, void* stream);

/* Stage "materialize_pairlist", first half (src/gpu_kernels.jl:315-333: count_neighbours_kernel!,
 * compute_pair_offsets, _scalar_getindex).
 *   X_sorted, perm, cell_offsets   the SortedCellList fields (from nl_build_cells or the CPU path)
 *   first            out  N+1 TI   CSR offsets in ORIGINAL atom order, first[0] = 1
 *   total_pairs_host out  host int64: P = first[N] - 1.  The call synchronises `stream`.
 * Leaves per-atom records in `ws` for nl_fill_pairs: pass the SAME ws, untouched, to it.          */
NL_API int nl_count_pairs(const nl_params* params, const void* X_sorted, int64_t N, const void* perm,
                   const void* cell_offsets, void* first, int64_t* total_pairs_host, void* ws,
                   size_t ws_bytes, void* stream);

/* Stage "materialize_pairlist", second half (src/gpu_kernels.jl:347-359: fill_pairs_kernel!).
 *   i_out, j_out  out  P TI      pair (i[n], j[n]); row m occupies first[m]..first[m+1]-1
 *   S_out         out  P x 3 TI  cell shifts, packed
 *   R_out         out  P x 3 T   X[j] - X[i] + cell' * S (what _getR recomputes,
 *                                src/cell_list.jl:525-531); MAY BE NULL (PairList stores no R)
 * Order inside a row is unspecified (the reference's tests sort before comparing).               */
NL_API int nl_fill_pairs(const nl_params* params, const void* X_sorted, int64_t N, const void* perm,
                  const void* cell_offsets, const void* first, void* i_out, void* j_out, void* S_out,
                  void* R_out, void* ws, size_t ws_bytes, void* stream);

/* nl_fill_pairs for a SHARD (multi-GPU slabs): local atoms are ordered owned-first, only the
 * first n_rows of them (the owned atoms) get rows, and i/j are written as index_map[local index - 1]
 * (the atoms' GLOBAL 1-based indices; N TI, may be NULL for local indices).  `first` comes from
 * nl_count_pairs on the same local set; the owned rows occupy first[0] .. first[n_rows]-1. */
NL_API int nl_fill_pairs_rows(const nl_params* params, const void* X_sorted, int64_t N, const void* perm,
                       const void* cell_offsets, const void* first, int64_t n_rows, const void* index_map,
                       void* i_out, void* j_out, void* S_out, void* R_out, void* ws, size_t ws_bytes, void* stream);

/* nl_count_pairs / nl_fill_pairs_rows for a slab shard whose slabs are cut along z (the slowest key axis): plane_active is a
 * HOST array of ncells[2] bytes; plane z may hold atoms of the local set (owned + halo) iff plane_active[z] != 0.  The caller
 * PROMISES that every other plane is empty; the library then launches only the tile layers that can hold atoms instead of the
 * whole global grid (1.1 ms per list at 8 ranks).  NULL = no promise (identical to the plain entry points).                 */
NL_API int nl_count_pairs_window(const nl_params* params, const void* X_sorted, int64_t N, const void* perm,
                                 const void* cell_offsets, void* first, int64_t* total_pairs_host,
                                 const uint8_t* plane_active, void* ws, size_t ws_bytes, void* stream);
NL_API int nl_fill_pairs_window(const nl_params* params, const void* X_sorted, int64_t N, const void* perm,
                                const void* cell_offsets, const void* first, int64_t n_rows, const void* index_map,
                                const uint8_t* plane_active, void* i_out, void* j_out, void* S_out, void* R_out, void* ws,
                                size_t ws_bytes, void* stream);

/* Stage (1) alone: cell_id_out[n] (N TI) = 1-based linear cell of atom n in the caller's order
 * (replaces _compute_cell_ids, src/gpu_kernels.jl:244-255,378).  The slab sharding bins atoms with it. */
NL_API int nl_cell_ids(const nl_params* params, const void* X, int64_t N, void* cell_id_out, void* stream);

/* Slab plan for multi-GPU sharding (host-only, no CUDA call): cuts the `nplanes` cell planes of the slab axis into
 * `nranks` slabs balanced by atom count.  plane_hist[nplanes] = atoms per plane (already summed over ranks);
 * halo = nxyz of the slab axis; every slab gets at least 2*halo+1 planes (1 if nranks == 1).
 * bounds_out[nranks+1]: rank r owns planes [bounds[r], bounds[r+1]).  NL_ERR_BAD_ARG if the planes do not suffice. */
NL_API int nl_shard_plan(const int64_t* plane_hist, int32_t nplanes, int32_t nranks, int32_t halo, int64_t* bounds_out);

/* ---- Multi-GPU: 1-D slabs of whole cell planes with a cutoff-wide halo.  One process per GPU; the reference has
 * no multi-GPU code.  The slab axis is the axis with the most cells (ties: z, the slowest key axis) of the reference's OWN
 * cell grid (src/cell_list.jl:83-86, widths :94-95), so the unchanged single-GPU stages run on every rank's local set with the
 * GLOBAL params.  Call sequence per list (all on `stream`, the current device being this rank's GPU):
 *     nl_shard_prepare   -> info (host): slab bounds, n_owned, halo sizes, every send / receive count.  ONE host read.
 *     caller allocates X_all, gidx_all with info.n_owned + info.n_halo_dn + info.n_halo_up rows
 *     nl_shard_exchange  -> X_all / gidx_all = [owned atoms | halo from the rank below | halo from the rank above]
 *     nl_build_cells(X_all) ; nl_count_pairs_window ; nl_fill_pairs_window(n_rows = info.n_owned, index_map = gidx_all)
 * Each rank then holds the CSR rows of its owned atoms with GLOBAL i / j and global shifts S; concatenating the ranks' rows by
 * global i reproduces the single-device list.  `comm` is an ncclComm_t of the NCCL instance loaded in this process (bound with
 * dlopen at first use: libnlcuda.so itself does not link NCCL); nl_nccl_* below create one when the host has none. */
#define NL_MAX_RANKS 64
typedef struct nl_shard_info {
  int32_t axis, halo, periodic, nranks, rank, nplanes; /* slab axis, halo width in planes (= nxyz[axis]), pbc[axis], ... */
  int32_t has_dn, has_up, dn_peer, up_peer;             /* neighbours along the slab axis (ring if periodic)             */
  int64_t n_local;                /* atoms this rank passed in                                                            */
  int64_t n_owned;                /* atoms of its slab after the redistribution                                           */
  int64_t n_halo_dn, n_halo_up;   /* halo atoms it receives from the rank below / above                                   */
  int64_t n_send_dn, n_send_up;   /* owned atoms it sends as halo to the rank below / above                               */
  int64_t bounds[NL_MAX_RANKS + 1];    /* rank r owns planes [bounds[r], bounds[r+1])                                     */
  int64_t send_count[NL_MAX_RANKS];    /* local atoms owned by rank d (send_count[rank]: atoms that stay)                 */
  int64_t recv_count[NL_MAX_RANKS];    /* atoms rank s holds that this rank owns                                          */
  /* what the PEER path (nl_shard_connect / nl_shard_exchange_peer) needs to know about the other ranks; all of it follows
   * from the all-gathered histograms, so every rank computes the same numbers:                                           */
  int64_t n_max_all;                   /* max over ranks of max(atoms passed in, atoms owned): the workspace capacity needed */
  int64_t src_offset[NL_MAX_RANKS];    /* where, in rank s's send buffer (atoms), the block for THIS rank starts          */
  int64_t halo_src_offset_dn, halo_src_offset_up; /* where, in the dn / up peer's halo buffer, this rank's halo starts   */
} nl_shard_info;

/* Scratch for nl_shard_prepare / nl_shard_exchange with at most n_max atoms on either side of the redistribution
 * (n_max >= max(n_local, info.n_owned)). */
NL_API size_t nl_shard_workspace_bytes(const nl_params* params, int64_t n_max, int32_t nranks);

/* Bins the n local atoms (X: n x 3 T, ANY subset of the system per rank) to cell planes of the slab axis, all-gathers the
 * per-rank plane histograms and plans the slabs (nl_shard_plan: balanced by atom count, >= 2 * halo + 1 planes each).
 * Synchronises `stream` once.  NL_ERR_BAD_ARG when the planes do not suffice for nranks slabs (use fewer ranks / replicas). */
NL_API int nl_shard_prepare(const nl_params* params, const void* X, int64_t n, void* comm, int32_t rank, int32_t nranks,
                            nl_shard_info* info_out, void* ws, size_t ws_bytes, void* stream);

/* Moves every atom (position + global 1-based index gidx: n TI) to its owner and exchanges the halos; fully asynchronous.
 * plane_active_out (HOST, ncells[2] bytes, may be NULL) receives the plane window for nl_*_window when the slab axis is z
 * (all ones otherwise). */
NL_API int nl_shard_exchange(const nl_params* params, const nl_shard_info* info, const void* X, const void* gidx, int64_t n,
                             void* comm, void* X_all, void* gidx_all, uint8_t* plane_active_out, void* ws, size_t ws_bytes,
                             void* stream);

/* ---- Peer path: the same exchange as direct NVLink COPIES between the ranks' workspaces instead of ncclSend / ncclRecv.
 * Every rank maps the other ranks' workspace into its address space once (CUDA IPC; one process per GPU on one node) and then
 * PULLS its blocks of the all-to-all-v and its halos straight from the peers' send buffers into X_all / gidx_all with
 * cudaMemcpyAsync (copy engines over NVLink / NVSwitch); the only collectives left per list are two 8-byte all-gathers used as
 * stream-ordered barriers ("every send buffer is ready").  Measured on 2 B200: all-to-all-v of 2 x 160 MB 1.09 ms with
 * ncclSend / ncclRecv, see DESIGN.md 6 for the peer path.
 *   nl_shard_connect     collective.  `ws` is THE workspace of this rank for every later exchange: nl_shard_workspace_bytes(params,
 *                        cap, nranks) bytes for the SAME cap on every rank (the layout must agree), alive until nl_shard_disconnect.
 *   nl_shard_exchange_peer  like nl_shard_exchange; NL_ERR_WORKSPACE when info->n_max_all > peers->cap (on every rank alike:
 *                        reconnect with a larger workspace) or when ws is not the connected one.
 *   nl_shard_disconnect  unmaps the peers' workspaces.                                                                      */
typedef struct nl_shard_peers {
  int32_t nranks, rank;
  int64_t cap;                   /* atoms the workspace layout is made for                                               */
  uint64_t ws_bytes;
  void* ws;                      /* this rank's workspace                                                                */
  void* peer_ws[NL_MAX_RANKS];   /* rank r's workspace in this process's address space (NULL for r == rank)             */
  void* peer_base[NL_MAX_RANKS]; /* what cudaIpcOpenMemHandle returned (for nl_shard_disconnect)                         */
} nl_shard_peers;
NL_API int nl_shard_connect(const nl_params* params, int64_t cap, void* comm, int32_t rank, int32_t nranks, void* ws,
                            size_t ws_bytes, nl_shard_peers* peers_out, void* stream);
NL_API int nl_shard_exchange_peer(const nl_params* params, const nl_shard_info* info, const void* X, const void* gidx,
                                  int64_t n, void* comm, const nl_shard_peers* peers, void* X_all, void* gidx_all,
                                  uint8_t* plane_active_out, void* ws, size_t ws_bytes, void* stream);
NL_API int nl_shard_disconnect(nl_shard_peers* peers);

/* Communicator helpers for hosts without an NCCL binding of their own: rank 0 calls nl_nccl_unique_id, ships the 128 bytes to
 * the other ranks by any means, then every rank calls nl_nccl_comm_init (collective). */
NL_API int nl_nccl_unique_id(void* id128_out);
NL_API int nl_nccl_comm_init(void** comm_out, int32_t nranks, const void* id128, int32_t rank);
NL_API int nl_nccl_comm_destroy(void* comm);

/* Lazy mode: fused for_each_neighbour traversals (src/cell_list.jl:779-801) with fixed sinks.
 * nl_lazy_count: counts_out[m] (N TI, original order) = count_neighbours(clist, m) (:808-814).
 * nl_lazy_lj_energy: *energy_out (DEVICE double) = sum over ordered pairs of
 *   4 eps ((sigma/r)^12 - (sigma/r)^6), r^2 = dot(R,R) evaluated in T, accumulated in double.      */
NL_API int nl_lazy_count(const nl_params* params, const void* X_sorted, int64_t N, const void* perm,
                  const void* cell_offsets, void* counts_out, void* ws, size_t ws_bytes, void* stream);
NL_API int nl_lazy_lj_energy(const nl_params* params, const void* X_sorted, int64_t N, const void* perm,
                      const void* cell_offsets, double eps, double sigma, double* energy_out, void* ws,
                      size_t ws_bytes, void* stream);

/* nl_lazy_lj_forces: the fused traversal with a force sink.  fe_out (DEVICE, N x 4 T, 16-byte aligned, ORIGINAL atom order,
 * overwritten) = (F_x, F_y, F_z, e) per atom with
 *   e_n = sum over n's neighbours of phi(r),  phi = 4 eps ((sigma/r)^12 - (sigma/r)^6)   (sum_n e_n = nl_lazy_lj_energy)
 *   F_n = -dE/dx_n for E = 1/2 sum_n e_n     = sum over ordered pairs (m, n) of 24 eps (2 (sigma/r)^12 - (sigma/r)^6) / r^2 * R_mn
 * Sums are accumulated in T with atomics (order not deterministic: compare to a tolerance).  Float32 runs on the packed
 * counting kernel with per-tile shared-memory accumulators; Float64 on the exact tiled traversal.                     */
NL_API int nl_lazy_lj_forces(const nl_params* params, const void* X_sorted, int64_t N, const void* perm,
                             const void* cell_offsets, double eps, double sigma, void* fe_out, void* ws,
                             size_t ws_bytes, void* stream);

/* ---- The callers either side of the hot path: device-side PairList accessors, the IsolatedCell
 * bounding box of the AtomsBase adapter, and the displacement check of a skin (Verlet) list. ------------------- */

/* _getR for a range of pairs without scalar indexing (src/cell_list.jl:525-531 as looped by neigss!, :583-592):
 *   R_out[p - p_lo] = (X[j[p]] - X[i[p]]) + cell' * S[p]   for p in [p_lo, p_hi)   (0-based, half-open)
 * X is the caller's ORIGINAL-order position array (PairList.X), N atoms; i, j, S are the PairList arrays.
 * Same expression and association as the fill pass, so R is bit-identical to nl_fill_pairs' R_out.  Passing NEW
 * positions refreshes R for an unchanged pair topology (skin list).                                            */
NL_API int nl_pairs_R(const nl_params* params, const void* X, int64_t N, const void* i, const void* j, const void* S,
                      int64_t p_lo, int64_t p_hi, void* R_out, void* stream);

/* maxneigs(nlist) (src/cell_list.jl:513) as a device reduction: *max_out (DEVICE int64) = max_n first[n+1]-first[n].
 * N == 0 is NL_ERR_BAD_ARG (the reference's maximum over an empty collection throws).                          */
NL_API int nl_max_neighbours(const nl_params* params, const void* first, int64_t N, int64_t* max_out, void* stream);

/* Neighbourhoods of a SET of atoms at once (the sites() loop, src/iterators.jl:27-40, over neigss!,
 * src/cell_list.jl:583-592) as fixed-width padded blocks:
 *   rows   in   n_sel TI            atoms (1-based)
 *   n_out  out  n_sel TI            nneigs(nlist, rows[s]) (src/cell_list.jl:523) -- the FULL count even if > width
 *   j_out  out  n_sel x width TI    neighbours, 0 in the padding
 *   S_out  out  n_sel x width x 3 TI   (may be NULL)
 *   R_out  out  n_sel x width x 3 T    (may be NULL) computed like nl_pairs_R                                */
NL_API int nl_rows_padded(const nl_params* params, const void* X, int64_t N, const void* first, const void* j,
                          const void* S, const void* rows, int64_t n_sel, int32_t width, void* n_out, void* j_out,
                          void* S_out, void* R_out, void* stream);

/* neighbours(clist, i) for a set of atoms straight from the cell list (src/cell_list.jl:821-833 over for_each_neighbour,
 * :779-801): no pair list is materialised.  X_orig = clist.X_orig (the caller's order), X_sorted / perm / cell_offsets the
 * other SortedCellList fields; atoms = n_sel TI (1-based).  Rows come out in the reference's own traversal order (dz, dy, dx,
 * then sorted slot).  Outputs exactly as nl_rows_padded (n_out = full count, blocks truncated to `width`, zero padding).   */
NL_API int nl_lazy_neighbours(const nl_params* params, const void* X_orig, const void* X_sorted, int64_t N, const void* perm,
                              const void* cell_offsets, const void* atoms, int64_t n_sel, int32_t width, void* n_out,
                              void* j_out, void* S_out, void* R_out, void* stream);

/* Scratch for the two reductions below (256-byte aligned device memory). */
#define NL_REDUCE_WS_BYTES 32768

/* Bounding box of the positions: minmax_out (DEVICE, 6 T) = (min x, min y, min z, max x, max y, max z).  The
 * IsolatedCell branch of _get_cell_matrix (ext/NeighbourListsAtomsBaseExt.jl:17-31) builds its cell as
 * diag(max - min + 1) from these.  N >= 1.                                                                     */
NL_API int nl_bounding_box(int32_t float_type, const void* X, int64_t N, void* minmax_out, void* ws, size_t ws_bytes,
                           void* stream);

/* d2_out (DEVICE, 1 T) = max_n |X[n] - X_ref[n]|^2, evaluated in T as (dx dx + dy dy) + dz dz.  A list built with
 * cutoff + skin stays complete for `cutoff` while sqrt(d2) < skin / 2 (absent in the reference, which rebuilds on
 * every call, src/cell_list.jl:906-916).                                                                       */
NL_API int nl_max_displacement2(int32_t float_type, const void* X, const void* X_ref, int64_t N, void* d2_out, void* ws,
                                size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Array(PairList): the whole materialised list into HOST memory -- what a host consumer of neighbour_list receives
 * (the reference's CPU path returns host Vectors, src/cell_list.jl:897-916; its GPU tests bring the device list back with
 * Array(...) field by field, test/test_utils.jl:127-131).  Transfer format: over the bus go `first` (n_rows + 1 TI), `j`
 * (P TI) and ONE BYTE per pair for S, code = (Sx+1) + 3 (Sy+1) + 9 (Sz+1) packed on the device; `i` is rebuilt from `first`
 * (i[p] = r for first[r] <= p+1 < first[r+1]) and S from the codes by `nthreads` host threads of the library (non-temporal
 * stores) while the copies are in flight: 5 B/pair instead of 20 B/pair of PCIe traffic.  A list with a shift component
 * outside {-1, 0, 1} sends S as it is (detected on the device, no caller involvement).
 *   first, j, S     in   DEVICE  the PairList arrays (S must be 16-byte aligned for the fast path; any alignment works)
 *   i, i_copy_from  in   DEVICE  pairs [0, i_copy_from) of i are rebuilt from first (a whole list: i[p] = row of p), pairs
 *                                [i_copy_from, P) are copied from the device array `i` (NULL allowed when i_copy_from == P).
 *                                Shard lists, whose i runs through an index map, pass 0.  A whole list may split i between
 *                                the host threads and the bus; i_copy_from == P was fastest where measured.
 *   row_index       in   DEVICE  n_rows TI or NULL: the value i takes for the pairs of row r (a shard's rows carry GLOBAL atom
 *                                indices: pass its index map and i_copy_from = P, and i never crosses the bus); NULL: r + 1.
 *                                row_index_host: n_rows TI of host memory for its copy (pinned; may be NULL with row_index)
 *   *_host          out  HOST    n_rows + 1, P, P, 3 P elements of TI; pinned memory for full PCIe speed
 *   dev_scratch / host_scratch   nl_to_host_scratch_bytes(P) bytes each, 16-byte aligned; host_scratch pinned
 *   nthreads        host worker threads (<= 0: half the hardware concurrency -- the decoders are bound by memory bandwidth)
 * Blocks until every host array is complete (it is a device -> host read); enqueues on `stream`.
 * NL_ERR_BAD_ARG if first[n_rows] - 1 != P.                                                                        */
NL_API size_t nl_to_host_scratch_bytes(int64_t P);
NL_API int nl_pairs_to_host(const nl_params* params, const void* first, int64_t n_rows, const void* i, int64_t i_copy_from,
                            const void* row_index, void* row_index_host, const void* j, const void* S, int64_t P, void* first_host, void* i_host, void* j_host, void* S_host,
                            void* dev_scratch, void* host_scratch, size_t scratch_bytes, int32_t nthreads, void* stream);
/* The same transfer in two steps, so that the `first` copy and the rebuild of i (a third of the host-side work) run WHILE the fill pass
 * is still on the GPU:
 *     nl_count_pairs(...)                       -> P
 *     nl_pairs_to_host_begin(first, n_rows, P, first_host, i_host, nthreads, stream, &job)   returns at once; own stream (ordered
 *                                                  after what `stream` holds at the time of the call) + host threads
 *     nl_fill_pairs(..., stream)
 *     nl_pairs_to_host_finish(job, j, S, j_host, S_host, scratch..., stream)          blocks until every host array is complete
 * For whole lists only (i is rebuilt from first).  Every job must be finished exactly once (finish releases it, also on error). */
typedef struct nl_to_host_job nl_to_host_job;
NL_API int nl_pairs_to_host_begin(const nl_params* params, const void* first, int64_t n_rows, int64_t P, void* first_host,
                                  void* i_host, int32_t nthreads, void* stream, nl_to_host_job** job_out);
NL_API int nl_pairs_to_host_finish(nl_to_host_job* job, const void* j, const void* S, void* j_host, void* S_host,
                                   void* dev_scratch, void* host_scratch, size_t scratch_bytes, void* stream);
/* The two host-side decoders of that format on their own (pure host code, no CUDA call): pairs [p_lo, p_hi), 0-based.  */
NL_API int nl_host_expand_rows(int32_t int_type, const void* first, const void* row_index, int64_t n_rows, int64_t p_lo, int64_t p_hi,
                               void* i_out);   /* row_index: HOST n_rows TI or NULL */
NL_API int nl_host_unpack_shifts(int32_t int_type, const uint8_t* codes, int64_t p_lo, int64_t p_hi, void* S_out);

#ifdef __cplusplus
}
#endif
#endif /* NLCUDA_H */
