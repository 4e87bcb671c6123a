// nl_build.cuh -- stage kernels of build_cell_list: cell binning, gather of the sorted copies and
// cell_offsets extraction.  The sort between them is nl_scan_sort.cuh.
#pragma once
#include "nl_common.cuh"

namespace nl {

// Stage (1): compute_cell_ids_kernel! (src/gpu_kernels.jl:108-125).  One thread per atom in the
// caller's order; key = 0-based linear cell id (x fastest, src/cell_list.jl:83-86).
template <class T>
__global__ void __launch_bounds__(256) k_bin(const T* __restrict__ X, long long n, Geo<T> g, uint32_t* __restrict__ keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
  int c[3];
  long long w[3];
  cell_of(g, x, y, z, c, w);
  keys[i] = (uint32_t)c[0] + (uint32_t)g.nc[0] * ((uint32_t)c[1] + (uint32_t)g.nc[1] * (uint32_t)c[2]);
}

// ---- bucket build (round 2): when the cell grid is not much larger than the atom count, the stable sort by cell id is a
// counting sort whose counters live in L2 (prod(ncells) x 4 B: 6.4 MB at the headline size):
//   k_bin_count        key = cell id; rank = arrival order inside the cell (atomicAdd with return on the cell's counter)
//   exclusive scan     of the counters IS cell_offsets (1-based) -- no lower-bound pass over the sorted keys
//   k_bucket_scatter   tmp[cell_offsets[key] - 1 + rank] = atom: every cell's atoms are now contiguous, in arrival order
//   k_finalize_buckets the place of an atom inside its cell = number of atoms of the cell with a smaller original index
//                      (a scan of the cell's ~6 entries, L2-resident), which is exactly the stable sortperm of the
//                      reference (src/cell_list.jl:711-718); perm, cell_id and X_sorted are written from there.
// 7 launches instead of 15 and one pass over the keys instead of three (radix: 0.44 ms + 0.05 ms of cell_offsets).
template <class T>
__global__ void __launch_bounds__(256) k_bin_count(const T* __restrict__ X, long long n, Geo<T> g, uint32_t* __restrict__ keys,
                                                   uint32_t* __restrict__ rank, uint32_t* __restrict__ cnt) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
  int c[3];
  long long w[3];
  cell_of(g, x, y, z, c, w);
  const uint32_t key = (uint32_t)c[0] + (uint32_t)g.nc[0] * ((uint32_t)c[1] + (uint32_t)g.nc[1] * (uint32_t)c[2]);
  keys[i] = key;
  rank[i] = atomicAdd(&cnt[key], 1u);
}
template <class TI>
__global__ void __launch_bounds__(256) k_bucket_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank, const TI* __restrict__ co,
                                                        long long n, uint32_t* __restrict__ tmp) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tmp[(long long)co[keys[i]] - 1 + (long long)rank[i]] = (uint32_t)i;
}
template <class T, class TI>
__global__ void __launch_bounds__(256) k_finalize_buckets(const uint32_t* __restrict__ tmp, const uint32_t* __restrict__ keys, const TI* __restrict__ co,
                                                          const T* __restrict__ X, long long n, T* __restrict__ Xs, TI* __restrict__ perm,
                                                          TI* __restrict__ cell_id) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t o = tmp[s];
  const uint32_t c = keys[o];
  const long long c0 = (long long)co[c] - 1, c1 = (long long)co[c + 1] - 1;
  const T x = X[3ll * o], y = X[3ll * o + 1], z = X[3ll * o + 2];
  long long d = c0;
  for (long long k = c0; k < c1; k++) d += tmp[k] < o ? 1 : 0;
  perm[d] = (TI)o + 1;
  cell_id[d] = (TI)c + 1;
  Xs[3 * d] = x;
  Xs[3 * d + 1] = y;
  Xs[3 * d + 2] = z;
}

// The same binning with the reference's output convention: 1-based linear cell id per atom in the caller's
// order (_compute_cell_ids, src/gpu_kernels.jl:244-255,378).  Used by the slab sharding.
template <class T, class TI>
__global__ void __launch_bounds__(256) k_cell_ids(const T* __restrict__ X, long long n, Geo<T> g, TI* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c[3];
  long long w[3];
  cell_of(g, X[3 * i], X[3 * i + 1], X[3 * i + 2], c, w);
  out[i] = (TI)c[0] + (TI)g.nc[0] * ((TI)c[1] + (TI)g.nc[1] * (TI)c[2]) + 1;
}

// After the sort: perm, cell_id (1-based, TI) and X_sorted = X[perm] (src/cell_list.jl:669-670).
template <class T, class TI>
__global__ void __launch_bounds__(256) k_finalize_sorted(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ svals,
                                                         const T* __restrict__ X, long long n, T* __restrict__ Xs,
                                                         TI* __restrict__ perm, TI* __restrict__ cell_id) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t o = svals[s];
  perm[s] = (TI)o + 1;
  cell_id[s] = (TI)skeys[s] + 1;
  T x = X[3ll * o], y = X[3ll * o + 1], z = X[3ll * o + 2];
  Xs[3 * s] = x;
  Xs[3 * s + 1] = y;
  Xs[3 * s + 2] = z;
}

// Stage (3): cell_offsets[c] = 1 + (number of sorted keys < c), c in [0, nct]; no atomics, no scan
// (replaces _histogram_kernel! + accumulate!, src/gpu_kernels.jl:191-197,262-285).  n == 0 gives all
// ones (src/gpu_kernels.jl:268-271).
template <class TI>
__global__ void __launch_bounds__(256) k_cell_offsets(const uint32_t* __restrict__ skeys, long long n, long long nct, TI* __restrict__ co) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nct) return;
  long long lo = 0, hi = n;  // lower_bound(skeys, c)
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if ((long long)skeys[mid] < c) lo = mid + 1; else hi = mid;
  }
  co[c] = (TI)(lo + 1);
}

}  // namespace nl
