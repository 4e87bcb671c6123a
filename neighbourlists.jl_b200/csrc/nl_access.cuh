// nl_access.cuh -- the callers either side of the hot path: PairList accessors without scalar
// indexing, the IsolatedCell bounding box and the displacement check of a skin (Verlet) list.  All of it is
// bandwidth-bound streaming / gather work: one pass, no intermediate arrays.
#pragma once
#include "nl_common.cuh"

namespace nl {

// ---------------------------------------------------------------------------------------------------------
// _getR for a whole range of pairs (src/cell_list.jl:525-531 as looped by neigss!, :583-592):
//   R[p] = (X[j[p]] - X[i[p]]) + cell' * S[p],  X in the caller's ORIGINAL order, i/j 1-based.
// One thread per pair; i/j/S are unit-stride, X[i] is row-constant (cache hit), X[j] one 24/12-byte gather.
// R is transposed through shared memory so that the stores are unit-stride as well.
template <class T, class TI>
__global__ void __launch_bounds__(256) k_pairs_R(const T* __restrict__ X, const TI* __restrict__ ia, const TI* __restrict__ ja,
                                                 const TI* __restrict__ Sa, long long p_lo, long long np, Geo<T> g, T* __restrict__ R) {
  __shared__ T st[256 * 3];
  const long long b0 = (long long)blockIdx.x * 256;
  const long long q = b0 + threadIdx.x;
  if (q < np) {
    const long long p = p_lo + q;
    const long long i = (long long)ia[p] - 1, j = (long long)ja[p] - 1;
    const T s0 = (T)Sa[3 * p], s1 = (T)Sa[3 * p + 1], s2 = (T)Sa[3 * p + 2];
    T c0, c1, c2;
    mtv(g.cell, s0, s1, s2, c0, c1, c2);
    st[3 * threadIdx.x] = add_rn(sub_rn(X[3 * j], X[3 * i]), c0);
    st[3 * threadIdx.x + 1] = add_rn(sub_rn(X[3 * j + 1], X[3 * i + 1]), c1);
    st[3 * threadIdx.x + 2] = add_rn(sub_rn(X[3 * j + 2], X[3 * i + 2]), c2);
  }
  __syncthreads();
  const long long nw = 3 * min((long long)256, np - b0);
  T* const out = R + 3 * b0;
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const int w = m * 256 + threadIdx.x;
    if (w < nw) out[w] = st[w];
  }
}

// maxneigs (src/cell_list.jl:513): max over rows of first[n+1] - first[n].  out must be zeroed by the caller.
template <class TI>
__global__ void __launch_bounds__(256) k_max_neighbours(const TI* __restrict__ first, long long n, unsigned long long* __restrict__ out) {
  unsigned long long m = 0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const unsigned long long d = (unsigned long long)(first[r + 1] - first[r]);
    m = d > m ? d : m;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  __shared__ unsigned long long sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) m = sm[k] > m ? sm[k] : m;
    if (m) atomicMax(out, m);
  }
}

// Batched neighbourhoods (the sites() loop of src/iterators.jl:27-40 / neigss!, src/cell_list.jl:583-592, for a
// SET of atoms at once): atom rows[s] (1-based) -> fixed-width blocks
//   n_out[s] = nneigs;  j_out[s*width + k], S_out[(s*width + k)*3 ..], R_out likewise for k < min(nneigs, width);
//   the padding k >= nneigs is j = 0, S = 0, R = 0.  One warp per selected atom.
template <class T, class TI>
__global__ void __launch_bounds__(256) k_rows_padded(const T* __restrict__ X, const TI* __restrict__ first, const TI* __restrict__ ja,
                                                     const TI* __restrict__ Sa, const TI* __restrict__ rows, long long n_sel, int width,
                                                     Geo<T> g, TI* __restrict__ n_out, TI* __restrict__ j_out, TI* __restrict__ S_out,
                                                     T* __restrict__ R_out) {
  const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= n_sel) return;
  const long long i = (long long)rows[s] - 1;
  const long long p0 = (long long)first[i] - 1;
  const long long nn = (long long)first[i + 1] - 1 - p0;
  if (lane == 0) n_out[s] = (TI)nn;
  const T xi = X[3 * i], yi = X[3 * i + 1], zi = X[3 * i + 2];
  const long long o0 = s * (long long)width;
  for (int k = lane; k < width; k += 32) {
    TI j = 0, S0 = 0, S1 = 0, S2 = 0;
    T R0 = 0, R1 = 0, R2 = 0;
    if (k < nn) {
      const long long p = p0 + k;
      j = ja[p];
      S0 = Sa[3 * p]; S1 = Sa[3 * p + 1]; S2 = Sa[3 * p + 2];
      if (R_out) {
        const long long jj = (long long)j - 1;
        T c0, c1, c2;
        mtv(g.cell, (T)S0, (T)S1, (T)S2, c0, c1, c2);
        R0 = add_rn(sub_rn(X[3 * jj], xi), c0);
        R1 = add_rn(sub_rn(X[3 * jj + 1], yi), c1);
        R2 = add_rn(sub_rn(X[3 * jj + 2], zi), c2);
      }
    }
    j_out[o0 + k] = j;
    if (S_out) { S_out[3 * (o0 + k)] = S0; S_out[3 * (o0 + k) + 1] = S1; S_out[3 * (o0 + k) + 2] = S2; }
    if (R_out) { R_out[3 * (o0 + k)] = R0; R_out[3 * (o0 + k) + 1] = R1; R_out[3 * (o0 + k) + 2] = R2; }
  }
}

// ---------------------------------------------------------------------------------------------------------
// neighbours(clist, i) for a SET of atoms straight from the cell list (src/cell_list.jl:821-833 over
// for_each_neighbour, :779-801 -> _for_each_neighbor_pair, src/gpu_kernels.jl:58-101): nothing is materialised.
// One warp per atom walks the stencil in the reference's own order (dz outermost, dx innermost, then sorted slot):
// lanes take 32 consecutive slots of a cell, hits are compacted with ballot/popc, so row CONTENT AND ORDER equal
// the reference's.  Output blocks as in k_rows_padded.
template <class T, class TI>
__global__ void __launch_bounds__(256) k_lazy_neighbours(const T* __restrict__ Xo, const T* __restrict__ Xs, const TI* __restrict__ perm,
                                                         const TI* __restrict__ co, Geo<T> g, const TI* __restrict__ atoms, long long n_sel,
                                                         int width, TI* __restrict__ n_out, TI* __restrict__ j_out, TI* __restrict__ S_out,
                                                         T* __restrict__ R_out) {
  const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= n_sel) return;
  const unsigned lt = (1u << lane) - 1u;
  const long long i = (long long)atoms[s] - 1;
  const T xi = Xo[3 * i], yi = Xo[3 * i + 1], zi = Xo[3 * i + 2];
  int ci[3];
  long long wi[3];
  cell_of(g, xi, yi, zi, ci, wi);
  const long long o0 = s * (long long)width;
  long long cnt = 0;
  for (int dz = -g.nxyz[2]; dz <= g.nxyz[2]; dz++) {
    int cz; long long sz = 0;
    { const long long v = (long long)ci[2] + dz;
      if (g.pbc[2]) wrap0(v, g.nc[2], cz, sz); else { if (v < 0 || v >= g.nc[2]) continue; cz = (int)v; } }
    for (int dy = -g.nxyz[1]; dy <= g.nxyz[1]; dy++) {
      int cy; long long sy = 0;
      { const long long v = (long long)ci[1] + dy;
        if (g.pbc[1]) wrap0(v, g.nc[1], cy, sy); else { if (v < 0 || v >= g.nc[1]) continue; cy = (int)v; } }
      for (int dx = -g.nxyz[0]; dx <= g.nxyz[0]; dx++) {
        int cx; long long sx = 0;
        { const long long v = (long long)ci[0] + dx;
          if (g.pbc[0]) wrap0(v, g.nc[0], cx, sx); else { if (v < 0 || v >= g.nc[0]) continue; cx = (int)v; } }
        const long long cl = (long long)cx + (long long)g.nc[0] * ((long long)cy + (long long)g.nc[1] * cz);
        const long long b0 = (long long)co[cl] - 1, b1 = (long long)co[cl + 1] - 1;
        const bool zero_shift = (sx == 0 && sy == 0 && sz == 0);
        for (long long t0 = b0; t0 < b1; t0 += 32) {
          const long long t = t0 + lane;
          bool hit = false;
          long long jo = 0, S[3] = {0, 0, 0};
          T R[3] = {0, 0, 0};
          if (t < b1) {
            jo = (long long)perm[t] - 1;
            if (!(jo == i && zero_shift)) {  // _is_self_interaction, src/gpu_kernels.jl:30-33
              const T xj = Xs[3 * t], yj = Xs[3 * t + 1], zj = Xs[3 * t + 2];
              int cj[3];
              long long wj[3];
              cell_of(g, xj, yj, zj, cj, wj);
              S[0] = sx + wi[0] - wj[0]; S[1] = sy + wi[1] - wj[1]; S[2] = sz + wi[2] - wj[2];
              hit = pair_r2(g, xi, yi, zi, xj, yj, zj, S, R) < g.cutoff_sq;
            }
          }
          const unsigned bal = __ballot_sync(0xffffffffu, hit);
          const long long pos = cnt + __popc(bal & lt);
          if (hit && pos < width) {
            j_out[o0 + pos] = (TI)(jo + 1);
            if (S_out) { S_out[3 * (o0 + pos)] = (TI)S[0]; S_out[3 * (o0 + pos) + 1] = (TI)S[1]; S_out[3 * (o0 + pos) + 2] = (TI)S[2]; }
            if (R_out) { R_out[3 * (o0 + pos)] = R[0]; R_out[3 * (o0 + pos) + 1] = R[1]; R_out[3 * (o0 + pos) + 2] = R[2]; }
          }
          cnt += __popc(bal);
        }
      }
    }
  }
  if (lane == 0) n_out[s] = (TI)cnt;
  for (long long k = cnt + lane; k < width; k += 32) {
    j_out[o0 + k] = 0;
    if (S_out) { S_out[3 * (o0 + k)] = 0; S_out[3 * (o0 + k) + 1] = 0; S_out[3 * (o0 + k) + 2] = 0; }
    if (R_out) { R_out[3 * (o0 + k)] = 0; R_out[3 * (o0 + k) + 1] = 0; R_out[3 * (o0 + k) + 2] = 0; }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Two-launch reductions over the atoms (partials in caller scratch, then one block).
constexpr int RED_BLOCKS = 592;  // 4 x 148 SMs

// min / max that PROPAGATE NaN, like Julia's minimum / maximum (a NaN position must not pass for a valid bounding box or
// for "nothing moved": the host then sees NaN and raises / rebuilds)
template <class T> __device__ __forceinline__ T nan_max(T a, T b) { return a != a ? a : (b != b ? b : (b > a ? b : a)); }
template <class T> __device__ __forceinline__ T nan_min(T a, T b) { return a != a ? a : (b != b ? b : (b < a ? b : a)); }
template <class T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) { const T t = __shfl_xor_sync(0xffffffffu, v, o); v = nan_min(v, t); }
  return v;
}
template <class T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) { const T t = __shfl_xor_sync(0xffffffffu, v, o); v = nan_max(v, t); }
  return v;
}

// block-wide (min, max) of 3 components -> part[6]: (min x, min y, min z, max x, max y, max z)
template <class T> __device__ __forceinline__ void block_minmax3(T mn[3], T mx[3], T* part) {
  __shared__ T sm[8][6];
  for (int k = 0; k < 3; k++) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) for (int k = 0; k < 3; k++) { sm[w][k] = mn[k]; sm[w][3 + k] = mx[k]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    T v = sm[0][threadIdx.x];
    for (int q = 1; q < (int)(blockDim.x >> 5); q++) {
      const T t = sm[q][threadIdx.x];
      v = threadIdx.x < 3 ? nan_min(v, t) : nan_max(v, t);
    }
    part[threadIdx.x] = v;
  }
}

// Bounding box of the positions (IsolatedCell branch of _get_cell_matrix, ext/NeighbourListsAtomsBaseExt.jl:17-31).
template <class T> __global__ void __launch_bounds__(256) k_bbox_partial(const T* __restrict__ X, long long n, T* __restrict__ part) {
  T mn[3], mx[3];
  const T x0 = X[0], y0 = X[1], z0 = X[2];  // n >= 1
  mn[0] = mx[0] = x0; mn[1] = mx[1] = y0; mn[2] = mx[2] = z0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const T v = X[3 * i + k];
      mn[k] = nan_min(mn[k], v);
      mx[k] = nan_max(mx[k], v);
    }
  }
  block_minmax3(mn, mx, part + 6 * blockIdx.x);
}
template <class T> __global__ void __launch_bounds__(256) k_bbox_final(const T* __restrict__ part, int nb, T* __restrict__ out) {
  T mn[3], mx[3];
  for (int k = 0; k < 3; k++) { mn[k] = part[k]; mx[k] = part[3 + k]; }
  for (int b = threadIdx.x; b < nb; b += blockDim.x)
    for (int k = 0; k < 3; k++) {
      const T a = part[6 * b + k], c = part[6 * b + 3 + k];
      mn[k] = nan_min(mn[k], a);
      mx[k] = nan_max(mx[k], c);
    }
  block_minmax3(mn, mx, out);
}

// max over atoms of |X - X_ref|^2 = (dx dx + dy dy) + dz dz, in T (skin-list validity check).
template <class T>
__global__ void __launch_bounds__(256) k_maxdisp_partial(const T* __restrict__ X, const T* __restrict__ Y, long long n, T* __restrict__ part) {
  T m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T dx = sub_rn(X[3 * i], Y[3 * i]), dy = sub_rn(X[3 * i + 1], Y[3 * i + 1]), dz = sub_rn(X[3 * i + 2], Y[3 * i + 2]);
    const T d2 = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
    m = nan_max(m, d2);
  }
  m = warp_max(m);
  __shared__ T sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); k++) m = nan_max(m, sm[k]);
    part[blockIdx.x] = m;
  }
}
template <class T> __global__ void __launch_bounds__(256) k_max_final(const T* __restrict__ part, int nb, T* __restrict__ out) {
  T m = 0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) m = nan_max(m, part[b]);
  m = warp_max(m);
  __shared__ T sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); k++) m = nan_max(m, sm[k]);
    out[0] = m;
  }
}

}  // namespace nl
