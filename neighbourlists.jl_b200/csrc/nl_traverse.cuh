// nl_traverse.cuh -- the neighbour traversal (_for_each_neighbor_pair, src/gpu_kernels.jl:58-101)
// with its three sinks: count (count_neighbours_kernel! :132-152), fill (fill_pairs_kernel!
// :159-180, plus optional R) and the fused Lennard-Jones energy of the lazy mode.
//
// Records: nl_count_pairs / nl_lazy_* first rewrite the SortedCellList into SoA "records" in the
// workspace (k_prep_records): px/py/pz (T), pidx (0-based original index) and pw (packed winding
// of the atom's raw cell index).  Because S = s_loop + w_i - w_j, carrying w per atom removes the
// reference's per-candidate recomputation of j's cell (a 3x3 mat-vec + 3 floors, :86-87).
#pragma once
#include "nl_common.cuh"

namespace nl {

enum { MODE_COUNT = 0, MODE_FILL = 1, MODE_LJ = 2, MODE_LJF = 3 };


template <class T> struct Records {
  const T* px;
  const T* py;
  const T* pz;
  const uint32_t* pidx;
  const uint32_t* pw;
};

template <class T, class TI> struct Sinks {
  uint32_t* counts;      // MODE_COUNT: per original atom
  const TI* first;       // MODE_FILL: 1-based CSR offsets, original order
  TI* io;                //            pair arrays
  TI* jo;
  TI* So;
  T* Ro;                 //            may be null
  long long n_rows;      //            only atoms with original index < n_rows get a row (sharding: owned atoms first)
  const TI* gmap;        //            may be null; else i/j are written as gmap[original index] (global, 1-based)
  const uint32_t* pgid0; //            with gmap: gmap[pidx[s]] - 1 per SORTED atom (one gather per atom instead of one per pair)
  double* energy;        // MODE_LJ: device scalar
  double lj_eps, lj_sigma2;
  int half;              // MODE_COUNT / MODE_FILL of a materialisation: keep one pair of each mirror couple (see half_keep)
  const uint8_t* plane_active;  // HOST pointer (read by the launcher only): per z plane of cells, may hold atoms; null = all
  T* fe;                 // MODE_LJF: N x 4 (force x, y, z, energy) per ORIGINAL atom, accumulated with atomics
};

// Lennard-Jones pair terms for squared distance r2: phi = 4 eps (s^12 - s^6), gg = 24 eps (2 s^12 - s^6) / r2 with
// s^2 = sigma^2 / r2.  A visit of the ordered pair (i, j) with R = x_j - x_i + shift adds gg * R to the force on j and
// phi to j's energy; every ordered pair is visited exactly once, so each atom ends up with its full force
// F_j = -dE/dx_j (E = half the ordered-pair sum) and with e_j = sum of phi over its neighbours.
__device__ __forceinline__ void lj_pair_terms(double eps, double sigma2, double r2, double& phi, double& gg) {
  const double inv = 1.0 / r2, s2 = sigma2 * inv, s6 = s2 * s2 * s2;
  phi = 4.0 * eps * (s6 * s6 - s6);
  gg = 24.0 * eps * (2.0 * s6 * s6 - s6) * inv;
}
template <class T, class TI>
__device__ __forceinline__ void ljf_add_global(const Sinks<T, TI>& out, uint32_t jo, T f0, T f1, T f2, T e) {
  T* d = out.fe + 4ll * jo;
  atomicAdd(d, f0); atomicAdd(d + 1, f1); atomicAdd(d + 2, f2); atomicAdd(d + 3, e);
}

// Half lists: of the mirror couple (i, j, S) / (j, i, -S) exactly one pair is kept -- the one whose SECOND atom comes
// later in cell-sorted order, or, for self images (same atom), the one with a lexicographically positive shift.  The rule
// only uses sorted indices and the shift, so every traversal route reaches the same verdict.
__device__ __forceinline__ bool shift_lex_positive(long long s0, long long s1, long long s2) {
  return s0 > 0 || (s0 == 0 && (s1 > 0 || (s1 == 0 && s2 > 0)));
}
__device__ __forceinline__ bool half_keep(long long sorted_i, long long sorted_j, long long s0, long long s1, long long s2) {
  return sorted_j > sorted_i || (sorted_j == sorted_i && shift_lex_positive(s0, s1, s2));
}

template <class T, class TI> __device__ __forceinline__ TI out_index(const Sinks<T, TI>& out, uint32_t orig) {
  return out.gmap ? out.gmap[orig] : (TI)orig + 1;
}

// AoS twin of the records: one 32-byte sector per atom (x, y, z, index to publish, packed winding).
template <class T> struct RecAoS;
template <> struct alignas(32) RecAoS<double> { double x, y, z; uint32_t idx, w; };
template <> struct alignas(32) RecAoS<float> { float x, y, z; uint32_t idx, w, pad0, pad1, pad2; };
static_assert(sizeof(RecAoS<double>) == 32 && sizeof(RecAoS<float>) == 32, "one sector per record");

template <class T, class TI>
__global__ void __launch_bounds__(256) k_prep_records(const T* __restrict__ Xs, const TI* __restrict__ perm, long long n, Geo<T> g,
                                                      T* __restrict__ px, T* __restrict__ py, T* __restrict__ pz,
                                                      uint32_t* __restrict__ pidx, uint32_t* __restrict__ pw,
                                                      RecAoS<T>* __restrict__ ra, uint32_t* __restrict__ pkey) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  T x = Xs[3 * s], y = Xs[3 * s + 1], z = Xs[3 * s + 2];
  int c[3];
  long long w[3];
  cell_of(g, x, y, z, c, w);
  const uint32_t idx = (uint32_t)(perm[s] - 1), pwv = pack_wind(w);
  px[s] = x;
  py[s] = y;
  pz[s] = z;
  pidx[s] = idx;
  pw[s] = pwv;
  if (ra) {  // AoS twin + cell key: only the original-order fill experiment (NL_FILL_ROWS=1) reads them
    RecAoS<T> r;
    r.x = x; r.y = y; r.z = z; r.idx = idx; r.w = pwv;
    ra[s] = r;
    pkey[s] = (uint32_t)c[0] + (uint32_t)g.nc[0] * ((uint32_t)c[1] + (uint32_t)g.nc[1] * (uint32_t)c[2]);
  }
}

template <class TI>
__global__ void __launch_bounds__(256) k_make_pgid(const uint32_t* __restrict__ pidx, const TI* __restrict__ gmap, long long n,
                                                   uint32_t* __restrict__ pgid0) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) pgid0[s] = (uint32_t)(gmap[pidx[s]] - 1);
}

// Generic traversal of ONE atom (sorted slot s): any nxyz >= 1, any ncells >= 1, any density.
// Visits cells in the reference's order (dz, dy, dx, then sorted slot), so a FILL row comes out in
// the reference's own traversal order.  Used where the tiled kernel's assumptions do not hold
// (tiny boxes, stencils wider than one cell, tiles denser than the shared-memory budget).
// Returns the atom's LJ energy contribution (MODE_LJ) or 0.
template <class T, class TI, int MODE>
__device__ __forceinline__ double generic_atom(long long s, const Records<T>& rec, const TI* __restrict__ co, const Geo<T>& g,
                                               const Sinks<T, TI>& out) {
  double e_acc = 0.0;
  const T xi = rec.px[s], yi = rec.py[s], zi = rec.pz[s];
  const uint32_t io = rec.pidx[s];
  int ci[3];
  long long wi[3];
  cell_of(g, xi, yi, zi, ci, wi);
  uint32_t cnt = 0;
  long long wpos = 0;
  if (MODE == MODE_FILL) {
    if ((long long)io >= out.n_rows) return 0.0;
    wpos = (long long)out.first[io] - 1;
  }
  for (int dz = -g.nxyz[2]; dz <= g.nxyz[2]; dz++) {
    int cz; long long sz = 0;
    { long long v = (long long)ci[2] + dz;
      if (g.pbc[2]) wrap0(v, g.nc[2], cz, sz); else { if (v < 0 || v >= g.nc[2]) continue; cz = (int)v; } }
    for (int dy = -g.nxyz[1]; dy <= g.nxyz[1]; dy++) {
      int cy; long long sy = 0;
      { long long v = (long long)ci[1] + dy;
        if (g.pbc[1]) wrap0(v, g.nc[1], cy, sy); else { if (v < 0 || v >= g.nc[1]) continue; cy = (int)v; } }
      for (int dx = -g.nxyz[0]; dx <= g.nxyz[0]; dx++) {
        int cx; long long sx = 0;
        { long long v = (long long)ci[0] + dx;
          if (g.pbc[0]) wrap0(v, g.nc[0], cx, sx); else { if (v < 0 || v >= g.nc[0]) continue; cx = (int)v; } }
        const long long cl = (long long)cx + (long long)g.nc[0] * ((long long)cy + (long long)g.nc[1] * cz);
        const long long b0 = (long long)co[cl] - 1, b1 = (long long)co[cl + 1] - 1;
        const bool zero_shift = (sx == 0 && sy == 0 && sz == 0);
        for (long long t = b0; t < b1; t++) {
          const uint32_t jo = rec.pidx[t];
          if (jo == io && zero_shift) continue;  // _is_self_interaction, src/gpu_kernels.jl:30-33
          if ((MODE == MODE_COUNT || MODE == MODE_FILL) && out.half && !half_keep(s, t, sx, sy, sz)) continue;
          const T xj = rec.px[t], yj = rec.py[t], zj = rec.pz[t];
          long long wj[3];
          const uint32_t pwj = rec.pw[t];
          if (pwj & WIND_OVERFLOW) { int cj[3]; cell_of(g, xj, yj, zj, cj, wj); } else unpack_wind(pwj, wj);
          const long long S[3] = {sx + wi[0] - wj[0], sy + wi[1] - wj[1], sz + wi[2] - wj[2]};
          T R[3];
          const T r2 = pair_r2(g, xi, yi, zi, xj, yj, zj, S, R);
          if (r2 < g.cutoff_sq) {
            if (MODE == MODE_COUNT) cnt++;
            if (MODE == MODE_FILL) {
              out.io[wpos] = out_index(out, io);
              out.jo[wpos] = out_index(out, jo);
              out.So[3 * wpos] = (TI)S[0];
              out.So[3 * wpos + 1] = (TI)S[1];
              out.So[3 * wpos + 2] = (TI)S[2];
              if (out.Ro) { out.Ro[3 * wpos] = R[0]; out.Ro[3 * wpos + 1] = R[1]; out.Ro[3 * wpos + 2] = R[2]; }
              wpos++;
            }
            if (MODE == MODE_LJ) {
              const double s2 = out.lj_sigma2 / (double)r2, s6 = s2 * s2 * s2;
              e_acc += 4.0 * out.lj_eps * (s6 * s6 - s6);
            }
            if (MODE == MODE_LJF) {
              double phi, gg;
              lj_pair_terms(out.lj_eps, out.lj_sigma2, (double)r2, phi, gg);
              ljf_add_global<T, TI>(out, jo, (T)(gg * (double)R[0]), (T)(gg * (double)R[1]), (T)(gg * (double)R[2]), (T)phi);
            }
          }
        }
      }
    }
  }
  if (MODE == MODE_COUNT) out.counts[io] = cnt;
  return e_acc;
}

template <class T, class TI, int MODE>
__global__ void __launch_bounds__(128) k_traverse_generic(Records<T> rec, const TI* __restrict__ co, long long n, Geo<T> g, Sinks<T, TI> out) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double e_acc = 0.0;
  if (s < n) e_acc = generic_atom<T, TI, MODE>(s, rec, co, g, out);
  if (MODE == MODE_LJ) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e_acc += __shfl_xor_sync(FULL, e_acc, o);
    if ((threadIdx.x & 31) == 0 && e_acc != 0.0) atomicAdd(out.energy, e_acc);
  }
}

}  // namespace nl
