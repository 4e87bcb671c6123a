// nl_oracle.cpp -- CPU restatement of NeighbourLists.jl's sort-based neighbour-list path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under oracle/ is part of the product.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library, and only as the checker / the timed CPU baseline.  The product (libnlcuda.so and the
// Python host mirror) never links, imports or calls it.
//
// PARITY STATUS: the reference is pure Julia and cannot be executed in this image (no julia, no
// juliacall), so this restatement is pinned against the reference's analytic known answers
// (tests/test_oracle_kats.py) and against an independent brute-force image
// enumeration; it is NOT pinned at the ulp level against a live run of the reference ("parity
// unpinned at ulp level": the association order of StaticArrays' 3x3 mat-vec / dot / inv is
// restated from StaticArrays 1.x's published generated code, which is not vendored in
// /root/reference).
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Build: see oracle/Makefile (g++ -O3 -ffp-contract=off -fopenmp; no fast-math: the arithmetic
// contract below forbids FMA contraction and reassociation).
//
// Conventions: matrices are 3x3 in Julia column-major order, m[r + 3*c] = M[r+1, c+1]; ROWS of
// `cell` are the lattice vectors (x = cell' * frac).  All integer outputs are 1-based like the
// reference's.  T in {float, double}, TI in {int32_t, int64_t}.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------- small static-array algebra
// StaticArrays-style unrolled, left-associated, non-FMA arithmetic (SURVEY.md 8a "arithmetic
// contract").
template <class T> struct V3 { T x, y, z; };

template <class T> inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class T> inline T dot(const V3<T>& a, const V3<T>& b) {
  return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
template <class T> inline T norm(const V3<T>& a) { return std::sqrt(dot(a, a)); }
template <class T> inline V3<T> col(const T* m, int c) { return {m[3 * c], m[3 * c + 1], m[3 * c + 2]}; }
template <class T> inline V3<T> row(const T* m, int r) { return {m[r], m[r + 3], m[r + 6]}; }

// det(::SMatrix{3,3}) = dot(col1, cross(col2, col3))
template <class T> inline T det3(const T* m) { return dot(col(m, 0), cross(col(m, 1), col(m, 2))); }

// inv(::SMatrix{3,3}) -- StaticArrays' adjugate formulation (src/inv.jl, Size (3,3)).
template <class T> inline void inv3(const T* m, T* out) {
  V3<T> x0 = col(m, 0), x1 = col(m, 1), x2 = col(m, 2);
  V3<T> y0 = cross(x1, x2);
  T d = dot(x0, y0);
  x0 = {x0.x / d, x0.y / d, x0.z / d};
  y0 = {y0.x / d, y0.y / d, y0.z / d};
  V3<T> y1 = cross(x2, x0);
  V3<T> y2 = cross(x0, x1);
  // column-major tuple (y0[1], y1[1], y2[1], y0[2], y1[2], y2[2], y0[3], y1[3], y2[3])
  out[0] = y0.x; out[1] = y1.x; out[2] = y2.x;
  out[3] = y0.y; out[4] = y1.y; out[5] = y2.y;
  out[6] = y0.z; out[7] = y1.z; out[8] = y2.z;
}

// lengths(C) -- src/cell_list.jl:94-95 (signed: det(C) ./ norms of row cross products)
template <class T> inline void lengths3(const T* c, T* out) {
  T d = det3(c);
  out[0] = d / norm(cross(row(c, 1), row(c, 2)));
  out[1] = d / norm(cross(row(c, 2), row(c, 0)));
  out[2] = d / norm(cross(row(c, 0), row(c, 1)));
}

// M' * v  (v floating): (M')[k,j] = M[j,k]  ->  out_k = (M[1,k] v1 + M[2,k] v2) + M[3,k] v3
template <class T> inline V3<T> mtv(const T* m, const V3<T>& v) {
  return {(m[0] * v.x + m[1] * v.y) + m[2] * v.z,
          (m[3] * v.x + m[4] * v.y) + m[5] * v.z,
          (m[6] * v.x + m[7] * v.y) + m[8] * v.z};
}
// M * v: out_k = (M[k,1] v1 + M[k,2] v2) + M[k,3] v3
template <class T> inline V3<T> mv(const T* m, const V3<T>& v) {
  return {(m[0] * v.x + m[3] * v.y) + m[6] * v.z,
          (m[1] * v.x + m[4] * v.y) + m[7] * v.z,
          (m[2] * v.x + m[5] * v.y) + m[8] * v.z};
}

// ---------------------------------------------------------------- cell index algebra
// wrap_and_shift, src/cell_list.jl:105-120 (closed form of the two while loops).
template <class TI> inline void wrap_and_shift(TI i, TI n, bool pbc, TI& wrapped, TI& shift) {
  if (!pbc) {
    wrapped = i < 1 ? TI(1) : (i > n ? n : i);
    shift = 0;
    return;
  }
  if (i >= 1 && i <= n) { wrapped = i; shift = 0; return; }  // neither while loop runs
  TI q = (i - 1) / n, r = (i - 1) % n;
  if (r < 0) { r += n; q -= 1; }
  wrapped = r + 1;
  shift = q;
}

// position_to_cell_index, src/cell_list.jl:66-74
template <class T, class TI>
inline void position_to_cell_index(const T* inv, const V3<T>& x, const TI* nc, TI* c) {
  V3<T> f = mtv(inv, x);
  c[0] = (TI)std::floor(f.x * (T)nc[0] + (T)1);
  c[1] = (TI)std::floor(f.y * (T)nc[1] + (T)1);
  c[2] = (TI)std::floor(f.z * (T)nc[2] + (T)1);
}

// _sub2ind, src/cell_list.jl:83-86
template <class TI> inline TI sub2ind(const TI* d, const TI* i) {
  return i[0] + (i[1] - 1) * d[0] + (i[2] - 1) * d[0] * d[1];
}

template <class T, class TI> struct Geo {
  T cell[9], inv[9], cutoff;
  TI nc[3], nxyz[3];
  bool pbc[3];
};

template <class T, class TI>
Geo<T, TI> make_geo(const T* cell, const T* inv, T cutoff, const int64_t* nc, const int64_t* nxyz,
                    const uint8_t* pbc) {
  Geo<T, TI> g;
  for (int k = 0; k < 9; k++) { g.cell[k] = cell[k]; g.inv[k] = inv[k]; }
  g.cutoff = cutoff;
  for (int k = 0; k < 3; k++) { g.nc[k] = (TI)nc[k]; g.nxyz[k] = nxyz ? (TI)nxyz[k] : 0; g.pbc[k] = pbc[k] != 0; }
  return g;
}

// ---------------------------------------------------------------- analyze_cell
// src/cell_list.jl:152-170 plus the nxyz formula of src/gpu_kernels.jl:315-316 /
// src/cell_list.jl:787-790.
template <class T>
int analyze_cell(const T* cell, T cutoff, T* inv, T* lens, int64_t* nc, int64_t* nxyz) {
  inv3(cell, inv);
  T l[3];
  lengths3(cell, l);
  for (int k = 0; k < 3; k++) {
    lens[k] = std::fabs(l[k]);
    int64_t t = (int64_t)std::floor(lens[k] / cutoff);
    nc[k] = t > 1 ? t : 1;
  }
  for (int k = 0; k < 3; k++) nxyz[k] = (int64_t)std::ceil(cutoff * ((T)nc[k] / lens[k]));
  T vol = std::fabs(det3(cell));
  return vol < (T)1e-12 ? 1 : 0;  // 1 = the reference would @warn "zero volume"
}

// ---------------------------------------------------------------- sort-based build
// _build_sorted_celllist, src/cell_list.jl:647-679 with the CPU stage variants :684-748.
template <class T, class TI>
void build_cells(const T* X, int64_t N, const Geo<T, TI>& g, T* Xs, TI* perm, TI* cell_id, TI* cell_offsets) {
  int64_t nct = (int64_t)g.nc[0] * g.nc[1] * g.nc[2];
  std::vector<TI> ids((size_t)N);
  // _compute_cell_ids (CPU), :684-701 -- serial loop in the reference
  for (int64_t i = 0; i < N; i++) {
    V3<T> x{X[3 * i], X[3 * i + 1], X[3 * i + 2]};
    TI c0[3], c[3], w;
    position_to_cell_index(g.inv, x, g.nc, c0);
    for (int k = 0; k < 3; k++) wrap_and_shift(c0[k], g.nc[k], g.pbc[k], c[k], w);  // bin_wrap_or_trunc :126-132
    ids[(size_t)i] = sub2ind(g.nc, c);
  }
  // _get_sortperm (CPU) = TI.(sortperm(cell_ids)), :706-708.  sortperm is stable, so the result is
  // the unique stable permutation; computed here with a stable counting sort.
  std::vector<int64_t> start((size_t)nct + 2, 0);
  for (int64_t i = 0; i < N; i++) start[(size_t)ids[(size_t)i] + 1]++;
  // _compute_cell_offsets (CPU), :724-748: histogram + cumulative sum, 1-based, nat==0 -> all ones
  for (int64_t c = 1; c <= nct + 1; c++) cell_offsets[c - 1] = 0;
  if (N == 0) {
    for (int64_t c = 0; c <= nct; c++) cell_offsets[c] = 1;
    return;
  }
  cell_offsets[0] = 1;
  for (int64_t c = 1; c <= nct; c++) cell_offsets[c] = (TI)(cell_offsets[c - 1] + (TI)start[(size_t)c + 1]);
  std::vector<int64_t> cursor((size_t)nct + 1);
  for (int64_t c = 1; c <= nct; c++) cursor[(size_t)c] = (int64_t)cell_offsets[c - 1] - 1;
  for (int64_t i = 0; i < N; i++) {
    int64_t s = cursor[(size_t)ids[(size_t)i]]++;
    perm[s] = (TI)(i + 1);
  }
  // sorted_cell_ids = cell_ids[perm]; sorted_X = X[perm], :669-670
  for (int64_t s = 0; s < N; s++) {
    int64_t i = (int64_t)perm[s] - 1;
    cell_id[s] = ids[(size_t)i];
    Xs[3 * s] = X[3 * i]; Xs[3 * s + 1] = X[3 * i + 1]; Xs[3 * s + 2] = X[3 * i + 2];
  }
}

// ---------------------------------------------------------------- traversal
// _for_each_neighbor_pair, src/gpu_kernels.jl:58-101 (with _is_cell_in_bounds :20-25,
// _get_neighbor_cell :39-47, _is_self_interaction :30-33).  `i` is the 1-based ORIGINAL index.
template <class T, class TI, class F>
inline void for_each_neighbor_pair(F&& f, TI i, const T* X, const TI* cell_offsets, const TI* perm,
                                   const Geo<T, TI>& g, T cutoff_sq) {
  V3<T> xi{X[3 * (i - 1)], X[3 * (i - 1) + 1], X[3 * (i - 1) + 2]};
  TI ci0[3], ci[3], wi[3];
  position_to_cell_index(g.inv, xi, g.nc, ci0);
  for (int k = 0; k < 3; k++) wrap_and_shift(ci0[k], g.nc[k], g.pbc[k], ci[k], wi[k]);
  for (TI dz = -g.nxyz[2]; dz <= g.nxyz[2]; dz++)
    for (TI dy = -g.nxyz[1]; dy <= g.nxyz[1]; dy++)
      for (TI dx = -g.nxyz[0]; dx <= g.nxyz[0]; dx++) {
        TI d[3] = {dx, dy, dz};
        bool inb = true;
        for (int k = 0; k < 3; k++) inb = inb && (g.pbc[k] || (1 <= ci[k] + d[k] && ci[k] + d[k] <= g.nc[k]));
        if (!inb) continue;
        TI cj[3], sl[3];
        for (int k = 0; k < 3; k++) wrap_and_shift((TI)(ci[k] + d[k]), g.nc[k], g.pbc[k], cj[k], sl[k]);
        TI cjl = sub2ind(g.nc, cj);
        TI i0 = cell_offsets[cjl - 1], i1 = cell_offsets[cjl] - 1;
        for (TI idx = i0; idx <= i1; idx++) {
          TI j = perm[idx - 1];
          if (i == j && sl[0] == 0 && sl[1] == 0 && sl[2] == 0) continue;
          V3<T> xj{X[3 * (j - 1)], X[3 * (j - 1) + 1], X[3 * (j - 1) + 2]};
          TI cj0[3], cjw, wj[3];
          position_to_cell_index(g.inv, xj, g.nc, cj0);
          for (int k = 0; k < 3; k++) wrap_and_shift(cj0[k], g.nc[k], g.pbc[k], cjw, wj[k]);
          TI S[3] = {(TI)(sl[0] + wi[0] - wj[0]), (TI)(sl[1] + wi[1] - wj[1]), (TI)(sl[2] + wi[2] - wj[2])};
          V3<T> Sf{(T)S[0], (T)S[1], (T)S[2]};
          V3<T> cs = mtv(g.cell, Sf);  // cell_mat' * S
          V3<T> R{(xj.x - xi.x) + cs.x, (xj.y - xi.y) + cs.y, (xj.z - xi.z) + cs.z};
          T r2 = dot(R, R);
          if (r2 < cutoff_sq) f(j, R, S);
        }
      }
}

// count_neighbours_kernel!, src/gpu_kernels.jl:132-152 (KA CPU backend: parallel over atoms)
template <class T, class TI>
void count_pairs(const T* X, int64_t N, const TI* perm, const TI* cell_offsets, const Geo<T, TI>& g,
                 int nthreads, TI* counts) {
  T csq = g.cutoff * g.cutoff;  // clist.cutoff^2, src/gpu_kernels.jl:317
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
  for (int64_t i = 1; i <= N; i++) {
    TI c = 0;
    for_each_neighbor_pair([&](TI, const V3<T>&, const TI*) { c += 1; }, (TI)i, X, cell_offsets, perm, g, csq);
    counts[i - 1] = c;
  }
}

// compute_pair_offsets (CPU), src/gpu_kernels.jl:211-219
template <class TI> void pair_offsets(const TI* counts, int64_t N, TI* offsets) {
  offsets[0] = 1;
  for (int64_t i = 0; i < N; i++) offsets[i + 1] = (TI)(offsets[i] + counts[i]);
}

// fill_pairs_kernel!, src/gpu_kernels.jl:159-180 (+ optional R, the quantity _getR recomputes,
// src/cell_list.jl:525-531: identical expression, so identical bits)
template <class T, class TI>
void fill_pairs(const T* X, int64_t N, const TI* perm, const TI* cell_offsets, const Geo<T, TI>& g,
                int nthreads, const TI* first, TI* io, TI* jo, TI* So, T* Ro) {
  T csq = g.cutoff * g.cutoff;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
  for (int64_t i = 1; i <= N; i++) {
    int64_t w = (int64_t)first[i - 1] - 1;
    for_each_neighbor_pair(
        [&](TI j, const V3<T>& R, const TI* S) {
          io[w] = (TI)i; jo[w] = j;
          So[3 * w] = S[0]; So[3 * w + 1] = S[1]; So[3 * w + 2] = S[2];
          if (Ro) { Ro[3 * w] = R.x; Ro[3 * w + 1] = R.y; Ro[3 * w + 2] = R.z; }
          w++;
        },
        (TI)i, X, cell_offsets, perm, g, csq);
  }
}

// Lennard-Jones sink over for_each_neighbour (BASELINE config 5): sum over ORDERED pairs of
// 4 eps ((sigma/r)^12 - (sigma/r)^6), R and r^2 in T, accumulation in double.
template <class T, class TI>
double lj_energy(const T* X, int64_t N, const TI* perm, const TI* cell_offsets, const Geo<T, TI>& g,
                 int nthreads, double eps, double sigma) {
  T csq = g.cutoff * g.cutoff;
  double total = 0.0;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads) reduction(+ : total)
  for (int64_t i = 1; i <= N; i++) {
    double e = 0.0;
    for_each_neighbor_pair(
        [&](TI, const V3<T>& R, const TI*) {
          double r2 = (double)dot(R, R);
          double s2 = sigma * sigma / r2, s6 = s2 * s2 * s2;
          e += 4.0 * eps * (s6 * s6 - s6);
        },
        (TI)i, X, cell_offsets, perm, g, csq);
    total += e;
  }
  return total;
}

// ---------------------------------------------------------------- legacy linked-list path (second oracle)
struct PairSet {
  std::vector<int64_t> i, j, S;  // S is 3 per pair
  std::vector<int64_t> first;    // N+1, 1-based
  std::vector<double> X, C;      // possibly "fixed" positions / cell (legacy) -- as double
  std::vector<double> R;         // 3 per pair (brute only)
};

// Julia mod(x, 1.0) for floats (Base: rem then sign fix)
template <class T> inline T jl_mod1(T x) {
  T r = std::fmod(x, (T)1);
  if (r == 0) return std::copysign(r, (T)1);
  if (r < 0) return r + (T)1;
  return r;
}

// _fix_cell_, src/cell_list.jl:459-505
template <class T> void fix_cell(std::vector<T>& X, T* C, const bool* pbc) {
  T invC[9];
  inv3(C, invC);
  int64_t N = (int64_t)X.size() / 3;
  double min_lam[3] = {0, 0, 0}, max_lam[3] = {1, 1, 1};
  for (int64_t n = 0; n < N; n++) {
    V3<T> x{X[3 * n], X[3 * n + 1], X[3 * n + 2]};
    V3<T> lv = mtv(invC, x);  // inv(C)' * x
    T lam[3] = {lv.x, lv.y, lv.z};
    bool upd = false;
    for (int k = 0; k < 3; k++) {
      if (!(0.0 <= lam[k] && lam[k] < 1.0)) {
        if (pbc[k]) { lam[k] = jl_mod1(lam[k]); upd = true; }
        else { min_lam[k] = std::min(min_lam[k], (double)lam[k]); max_lam[k] = std::max(max_lam[k], (double)lam[k]); }
      }
    }
    if (upd) {
      V3<T> nx = mtv(C, V3<T>{lam[0], lam[1], lam[2]});  // C' * lam
      X[3 * n] = nx.x; X[3 * n + 1] = nx.y; X[3 * n + 2] = nx.z;
    }
  }
  double mn = std::min({min_lam[0], min_lam[1], min_lam[2]}), mx = std::max({max_lam[0], max_lam[1], max_lam[2]});
  if (mn < 0.0 || mx > 1.0) {
    // t = -C' * min_lam  (min_lam is Float64 in the reference; T==double in the tests that use this)
    V3<T> t = mtv(C, V3<T>{(T)min_lam[0], (T)min_lam[1], (T)min_lam[2]});
    for (int64_t n = 0; n < N; n++) { X[3 * n] += -t.x; X[3 * n + 1] += -t.y; X[3 * n + 2] += -t.z; }
    for (int k = 0; k < 3; k++) { max_lam[k] -= min_lam[k]; min_lam[k] = 0.0; }
    for (int k = 0; k < 3; k++) if (max_lam[k] > 1) max_lam[k] *= 1.01;
    // C = hcat(max_lam[k] * C[k,:])'  -> row k scaled
    for (int k = 0; k < 3; k++) for (int c = 0; c < 3; c++) C[k + 3 * c] = (T)(max_lam[k] * (double)C[k + 3 * c]);
  }
}

// _pairlist_(X, cell, pbc, cutoff, TI, fixcell) and everything below it:
// src/cell_list.jl:369-384, _celllist_ :185-230, _pairlist_(clist) :233-297,
// _find_neighbours_! :301-366, get_first :401-418, sort_neigs! :427-446.
template <class T>
PairSet* legacy_pairlist(const T* Xin, int64_t N, const T* cell_in, const uint8_t* pbc8, T cutoff, bool fixcell) {
  typedef int64_t TI;
  std::vector<T> X(Xin, Xin + 3 * N);
  T C[9];
  for (int k = 0; k < 9; k++) C[k] = cell_in[k];
  bool pbc[3] = {pbc8[0] != 0, pbc8[1] != 0, pbc8[2] != 0};
  if (fixcell) fix_cell(X, C, pbc);
  T inv[9], lens[3];
  int64_t nc64[3], nxyz64[3];
  analyze_cell(C, cutoff, inv, lens, nc64, nxyz64);
  TI ns[3] = {nc64[0], nc64[1], nc64[2]};
  TI nxyz[3] = {nxyz64[0], nxyz64[1], nxyz64[2]};
  TI ncells = ns[0] * ns[1] * ns[2];
  std::vector<TI> seed((size_t)ncells, -1), last((size_t)ncells, 0), next((size_t)N, -1);
  for (TI i = 1; i <= N; i++) {
    V3<T> x{X[3 * (i - 1)], X[3 * (i - 1) + 1], X[3 * (i - 1) + 2]};
    TI c0[3], c[3], w;
    position_to_cell_index(inv, x, ns, c0);
    for (int k = 0; k < 3; k++) wrap_and_shift(c0[k], ns[k], pbc[k], c[k], w);
    TI ci = sub2ind(ns, c);
    if (seed[(size_t)ci - 1] < 0) { next[(size_t)i - 1] = -1; seed[(size_t)ci - 1] = i; last[(size_t)ci - 1] = i; }
    else { next[(size_t)i - 1] = -1; next[(size_t)last[(size_t)ci - 1] - 1] = i; last[(size_t)ci - 1] = i; }
  }
  // bins[:, k] = cell[k, :] / ns[k]   (:263)
  T bins[9];
  for (int k = 0; k < 3; k++) for (int r = 0; r < 3; r++) bins[r + 3 * k] = C[k + 3 * r] / (T)ns[k];
  T csq = cutoff * cutoff;
  PairSet* ps = new PairSet();
  auto btrunc = [&](TI i, int k) -> TI { return pbc[k] ? i : (i <= 0 ? TI(1) : (i > ns[k] ? ns[k] : i)); };
  for (TI i = 1; i <= N; i++) {
    V3<T> xi{X[3 * (i - 1)], X[3 * (i - 1) + 1], X[3 * (i - 1) + 2]};
    TI ci0[3], ci[3], w;
    position_to_cell_index(inv, xi, ns, ci0);
    TI ct[3] = {btrunc(ci0[0], 0), btrunc(ci0[1], 1), btrunc(ci0[2], 2)};
    V3<T> o = mv(bins, V3<T>{(T)(ct[0] - 1), (T)(ct[1] - 1), (T)(ct[2] - 1)});
    V3<T> dxi{xi.x - o.x, xi.y - o.y, xi.z - o.z};
    for (int k = 0; k < 3; k++) wrap_and_shift(ci0[k], ns[k], pbc[k], ci[k], w);
    for (TI dz = -nxyz[2]; dz <= nxyz[2]; dz++)
      for (TI dy = -nxyz[1]; dy <= nxyz[1]; dy++)
        for (TI dx = -nxyz[0]; dx <= nxyz[0]; dx++) {  // CartesianIndices: first index fastest
          TI xyz[3] = {dx, dy, dz}, cj[3];
          bool ok = true;
          for (int k = 0; k < 3; k++) {
            TI v = ci[k] + xyz[k];
            if (pbc[k]) { TI s; wrap_and_shift(v, ns[k], true, cj[k], s); } else cj[k] = v;
            ok = ok && (1 <= cj[k] && cj[k] <= ns[k]);
          }
          if (!ok) continue;
          TI ncj = sub2ind(ns, cj);
          V3<T> off = mv(bins, V3<T>{(T)dx, (T)dy, (T)dz});
          TI j = seed[(size_t)ncj - 1];
          while (j > 0) {
            if (i != j || dx != 0 || dy != 0 || dz != 0) {
              V3<T> xj{X[3 * (j - 1)], X[3 * (j - 1) + 1], X[3 * (j - 1) + 2]};
              TI cjr[3];
              position_to_cell_index(inv, xj, ns, cjr);
              TI cjt[3] = {btrunc(cjr[0], 0), btrunc(cjr[1], 1), btrunc(cjr[2], 2)};
              V3<T> oj = mv(bins, V3<T>{(T)(cjt[0] - 1), (T)(cjt[1] - 1), (T)(cjt[2] - 1)});
              V3<T> dxj{xj.x - oj.x, xj.y - oj.y, xj.z - oj.z};
              V3<T> d{(dxj.x - dxi.x) + off.x, (dxj.y - dxi.y) + off.y, (dxj.z - dxi.z) + off.z};
              if (dot(d, d) < csq) {
                ps->i.push_back(i); ps->j.push_back(j);
                for (int k = 0; k < 3; k++) ps->S.push_back((ci0[k] - cjt[k] + xyz[k]) / ns[k]);  // .÷ truncates
              }
            }
            j = next[(size_t)j - 1];
          }
        }
  }
  // get_first + sort_neigs! (stable by j within each row)
  int64_t P = (int64_t)ps->i.size();
  ps->first.assign((size_t)N + 1, P + 1);
  {
    int64_t idx = 1, n = 1;
    while (n <= N && idx <= P) {
      ps->first[(size_t)n - 1] = idx;
      while (idx <= P && ps->i[(size_t)idx - 1] == n) idx++;
      n++;
    }
    for (int64_t m = n; m <= N + 1; m++) ps->first[(size_t)m - 1] = P + 1;
  }
  for (int64_t n = 0; n < N; n++) {
    int64_t a = ps->first[(size_t)n] - 1, b = ps->first[(size_t)n + 1] - 1;
    if (b - a < 2) continue;
    std::vector<int64_t> ord((size_t)(b - a));
    for (int64_t k = 0; k < b - a; k++) ord[(size_t)k] = a + k;
    std::stable_sort(ord.begin(), ord.end(), [&](int64_t p, int64_t q) { return ps->j[(size_t)p] < ps->j[(size_t)q]; });
    std::vector<int64_t> jj, ss;
    for (int64_t p : ord) { jj.push_back(ps->j[(size_t)p]); for (int k = 0; k < 3; k++) ss.push_back(ps->S[(size_t)(3 * p + k)]); }
    for (int64_t k = 0; k < b - a; k++) { ps->j[(size_t)(a + k)] = jj[(size_t)k]; for (int c = 0; c < 3; c++) ps->S[(size_t)(3 * (a + k) + c)] = ss[(size_t)(3 * k + c)]; }
  }
  ps->X.assign(X.begin(), X.end());
  ps->C.assign(C, C + 9);
  return ps;
}

// ---------------------------------------------------------------- brute force (third oracle)
// The set-level invariant of SURVEY.md 8a: {(i,j,S): |X[j]-X[i]+C'S| < rc, (i,j,S) != (i,i,0)},
// S over Z on periodic axes and {0} on open axes; distance arithmetic follows the contract.
template <class T>
PairSet* brute_pairlist(const T* X, int64_t N, const T* cell, const uint8_t* pbc8, T cutoff) {
  T inv[9], l[3];
  inv3(cell, inv);
  lengths3(cell, l);
  T csq = cutoff * cutoff;
  PairSet* ps = new PairSet();
  ps->first.assign((size_t)N + 1, 1);
  for (int64_t i = 1; i <= N; i++) {
    ps->first[(size_t)i - 1] = (int64_t)ps->i.size() + 1;
    V3<T> xi{X[3 * (i - 1)], X[3 * (i - 1) + 1], X[3 * (i - 1) + 2]};
    for (int64_t j = 1; j <= N; j++) {
      V3<T> xj{X[3 * (j - 1)], X[3 * (j - 1) + 1], X[3 * (j - 1) + 2]};
      V3<T> d{xj.x - xi.x, xj.y - xi.y, xj.z - xi.z};
      V3<T> f = mtv(inv, d);
      T fr[3] = {f.x, f.y, f.z};
      int64_t lo[3], hi[3];
      for (int k = 0; k < 3; k++) {
        if (pbc8[k]) {
          double reach = (double)cutoff / std::fabs((double)l[k]);
          lo[k] = (int64_t)std::floor(-(double)fr[k] - reach) - 1;
          hi[k] = (int64_t)std::ceil(-(double)fr[k] + reach) + 1;
        } else lo[k] = hi[k] = 0;
      }
      for (int64_t s3 = lo[2]; s3 <= hi[2]; s3++)
        for (int64_t s2 = lo[1]; s2 <= hi[1]; s2++)
          for (int64_t s1 = lo[0]; s1 <= hi[0]; s1++) {
            if (i == j && s1 == 0 && s2 == 0 && s3 == 0) continue;
            V3<T> cs = mtv(cell, V3<T>{(T)s1, (T)s2, (T)s3});
            V3<T> R{d.x + cs.x, d.y + cs.y, d.z + cs.z};
            if (dot(R, R) < csq) {
              ps->i.push_back(i); ps->j.push_back(j);
              ps->S.push_back(s1); ps->S.push_back(s2); ps->S.push_back(s3);
              ps->R.push_back((double)R.x); ps->R.push_back((double)R.y); ps->R.push_back((double)R.z);
            }
          }
    }
  }
  ps->first[(size_t)N] = (int64_t)ps->i.size() + 1;
  return ps;
}

}  // namespace

// ================================================================ C ABI (ctypes)
extern "C" {

int nlo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

#define NLO_ANALYZE(SUF, T)                                                                                  \
  int nlo_analyze_cell_##SUF(const T* cell, T cutoff, T* inv, T* lens, int64_t* nc, int64_t* nxyz) {         \
    return analyze_cell<T>(cell, cutoff, inv, lens, nc, nxyz);                                               \
  }
NLO_ANALYZE(f32, float)
NLO_ANALYZE(f64, double)

#define NLO_SORTBASED(SUF, T, TI)                                                                            \
  void nlo_build_cells_##SUF(const T* X, int64_t N, const T* cell, const T* inv, T cutoff,                   \
                             const int64_t* nc, const uint8_t* pbc, T* Xs, TI* perm, TI* cell_id,            \
                             TI* cell_offsets) {                                                             \
    Geo<T, TI> g = make_geo<T, TI>(cell, inv, cutoff, nc, nullptr, pbc);                                     \
    build_cells<T, TI>(X, N, g, Xs, perm, cell_id, cell_offsets);                                            \
  }                                                                                                          \
  void nlo_count_pairs_##SUF(const T* X, int64_t N, const TI* perm, const TI* cell_offsets, const T* cell,   \
                             const T* inv, T cutoff, const int64_t* nc, const int64_t* nxyz,                 \
                             const uint8_t* pbc, int nthreads, TI* counts) {                                 \
    Geo<T, TI> g = make_geo<T, TI>(cell, inv, cutoff, nc, nxyz, pbc);                                        \
    count_pairs<T, TI>(X, N, perm, cell_offsets, g, nthreads, counts);                                       \
  }                                                                                                          \
  void nlo_pair_offsets_##SUF(const TI* counts, int64_t N, TI* offsets) { pair_offsets<TI>(counts, N, offsets); } \
  void nlo_fill_pairs_##SUF(const T* X, int64_t N, const TI* perm, const TI* cell_offsets, const T* cell,    \
                            const T* inv, T cutoff, const int64_t* nc, const int64_t* nxyz,                  \
                            const uint8_t* pbc, int nthreads, const TI* first, TI* io, TI* jo, TI* So,       \
                            T* Ro) {                                                                         \
    Geo<T, TI> g = make_geo<T, TI>(cell, inv, cutoff, nc, nxyz, pbc);                                        \
    fill_pairs<T, TI>(X, N, perm, cell_offsets, g, nthreads, first, io, jo, So, Ro);                         \
  }                                                                                                          \
  double nlo_lj_energy_##SUF(const T* X, int64_t N, const TI* perm, const TI* cell_offsets, const T* cell,   \
                             const T* inv, T cutoff, const int64_t* nc, const int64_t* nxyz,                 \
                             const uint8_t* pbc, int nthreads, double eps, double sigma) {                   \
    Geo<T, TI> g = make_geo<T, TI>(cell, inv, cutoff, nc, nxyz, pbc);                                        \
    return lj_energy<T, TI>(X, N, perm, cell_offsets, g, nthreads, eps, sigma);                              \
  }
NLO_SORTBASED(f32_i32, float, int32_t)
NLO_SORTBASED(f32_i64, float, int64_t)
NLO_SORTBASED(f64_i32, double, int32_t)
NLO_SORTBASED(f64_i64, double, int64_t)

void* nlo_legacy_f64(const double* X, int64_t N, const double* cell, const uint8_t* pbc, double cutoff, int fixcell) {
  return legacy_pairlist<double>(X, N, cell, pbc, cutoff, fixcell != 0);
}
void* nlo_legacy_f32(const float* X, int64_t N, const float* cell, const uint8_t* pbc, float cutoff, int fixcell) {
  return legacy_pairlist<float>(X, N, cell, pbc, cutoff, fixcell != 0);
}
void* nlo_brute_f64(const double* X, int64_t N, const double* cell, const uint8_t* pbc, double cutoff) {
  return brute_pairlist<double>(X, N, cell, pbc, cutoff);
}
void* nlo_brute_f32(const float* X, int64_t N, const float* cell, const uint8_t* pbc, float cutoff) {
  return brute_pairlist<float>(X, N, cell, pbc, cutoff);
}
int64_t nlo_pairset_npairs(void* h) { return (int64_t)((PairSet*)h)->i.size(); }
int64_t nlo_pairset_nsites(void* h) { return (int64_t)((PairSet*)h)->first.size() - 1; }
void nlo_pairset_copy(void* h, int64_t* i, int64_t* j, int64_t* S, int64_t* first, double* R, double* X, double* C) {
  PairSet* p = (PairSet*)h;
  if (i) std::copy(p->i.begin(), p->i.end(), i);
  if (j) std::copy(p->j.begin(), p->j.end(), j);
  if (S) std::copy(p->S.begin(), p->S.end(), S);
  if (first) std::copy(p->first.begin(), p->first.end(), first);
  if (R) std::copy(p->R.begin(), p->R.end(), R);
  if (X) std::copy(p->X.begin(), p->X.end(), X);
  if (C) std::copy(p->C.begin(), p->C.end(), C);
}
void nlo_pairset_free(void* h) { delete (PairSet*)h; }

}  // extern "C"
