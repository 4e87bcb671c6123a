// nl_mask.cuh -- the fast materialisation path: a counting pass that records, per atom, a BITMASK
// of which candidates of its 27-cell stencil are neighbours, and a fill pass that expands the
// masks into (i, j, S, R) rows without repeating a single distance test.
//
// Candidate numbering.  For a home cell the 27 neighbour cells are visited as 9 x-rows
// (dz outer, dy inner) of 3 x-adjacent cells; concatenating their atoms in sorted order gives the
// "flat" candidate list of that cell (independent of tile shape).  Bit f of an atom's mask says
// whether flat candidate f is within the cutoff.  Cells with more than 256 candidates take the
// generic per-atom route in both passes.
//
// Counting pass, Float64: deciding r2 < rc2 with the contract's Float64 arithmetic for all ~173
// candidates per atom would make the pass FP64-bound.  Instead every staged atom carries a Float32
// position RELATIVE TO THE TILE ORIGIN of its periodic IMAGE,
//     q = fl32( (x - O) + cell' * (s_loop - w) ),
// so that q_j - q_i approximates R = x_j - x_i + cell' * (s_loop + w_i - w_j) with an absolute error
// bounded by delta = 2^-22 * D + 1e-7 (D = tile extent; see mask_thresholds()).  A pair is a sure hit if
// r2~ < rc2 - E, a sure miss if r2~ > rc2 + E, and only pairs inside the band (about 1 in 10^5) are
// re-evaluated with the exact Float64 contract (exact_pair_hit).  Slots whose assumptions fail
// (|x| > 1e5, outside the tile extent on an open axis, winding overflow) are flagged and always take
// the exact path.  The decisions are therefore identical to the contract's.  The arithmetic runs on
// Blackwell's packed Float32 pipe (FFMA2 / FADD2: two home atoms per instruction).
// Float32 inputs evaluate the contract itself (packed, unfused) whenever the home atoms of a group
// share one winding number, and fall back to exact_pair_hit otherwise.
#pragma once
#include <cmath>
#include <type_traits>

#include "nl_tiled.cuh"

namespace nl {

constexpr int MASK_WORDS = 8;                  // 256 candidates
constexpr int MASK_MAXCAND = 32 * MASK_WORDS;

struct MaskThresholds { float mid, hw, dguard; int ok; };

// Error budget of the Float32 pre-filter (derivation in the header comment / DESIGN.md):
//   |q - q*| <= 2^-24 D + 2e-8 per coordinate (q* = exact image position), same for the home atom,
//   the Float32 subtraction adds 2^-24 * 2D, the contract's own Float64 rounding < 2e-8 (|x| <= 1e5):
//   delta = 2^-22 D + 1e-7;  |r2~ - r2_contract| <= 2 sqrt(3) rc delta + 3 delta^2 + 2^-22 rc2.  Doubled for margin.
inline MaskThresholds mask_thresholds(const double cell[9], const int nc[3], const TileShape& ts, double cutoff_sq) {
  MaskThresholds m;
  const int t[3] = {ts.tx, ts.ty, ts.tz};
  double D = 0;
  for (int k = 0; k < 3; k++) {
    double dk = 0;
    for (int r = 0; r < 3; r++) dk += std::fabs(cell[r + 3 * k]) * (double)(t[r] + 3) / (double)nc[r];
    D = dk > D ? dk : D;
  }
  const double rc = std::sqrt(cutoff_sq);
  const double delta = std::ldexp(D, -22) + 1e-7;
  const double E = 2.0 * (2.0 * std::sqrt(3.0) * rc * delta + 3.0 * delta * delta + std::ldexp(cutoff_sq, -21));
  // the kernel evaluates t = r2~ - mid in one FMA chain: sure hit <=> t < -hw, sure miss <=> t > hw
  m.mid = (float)cutoff_sq;
  m.hw = std::nextafterf((float)(E + std::fabs((double)m.mid - cutoff_sq)), INFINITY);
  m.dguard = (float)D;
  m.ok = (E < 0.01 * cutoff_sq && std::isfinite(D) && std::isfinite(E)) ? 1 : 0;
  return m;
}

template <class T, class TI> struct MaskArgs {
  Records<T> rec;
  const TI* co;
  long long n;
  Geo<T> g;
  Sinks<T, TI> out;
  uint32_t* masks;    // n * MASK_WORDS, sorted order
  uint8_t* cellflag;  // per cell: 1 if the count pass stored masks for its atoms
  int tx, ty, tz, ntx, nty, ntz;
  float mid, hw, dguard;
  const void* srow;   // fill pass: row start per sorted atom (k_row_starts)
  const int* zlayers; // launched tile layers along z (slab shards: the layers that can hold atoms); null = all ntz layers
  const MaskArgs<T, TI>* self;  // this struct in GLOBAL memory: the rare out-of-line paths read their inputs from there, so the
                                // kernels never copy their parameters onto the local-memory stack (measured: that copy cost
                                // 77 KB of local stores per CTA, 9 GB per launch on a slab of an 8x larger global grid)
};

__device__ __forceinline__ int pack_shift(int s0, int s1, int s2) { return (s0 + 1) | ((s1 + 1) << 2) | ((s2 + 1) << 4); }
__device__ __forceinline__ void unpack_shift(int p, long long s[3]) { s[0] = (p & 3) - 1; s[1] = ((p >> 2) & 3) - 1; s[2] = ((p >> 4) & 3) - 1; }

// The exact contract for one pair given global sorted indices; returns r2 < cutoff_sq.
template <class T, class TI>
__device__ __noinline__ bool exact_pair_hit(const MaskArgs<T, TI>* ad, long long gi, long long gj, int shp, T* r2_out = nullptr) {
  const Geo<T>& g = ad->g;
  const Records<T>& rec = ad->rec;
  const T xi = rec.px[gi], yi = rec.py[gi], zi = rec.pz[gi];
  const T xj = rec.px[gj], yj = rec.py[gj], zj = rec.pz[gj];
  long long wi[3], wj[3], sl[3];
  int cc[3];
  const uint32_t pwi = rec.pw[gi], pwj = rec.pw[gj];
  if (pwi & WIND_OVERFLOW) cell_of(g, xi, yi, zi, cc, wi); else unpack_wind(pwi, wi);
  if (pwj & WIND_OVERFLOW) cell_of(g, xj, yj, zj, cc, wj); else unpack_wind(pwj, wj);
  unpack_shift(shp, sl);
  const long long S[3] = {sl[0] + wi[0] - wj[0], sl[1] + wi[1] - wj[1], sl[2] + wi[2] - wj[2]};
  T R[3];
  const T r2 = pair_r2(g, xi, yi, zi, xj, yj, zj, S, R);
  if (r2_out) *r2_out = r2;
  return r2 < g.cutoff_sq;
}

// The same, also returning R (fused force sink, rare path).
template <class T, class TI>
__device__ __noinline__ bool exact_pair_R(const MaskArgs<T, TI>* ad, long long gi, long long gj, int shp, T* r2_out, T* R_out) {
  const Geo<T>& g = ad->g;
  const Records<T>& rec = ad->rec;
  const T xi = rec.px[gi], yi = rec.py[gi], zi = rec.pz[gi];
  const T xj = rec.px[gj], yj = rec.py[gj], zj = rec.pz[gj];
  long long wi[3], wj[3], sl[3];
  int cc[3];
  const uint32_t pwi = rec.pw[gi], pwj = rec.pw[gj];
  if (pwi & WIND_OVERFLOW) cell_of(g, xi, yi, zi, cc, wi); else unpack_wind(pwi, wi);
  if (pwj & WIND_OVERFLOW) cell_of(g, xj, yj, zj, cc, wj); else unpack_wind(pwj, wj);
  unpack_shift(shp, sl);
  const long long S[3] = {sl[0] + wi[0] - wj[0], sl[1] + wi[1] - wj[1], sl[2] + wi[2] - wj[2]};
  T R[3];
  const T r2 = pair_r2(g, xi, yi, zi, xj, yj, zj, S, R);
  *r2_out = r2;
  R_out[0] = R[0]; R_out[1] = R[1]; R_out[2] = R[2];
  return r2 < g.cutoff_sq;
}

// Float32 Lennard-Jones pair terms (fused force sink): phi = 4 eps (s^12 - s^6), gg = 24 eps (2 s^12 - s^6) / r2.
__device__ __forceinline__ void lj_pair_terms_f32(float eps4, float eps24, float sigma2, float r2, float& phi, float& gg) {
  float rc = __frcp_rn(r2);
  rc = rc * (2.0f - r2 * rc);  // one Newton step: ~1 ulp
  const float s2 = sigma2 * rc, s6 = s2 * s2 * s2, s12 = s6 * s6;
  phi = eps4 * (s12 - s6);
  gg = eps24 * (2.0f * s12 - s6) * rc;
}

// Shared tile prologue: virtual cell table (slot starts, global starts, packed shifts).
// Returns the total number of staged slots.
template <class T, class TI>
__device__ __forceinline__ int tile_table(const Geo<T>& g, const TI* __restrict__ co, int hx0, int hy0, int hz0, int VX, int VY, int NV,
                                          int* vstart, int* vgs, int* vsh, int* scan_sm) {
  const int tid = threadIdx.x;
  int cnt = 0, gs = 0, sh = 0;
  if (tid < NV) {
    int cx, cy, cz, s0, s1, s2;
    bool ok = map_virtual(hx0 + tid % VX - 1, g.nc[0], g.pbc[0], cx, s0);
    ok = map_virtual(hy0 + (tid / VX) % VY - 1, g.nc[1], g.pbc[1], cy, s1) && ok;
    ok = map_virtual(hz0 + tid / (VX * VY) - 1, g.nc[2], g.pbc[2], cz, s2) && ok;
    if (ok) {
      const long long cl = (long long)cx + (long long)g.nc[0] * ((long long)cy + (long long)g.nc[1] * cz);
      const long long c0 = (long long)co[cl], c1 = (long long)co[cl + 1];
      gs = (int)(c0 - 1);
      cnt = (int)(c1 - c0);
      sh = pack_shift(s0, s1, s2);
    }
  }
  int total;
  const int excl = block_excl_scan<int, TILE_NT>(cnt, scan_sm, &total);
  if (tid < NV) { vstart[tid] = excl; vgs[tid] = gs; vsh[tid] = sh; }
  if (tid == NV) vstart[NV] = total;
  return total;
}

__device__ __forceinline__ int find_vcell(const int* vstart, int NV, int sl) {
  int lo = 0, hi = NV;  // last v with vstart[v] <= sl
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (vstart[mid] <= sl) lo = mid; else hi = mid;
  }
  return lo;
}

// Rare path of the fill pass: atoms i and j carry different (or overflowed) winding numbers, so the
// shift is S = s_loop + w_i - w_j and cell' * S is not the per-cell table entry.  Kept out of line so
// that none of it is hoisted into the common path.
template <class T, class TI>
__device__ __noinline__ void slow_shift_and_R(const MaskArgs<T, TI>* ad, T xi, T yi, T zi, T xj, T yj, T zj, uint32_t wi, uint32_t wj, int* S012,
                                              T* R012) {
  const Geo<T>& g = ad->g;
  long long w_i[3], w_j[3];
  int cc[3];
  if (wi & WIND_OVERFLOW) cell_of(g, xi, yi, zi, cc, w_i); else unpack_wind(wi, w_i);
  if (wj & WIND_OVERFLOW) cell_of(g, xj, yj, zj, cc, w_j); else unpack_wind(wj, w_j);
  const long long S[3] = {S012[0] + w_i[0] - w_j[0], S012[1] + w_i[1] - w_j[1], S012[2] + w_i[2] - w_j[2]};
  T R[3];
  pair_r2(g, xi, yi, zi, xj, yj, zj, S, R);
  R012[0] = R[0]; R012[1] = R[1]; R012[2] = R[2];
  S012[0] = (int)S[0]; S012[1] = (int)S[1]; S012[2] = (int)S[2];
}

// Generic per-atom route for the atoms [g0, g0 + n) of one cell (out of line: rare).
template <class T, class TI, int MODE>
__device__ __noinline__ double generic_cell(const MaskArgs<T, TI>* ad, long long g0, int n, int lane) {
  double e = 0.0;
  for (int k = lane; k < n; k += 32) e += generic_atom<T, TI, MODE>(g0 + k, ad->rec, ad->co, ad->g, ad->out);
  return e;
}

// 4 eps ((sigma^2/r2)^6 - (sigma^2/r2)^3) in double; 1/r2 from the Float32 reciprocal plus one Newton step
// (relative error ~1e-14), which is much cheaper than an IEEE double division in the inner loop.
__device__ __forceinline__ double lj_term(double eps, double sigma2, float r2) {
  const double r2d = (double)r2;
  double rc = (double)__frcp_rn(r2);
  rc = rc * (2.0 - r2d * rc);
  const double s2 = sigma2 * rc, s6 = s2 * s2 * s2;
  return 4.0 * eps * (s6 * s6 - s6);
}

// Per-warp candidate tables of one home cell: flat candidate index -> staged slot / virtual cell.
struct CellTables {
  uint16_t* cslot;  // [MASK_MAXCAND]
  uint8_t* cv;      // [MASK_MAXCAND] virtual cell id (count pass) or stencil cell index 0..26 (fill pass)
};
constexpr int CELLTAB_BYTES = MASK_MAXCAND * 3;

// Builds the tables for home cell (lx, ly, lz) of the tile; returns the number of candidates
// (tables are valid only if it is <= MASK_MAXCAND).  Lane c < 27 owns neighbour cell
// c = (dz+1)*9 + (dy+1)*3 + (dx+1); flat order = cell order, then sorted order inside the cell.
template <bool STORE_STENCIL_INDEX, bool STORE_CV = true>
__device__ __forceinline__ int build_cell_tables(const int* vstart, int VX, int VY, int lx, int ly, int lz, int lane, const CellTables& t,
                                                 int cap = MASK_MAXCAND) {
  int v = 0, st = 0, cn = 0;
  if (lane < 27) {
    v = ((lz + lane / 9) * VY + (ly + (lane / 3) % 3)) * VX + (lx + lane % 3);
    st = vstart[v];
    cn = vstart[v + 1] - st;
  }
  const int incl = warp_incl_scan(cn, lane);
  const int ncand = __shfl_sync(FULL, incl, 31);
  if (ncand <= cap) {
    const int pre = incl - cn;
    const int mx = __reduce_max_sync(FULL, cn);
    for (int j = 0; j < mx; j++)
      if (j < cn) {
        t.cslot[pre + j] = (uint16_t)(st + j);
        if (STORE_CV) t.cv[pre + j] = (uint8_t)(STORE_STENCIL_INDEX ? lane : v);
      }
  }
  __syncwarp();
  return ncand;
}

__device__ __forceinline__ long long cell_linear(const int nc[3], int cx, int cy, int cz) {
  return (long long)cx + (long long)nc[0] * ((long long)cy + (long long)nc[1] * cz);
}

// ------------------------------------------------------------------------------------------------
// Counting pass, three sinks (CM):
//   CM_MASK  counts + hit masks + per-cell flags for the fill pass (candidate lists up to 256)
//   CM_COUNT counts only: the lazy count_neighbours sink (candidate lists up to 512, so denser systems stay on this path)
//   CM_LJ    fused Lennard-Jones energy over the hits, Float32 positions (BASELINE config 5); no counts, no masks
//   CM_LJF   fused Lennard-Jones forces + per-atom energies, Float32: every lane sums what the home atoms do to ITS candidate,
//            adds that to a per-slot accumulator in shared memory, and the tile flushes each slot with one vector atomic
enum { CM_MASK = 0, CM_COUNT = 1, CM_LJ = 2, CM_LJF = 3 };
#ifndef NL_LJF_SMEM_KB
#define NL_LJF_SMEM_KB 72
#endif
#ifndef NL_LJF_MINB
#define NL_LJF_MINB 3
#endif
__host__ __device__ constexpr int cm_tabcap(int cm) { return cm == CM_MASK ? MASK_MAXCAND : 2 * MASK_MAXCAND; }
__host__ __device__ constexpr int cm_smem_bytes(int cm) { return cm == CM_MASK ? 48 * 1024 : (cm == CM_LJF ? NL_LJF_SMEM_KB * 1024 : 64 * 1024); }
__host__ __device__ constexpr int cm_slot_bytes(int cm) { return cm == CM_LJF ? 32 : 16; }  // float4 position (+ float4 force/energy accumulator)
__host__ __device__ constexpr int cm_warp_bytes(int cm) { return cm_tabcap(cm) * 3 + MASK_WORDS * 34 * 4 + 16 * 32; }
__host__ __device__ constexpr int cm_fixed_bytes(int cm) { return 3 * TILE_VPAD * 4 + 64 * 4 + (TILE_NT / 32) * cm_warp_bytes(cm); }
__host__ __device__ constexpr int cm_cap(int cm) { return (cm_smem_bytes(cm) - cm_fixed_bytes(cm)) / cm_slot_bytes(cm) / 8 * 8; }
static_assert(cm_warp_bytes(CM_MASK) % 16 == 0 && cm_fixed_bytes(CM_MASK) % 16 == 0 && cm_warp_bytes(CM_COUNT) % 16 == 0, "alignment");
//
// Shared memory: tile tables | per-warp {cell tables, chunk-major mask words, home-atom pair buffer} |
// float4 per staged slot: Float64 -> (q.xyz = Float32 image-relative position, w = 1 if the slot needs
// the exact path); Float32 -> (absolute x, y, z, packed winding).
constexpr int CNT_CAP2 = cm_cap(CM_MASK);

constexpr float CAND_FAR = 1.0e18f;   // lanes beyond the candidate list: finite "nowhere" (squares stay finite in Float32)
constexpr float SLOT_FAR = -1.0e18f;  // staged slots that need the exact path: far from everything, including CAND_FAR

#ifndef NL_CNT_MINB
#define NL_CNT_MINB 4
#endif
template <class T, class TI, int CM>
__global__ void __launch_bounds__(TILE_NT, CM == CM_MASK ? NL_CNT_MINB : (CM == CM_LJF ? NL_LJF_MINB : 3)) k_count_mask(const MaskArgs<T, TI> a) {
  constexpr bool WANT_MASK = CM == CM_MASK;
  constexpr int TABCAP = cm_tabcap(CM);
  constexpr int CNT_WARP_BYTES = cm_warp_bytes(CM);
  constexpr int CAPSLOTS = cm_cap(CM);
  static_assert((CM != CM_LJ && CM != CM_LJF) || sizeof(T) == 4, "the fused LJ sinks of this kernel are Float32 only");
  constexpr bool LJ_ANY = CM == CM_LJ || CM == CM_LJF;
  // Hit words of a chunk: CM_MASK keeps them in the shared-memory rows it has to write out anyway (measured 2.75 vs 2.79 ms);
  // CM_COUNT needs no rows at all and keeps them in registers (lazy count 3.64 -> 3.39 ms at C5)
  constexpr bool REGW = CM == CM_COUNT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* vstart = (int*)smem_raw;
  int* vgs = vstart + TILE_VPAD;
  int* vsh = vgs + TILE_VPAD;
  int* hcell = vsh + TILE_VPAD;  // [64] packed (lx, ly, lz) of each home cell
  unsigned char* wbase = (unsigned char*)(hcell + 64);
  float4* sq = (float4*)(wbase + (TILE_NT / 32) * CNT_WARP_BYTES);
  float4* sF = sq + CAPSLOTS;  // CM_LJF: (force x, y, z, energy) accumulated per staged slot
  __shared__ int scan_sm[33];
  __shared__ int s_next;

  const Geo<T>& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned char* wb = wbase + wid * CNT_WARP_BYTES;
  CellTables tab;
  tab.cslot = (uint16_t*)wb;
  tab.cv = (uint8_t*)(wb + TABCAP * 2);
  uint32_t* mkT = (uint32_t*)(wb + TABCAP * 3);                             // [MASK_WORDS][34]: word (kc & 7) of home atom aa at (kc & 7)*34 + aa
  float* hb = (float*)(wb + TABCAP * 3 + MASK_WORDS * 34 * 4);              // [16 pairs][8]: x0 x1 y0 y1 z0 z1 f0 f1
  double e_acc = 0.0;                                                       // CM_LJ: this lane's share of the energy

  const int b = blockIdx.x;
  const int bz = b / (a.ntx * a.nty);
  const int hx0 = (b % a.ntx) * a.tx, hy0 = ((b / a.ntx) % a.nty) * a.ty, hz0 = (a.zlayers ? a.zlayers[bz] : bz) * a.tz;
  const int hxn = min(a.tx, g.nc[0] - hx0), hyn = min(a.ty, g.nc[1] - hy0), hzn = min(a.tz, g.nc[2] - hz0);
  const int VX = hxn + 2, VY = hyn + 2, VZ = hzn + 2, NV = VX * VY * VZ;
  {  // quick reject: tiles without a single home atom (a slab shard sees the global grid, mostly empty) cost two loads per row
    int nonempty = 0;
    if (tid < hyn * hzn) {
      const long long c0 = (long long)hx0 + (long long)g.nc[0] * ((long long)(hy0 + tid % hyn) + (long long)g.nc[1] * (hz0 + tid / hyn));
      nonempty = (long long)a.co[c0 + hxn] > (long long)a.co[c0];
    }
    if (!__syncthreads_or(nonempty)) return;
  }
  const int total = tile_table<T, TI>(g, a.co, hx0, hy0, hz0, VX, VY, NV, vstart, vgs, vsh, scan_sm);
  const int nhome = hxn * hyn * hzn;
  if (tid < nhome) hcell[tid] = (tid % hxn) | (((tid / hxn) % hyn) << 8) | ((tid / (hxn * hyn)) << 16);
  if (tid == 0) s_next = 0;
  __syncthreads();

  constexpr int GMODE = CM == CM_LJ ? MODE_LJ : (CM == CM_LJF ? MODE_LJF : MODE_COUNT);
  const bool tile_generic = total > CAPSLOTS;   // denser than the staging capacity: the whole tile takes the generic route
  if (tile_generic) {
    for (int hc = wid; hc < nhome; hc += TILE_NT / 32) {
      const int lx = hcell[hc] & 255, ly = (hcell[hc] >> 8) & 255, lz = hcell[hc] >> 16;
      const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
      const int nh = vstart[vh + 1] - vstart[vh];
      e_acc += generic_cell<T, TI, GMODE>(a.self, (long long)vgs[vh], nh, lane);
      if (WANT_MASK && lane == 0) a.cellflag[cell_linear(g.nc, hx0 + lx, hy0 + ly, hz0 + lz)] = 0;
    }
  }
  if (!tile_generic) {

  // ---- stage
  if constexpr (sizeof(T) == 8) {
    double O[3];  // tile origin: the corner of the home block, O = cell' * (h0 / n)
    {
      const double f0 = (double)hx0 / g.nc[0], f1 = (double)hy0 / g.nc[1], f2 = (double)hz0 / g.nc[2];
      O[0] = g.cell[0] * f0 + g.cell[1] * f1 + g.cell[2] * f2;
      O[1] = g.cell[3] * f0 + g.cell[4] * f1 + g.cell[5] * f2;
      O[2] = g.cell[6] * f0 + g.cell[7] * f1 + g.cell[8] * f2;
    }
    const double dg = (double)a.dguard;
    for (int sl = tid; sl < total; sl += TILE_NT) {
      const int v = find_vcell(vstart, NV, sl);
      const long long src = (long long)vgs[v] + (sl - vstart[v]);
      const double x = a.rec.px[src], y = a.rec.py[src], z = a.rec.pz[src];
      const uint32_t pw = a.rec.pw[src];
      const int sh = vsh[v];
      const double m0 = (double)(((sh & 3) - 1) - ((int)(pw & 1023u) - 512));
      const double m1 = (double)((((sh >> 2) & 3) - 1) - ((int)((pw >> 10) & 1023u) - 512));
      const double m2 = (double)((((sh >> 4) & 3) - 1) - ((int)((pw >> 20) & 1023u) - 512));
      const double q0 = (x - O[0]) + ((g.cell[0] * m0 + g.cell[1] * m1) + g.cell[2] * m2);
      const double q1 = (y - O[1]) + ((g.cell[3] * m0 + g.cell[4] * m1) + g.cell[5] * m2);
      const double q2 = (z - O[2]) + ((g.cell[6] * m0 + g.cell[7] * m1) + g.cell[8] * m2);
      const bool good = !(pw & WIND_OVERFLOW) && fabs(x) <= 1e5 && fabs(y) <= 1e5 && fabs(z) <= 1e5 && fabs(q0) <= dg && fabs(q1) <= dg &&
                        fabs(q2) <= dg;  // false for NaN too
      sq[sl] = good ? make_float4((float)q0, (float)q1, (float)q2, 0.f) : make_float4(SLOT_FAR, SLOT_FAR, SLOT_FAR, 1.f);
    }
  } else {
    for (int sl = tid; sl < total; sl += TILE_NT) {
      const int v = find_vcell(vstart, NV, sl);
      const long long src = (long long)vgs[v] + (sl - vstart[v]);
      sq[sl] = make_float4((float)a.rec.px[src], (float)a.rec.py[src], (float)a.rec.pz[src], __uint_as_float(a.rec.pw[src]));
      if (CM == CM_LJF) sF[sl] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();

  // Float64 pre-filter constants: t = r2~ - mid; sure hit <=> t < -hw; in the band <=> |t| <= hw
  const float mid = a.mid, hw = a.hw;
  const float2 nmid2 = make_float2(-mid, -mid);
  const float csqf = (float)g.cutoff_sq;
  const float lj_s2 = (float)a.out.lj_sigma2, lj_e4 = (float)(4.0 * a.out.lj_eps), lj_e24 = (float)(24.0 * a.out.lj_eps);

  while (true) {
    int hc = 0;
    if (lane == 0) hc = atomicAdd(&s_next, 1);
    hc = __shfl_sync(FULL, hc, 0);
    if (hc >= nhome) break;
    const int lx = hcell[hc] & 255, ly = (hcell[hc] >> 8) & 255, lz = hcell[hc] >> 16;
    const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
    const int hstart = vstart[vh], nh = vstart[vh + 1] - hstart;
    if (nh == 0) continue;
    const long long hg0 = vgs[vh];
    const bool need_gj = (sizeof(T) == 8) && WANT_MASK && a.out.half;  // half lists compare global sorted indices per candidate
    const int ncand = need_gj ? build_cell_tables<false, true>(vstart, VX, VY, lx, ly, lz, lane, tab, TABCAP)
                              : build_cell_tables<false, sizeof(T) == 4>(vstart, VX, VY, lx, ly, lz, lane, tab, TABCAP);
    if (WANT_MASK && lane == 0) a.cellflag[cell_linear(g.nc, hx0 + lx, hy0 + ly, hz0 + lz)] = ncand <= TABCAP ? 1 : 0;
    if (ncand > TABCAP) {  // too many candidates for the tables / the 256-bit masks: generic route (the fill pass does the same)
      e_acc += generic_cell<T, TI, GMODE>(a.self, hg0, nh, lane);
      continue;
    }
    const int nchunk = (ncand + 31) >> 5;
    int fh;  // flat index of home atom 0 as a candidate (cell 13 of the stencil): used to drop the self pair
    {
      int cn = 0;
      if (lane < 13) {
        const int v = ((lz + lane / 9) * VY + (ly + (lane / 3) % 3)) * VX + (lx + lane % 3);
        cn = vstart[v + 1] - vstart[v];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cn += __shfl_xor_sync(FULL, cn, o);
      fh = cn;
    }

    for (int g0 = 0; g0 < nh; g0 += 32) {
      const int ng = min(32, nh - g0);
      const int npair = (ng + 1) >> 1;
      __syncwarp();
      // home atoms -> pair-interleaved buffer; group-uniformity of the winding (Float32 path)
      bool my_bad = false;
      uint32_t my_w = 0;
      {
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < ng) p = sq[hstart + g0 + lane];
        float* d = hb + (lane >> 1) * 8 + (lane & 1);
        d[0] = p.x; d[2] = p.y; d[4] = p.z;
        if constexpr (sizeof(T) == 8) my_bad = (lane < ng) && p.w != 0.f; else my_w = __float_as_uint(p.w);
      }
      bool slow_group = false;   // every pair of this group must take the exact path
      uint32_t W0 = 0;
      if constexpr (sizeof(T) == 4) {
        W0 = __shfl_sync(FULL, my_w, 0);
        slow_group = __any_sync(FULL, lane < ng && (my_w != W0 || (my_w & WIND_OVERFLOW)));
      }
      const unsigned hbad = (sizeof(T) == 8) ? __ballot_sync(FULL, my_bad) : 0u;
      uint32_t cnt_acc = 0;  // lane aa < ng: neighbours of home atom aa so far
      __syncwarp();

      for (int kc = 0; kc < nchunk; kc++) {
        const int f = kc * 32 + lane;
        const bool valid = f < ncand;
        int gj = 0, shp = 0;
        float qx = CAND_FAR, qy = CAND_FAR, qz = CAND_FAR;
        float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f;
        bool cand_bad = false;
        if (valid) {
          const int slot = tab.cslot[f];
          const float4 q = sq[slot];
          qx = q.x; qy = q.y; qz = q.z;
          if constexpr (sizeof(T) == 8) {
            cand_bad = q.w != 0.f;  // the candidate's virtual cell (global index, shift) is only looked up on the rare paths below
          } else {
            const int v = tab.cv[f];
            gj = vgs[v] + (slot - vstart[v]);
            shp = vsh[v];
            const uint32_t wj = __float_as_uint(q.w);
            cand_bad = (wj & WIND_OVERFLOW) != 0;
            long long s[3], w0[3], w1[3];
            unpack_shift(shp, s);
            unpack_wind(W0, w0);
            unpack_wind(wj, w1);
            mtv(g.cell, (T)(s[0] + w0[0] - w1[0]), (T)(s[1] + w0[1] - w1[1]), (T)(s[2] + w0[2] - w1[2]), cs0, cs1, cs2);
          }
        }
        float lf0 = 0.f, lf1 = 0.f, lf2 = 0.f, lfe = 0.f;  // CM_LJF: what this group's home atoms do to the lane's candidate
        float bmin = 3.0e38f;  // min |t| over this lane's pairs: <= hw means some pair fell in the uncertainty band
        uint32_t* mrow = mkT + (kc & (MASK_WORDS - 1)) * 34;
        uint32_t myw = 0;  // REGW: lane aa keeps the hit word of home atom aa for this chunk in a register
        const int self_aa = valid ? (int)tab.cslot[f] - (hstart + g0) : -1;  // the home atom this candidate IS under zero shift (CM_LJ)

        if constexpr (sizeof(T) == 8) {
          const float2 nqx = make_float2(-qx, -qx), nqy = make_float2(-qy, -qy), nqz = make_float2(-qz, -qz);
#pragma unroll 2
          for (int pr = 0; pr < npair; pr++) {
            const float4 A = *(const float4*)(hb + pr * 8);
            const float2 Z = *(const float2*)(hb + pr * 8 + 4);
            const float2 dx = add2_rn(make_float2(A.x, A.y), nqx);
            const float2 dy = add2_rn(make_float2(A.z, A.w), nqy);
            const float2 dz = add2_rn(Z, nqz);
            float2 t = fma2_rn(dx, dx, nmid2);
            t = fma2_rn(dy, dy, t);
            t = fma2_rn(dz, dz, t);
            bmin = fminf(bmin, fminf(fabsf(t.x), fabsf(t.y)));
            const unsigned b0 = __ballot_sync(FULL, t.x < -hw);
            const unsigned b1 = __ballot_sync(FULL, t.y < -hw);
            if constexpr (REGW) myw = (lane >> 1) == pr ? ((lane & 1) ? b1 : b0) : myw;
            else *(uint2*)(mrow + 2 * pr) = make_uint2(b0, b1);  // every lane stores the same words: no branch
          }
        } else {
          const float2 qx2 = make_float2(qx, qx), qy2 = make_float2(qy, qy), qz2 = make_float2(qz, qz);
          const float2 c0 = make_float2(cs0, cs0), c1 = make_float2(cs1, cs1), c2 = make_float2(cs2, cs2);
#pragma unroll 2
          for (int pr = 0; pr < npair; pr++) {
            const float4 A = *(const float4*)(hb + pr * 8);
            const float2 Z = *(const float2*)(hb + pr * 8 + 4);
            // contract: R = (xj - xi) + cs ; r2 = (R0 R0 + R1 R1) + R2 R2, no fused operations
            const float2 R0 = add2_rn(add2_rn(qx2, make_float2(-A.x, -A.y)), c0);
            const float2 R1 = add2_rn(add2_rn(qy2, make_float2(-A.z, -A.w)), c1);
            const float2 R2 = add2_rn(add2_rn(qz2, make_float2(-Z.x, -Z.y)), c2);
            // packed multiplies, SCALAR adds: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with
            // explicit .rn and -fmad=false (verified in SASS), which would break the unfused contract
            const float2 p0 = mul2_rn(R0, R0), p1 = mul2_rn(R1, R1), p2 = mul2_rn(R2, R2);
            const float r2x = __fadd_rn(__fadd_rn(p0.x, p1.x), p2.x), r2y = __fadd_rn(__fadd_rn(p0.y, p1.y), p2.y);
            if (CM == CM_LJ) {
              if (!slow_group && !cand_bad) {  // otherwise the exact pass below accumulates this chunk
                if (valid && r2x < csqf && 2 * pr != self_aa) e_acc += lj_term(a.out.lj_eps, a.out.lj_sigma2, r2x);
                if (valid && r2y < csqf && 2 * pr + 1 != self_aa && 2 * pr + 1 < ng) e_acc += lj_term(a.out.lj_eps, a.out.lj_sigma2, r2y);
              }
            } else if (CM == CM_LJF) {
              if (!slow_group && !cand_bad) {
                if (valid && r2x < csqf && 2 * pr != self_aa) {
                  float phi, gg;
                  lj_pair_terms_f32(lj_e4, lj_e24, lj_s2, r2x, phi, gg);
                  lf0 += gg * R0.x; lf1 += gg * R1.x; lf2 += gg * R2.x; lfe += phi;
                }
                if (valid && r2y < csqf && 2 * pr + 1 != self_aa && 2 * pr + 1 < ng) {
                  float phi, gg;
                  lj_pair_terms_f32(lj_e4, lj_e24, lj_s2, r2y, phi, gg);
                  lf0 += gg * R0.y; lf1 += gg * R1.y; lf2 += gg * R2.y; lfe += phi;
                }
              }
            } else {
              const unsigned b0 = __ballot_sync(FULL, valid && r2x < csqf);
              const unsigned b1 = __ballot_sync(FULL, valid && r2y < csqf);
              if constexpr (REGW) myw = (lane >> 1) == pr ? ((lane & 1) ? b1 : b0) : myw;
            else *(uint2*)(mrow + 2 * pr) = make_uint2(b0, b1);  // every lane stores the same words: no branch
            }
          }
        }
        __syncwarp();
        // ---- rare: redo this chunk with the exact contract wherever the fast loop cannot be trusted
        const bool rare = __any_sync(FULL, bmin <= hw || cand_bad) || hbad != 0 || slow_group;
        if constexpr (sizeof(T) == 8) {
          if ((rare || need_gj) && valid) {
            const int slot = tab.cslot[f], v = need_gj ? (int)tab.cv[f] : find_vcell(vstart, NV, slot);
            gj = vgs[v] + (slot - vstart[v]);
            shp = vsh[v];
          }
        }
        if (rare) {
          for (int aa = 0; aa < ng; aa++) {
            bool hit;
            if constexpr (sizeof(T) == 8) {
              const float px = hb[(aa >> 1) * 8 + (aa & 1)], py = hb[(aa >> 1) * 8 + 2 + (aa & 1)], pz = hb[(aa >> 1) * 8 + 4 + (aa & 1)];
              const float dx = px - qx, dy = py - qy, dz = pz - qz;
              const float t = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmaf_rn(dx, dx, -mid)));
              hit = valid && t < -hw;
              if (valid && (cand_bad || ((hbad >> aa) & 1u) || fabsf(t) <= hw)) hit = exact_pair_hit<T, TI>(a.self, hg0 + g0 + aa, gj, shp);
            } else if (CM == CM_LJF) {
              hit = false;
              if (valid && (slow_group || cand_bad) && aa != self_aa) {
                T r2e, Re[3];
                if (exact_pair_R<T, TI>(a.self, hg0 + g0 + aa, gj, shp, &r2e, Re)) {
                  float phi, gg;
                  lj_pair_terms_f32(lj_e4, lj_e24, lj_s2, (float)r2e, phi, gg);
                  lf0 += gg * (float)Re[0]; lf1 += gg * (float)Re[1]; lf2 += gg * (float)Re[2]; lfe += phi;
                }
              }
            } else if (CM == CM_LJ) {
              hit = false;
              if (valid && (slow_group || cand_bad) && aa != self_aa) {  // exactly the pairs the fast loop skipped
                T r2e;
                if (exact_pair_hit<T, TI>(a.self, hg0 + g0 + aa, gj, shp, &r2e)) {
                  const double s2 = a.out.lj_sigma2 / (double)r2e, s6 = s2 * s2 * s2;
                  e_acc += 4.0 * a.out.lj_eps * (s6 * s6 - s6);
                }
              }
            } else {
              hit = valid && exact_pair_hit<T, TI>(a.self, hg0 + g0 + aa, gj, shp);
            }
            if (!LJ_ANY) {
              const unsigned bal = __ballot_sync(FULL, hit);
              if (REGW) { if (lane == aa) myw = bal; } else if (lane == 0) mrow[aa] = bal;
            }
          }
          __syncwarp();
        }
        if (WANT_MASK && a.out.half) {
          // half list (half_keep): a candidate later than the whole home group in sorted order is kept for every home atom,
          // an earlier one for none; candidates INSIDE the group (d = sorted distance to its first atom) are settled one by one
          const int d = valid ? gj - (int)(hg0 + g0) : -1;
          const int pos = ((shp & 3) > 1 || ((shp & 3) == 1 && (((shp >> 2) & 3) > 1 || (((shp >> 2) & 3) == 1 && ((shp >> 4) & 3) > 1)))) ? 1 : 0;
          unsigned word = __ballot_sync(FULL, valid && d >= ng);
          unsigned ingrp = __ballot_sync(FULL, valid && d >= 0 && d < ng);
          while (ingrp) {
            const int l = __ffs(ingrp) - 1;
            ingrp &= ingrp - 1;
            const int dl = __shfl_sync(FULL, d, l), pl = __shfl_sync(FULL, pos, l);
            if (dl > lane || (dl == lane && pl)) word |= 1u << l;
          }
          if (REGW) myw &= word; else if (lane < ng) mrow[lane] &= word;
        }
        // drop the self pair (same atom, zero shift): flat index fh + g0 + aa of home atom aa; accumulate the counts
        if (CM == CM_LJF && valid && (lf0 != 0.f || lf1 != 0.f || lf2 != 0.f || lfe != 0.f)) {
          float* d = (float*)(sF + tab.cslot[f]);  // lanes hold distinct slots; other warps may hit the same slot
          atomicAdd(d, lf0); atomicAdd(d + 1, lf1); atomicAdd(d + 2, lf2); atomicAdd(d + 3, lfe);
        }
        if (!LJ_ANY) {
          const int fs = fh + g0 + lane;
          if constexpr (REGW) {
            if ((fs >> 5) == kc) myw &= ~(1u << (fs & 31));
            if (lane >= ng) myw = 0;  // padding lanes (odd group sizes leave a phantom second atom in the last pair)
            cnt_acc += __popc(myw);
          } else if (lane < ng) {
            if ((fs >> 5) == kc) mrow[lane] &= ~(1u << (fs & 31));
            cnt_acc += __popc(mrow[lane]);
          }
        }
        if (CM != CM_MASK) __syncwarp();  // the mask rows are a ring of MASK_WORDS chunks
      }
      __syncwarp();
      // per-atom counts; masks to global (atom-major, MASK_WORDS per atom)
      if (!LJ_ANY && lane < ng) a.out.counts[a.rec.pidx[hg0 + g0 + lane]] = cnt_acc;
      if (WANT_MASK) {
        uint32_t* dst = a.masks + (hg0 + g0) * MASK_WORDS;
        for (int w = lane; w < ng * MASK_WORDS; w += 32) dst[w] = (w & 7) < nchunk ? mkT[(w & 7) * 34 + (w >> 3)] : 0u;
      }
    }
  }
  if (CM == CM_LJF) {
    // flush: one vector atomic per staged slot (home AND halo slots: a halo atom's other contributions come from its own tiles)
    __syncthreads();
    for (int sl = tid; sl < total; sl += TILE_NT) {
      const float4 f = sF[sl];
      if (f.x == 0.f && f.y == 0.f && f.z == 0.f && f.w == 0.f) continue;
      const int v = find_vcell(vstart, NV, sl);
      const uint32_t jo = a.rec.pidx[(long long)vgs[v] + (sl - vstart[v])];
      float* dst = (float*)a.out.fe + 4ll * jo;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w) : "memory");
    }
  }
  }  // !tile_generic
  if (CM == CM_LJ) {
    __shared__ double s_energy[TILE_NT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e_acc += __shfl_xor_sync(FULL, e_acc, o);
    if (lane == 0) s_energy[wid] = e_acc;
    __syncthreads();
    if (tid == 0) {
      double e = 0.0;
      for (int w = 0; w < TILE_NT / 32; w++) e += s_energy[w];
      if (e != 0.0) atomicAdd(a.out.energy, e);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fill pass: expands the masks.  Stages full records (positions in T, original index, winding).
// Four atoms' masks are compacted at once (lane = 8*atom + word); each atom's row is then produced
// by the whole warp, transposed through shared memory and written with contiguous full-sector stores.
#ifndef NL_FILL_SMEM_KB
#define NL_FILL_SMEM_KB 100
#endif
constexpr int FILL_SMEM_BYTES = NL_FILL_SMEM_KB * 1024;
constexpr int FILL_WARP_BYTES = CELLTAB_BYTES + 4 * MASK_MAXCAND + 32 * 3 * 4 + 32 * 3 * 8 + 28 * 3 * 8;  // tables | 4 hit lists | S stage | R stage | cs table
constexpr int FILL_FIXED_BYTES = 3 * TILE_VPAD * 4 + 64 * 4 + (TILE_NT / 32) * FILL_WARP_BYTES;
template <class T, class TI> __host__ __device__ constexpr int fill_cap() {
  return (FILL_SMEM_BYTES - FILL_FIXED_BYTES) / (TileRecBytes<T>::value + 4 + (int)sizeof(TI)) / 8 * 8;
}
static_assert(FILL_WARP_BYTES % 16 == 0 && FILL_FIXED_BYTES % 16 == 0, "alignment");

#ifndef NL_FILL_MINB
#define NL_FILL_MINB 2
#endif
template <class TI> struct FillBase { typedef typename std::conditional<sizeof(TI) == 4, uint32_t, unsigned long long>::type type; };

// 0-based start of the row of every SORTED atom (all ones: the atom gets no row -- halo atoms of a shard).
template <class T, class TI>
__global__ void __launch_bounds__(256) k_row_starts(const uint32_t* __restrict__ pidx, const TI* __restrict__ first, long long n, long long n_rows,
                                                    typename FillBase<TI>::type* __restrict__ srow) {
  typedef typename FillBase<TI>::type BaseT;
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t io = pidx[s];
  srow[s] = (long long)io < n_rows ? (BaseT)(first[io] - 1) : ~(BaseT)0;
}
template <class T, class TI>
__global__ void __launch_bounds__(TILE_NT, NL_FILL_MINB) k_fill_mask(const MaskArgs<T, TI> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CAP = fill_cap<T, TI>();
  int* vstart = (int*)smem_raw;
  int* vgs = vstart + TILE_VPAD;
  int* vsh = vgs + TILE_VPAD;
  int* hcell = vsh + TILE_VPAD;
  unsigned char* wbase = (unsigned char*)(hcell + 64);
  T* sx = (T*)(wbase + (TILE_NT / 32) * FILL_WARP_BYTES);
  T* sy = sx + CAP;
  T* sz = sy + CAP;
  uint32_t* sidx = (uint32_t*)(sz + CAP);
  uint32_t* sw = sidx + CAP;
  uint32_t* sgid = sw + CAP;  // shard mode: global index - 1 of each staged atom
  typedef typename FillBase<TI>::type BaseT;
  BaseT* sbase = (BaseT*)(sgid + CAP);  // home slots: 0-based start of the atom's row (all ones: no row); gathered once per tile
  __shared__ int scan_sm[33];
  __shared__ int s_next;

  const Geo<T>& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int grp = lane >> 3, sub = lane & 7;
  unsigned char* wb = wbase + wid * FILL_WARP_BYTES;
  CellTables tab;
  tab.cslot = (uint16_t*)wb;
  tab.cv = (uint8_t*)(wb + MASK_MAXCAND * 2);
  uint8_t* lists = wb + CELLTAB_BYTES;                                        // [4][MASK_MAXCAND]
  int* stS = (int*)(wb + CELLTAB_BYTES + 4 * MASK_MAXCAND);                   // [32][3] shifts of the current row segment
  T* stR = (T*)(wb + CELLTAB_BYTES + 4 * MASK_MAXCAND + 32 * 3 * 4);          // [32][3] R of the current row segment
  T* cst = (T*)(wb + CELLTAB_BYTES + 4 * MASK_MAXCAND + 32 * 3 * 4 + 32 * 3 * 8);  // [27][3] cell' * s_loop per stencil cell

  const int b = blockIdx.x;
  const int bz = b / (a.ntx * a.nty);
  const int hx0 = (b % a.ntx) * a.tx, hy0 = ((b / a.ntx) % a.nty) * a.ty, hz0 = (a.zlayers ? a.zlayers[bz] : bz) * a.tz;
  const int hxn = min(a.tx, g.nc[0] - hx0), hyn = min(a.ty, g.nc[1] - hy0), hzn = min(a.tz, g.nc[2] - hz0);
  const int VX = hxn + 2, VY = hyn + 2, VZ = hzn + 2, NV = VX * VY * VZ;
  {  // quick reject: tiles without a single home atom (a slab shard sees the global grid, mostly empty) cost two loads per row
    int nonempty = 0;
    if (tid < hyn * hzn) {
      const long long c0 = (long long)hx0 + (long long)g.nc[0] * ((long long)(hy0 + tid % hyn) + (long long)g.nc[1] * (hz0 + tid / hyn));
      nonempty = (long long)a.co[c0 + hxn] > (long long)a.co[c0];
    }
    if (!__syncthreads_or(nonempty)) return;
  }
  const int total = tile_table<T, TI>(g, a.co, hx0, hy0, hz0, VX, VY, NV, vstart, vgs, vsh, scan_sm);
  const int nhome = hxn * hyn * hzn;
  if (tid < nhome) hcell[tid] = (tid % hxn) | (((tid / hxn) % hyn) << 8) | ((tid / (hxn * hyn)) << 16);
  if (tid == 0) s_next = 0;
  __syncthreads();

  if (total > CAP) {
    // denser than the staging capacity: the generic route needs no masks
    for (int hc = wid; hc < nhome; hc += TILE_NT / 32) {
      const int lx = hcell[hc] & 255, ly = (hcell[hc] >> 8) & 255, lz = hcell[hc] >> 16;
      const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
      const int nh = vstart[vh + 1] - vstart[vh];
      generic_cell<T, TI, MODE_FILL>(a.self, (long long)vgs[vh], nh, lane);
    }
    return;
  }
  for (int sl = tid; sl < total; sl += TILE_NT) {
    const int v = find_vcell(vstart, NV, sl);
    const long long src = (long long)vgs[v] + (sl - vstart[v]);
    sx[sl] = a.rec.px[src];
    sy[sl] = a.rec.py[src];
    sz[sl] = a.rec.pz[src];
    sidx[sl] = a.rec.pidx[src];
    sw[sl] = a.rec.pw[src];
    if (a.out.pgid0) sgid[sl] = a.out.pgid0[src];
    sbase[sl] = ((const BaseT*)a.srow)[src];  // row start (k_row_starts): coalesced, like the rest of the record
  }
  if (tid < nhome) {  // per-cell "masks valid" flags, fetched once per tile
    const int lx = hcell[tid] & 255, ly = (hcell[tid] >> 8) & 255, lz = (hcell[tid] >> 16) & 255;
    if (a.cellflag[cell_linear(g.nc, hx0 + lx, hy0 + ly, hz0 + lz)]) hcell[tid] |= 1 << 24;
  }
  __syncthreads();
  const bool use_gid = a.out.pgid0 != nullptr;

  while (true) {
    int hc = 0;
    if (lane == 0) hc = atomicAdd(&s_next, 1);
    hc = __shfl_sync(FULL, hc, 0);
    if (hc >= nhome) break;
    const int lx = hcell[hc] & 255, ly = (hcell[hc] >> 8) & 255, lz = (hcell[hc] >> 16) & 255;
    const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
    const int hstart = vstart[vh], nh = vstart[vh + 1] - hstart;
    if (nh == 0) continue;
    const long long hg0 = vgs[vh];
    if (!(hcell[hc] >> 24)) {
      generic_cell<T, TI, MODE_FILL>(a.self, hg0, nh, lane);
      continue;
    }
    // mask words of the first two passes (lane = 8 * atom + word): issued before the table building so that it hides them
    uint32_t nx_word = 0, nx2_word = 0;
    if (grp < nh && sbase[hstart + grp] != ~(BaseT)0) nx_word = a.masks[(hg0 + grp) * MASK_WORDS + sub];
    if (4 + grp < nh && sbase[hstart + 4 + grp] != ~(BaseT)0) nx2_word = a.masks[(hg0 + 4 + grp) * MASK_WORDS + sub];
    build_cell_tables<true>(vstart, VX, VY, lx, ly, lz, lane, tab);
    // lane c < 27: packed periodic shift of stencil cell c and its shift vector cs = cell' * s_loop (contract arithmetic)
    int my_shp = 0;
    if (lane < 27) {
      const int v = ((lz + lane / 9) * VY + (ly + (lane / 3) % 3)) * VX + (lx + lane % 3);
      my_shp = vsh[v];
      T c0, c1, c2;
      mtv(g.cell, (T)((my_shp & 3) - 1), (T)(((my_shp >> 2) & 3) - 1), (T)(((my_shp >> 4) & 3) - 1), c0, c1, c2);
      cst[3 * lane] = c0; cst[3 * lane + 1] = c1; cst[3 * lane + 2] = c2;
    }
    __syncwarp();


    for (int a0 = 0; a0 < nh; a0 += 4) {
      // ---- four atoms at once: lane = 8 * atom + mask word
      uint32_t word = nx_word;
      uint32_t my_io = 0;
      long long my_base = 0;
      if (a0 + grp < nh) {
        my_io = sidx[hstart + a0 + grp];
        my_base = (long long)sbase[hstart + a0 + grp];   // all ones (no row) comes with word == 0
      }
      nx_word = nx2_word;
      nx2_word = 0;
      if (a0 + 8 + grp < nh && sbase[hstart + a0 + 8 + grp] != ~(BaseT)0)  // two passes ahead
        nx2_word = a.masks[(hg0 + a0 + 8 + grp) * MASK_WORDS + sub];
      const int pc = __popc(word);
      int incl = pc;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o, 8);
        if (sub >= o) incl += t;
      }
      const int my_nhit = __shfl_sync(FULL, incl, 7, 8);
      __syncwarp();
      {
        uint8_t* L = lists + grp * MASK_MAXCAND + (incl - pc);
        const int fb = sub * 32;
        while (word) {
          const int bit = __ffs(word) - 1;
          word &= word - 1;
          *L++ = (uint8_t)(fb + bit);
        }
      }
      __syncwarp();

      const int na = min(4, nh - a0);
      for (int q = 0; q < na; q++) {
        const int hs = hstart + a0 + q;
        const int nhit = __shfl_sync(FULL, my_nhit, q * 8);
        if (nhit == 0) continue;
        const uint32_t io = __shfl_sync(FULL, my_io, q * 8);
        const long long base = __shfl_sync(FULL, my_base, q * 8);
        const TI io_out = use_gid ? (TI)sgid[hs] + 1 : (TI)io + 1;
        const T xi = sx[hs], yi = sy[hs], zi = sz[hs];
        const uint32_t wi = sw[hs];
        const uint8_t* L = lists + q * MASK_MAXCAND;
        TI* const io_row = a.out.io + base;
        TI* const jo_row = a.out.jo + base;
        TI* const So_row = a.out.So + 3 * base;
        T* const Ro_row = a.out.Ro ? a.out.Ro + 3 * base : nullptr;

        for (int r0 = 0; r0 < nhit; r0 += 32) {
          const int r = r0 + lane;
          const int nr = min(32, nhit - r0);
          const bool act = r < nhit;
          const int f = act ? (int)L[r] : 0;
          const int slot = tab.cslot[f], c = tab.cv[f];  // c: stencil cell index (dz+1)*9 + (dy+1)*3 + (dx+1)
          const int shp = __shfl_sync(FULL, my_shp, c);
          if (act) {
            const T xj = sx[slot], yj = sy[slot], zj = sz[slot];
            const uint32_t wj = sw[slot];
            int S0 = (shp & 3) - 1, S1 = ((shp >> 2) & 3) - 1, S2 = ((shp >> 4) & 3) - 1;
            T R0, R1, R2;
            if (wi == wj && !(wi & WIND_OVERFLOW)) {
              R0 = add_rn(sub_rn(xj, xi), cst[3 * c]);
              R1 = add_rn(sub_rn(yj, yi), cst[3 * c + 1]);
              R2 = add_rn(sub_rn(zj, zi), cst[3 * c + 2]);
            } else {
              int S3[3] = {S0, S1, S2};
              T R3[3];
              slow_shift_and_R<T, TI>(a.self, xi, yi, zi, xj, yj, zj, wi, wj, S3, R3);
              S0 = S3[0]; S1 = S3[1]; S2 = S3[2];
              R0 = R3[0]; R1 = R3[1]; R2 = R3[2];
            }
            stS[3 * lane] = S0; stS[3 * lane + 1] = S1; stS[3 * lane + 2] = S2;
            if (Ro_row) { stR[3 * lane] = R0; stR[3 * lane + 1] = R1; stR[3 * lane + 2] = R2; }
            io_row[r] = io_out;
            jo_row[r] = use_gid ? (TI)sgid[slot] + 1 : (TI)sidx[slot] + 1;
          }
          __syncwarp();
          // transposed, contiguous stores of the row segment [r0, r0 + nr): 3 nr words of S, 3 nr of R
          const int nw = 3 * nr;
#pragma unroll
          for (int m = 0; m < 3; m++) {
            const int w = m * 32 + lane;
            if (w < nw) {
              So_row[3 * r0 + w] = (TI)stS[w];
              if (Ro_row) Ro_row[3 * r0 + w] = stR[w];
            }
          }
          __syncwarp();
        }
      }
    }
  }
}

template <class T, class TI>
inline void mask_args(MaskArgs<T, TI>& a, int64_t n, const TI* co, const Records<T>& rec, const Geo<T>& g, const Sinks<T, TI>& sk,
                      const TileShape& ts, uint32_t* masks) {
  a.rec = rec; a.co = co; a.n = n; a.g = g; a.out = sk; a.masks = masks; a.cellflag = nullptr; a.self = nullptr; a.srow = nullptr; a.zlayers = nullptr;
  a.tx = ts.tx; a.ty = ts.ty; a.tz = ts.tz;
  a.ntx = (g.nc[0] + ts.tx - 1) / ts.tx; a.nty = (g.nc[1] + ts.ty - 1) / ts.ty; a.ntz = (g.nc[2] + ts.tz - 1) / ts.tz;
  a.mid = a.hw = a.dguard = 0.f;
}

}  // namespace nl
