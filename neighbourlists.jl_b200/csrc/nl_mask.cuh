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
// (|x| > 1e5, outside the tile extent on an open axis, winding overflow) are staged as NaN, which
// lands every comparison in the band.  The decisions are therefore identical to the contract's.
// Float32 inputs evaluate the contract directly (Float32 is already the native fast path).
#pragma once
#include <cmath>

#include "nl_tiled.cuh"

namespace nl {

constexpr int MASK_WORDS = 8;                  // 256 candidates
constexpr int MASK_MAXCAND = 32 * MASK_WORDS;
constexpr int CNT_SMEM_BYTES = 48 * 1024;      // count pass: 16 B per staged slot
constexpr int CNT_CAP = (CNT_SMEM_BYTES - 3 * TILE_VPAD * 4 - (TILE_NT / 32) * 32 * MASK_WORDS * 4) / 16 / 8 * 8;

struct MaskThresholds { float lo, hi, dguard; int ok; };

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
  m.lo = std::nextafterf((float)(cutoff_sq - E), -INFINITY);
  m.hi = std::nextafterf((float)(cutoff_sq + E), INFINITY);
  m.dguard = (float)D;
  m.ok = (E < 0.01 * cutoff_sq && std::isfinite(D) && m.lo > 0.0f) ? 1 : 0;
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
  float lo, hi, dguard;
};

__device__ __forceinline__ int pack_shift(int s0, int s1, int s2) { return (s0 + 1) | ((s1 + 1) << 2) | ((s2 + 1) << 4); }
__device__ __forceinline__ void unpack_shift(int p, long long s[3]) { s[0] = (p & 3) - 1; s[1] = ((p >> 2) & 3) - 1; s[2] = ((p >> 4) & 3) - 1; }

// The exact contract for one pair given global sorted indices; returns r2 < cutoff_sq.
template <class T>
__device__ __noinline__ bool exact_pair_hit(const Geo<T>& g, const Records<T>& rec, long long gi, long long gj, int shp) {
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
  return pair_r2(g, xi, yi, zi, xj, yj, zj, S, R) < g.cutoff_sq;
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

// ------------------------------------------------------------------------------------------------
// Counting pass.  WANT_MASK: also store the hit masks for the fill pass.
template <class T, class TI, bool WANT_MASK>
__global__ void __launch_bounds__(TILE_NT, 4) k_count_mask(const MaskArgs<T, TI> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* vstart = (int*)smem_raw;
  int* vgs = vstart + TILE_VPAD;
  int* vsh = vgs + TILE_VPAD;
  uint32_t* mk_all = (uint32_t*)(vsh + TILE_VPAD);                   // [8 warps][32 atoms][MASK_WORDS]
  float4* sq = (float4*)(mk_all + (TILE_NT / 32) * 32 * MASK_WORDS);  // [CNT_CAP]
  __shared__ int scan_sm[33];
  __shared__ int s_next;

  const Geo<T>& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t* mk = mk_all + wid * 32 * MASK_WORDS;

  const int b = blockIdx.x;
  const int hx0 = (b % a.ntx) * a.tx, hy0 = ((b / a.ntx) % a.nty) * a.ty, hz0 = (b / (a.ntx * a.nty)) * a.tz;
  const int hxn = min(a.tx, g.nc[0] - hx0), hyn = min(a.ty, g.nc[1] - hy0), hzn = min(a.tz, g.nc[2] - hz0);
  const int VX = hxn + 2, VY = hyn + 2, VZ = hzn + 2, NV = VX * VY * VZ;
  const int total = tile_table<T, TI>(g, a.co, hx0, hy0, hz0, VX, VY, NV, vstart, vgs, vsh, scan_sm);
  if (tid == 0) s_next = 0;
  __syncthreads();
  const int nhome = hxn * hyn * hzn;

  if (total > CNT_CAP) {
    for (int hc = wid; hc < nhome; hc += TILE_NT / 32) {
      const int lx = hc % hxn, ly = (hc / hxn) % hyn, lz = hc / (hxn * hyn);
      const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
      const int nh = vstart[vh + 1] - vstart[vh];
      for (int k = lane; k < nh; k += 32) generic_atom<T, TI, MODE_COUNT>((long long)vgs[vh] + k, a.rec, a.co, g, a.out);
      if (WANT_MASK && lane == 0) a.cellflag[(long long)(hx0 + lx) + (long long)g.nc[0] * ((hy0 + ly) + (long long)g.nc[1] * (hz0 + lz))] = 0;
    }
    return;
  }

  // ---- stage
  if (sizeof(T) == 8) {
    // tile origin: the corner of the home block, O = cell' * (h0 / n)
    double O[3];
    {
      const double f0 = (double)hx0 / g.nc[0], f1 = (double)hy0 / g.nc[1], f2 = (double)hz0 / g.nc[2];
      for (int k = 0; k < 3; k++) O[k] = (double)g.cell[3 * k] * f0 + (double)g.cell[3 * k + 1] * f1 + (double)g.cell[3 * k + 2] * f2;
    }
    const float nanf_ = __int_as_float(0x7fc00000);
    for (int sl = tid; sl < total; sl += TILE_NT) {
      const int v = find_vcell(vstart, NV, sl);
      const long long src = (long long)vgs[v] + (sl - vstart[v]);
      const double x = (double)a.rec.px[src], y = (double)a.rec.py[src], z = (double)a.rec.pz[src];
      const uint32_t pw = a.rec.pw[src];
      long long w[3], s[3];
      unpack_wind(pw, w);
      unpack_shift(vsh[v], s);
      const double m0 = (double)(s[0] - w[0]), m1 = (double)(s[1] - w[1]), m2 = (double)(s[2] - w[2]);
      const double q0 = (x - O[0]) + (((double)g.cell[0] * m0 + (double)g.cell[1] * m1) + (double)g.cell[2] * m2);
      const double q1 = (y - O[1]) + (((double)g.cell[3] * m0 + (double)g.cell[4] * m1) + (double)g.cell[5] * m2);
      const double q2 = (z - O[2]) + (((double)g.cell[6] * m0 + (double)g.cell[7] * m1) + (double)g.cell[8] * m2);
      const double dg = (double)a.dguard;
      const bool good = !(pw & WIND_OVERFLOW) && fabs(x) <= 1e5 && fabs(y) <= 1e5 && fabs(z) <= 1e5 && fabs(q0) <= dg && fabs(q1) <= dg &&
                        fabs(q2) <= dg;
      sq[sl] = good ? make_float4((float)q0, (float)q1, (float)q2, 0.f) : make_float4(nanf_, nanf_, nanf_, 0.f);
    }
  } else {
    for (int sl = tid; sl < total; sl += TILE_NT) {
      const int v = find_vcell(vstart, NV, sl);
      const long long src = (long long)vgs[v] + (sl - vstart[v]);
      sq[sl] = make_float4((float)a.rec.px[src], (float)a.rec.py[src], (float)a.rec.pz[src], __uint_as_float(a.rec.pw[src]));
    }
  }
  __syncthreads();

  const float inff_ = __int_as_float(0x7f800000);

  while (true) {
    int hc = 0;
    if (lane == 0) hc = atomicAdd(&s_next, 1);
    hc = __shfl_sync(FULL, hc, 0);
    if (hc >= nhome) break;
    const int lx = hc % hxn, ly = (hc / hxn) % hyn, lz = hc / (hxn * hyn);
    const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
    const int hstart = vstart[vh], nh = vstart[vh + 1] - hstart;
    if (nh == 0) continue;
    const long long hg0 = vgs[vh];

    int rstart = 0, rlen = 0;
    if (lane < 9) {
      const int vrow = ((lz + lane / 3) * VY + (ly + lane % 3)) * VX + lx;
      rstart = vstart[vrow];
      rlen = vstart[vrow + 3] - rstart;
    }
    const int rincl = warp_incl_scan(rlen, lane);
    const int ncand = __shfl_sync(FULL, rincl, 8);

    if (WANT_MASK && lane == 0)
      a.cellflag[(long long)(hx0 + lx) + (long long)g.nc[0] * ((hy0 + ly) + (long long)g.nc[1] * (hz0 + lz))] = ncand <= MASK_MAXCAND ? 1 : 0;
    if (ncand > MASK_MAXCAND) {  // too many candidates for a 256-bit mask: generic route (the fill pass does the same)
      for (int k = lane; k < nh; k += 32) generic_atom<T, TI, MODE_COUNT>(hg0 + k, a.rec, a.co, g, a.out);
      continue;
    }
    const int nchunk = (ncand + 31) >> 5;

    for (int g0 = 0; g0 < nh; g0 += 32) {
      const int ng = min(32, nh - g0);
      __syncwarp();
      for (int k0 = 0, kc = 0; k0 < ncand; k0 += 32, kc++) {
        const int f = k0 + lane;
        const bool valid = f < ncand;
        int rr = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) rr += (f >= __shfl_sync(FULL, rincl, r)) ? 1 : 0;
        const int r_incl = __shfl_sync(FULL, rincl, rr);
        const int r_len = __shfl_sync(FULL, rlen, rr);
        const int r_start = __shfl_sync(FULL, rstart, rr);
        int slot = -1, gj = 0, shp = 0;
        float qx = inff_, qy = inff_, qz = inff_;  // +inf: an invalid lane is a sure miss
        float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f;
        uint32_t wj = 0;
        if (valid) {
          slot = r_start + (f - (r_incl - r_len));
          const int vrow = ((lz + rr / 3) * VY + (ly + rr % 3)) * VX + lx;
          const int v = vrow + (slot >= vstart[vrow + 1] ? 1 : 0) + (slot >= vstart[vrow + 2] ? 1 : 0);
          gj = vgs[v] + (slot - vstart[v]);
          shp = vsh[v];
          const float4 q = sq[slot];
          qx = q.x; qy = q.y; qz = q.z;
          if constexpr (sizeof(T) == 4) {
            wj = __float_as_uint(q.w);
            long long s[3];
            unpack_shift(shp, s);
            mtv(g.cell, (float)s[0], (float)s[1], (float)s[2], cs0, cs1, cs2);
          }
        }
        const int self_aa = slot - (hstart + g0);  // the home atom this lane's candidate IS (zero shift), if in [0, ng)
        uint32_t umask = 0;                        // home atoms whose pair with this candidate needs the exact test

#pragma unroll 4
        for (int aa = 0; aa < ng; aa++) {
          const float4 p = sq[hstart + g0 + aa];
          bool sure, unsure;
          if (sizeof(T) == 8) {
            const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
            const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
            const bool ge = !(r2 < a.lo);          // true for r2 >= lo and for NaN
            unsure = ge && !(r2 > a.hi);
            sure = !ge && (aa != self_aa);
          } else {
            const uint32_t wi = __float_as_uint(p.w);
            const float R0 = __fadd_rn(__fsub_rn(qx, p.x), cs0), R1 = __fadd_rn(__fsub_rn(qy, p.y), cs1), R2 = __fadd_rn(__fsub_rn(qz, p.z), cs2);
            const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(R0, R0), __fmul_rn(R1, R1)), __fmul_rn(R2, R2));
            const bool same_w = (wi == wj) && !(wi & WIND_OVERFLOW);
            unsure = valid && !same_w;
            sure = same_w && (r2 < (float)g.cutoff_sq) && (aa != self_aa);
          }
          if (unsure) umask |= 1u << aa;
          const unsigned bal = __ballot_sync(FULL, sure);
          if (lane == 0) mk[aa * MASK_WORDS + kc] = bal;
        }
        // deferred exact evaluations (rare): OR the confirmed hits into the masks
        if (__any_sync(FULL, umask != 0)) {
          __syncwarp();
          while (umask) {
            const int aa = __ffs(umask) - 1;
            umask &= umask - 1;
            if (valid && aa != self_aa && exact_pair_hit<T>(g, a.rec, hg0 + g0 + aa, gj, shp)) atomicOr(&mk[aa * MASK_WORDS + kc], 1u << lane);
          }
        }
      }
      __syncwarp();
      // per-atom counts from the masks; masks to global
      if (lane < ng) {
        uint32_t c = 0;
        for (int k = 0; k < nchunk; k++) c += __popc(mk[lane * MASK_WORDS + k]);
        a.out.counts[a.rec.pidx[hg0 + g0 + lane]] = c;
      }
      if (WANT_MASK) {
        uint32_t* dst = a.masks + (hg0 + g0) * MASK_WORDS;
        for (int w = lane; w < ng * MASK_WORDS; w += 32) dst[w] = mk[w];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fill pass: expands the masks.  Stages full records (positions in T, original index, winding).
constexpr int FILL_SMEM_BYTES = 56 * 1024;
template <class T> __host__ __device__ constexpr int fill_cap() { return (FILL_SMEM_BYTES - 3 * TILE_VPAD * 4) / TileRecBytes<T>::value / 8 * 8; }

template <class T, class TI>
__global__ void __launch_bounds__(TILE_NT, 3) k_fill_mask(const MaskArgs<T, TI> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CAP = fill_cap<T>();
  int* vstart = (int*)smem_raw;
  int* vgs = vstart + TILE_VPAD;
  int* vsh = vgs + TILE_VPAD;
  T* sx = (T*)(vsh + TILE_VPAD);
  T* sy = sx + CAP;
  T* sz = sy + CAP;
  uint32_t* sidx = (uint32_t*)(sz + CAP);
  uint32_t* sw = sidx + CAP;
  __shared__ int scan_sm[33];
  __shared__ int s_next;
  __shared__ uint32_t s_mk[TILE_NT / 32][32 * MASK_WORDS];
  __shared__ uint8_t s_list[TILE_NT / 32][MASK_MAXCAND];

  const Geo<T>& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;

  const int b = blockIdx.x;
  const int hx0 = (b % a.ntx) * a.tx, hy0 = ((b / a.ntx) % a.nty) * a.ty, hz0 = (b / (a.ntx * a.nty)) * a.tz;
  const int hxn = min(a.tx, g.nc[0] - hx0), hyn = min(a.ty, g.nc[1] - hy0), hzn = min(a.tz, g.nc[2] - hz0);
  const int VX = hxn + 2, VY = hyn + 2, VZ = hzn + 2, NV = VX * VY * VZ;
  const int total = tile_table<T, TI>(g, a.co, hx0, hy0, hz0, VX, VY, NV, vstart, vgs, vsh, scan_sm);
  if (tid == 0) s_next = 0;
  __syncthreads();
  const int nhome = hxn * hyn * hzn;

  if (total > CAP) {
    // denser than the staging capacity: the generic route needs no masks
    for (int hc = wid; hc < nhome; hc += TILE_NT / 32) {
      const int vh = ((hc / (hxn * hyn) + 1) * VY + ((hc / hxn) % hyn + 1)) * VX + (hc % hxn + 1);
      const int nh = vstart[vh + 1] - vstart[vh];
      for (int k = lane; k < nh; k += 32) generic_atom<T, TI, MODE_FILL>((long long)vgs[vh] + k, a.rec, a.co, g, a.out);
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
  }
  __syncthreads();

  while (true) {
    int hc = 0;
    if (lane == 0) hc = atomicAdd(&s_next, 1);
    hc = __shfl_sync(FULL, hc, 0);
    if (hc >= nhome) break;
    const int lx = hc % hxn, ly = (hc / hxn) % hyn, lz = hc / (hxn * hyn);
    const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
    const int hstart = vstart[vh], nh = vstart[vh + 1] - hstart;
    if (nh == 0) continue;
    const long long hg0 = vgs[vh];

    int rstart = 0, rlen = 0;
    if (lane < 9) {
      const int vrow = ((lz + lane / 3) * VY + (ly + lane % 3)) * VX + lx;
      rstart = vstart[vrow];
      rlen = vstart[vrow + 3] - rstart;
    }
    const int rincl = warp_incl_scan(rlen, lane);
    const int ncand = __shfl_sync(FULL, rincl, 8);
    const int have_mask = a.cellflag[(long long)(hx0 + lx) + (long long)g.nc[0] * ((hy0 + ly) + (long long)g.nc[1] * (hz0 + lz))];
    if (!have_mask) {
      for (int k = lane; k < nh; k += 32) generic_atom<T, TI, MODE_FILL>(hg0 + k, a.rec, a.co, g, a.out);
      continue;
    }
    const int nchunk = (ncand + 31) >> 5;

    for (int g0 = 0; g0 < nh; g0 += 32) {
      const int ng = min(32, nh - g0);
      // masks of the whole group -> shared memory (coalesced), row bases gathered in parallel
      __syncwarp();
      {
        const uint32_t* src = a.masks + (hg0 + g0) * MASK_WORDS;
        for (int w = lane; w < ng * MASK_WORDS; w += 32) s_mk[wid][w] = src[w];
      }
      uint32_t my_io = 0;
      long long my_base = 0;
      if (lane < ng) {
        my_io = sidx[hstart + g0 + lane];
        my_base = (long long)a.out.first[my_io] - 1;
      }
      __syncwarp();

      for (int aa = 0; aa < ng; aa++) {
        const int hs = hstart + g0 + aa;
        // ---- this atom's mask -> dense list of flat candidate indices
        const uint32_t word = lane < nchunk ? s_mk[wid][aa * MASK_WORDS + lane] : 0u;
        const int pc = __popc(word);
        const int incl = warp_incl_scan(pc, lane);
        const int nhit = __shfl_sync(FULL, incl, 31);
        if (nhit == 0) continue;
        __syncwarp();
        for (int k = 0; k < nchunk; k++) {
          const uint32_t wk = __shfl_sync(FULL, word, k);
          const int pre = __shfl_sync(FULL, incl - pc, k);
          if ((wk >> lane) & 1u) s_list[wid][pre + __popc(wk & lt)] = (uint8_t)(k * 32 + lane);
        }
        __syncwarp();

        const T xi = sx[hs], yi = sy[hs], zi = sz[hs];
        const uint32_t wi = sw[hs];
        const uint32_t io = __shfl_sync(FULL, my_io, aa);
        const long long base = __shfl_sync(FULL, my_base, aa);

        for (int r0 = 0; r0 < nhit; r0 += 32) {
          const int r = r0 + lane;
          const bool act = r < nhit;
          // the row search uses warp shuffles: every lane executes it
          const int f = act ? (int)s_list[wid][r] : 0;
          int rr = 0;
#pragma unroll
          for (int q = 0; q < 8; q++) rr += (f >= __shfl_sync(FULL, rincl, q)) ? 1 : 0;
          const int r_incl = __shfl_sync(FULL, rincl, rr);
          const int r_len = __shfl_sync(FULL, rlen, rr);
          const int r_start = __shfl_sync(FULL, rstart, rr);
          if (act) {
            const int slot = r_start + (f - (r_incl - r_len));
            const int vrow = ((lz + rr / 3) * VY + (ly + rr % 3)) * VX + lx;
            const int v = vrow + (slot >= vstart[vrow + 1] ? 1 : 0) + (slot >= vstart[vrow + 2] ? 1 : 0);
            const T xj = sx[slot], yj = sy[slot], zj = sz[slot];
            const uint32_t wj = sw[slot];
            long long S[3];
            unpack_shift(vsh[v], S);
            if (wi != wj || (wi & WIND_OVERFLOW)) {
              long long w_i[3], w_j[3];
              int cc[3];
              if (wi & WIND_OVERFLOW) cell_of(g, xi, yi, zi, cc, w_i); else unpack_wind(wi, w_i);
              if (wj & WIND_OVERFLOW) cell_of(g, xj, yj, zj, cc, w_j); else unpack_wind(wj, w_j);
              S[0] += w_i[0] - w_j[0]; S[1] += w_i[1] - w_j[1]; S[2] += w_i[2] - w_j[2];
            }
            const long long pos = base + r;
            a.out.io[pos] = (TI)io + 1;
            a.out.jo[pos] = (TI)sidx[slot] + 1;
            a.out.So[3 * pos] = (TI)S[0];
            a.out.So[3 * pos + 1] = (TI)S[1];
            a.out.So[3 * pos + 2] = (TI)S[2];
            if (a.out.Ro) {
              T R[3];
              pair_r2(g, xi, yi, zi, xj, yj, zj, S, R);
              a.out.Ro[3 * pos] = R[0];
              a.out.Ro[3 * pos + 1] = R[1];
              a.out.Ro[3 * pos + 2] = R[2];
            }
          }
        }
      }
    }
  }
}

template <class T, class TI>
inline void mask_args(MaskArgs<T, TI>& a, int64_t n, const TI* co, const Records<T>& rec, const Geo<T>& g, const Sinks<T, TI>& sk,
                      const TileShape& ts, uint32_t* masks) {
  a.rec = rec; a.co = co; a.n = n; a.g = g; a.out = sk; a.masks = masks; a.cellflag = nullptr;
  a.tx = ts.tx; a.ty = ts.ty; a.tz = ts.tz;
  a.ntx = (g.nc[0] + ts.tx - 1) / ts.tx; a.nty = (g.nc[1] + ts.ty - 1) / ts.ty; a.ntz = (g.nc[2] + ts.tz - 1) / ts.tz;
  a.lo = a.hi = a.dguard = 0.f;
}

}  // namespace nl
