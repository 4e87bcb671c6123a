// nl_tiled.cuh -- tiled, shared-memory-staged traversal for the common geometry (nxyz == 1).
//
// One CTA owns a TX x TY x TZ block of "home" cells.  It stages the records of those cells plus a
// one-cell halo -- (TX+2)(TY+2)(TZ+2) VIRTUAL cells, each a (real cell, periodic image shift) pair
// or empty beyond an open boundary -- into shared memory as SoA, with coalesced loads: cells are
// sorted x-fastest, so every x-row of virtual cells is one contiguous run of sorted records.
// Virtual cells make every wrap case uniform: a box only 1 or 2 cells wide simply stages the same
// real cell several times under different shifts, exactly the bijection d -> (cell, s_loop) of
// _get_neighbor_cell (src/gpu_kernels.jl:39-47).
//
// One warp then takes one home cell at a time.  Its 27 neighbour cells are 9 contiguous slot ranges
// (3 x-adjacent cells each), flattened into one candidate list; LANES RUN OVER CANDIDATES (held in
// registers) while the home atoms are broadcast from shared memory, and hits are compacted with
// __ballot_sync / __popc.  The distance arithmetic is the contract of nl_common.cuh.
#pragma once
#include "../../include/nlcuda.h"
#include "nl_scan_sort.cuh"
#include <cmath>

#include "nl_traverse.cuh"

namespace nl {

constexpr int TILE_NT = 256;          // 8 warps per CTA
constexpr int TILE_MAXT = 4;          // max home cells per axis
constexpr int TILE_MAXV = (TILE_MAXT + 2) * (TILE_MAXT + 2) * (TILE_MAXT + 2);  // 216 virtual cells
constexpr int TILE_VPAD = 224;        // table stride
constexpr int TILE_SMEM_BYTES = 56 * 1024;  // dynamic shared memory per CTA -> 4 CTAs / SM

template <class T> struct TileRecBytes { static constexpr int value = 3 * (int)sizeof(T) + 8; };
template <class T> __host__ __device__ constexpr int tile_cap() { return (TILE_SMEM_BYTES - 2 * TILE_VPAD * 4) / TileRecBytes<T>::value / 8 * 8; }

struct TileShape { int tx, ty, tz; };

template <class T, class TI> struct TiledArgs {
  Records<T> rec;
  const TI* co;
  long long n;
  Geo<T> g;
  Sinks<T, TI> out;
  int tx, ty, tz;     // home cells per tile and axis
  int ntx, nty, ntz;  // tiles per axis
};

// Picks the largest tile whose expected staged population fits the shared-memory capacity.
template <class T> inline bool pick_tile(const Geo<T>& g, long long n, int cap, TileShape& best) {
  const double dens = (double)n / (double)g.nct;
  long long best_home = 0, best_staged = 0;
  for (int tz = 1; tz <= TILE_MAXT; tz++)
    for (int ty = 1; ty <= TILE_MAXT; ty++)
      for (int tx = 1; tx <= TILE_MAXT; tx++) {
        int hx = tx < g.nc[0] ? tx : g.nc[0], hy = ty < g.nc[1] ? ty : g.nc[1], hz = tz < g.nc[2] ? tz : g.nc[2];
        if (hx != tx || hy != ty || hz != tz) continue;
        long long home = (long long)tx * ty * tz, staged = (long long)(tx + 2) * (ty + 2) * (tz + 2);
        const double expect = (double)staged * dens;
        if (expect + 6.0 * sqrt(expect) + 8.0 > (double)cap) continue;  // mean + 6 sigma (Poisson) must fit
        if (home > best_home || (home == best_home && staged < best_staged)) {
          best_home = home; best_staged = staged; best = {tx, ty, tz};
        }
      }
  return best_home > 0;
}

// hit masks (8 words per atom) + one flag byte per cell
inline size_t tiled_scratch_bytes(const nl_params* p, int64_t n) {
  const size_t nct = (size_t)p->ncells[0] * p->ncells[1] * p->ncells[2];
  return 512 + (((size_t)(n > 0 ? n : 1) * 32 + 255) & ~(size_t)255) + nct;
}

template <class T> inline bool tiled_applicable(const nl_params* p, const Geo<T>& g, long long n, int cap, TileShape& ts) {
  if (p->nxyz[0] != 1 || p->nxyz[1] != 1 || p->nxyz[2] != 1) return false;
  return pick_tile<T>(g, n, cap, ts);
}

// 0-based floor-div / mod of a virtual cell coordinate; returns false for a cell beyond an open boundary.
__device__ __forceinline__ bool map_virtual(int v, int n, int pbc, int& c, int& s) {
  if (pbc) {
    int q = v / n, r = v % n;
    if (r < 0) { r += n; q -= 1; }
    c = r; s = q;
    return true;
  }
  c = v; s = 0;
  return v >= 0 && v < n;
}

template <class T, class TI, int MODE>
__global__ void __launch_bounds__(TILE_NT, 3) k_tiled(const TiledArgs<T, TI> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CAP = tile_cap<T>();
  int* vstart = (int*)smem_raw;            // [NV + 1] first staged slot of each virtual cell
  int* vgs = vstart + TILE_VPAD;           // [NV] first global sorted index of each virtual cell
  T* sx = (T*)(vgs + TILE_VPAD);
  T* sy = sx + CAP;
  T* sz = sy + CAP;
  uint32_t* sidx = (uint32_t*)(sz + CAP);
  uint32_t* sw = sidx + CAP;
  __shared__ int scan_sm[33];
  __shared__ int s_next;
  __shared__ double s_energy[TILE_NT / 32];

  const Geo<T>& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;

  // ---- tile geometry
  const int b = blockIdx.x;
  const int hx0 = (b % a.ntx) * a.tx, hy0 = ((b / a.ntx) % a.nty) * a.ty, hz0 = (b / (a.ntx * a.nty)) * a.tz;
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

  // ---- 1. virtual cell table
  int cnt = 0, gs = 0;
  if (tid < NV) {
    int cx, cy, cz, s0, s1, s2;
    bool ok = map_virtual(hx0 + tid % VX - 1, g.nc[0], g.pbc[0], cx, s0);
    ok = map_virtual(hy0 + (tid / VX) % VY - 1, g.nc[1], g.pbc[1], cy, s1) && ok;
    ok = map_virtual(hz0 + tid / (VX * VY) - 1, g.nc[2], g.pbc[2], cz, s2) && ok;
    if (ok) {
      const long long cl = (long long)cx + (long long)g.nc[0] * ((long long)cy + (long long)g.nc[1] * cz);
      const long long c0 = (long long)a.co[cl], c1 = (long long)a.co[cl + 1];
      gs = (int)(c0 - 1);
      cnt = (int)(c1 - c0);
    }
  }
  int total;
  const int excl = block_excl_scan<int, TILE_NT>(cnt, scan_sm, &total);
  if (tid < NV) { vstart[tid] = excl; vgs[tid] = gs; }
  if (tid == NV) vstart[NV] = total;  // NV <= 216 < TILE_NT
  if (tid == 0) s_next = 0;
  __syncthreads();

  const int nhome = hxn * hyn * hzn;
  double e_acc = 0.0;

  if (total > CAP) {
    // ---- denser than the staging capacity: this tile takes the generic per-atom route
    for (int hc = wid; hc < nhome; hc += TILE_NT / 32) {
      const int vh = ((hc / (hxn * hyn) + 1) * VY + ((hc / hxn) % hyn + 1)) * VX + (hc % hxn + 1);
      const int nh = vstart[vh + 1] - vstart[vh];
      for (int k = lane; k < nh; k += 32) e_acc += generic_atom<T, TI, MODE>((long long)vgs[vh] + k, a.rec, a.co, g, a.out);
    }
  } else {
    // ---- 2. stage the records of every virtual cell (coalesced: slots follow sorted order row by row)
    for (int sl = tid; sl < total; sl += TILE_NT) {
      int lo = 0, hi = NV;  // last v with vstart[v] <= sl
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (vstart[mid] <= sl) lo = mid; else hi = mid;
      }
      const long long src = (long long)vgs[lo] + (sl - vstart[lo]);
      sx[sl] = a.rec.px[src];
      sy[sl] = a.rec.py[src];
      sz[sl] = a.rec.pz[src];
      sidx[sl] = a.rec.pidx[src];
      sw[sl] = a.rec.pw[src];
    }
    __syncthreads();

    // ---- 3. one warp per home cell, dynamically scheduled
    while (true) {
      int hc = 0;
      if (lane == 0) hc = atomicAdd(&s_next, 1);
      hc = __shfl_sync(FULL, hc, 0);
      if (hc >= nhome) break;
      const int lx = hc % hxn, ly = (hc / hxn) % hyn, lz = hc / (hxn * hyn);
      const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
      const int hstart = vstart[vh], nh = vstart[vh + 1] - hstart;
      if (nh == 0) continue;

      // 9 x-rows of 3 virtual cells each: lane r < 9 owns row r = (dz+1)*3 + (dy+1)
      int rstart = 0, rlen = 0;
      if (lane < 9) {
        const int vrow = ((lz + lane / 3) * VY + (ly + lane % 3)) * VX + lx;  // dx = -1 cell of the row
        rstart = vstart[vrow];
        rlen = vstart[vrow + 3] - rstart;
      }
      const int rincl = warp_incl_scan(rlen, lane);
      const int ncand = __shfl_sync(FULL, rincl, 8);

      for (int g0 = 0; g0 < nh; g0 += 32) {
        const int ng = min(32, nh - g0);
        const int my_home = hstart + g0 + lane;  // lane a < ng owns home atom a of this group
        uint32_t my_cnt = 0;
        long long my_base = 0;
        uint32_t my_io = 0;
        if (lane < ng) {
          my_io = sidx[my_home];
          if (MODE == MODE_FILL && (long long)my_io < a.out.n_rows) my_base = (long long)a.out.first[my_io] - 1;
        }

        for (int k0 = 0; k0 < ncand; k0 += 32) {
          const int f = k0 + lane;
          const bool valid = f < ncand;
          // which row does flat index f fall in?
          int rr = 0;
#pragma unroll
          for (int r = 0; r < 8; r++) rr += (f >= __shfl_sync(FULL, rincl, r)) ? 1 : 0;
          const int r_incl = __shfl_sync(FULL, rincl, rr);
          const int r_len = __shfl_sync(FULL, rlen, rr);
          const int r_start = __shfl_sync(FULL, rstart, rr);
          int slot = 0;
          T xj = 0, yj = 0, zj = 0, cs0 = 0, cs1 = 0, cs2 = 0;
          uint32_t wj = 0, jo = 0;
          long long sl0 = 0, sl1 = 0, sl2 = 0;
          long long gsj = 0;  // global sorted index of the candidate (half lists)
          if (valid) {
            slot = r_start + (f - (r_incl - r_len));
            // x offset of the candidate's cell inside its row
            const int vrow = ((lz + rr / 3) * VY + (ly + rr % 3)) * VX + lx;
            const int dxi = (slot >= vstart[vrow + 1] ? 1 : 0) + (slot >= vstart[vrow + 2] ? 1 : 0);
            gsj = (long long)vgs[vrow + dxi] + (slot - vstart[vrow + dxi]);
            int c, s;
            map_virtual(hx0 + lx + dxi - 1, g.nc[0], g.pbc[0], c, s); sl0 = s;
            map_virtual(hy0 + ly + rr % 3 - 1, g.nc[1], g.pbc[1], c, s); sl1 = s;
            map_virtual(hz0 + lz + rr / 3 - 1, g.nc[2], g.pbc[2], c, s); sl2 = s;
            xj = sx[slot]; yj = sy[slot]; zj = sz[slot];
            wj = sw[slot];
            jo = sidx[slot];
            mtv(g.cell, (T)sl0, (T)sl1, (T)sl2, cs0, cs1, cs2);  // cell' * s_loop, shared by every home atom with w_i == w_j
          }

          T lf0 = 0, lf1 = 0, lf2 = 0, lfe = 0;  // MODE_LJF: force / energy this chunk's home atoms put on the lane's candidate
          for (int aa = 0; aa < ng; aa++) {
            const int hs = hstart + g0 + aa;
            const T xi = sx[hs], yi = sy[hs], zi = sz[hs];
            const uint32_t wi = sw[hs];
            bool hit = false;
            T R0 = 0, R1 = 0, R2 = 0;
            long long S0 = sl0, S1 = sl1, S2 = sl2;
            const bool dropped = (MODE == MODE_COUNT || MODE == MODE_FILL) && a.out.half &&
                                 !half_keep((long long)vgs[vh] + g0 + aa, gsj, sl0, sl1, sl2);
            if (valid && slot != hs && !dropped) {  // slot == hs <=> same atom under zero shift (_is_self_interaction)
              T r2;
              if (wi == wj && !(wi & WIND_OVERFLOW)) {
                R0 = add_rn(sub_rn(xj, xi), cs0);
                R1 = add_rn(sub_rn(yj, yi), cs1);
                R2 = add_rn(sub_rn(zj, zi), cs2);
                r2 = add_rn(add_rn(mul_rn(R0, R0), mul_rn(R1, R1)), mul_rn(R2, R2));
              } else {
                long long w_i[3], w_j[3];
                int cc[3];
                if (wi & WIND_OVERFLOW) cell_of(g, xi, yi, zi, cc, w_i); else unpack_wind(wi, w_i);
                if (wj & WIND_OVERFLOW) cell_of(g, xj, yj, zj, cc, w_j); else unpack_wind(wj, w_j);
                const long long S[3] = {sl0 + w_i[0] - w_j[0], sl1 + w_i[1] - w_j[1], sl2 + w_i[2] - w_j[2]};
                T R[3];
                r2 = pair_r2(g, xi, yi, zi, xj, yj, zj, S, R);
                R0 = R[0]; R1 = R[1]; R2 = R[2];
                S0 = S[0]; S1 = S[1]; S2 = S[2];
              }
              hit = r2 < g.cutoff_sq;
              if (MODE == MODE_LJ && hit) {
                const double s2 = a.out.lj_sigma2 / (double)r2, s6 = s2 * s2 * s2;
                e_acc += 4.0 * a.out.lj_eps * (s6 * s6 - s6);
              }
              if (MODE == MODE_LJF && hit) {
                double phi, gg;
                lj_pair_terms(a.out.lj_eps, a.out.lj_sigma2, (double)r2, phi, gg);
                lf0 += (T)(gg * (double)R0); lf1 += (T)(gg * (double)R1); lf2 += (T)(gg * (double)R2); lfe += (T)phi;
              }
            }
            if (MODE == MODE_COUNT || MODE == MODE_FILL) {
              const unsigned bal = __ballot_sync(FULL, hit);
              if (MODE == MODE_FILL) {
                const long long base = __shfl_sync(FULL, my_base, aa) + __shfl_sync(FULL, my_cnt, aa);
                const uint32_t io_b = __shfl_sync(FULL, my_io, aa);
                if (hit && (long long)io_b < a.out.n_rows) {
                  const long long pos = base + __popc(bal & lt);
                  a.out.io[pos] = out_index(a.out, io_b);
                  a.out.jo[pos] = out_index(a.out, jo);
                  a.out.So[3 * pos] = (TI)S0;
                  a.out.So[3 * pos + 1] = (TI)S1;
                  a.out.So[3 * pos + 2] = (TI)S2;
                  if (a.out.Ro) { a.out.Ro[3 * pos] = R0; a.out.Ro[3 * pos + 1] = R1; a.out.Ro[3 * pos + 2] = R2; }
                }
              }
              if (lane == aa) my_cnt += __popc(bal);
            }
          }
          if (MODE == MODE_LJF && valid && (lf0 != 0 || lf1 != 0 || lf2 != 0 || lfe != 0)) ljf_add_global<T, TI>(a.out, jo, lf0, lf1, lf2, lfe);
        }
        if (MODE == MODE_COUNT && lane < ng) a.out.counts[my_io] = my_cnt;
      }
    }
  }

  if (MODE == MODE_LJ) {
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

template <class T, class TI, int MODE>
inline int tiled_traverse(const nl_params*, int64_t n, const TI* co, const Records<T>& rec, const Geo<T>& g, const Sinks<T, TI>& sk,
                          const TileShape& ts, void*, cudaStream_t st) {
  static bool attr_set[64] = {};  // per device: the attribute belongs to the device's instance of the kernel
  int dev = 0;
  {
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) { last_cuda_slot() = (int)e0; return NL_ERR_CUDA; }
  }
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_tiled<T, TI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES);
    if (e != cudaSuccess) { last_cuda_slot() = (int)e; return NL_ERR_CUDA; }
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  TiledArgs<T, TI> a;
  a.rec = rec; a.co = co; a.n = n; a.g = g; a.out = sk;
  a.tx = ts.tx; a.ty = ts.ty; a.tz = ts.tz;
  a.ntx = (g.nc[0] + ts.tx - 1) / ts.tx; a.nty = (g.nc[1] + ts.ty - 1) / ts.ty; a.ntz = (g.nc[2] + ts.tz - 1) / ts.tz;
  const long long nblk = (long long)a.ntx * a.nty * a.ntz;
  k_tiled<T, TI, MODE><<<(unsigned)nblk, TILE_NT, TILE_SMEM_BYTES, st>>>(a);
  note_launch(1);
  return NL_OK;
}

}  // namespace nl
