// nl_fill3.cuh -- the fill pass of the mask route, round 2, second cut: same tile, same staging and same mask expansion as
// k_fill_park<.., false> (nl_fill2.cuh), rebuilt around what the profile of that kernel showed (profiles/r02_fill_blocks_*.txt):
// it was bound by instruction issue (3.3 G warp instructions, 165 per row chunk), not by its stores.
//
//   * S is zeroed up front by the streaming i kernel (k_expand_rows), so a row that crosses no periodic boundary writes
//     only j and R.
//   * PLAIN rows.  A tile whose staged atoms all carry the zero winding number (every atom inside the box: the usual case)
//     and a home cell whose 27 stencil cells all have a zero periodic shift need no per-pair shift bookkeeping at all:
//     S = 0, R = (x_j - x_i) + cell' * 0.  That path looks up a slot, subtracts, and stores -- no shift decode, no winding
//     compare, no per-cell shift table, no S staging.
//   * Everything else (cells at the periodic boundary, atoms outside the box, windings beyond +-511) takes the GENERAL row
//     body, which is the round-2 body unchanged.
//   * HAS_R is a template parameter (the reference's PairList layout has no R).
#pragma once
#include "nl_fill2.cuh"

namespace nl {

#ifndef NL_F3_LIST2
#define NL_F3_LIST2 1   // two bits per trip of the mask -> list loop: fill stage -0.10 ms (A/B with scripts/ab_build.sh)
#endif
#ifndef NL_F3_RDIRECT
#define NL_F3_RDIRECT 0   // experiment, measured SLOWER (fill stage 6.29 ms against 4.84): see f3_row_plain_direct
#endif

// Byte offset of component c (0, 1, 2) of a staged record from sA + 16 * slot (layout of f2_store).
template <class T> __device__ __forceinline__ int f3_comp_off(int c, int cap) {
  if constexpr (sizeof(T) == 8) return c == 0 ? 0 : (c == 1 ? 8 : cap * 16);
  else return 4 * c;
}

// One row, plain case, R WITHOUT the shared-memory transposition: the R row is 3 * nhit contiguous values, lane l of pass m
// writes element e = l + 32 m = component e % 3 of hit e / 3, which it fetches itself (hit list -> slot -> one component of the
// staged record).  Three independent passes per chunk, no staging stores, no warp barriers; e % 3 = (l % 3 + 2 m) % 3 is a
// per-lane constant of each pass, so the home atom's components and the cell' * 0 terms are picked once per row.
// MEASURED (-DNL_F3_RDIRECT=1, headline): parity green, but the fill stage takes 6.29 ms instead of 4.84: the three 8-byte fetches
// per lane from random 16-byte slots cost more shared-memory wavefronts than the two 16-byte fetches + conflict-free transposition
// of f3_row_plain.  Kept as a switch for the record; not the default.
template <class T, class TI, bool HAS_R>
__device__ __forceinline__ void f3_row_plain_direct(TI* __restrict__ jo, T* __restrict__ Ro, unsigned wofs, int q, T cz0, T cz1, T cz2, int lane, int nhit,
                                                    long long base, int hs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CAP = f2_cap<T, TI>();
  constexpr int WB = f2_warp_bytes<T, TI>();
  constexpr int OFF_SA = 3 * TILE_VPAD * 4 + 64 * 4 + 64 * 4 * (int)sizeof(T) + F2_NW * WB;
  const unsigned char* const sA = smem_raw + OFF_SA;
  const int c0 = lane % 3;
  T xim[3], czm[3];
#pragma unroll
  for (int m = 0; m < 3; m++) {
    const int c = (c0 + 2 * m) % 3;
    xim[m] = *(const T*)(sA + 16 * hs + f3_comp_off<T>(c, CAP));
    czm[m] = c == 0 ? cz0 : (c == 1 ? cz1 : cz2);
  }
  TI* const jrow = jo + base + lane;
  T* const Rrow = Ro + 3 * base + lane;
#pragma unroll 1
  for (int r0 = 0; r0 < nhit; r0 += 32) {
    unsigned wo = wofs;
    asm volatile("" : "+r"(wo));
    const int nr = min(32, nhit - r0);
    const uint8_t* const Lq = smem_raw + wo + 2 * MASK_MAXCAND + q * MASK_MAXCAND + r0;
    const uint16_t* const tabp = (const uint16_t*)(smem_raw + wo);
    if (lane < nr) {
      const int slot = (int)(tabp[Lq[lane]] & 2047u);
      const uint32_t jv = sizeof(T) == 8 ? *(const uint32_t*)(sA + CAP * 16 + 16 * slot + 8) : *(const uint32_t*)(sA + 16 * slot + 12);
      jrow[r0] = (TI)jv + 1;
    }
    if (HAS_R) {
      const int nw = 3 * nr;
#pragma unroll
      for (int m = 0; m < 3; m++) {
        const int e = lane + 32 * m;
        if (e < nw) {
          const int h = (e * 171) >> 9;  // e / 3 for e < 96
          const int slot = (int)(tabp[Lq[h]] & 2047u);
          const int c = (c0 + 2 * m) % 3;
          const T v = *(const T*)(sA + 16 * slot + f3_comp_off<T>(c, CAP));
          Rrow[3 * r0 + 32 * m] = add_rn(sub_rn(v, xim[m]), czm[m]);
        }
      }
    }
  }
}

// One row, plain case: nhit hits of the home atom staged at slot hs; L = its hit list (flat candidate numbers), tab = flat
// candidate -> staged slot.  j straight from the lane, R transposed through shared memory so that every store instruction
// writes one contiguous run.
template <class T, class TI, bool HAS_R>
__device__ __forceinline__ void f3_row_plain(TI* __restrict__ jo, T* __restrict__ Ro, unsigned wofs, int q, T cz0, T cz1, T cz2, int lane, int nhit,
                                             long long base, int hs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CAP = f2_cap<T, TI>();
  constexpr int WB = f2_warp_bytes<T, TI>();
  constexpr int OFF_SA = 3 * TILE_VPAD * 4 + 64 * 4 + 64 * 4 * (int)sizeof(T) + F2_NW * WB;   // staged records (two 16-byte halves)
  constexpr int OFF_BR = 6 * MASK_MAXCAND + 32 + f2_bufJ<T, TI>() + f2_bufS<T, TI>();          // per warp: R staging of one chunk
  const unsigned char* const sA = smem_raw + OFF_SA;
  const unsigned char* const sB = sA + CAP * 16;
  T xi, yi, zi;
  uint32_t idx_i, wi;
  f2_load<T>(sA, sB, hs, xi, yi, zi, idx_i, wi);
  TI* const jrow = jo + base + lane;
  T* const Rrow = Ro + 3 * base + lane;
#pragma unroll 1
  for (int r0 = 0; r0 < nhit; r0 += 32) {
    // the lane-dependent shared-memory addresses are rebuilt from ONE opaque 32-bit offset per chunk: hoisted out of the row
    // loops they do not fit the register budget of 2 CTAs / SM and get spilled (measured: the reloads sit on the critical path)
    unsigned wo = wofs;
    asm volatile("" : "+r"(wo));
    const int nr = min(32, nhit - r0);
    T* const bR = (T*)(smem_raw + wo + OFF_BR);
    if (lane < nr) {
      const int f = (int)smem_raw[wo + 2 * MASK_MAXCAND + q * MASK_MAXCAND + r0 + lane];
      const int slot = (int)(((const uint16_t*)(smem_raw + wo))[f] & 2047u);
      T xj, yj, zj;
      uint32_t jv, wj;
      f2_load<T>(sA, sB, slot, xj, yj, zj, jv, wj);
      jrow[r0] = (TI)jv + 1;
      if (HAS_R) {
        bR[3 * lane] = add_rn(sub_rn(xj, xi), cz0);
        bR[3 * lane + 1] = add_rn(sub_rn(yj, yi), cz1);
        bR[3 * lane + 2] = add_rn(sub_rn(zj, zi), cz2);
      }
    }
    if (HAS_R) {
      __syncwarp();
      const int nw = 3 * nr - lane;  // words left for this lane's column
      T* const d = Rrow + 3 * r0;
      if (nw > 0) d[0] = bR[lane];
      if (nw > 32) d[32] = bR[lane + 32];
      if (nw > 64) d[64] = bR[lane + 64];
      __syncwarp();
    }
  }
}

// One row, general case (the round-2 body of k_fill_park<.., false> with the S stream zeroed up front).
template <class T, class TI, bool HAS_R>
__device__ __forceinline__ void f3_row_general(const MaskArgs<T, TI>& a, const unsigned char* sA, const unsigned char* sB, const uint16_t* tab,
                                               const uint8_t* L, const uint8_t* shc, const T* cstab, TI* bS, T* bR, bool fastcell, int lane, int nhit,
                                               long long base, int hs) {
  T xi, yi, zi;
  uint32_t idx_i, wi;
  f2_load<T>(sA, sB, hs, xi, yi, zi, idx_i, wi);
  const bool wi_ok = !(wi & WIND_OVERFLOW);
#pragma unroll 1
  for (int r0 = 0; r0 < nhit; r0 += 32) {
    const long long p0 = base + r0;
    const int nr = min(32, nhit - r0);
    const bool act = lane < nr;
    bool zs = true;
    int S0 = 0, S1 = 0, S2 = 0;
    T R0 = 0, R1 = 0, R2 = 0;
    if (act) {
      const int f = (int)L[r0 + lane];
      const unsigned t = tab[f];
      const int slot = (int)(t & 2047u), c = (int)(t >> 11);
      T xj, yj, zj;
      uint32_t jv, wj;
      f2_load<T>(sA, sB, slot, xj, yj, zj, jv, wj);
      const int p = fastcell ? SHP_ZERO : (int)shc[c];
      S0 = (p & 3) - 1; S1 = ((p >> 2) & 3) - 1; S2 = ((p >> 4) & 3) - 1;
      if (wi == wj && wi_ok) {
        const T* cs = cstab + 4 * p;
        R0 = add_rn(sub_rn(xj, xi), cs[0]);
        R1 = add_rn(sub_rn(yj, yi), cs[1]);
        R2 = add_rn(sub_rn(zj, zi), cs[2]);
      } else {
        int S3[3] = {S0, S1, S2};
        T R3[3];
        slow_shift_and_R<T, TI>(a.self, xi, yi, zi, xj, yj, zj, wi, wj, S3, R3);
        S0 = S3[0]; S1 = S3[1]; S2 = S3[2];
        R0 = R3[0]; R1 = R3[1]; R2 = R3[2];
      }
      zs = (S0 | S1 | S2) == 0;
      a.out.jo[p0 + lane] = (TI)jv + 1;
      if (HAS_R) { bR[3 * lane] = R0; bR[3 * lane + 1] = R1; bR[3 * lane + 2] = R2; }
    }
    const bool zeroS = __all_sync(FULL, zs);   // S was zeroed up front: only chunks with a shift write it
    if (act && !zeroS) { bS[3 * lane] = (TI)S0; bS[3 * lane + 1] = (TI)S1; bS[3 * lane + 2] = (TI)S2; }
    __syncwarp();
    const int nw = 3 * nr;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int w = m * 32 + lane;
      if (w < nw) {
        if (!zeroS) a.out.So[3 * p0 + w] = bS[w];
        if (HAS_R) a.out.Ro[3 * p0 + w] = bR[w];
      }
    }
    __syncwarp();
  }
}

template <class T, class TI, bool HAS_R>
__global__ void __launch_bounds__(F2_NT, 2) k_fill3(const MaskArgs<T, TI> a, int prefetch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CAP = f2_cap<T, TI>();
  constexpr int WB = f2_warp_bytes<T, TI>();
  typedef typename FillBase<TI>::type BaseT;
  constexpr BaseT NOROW = ~(BaseT)0;
  int* vstart = (int*)smem_raw;
  int* vgs = vstart + TILE_VPAD;
  int* vsh = vgs + TILE_VPAD;
  int* hcell = vsh + TILE_VPAD;
  T* cstab = (T*)(hcell + 64);                          // [64][4]: cell' * s_loop for every packed shift (contract arithmetic)
  unsigned char* wbase = (unsigned char*)(cstab + 256);
  unsigned char* sA = wbase + F2_NW * WB;
  unsigned char* sB = sA + CAP * 16;
  __shared__ int scan_sm[33];
  __shared__ int s_next;

  const Geo<T>& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int grp = lane >> 3, sub = lane & 7;
  unsigned char* wb = wbase + wid * WB;
  const unsigned wofs = (unsigned)(wb - smem_raw);
  uint16_t* tab = (uint16_t*)wb;                         // [256] staged slot | stencil cell << 11
  uint8_t* lists = wb + 2 * MASK_MAXCAND;                // [4][256]
  uint8_t* shc = lists + 4 * MASK_MAXCAND;               // [27] packed shift of each stencil cell
  char* bufJ = (char*)(shc + 32);
  TI* bS = (TI*)(bufJ + f2_bufJ<T, TI>());
  T* bR = (T*)((char*)bS + f2_bufS<T, TI>());

  const int b = blockIdx.x;
  const int bz = b / (a.ntx * a.nty);
  const int hx0 = (b % a.ntx) * a.tx, hy0 = ((b / a.ntx) % a.nty) * a.ty, hz0 = (a.zlayers ? a.zlayers[bz] : bz) * a.tz;
  const int hxn = min(a.tx, g.nc[0] - hx0), hyn = min(a.ty, g.nc[1] - hy0), hzn = min(a.tz, g.nc[2] - hz0);
  const int VX = hxn + 2, VY = hyn + 2, VZ = hzn + 2, NV = VX * VY * VZ;
  {
    int nonempty = 0;
    if (tid < hyn * hzn) {
      const long long c0 = (long long)hx0 + (long long)g.nc[0] * ((long long)(hy0 + tid % hyn) + (long long)g.nc[1] * (hz0 + tid / hyn));
      nonempty = (long long)a.co[c0 + hxn] > (long long)a.co[c0];
    }
    if (!__syncthreads_or(nonempty)) return;
  }
  const int total = tile_table_nt<F2_NT>(g.nc, g.pbc, a.co, sizeof(TI) == 8, hx0, hy0, hz0, VX, VY, NV, vstart, vgs, vsh, scan_sm);
  const int nhome = hxn * hyn * hzn;
  if (tid < nhome) hcell[tid] = (tid % hxn) | (((tid / hxn) % hyn) << 8) | ((tid / (hxn * hyn)) << 16);
  if (tid < 64 && (tid & 3) < 3 && ((tid >> 2) & 3) < 3 && (tid >> 4) < 3) {
    T c0, c1, c2;
    mtv(g.cell, (T)((tid & 3) - 1), (T)(((tid >> 2) & 3) - 1), (T)((tid >> 4) - 1), c0, c1, c2);
    cstab[4 * tid] = c0; cstab[4 * tid + 1] = c1; cstab[4 * tid + 2] = c2; cstab[4 * tid + 3] = (T)0;
  }
  if (tid == 0) s_next = 0;
  __syncthreads();

  if (total > CAP) {  // denser than the staging capacity: generic route, in place
    for (int hc = wid; hc < nhome; hc += F2_NW) {
      const int lx = hcell[hc] & 255, ly = (hcell[hc] >> 8) & 255, lz = hcell[hc] >> 16;
      const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
      generic_cell<T, TI, MODE_FILL>(a.self, (long long)vgs[vh], vstart[vh + 1] - vstart[vh], lane);
    }
    return;
  }
  const uint32_t* __restrict__ idsrc = a.out.pgid0 ? a.out.pgid0 : a.rec.pidx;  // what j publishes (shard mode: global index - 1)
  int all_zero_wind = 1;
  for (int sl = tid; sl < total; sl += F2_NT) {
    const int v = find_vcell(vstart, NV, sl);
    const long long src = (long long)vgs[v] + (sl - vstart[v]);
    const uint32_t pw = a.rec.pw[src];
    all_zero_wind &= (pw == WIND_ZERO) ? 1 : 0;
    f2_store<T>(sA, sB, sl, a.rec.px[src], a.rec.py[src], a.rec.pz[src], idsrc[src], pw);
  }
  if (tid < nhome) {
    const int lx = hcell[tid] & 255, ly = (hcell[tid] >> 8) & 255, lz = (hcell[tid] >> 16) & 255;
    if (a.cellflag[cell_linear(g.nc, hx0 + lx, hy0 + ly, hz0 + lz)]) hcell[tid] |= 1 << 24;
  }
  // every staged atom lies inside the box (winding 0): no per-pair compare.  Kept in shared memory, like every other tile-wide
  // value the cell loop needs only once per cell: in registers they get spilled, and a spill reload at the head of a cell
  // was 8 % of this kernel's stall samples (the L1 left beside 2 x 104 KB of shared memory does not hold the stacks)
  __shared__ int s_tile_plain;
  {
    const int tp = __syncthreads_and(all_zero_wind);
    if (tid == 0) s_tile_plain = tp;
  }
  __syncthreads();
  const BaseT* __restrict__ srow = (const BaseT*)a.srow;
  const T cz0 = cstab[4 * SHP_ZERO], cz1 = cstab[4 * SHP_ZERO + 1], cz2 = cstab[4 * SHP_ZERO + 2];  // cell' * 0 under the contract (+0)

  while (true) {
    int hc = 0;
    if (lane == 0) hc = atomicAdd(&s_next, 1);
    hc = __shfl_sync(FULL, hc, 0);
    if (hc >= nhome) break;
    const int hcv = hcell[hc];
    const int lx = hcv & 255, ly = (hcv >> 8) & 255, lz = (hcv >> 16) & 255;
    const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
    const int hstart = vstart[vh], nh = vstart[vh + 1] - hstart;
    if (nh == 0) continue;
    const long long hg0 = vgs[vh];
    if (!(hcv >> 24)) {
      generic_cell<T, TI, MODE_FILL>(a.self, hg0, nh, lane);
      continue;
    }
    // row data of the first two passes (lane = 8 * atom + mask word), issued before the table building so that it hides them
    uint32_t nx_word = 0, nx2_word = 0;
    BaseT nx_base = NOROW, nx2_base = NOROW;
    if (grp < nh) { nx_word = a.masks[(hg0 + grp) * MASK_WORDS + sub]; nx_base = srow[hg0 + grp]; }
    if (4 + grp < nh) { nx2_word = a.masks[(hg0 + 4 + grp) * MASK_WORDS + sub]; nx2_base = srow[hg0 + 4 + grp]; }
    // candidate table: flat candidate -> staged slot | stencil cell; packed shift per stencil cell
    bool fastcell;
    {
      int st = 0, cn = 0, shp = SHP_ZERO;
      int l2 = lane;
      unsigned wo = wofs;
      asm volatile("" : "+r"(l2), "+r"(wo));  // rebuild the stencil coordinates and this warp's shared-memory pointers here
                                              // instead of carrying (and spilling) them across the cell loop
      uint16_t* const tab = (uint16_t*)(smem_raw + wo);
      uint8_t* const shc = smem_raw + wo + 6 * MASK_MAXCAND;
      if (l2 < 27) {
        const int v = ((lz + l2 / 9) * VY + (ly + (l2 / 3) % 3)) * VX + (lx + l2 % 3);
        st = vstart[v];
        cn = vstart[v + 1] - st;
        if (cn > 0) shp = vsh[v];
        shc[l2] = (uint8_t)shp;
      }
      const int incl = warp_incl_scan(cn, lane);
      const int pre = incl - cn;
      const int mx = __reduce_max_sync(FULL, cn);
      const unsigned tag = (unsigned)lane << 11;
      for (int j = 0; j < mx; j++)
        if (j < cn) tab[pre + j] = (uint16_t)(tag | (unsigned)(st + j));
      fastcell = __all_sync(FULL, shp == SHP_ZERO);
    }
    __syncwarp();
    const bool plain = fastcell && s_tile_plain != 0;

    for (int a0 = 0; a0 < nh; a0 += 4) {
      uint32_t word = nx_word;
      const BaseT my_base = nx_base;
      nx_word = nx2_word; nx_base = nx2_base;
      nx2_word = 0; nx2_base = NOROW;
      if (a0 + 8 + grp < nh) {
        nx2_word = a.masks[(hg0 + a0 + 8 + grp) * MASK_WORDS + sub];
        nx2_base = srow[hg0 + a0 + 8 + grp];
      }
      if (my_base == NOROW) word = 0;  // the atom gets no row (halo atom of a shard)
      const int pc = __popc(word);
      int incl = pc;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o, 8);
        if (sub >= o) incl += t;
      }
      const int my_nhit = __shfl_sync(FULL, incl, 7, 8);
      if (prefetch && my_nhit > 0 && sub < (HAS_R ? 4 : 2)) {
        // In place, the partial 32-byte sectors at both ends of a row segment are shared with the neighbouring rows, which
        // other warps write at other times: evicted half-written, each costs a DRAM read-modify-write.  Prefetching them
        // into L2 now makes the partial write land on a fully valid sector, which is later written back whole
        // (experiments/microbench_rows2.cu).  sub = 0, 1: head / tail of the j row; 2, 3: of the R row.
        const long long eb = (sub & 2) ? 3ll * (long long)sizeof(T) : (long long)sizeof(TI);
        const char* gb = (sub & 2) ? (const char*)a.out.Ro : (const char*)a.out.jo;
        const long long B = ((long long)my_base + ((sub & 1) ? my_nhit : 0)) * eb;
        if (B & 31) asm volatile("prefetch.global.L2 [%0];" ::"l"(gb + ((sub & 1) ? ((B - 1) & ~31ll) : (B & ~31ll))));
      }
      __syncwarp();
      {
        unsigned wo = wofs;
        asm volatile("" : "+r"(wo));
        uint8_t* Lw = smem_raw + wo + 2 * MASK_MAXCAND + grp * MASK_MAXCAND + (incl - pc);
        const int fb = sub * 32;
#if NL_F3_LIST2
        // the lane with the fullest word sets the pace of the whole warp (~14 hits against ~3 on average: the chunks of the
        // cells next to the home cell), so every trip takes the lowest AND the highest set bit: half the trips
        uint8_t* Lh = Lw + pc - 1;
        while (word) {
          const int lo = __ffs(word) - 1, hi = 31 - __clz(word);
          *Lw++ = (uint8_t)(fb + lo);
          *Lh-- = (uint8_t)(fb + hi);   // lo == hi on the last trip of an odd count: same value, same place
          word &= word - 1;
          word &= ~(1u << hi);
        }
#else
        while (word) {
          const int bit = __ffs(word) - 1;
          word &= word - 1;
          *Lw++ = (uint8_t)(fb + bit);
        }
#endif
      }
      __syncwarp();

      const int na = min(4, nh - a0);
      for (int q = 0; q < na; q++) {
        const int nhit = __shfl_sync(FULL, my_nhit, q * 8);
        if (nhit == 0) continue;
        const long long base = (long long)__shfl_sync(FULL, my_base, q * 8);
        const uint8_t* L = lists + q * MASK_MAXCAND;
#if NL_F3_RDIRECT
        if (plain) f3_row_plain_direct<T, TI, HAS_R>(a.out.jo, a.out.Ro, wofs, q, cz0, cz1, cz2, lane, nhit, base, hstart + a0 + q);
#else
        if (plain) f3_row_plain<T, TI, HAS_R>(a.out.jo, a.out.Ro, wofs, q, cz0, cz1, cz2, lane, nhit, base, hstart + a0 + q);
#endif
        else f3_row_general<T, TI, HAS_R>(a, sA, sB, tab, L, shc, cstab, bS, bR, fastcell, lane, nhit, base, hstart + a0 + q);
      }
    }
  }
}

}  // namespace nl
