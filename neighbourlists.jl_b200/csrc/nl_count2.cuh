// nl_count2.cuh -- round-2 counting pass of the mask route for Float64 positions (full lists).
//
// Same tile, same staging, same decisions and same outputs (counts, 256-bit hit masks, per-cell flags) as
// k_count_mask<double, TI, CM_MASK> (nl_mask.cuh); what changes is the loop order inside a home cell.
//
// k_count_mask walks the candidate list chunk by chunk (32 candidates per chunk, one per lane) and, inside a chunk, over the
// pairs of home atoms.  A home cell holds ~6 atoms, so the inner loop runs 3-4 times per chunk while everything around it --
// the candidate fetch, the band test, the self-pair removal, the popcounts -- is paid per chunk: measured, the packed distance
// loop was < 30 % of the 2.3 G warp instructions of the pass.
//
// Here a lane keeps ALL its candidates of the home cell in registers (<= 8 chunks x 3 floats), the loop over pairs of home
// atoms is the outer one, and the per-chunk work left is the distance test itself (3 FADD2 + 3 FFMA2 for two home atoms),
// two ballots and one shared-memory store of the two hit words.  The band test, the self pair and the popcounts are done
// once per home group.  The kernel body is instantiated for every chunk count 1..8, so the chunk loop is fully unrolled and
// the candidate registers are never indexed dynamically.
#pragma once
#include "nl_mask.cuh"

namespace nl {

#ifndef NL_C2_MINB
#define NL_C2_MINB 4
#endif

// Half lists (NL_FLAG_HALF; half_keep in nl_traverse.cuh): of every mirror couple keep the pair whose second atom comes later in
// cell-sorted order; self images keep the lexicographically positive shift.  For one chunk of 32 candidates (lane = candidate:
// d = its sorted index minus that of the group's first home atom, shp = packed loop shift of its stencil cell) returns, in lane
// aa < ng, the word of candidates home atom aa keeps: a candidate later than the whole group is kept by every home atom, an
// earlier one by none, candidates INSIDE the group are settled one by one.
__device__ __forceinline__ unsigned c2_half_word(bool valid, int d, int shp, int ng, int lane) {
  const int pos = ((shp & 3) > 1 || ((shp & 3) == 1 && (((shp >> 2) & 3) > 1 || (((shp >> 2) & 3) == 1 && ((shp >> 4) & 3) > 1)))) ? 1 : 0;
  unsigned word = __ballot_sync(FULL, valid && d >= ng);
  unsigned ingrp = __ballot_sync(FULL, valid && d >= 0 && d < ng);
  while (ingrp) {
    const int l = __ffs(ingrp) - 1;
    ingrp &= ingrp - 1;
    const int dl = __shfl_sync(FULL, d, l), pl = __shfl_sync(FULL, pos, l);
    if (dl > lane || (dl == lane && pl)) word |= 1u << l;
  }
  return word;
}

// Rare path, one whole home cell (deferred to the end of the tile, so that the hot loop never makes a call): the same table,
// the same chunk order, every decision recomputed with the scalar chain (bit-identical to the packed one) and re-taken with
// the exact Float64 contract wherever the pre-filter cannot be trusted (band, flagged slots).  Overwrites what the fast
// pass stored for this cell.
template <class TI>
__device__ __noinline__ void c2_cell_slow(const MaskArgs<double, TI>* ad, const float4* sq, uint16_t* cslot, uint32_t* mkT, const int* vstart,
                                          const int* vgs, const int* vsh, int NV, int VX, int VY, int hcv) {
  const int lane = threadIdx.x & 31;
  const float mid = ad->mid, hw = ad->hw;
  const int lx = hcv & 255, ly = (hcv >> 8) & 255, lz = hcv >> 16;
  const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
  const int hstart = vstart[vh], nh = vstart[vh + 1] - hstart;
  const long long hg0 = vgs[vh];
  int ncand, fh;
  {
    int st = 0, cn = 0;
    if (lane < 27) {
      const int v = ((lz + lane / 9) * VY + (ly + (lane / 3) % 3)) * VX + (lx + lane % 3);
      st = vstart[v];
      cn = vstart[v + 1] - st;
    }
    const int incl = warp_incl_scan(cn, lane);
    ncand = __shfl_sync(FULL, incl, 31);
    fh = __shfl_sync(FULL, incl, 12);
    const int pre = incl - cn;
    const int mx = __reduce_max_sync(FULL, cn);
    for (int j = 0; j < mx; j++)
      if (j < cn) cslot[pre + j] = (uint16_t)(st + j);
    __syncwarp();
  }
  const int nchunk = (ncand + 31) >> 5;
  for (int g0 = 0; g0 < nh; g0 += 32) {
    const int ng = min(32, nh - g0);
    __syncwarp();
    for (int kc = 0; kc < nchunk; kc++) {
      const int f = kc * 32 + lane;
      const bool valid = f < ncand;
      int gj = 0, shp = 0;
      float4 q = make_float4(CAND_FAR, CAND_FAR, CAND_FAR, 0.f);
      if (valid) {
        const int slot = cslot[f];
        q = sq[slot];
        const int v = find_vcell(vstart, NV, slot);
        gj = vgs[v] + (slot - vstart[v]);
        shp = vsh[v];
      }
      for (int aa = 0; aa < ng; aa++) {
        const float4 p = sq[hstart + g0 + aa];
        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        const float t = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmaf_rn(dx, dx, -mid)));
        bool hit = valid && t < -hw;
        if (valid && (q.w != 0.f || p.w != 0.f || fabsf(t) <= hw)) hit = exact_pair_hit<double, TI>(ad, hg0 + g0 + aa, gj, shp);
        const unsigned bal = __ballot_sync(FULL, hit);
        if (lane == 0) mkT[kc * 34 + aa] = bal;
      }
      if (ad->out.half) {  // half list: the rule of c2_half_word, applied to this chunk's words
        __syncwarp();
        const unsigned word = c2_half_word(valid, valid ? gj - (int)(hg0 + g0) : -1, shp, ng, lane);
        if (lane < ng) mkT[kc * 34 + lane] &= word;
      }
    }
    __syncwarp();
    if (lane < ng) {
      const int fs = fh + g0 + lane;
      mkT[(fs >> 5) * 34 + lane] &= ~(1u << (fs & 31));
    }
    __syncwarp();
    uint32_t* dst = ad->masks + (hg0 + g0) * MASK_WORDS;
    const int nw = ng * MASK_WORDS;
    for (int w0 = 0; w0 < nw; w0 += 32) {
      const int w = w0 + lane;
      const int kk = w & 7, aa = w >> 3;
      uint32_t val = 0;
      if (w < nw && kk < nchunk) val = mkT[kk * 34 + aa];
      if (w < nw) dst[w] = val;
      int c = __popc(val);
      c += __shfl_xor_sync(FULL, c, 1);
      c += __shfl_xor_sync(FULL, c, 2);
      c += __shfl_xor_sync(FULL, c, 4);
      if (w < nw && kk == 0) ad->out.counts[ad->rec.pidx[hg0 + g0 + aa]] = (uint32_t)c;
    }
  }
}

// One home cell with NCH chunks of candidates (NCH = ceil(ncand / 32) exactly).  Returns true if some decision of the cell
// could not be trusted (band, flagged slot): the caller queues the cell for c2_cell_slow.
template <class TI, int NCH, bool HALF>
__device__ __forceinline__ bool c2_cell(const MaskArgs<double, TI>& a, unsigned wofs, int lane, int hstart, int nh, long long hg0, int ncand, int fh,
                                        int hbase, int hsh) {
  // HALF: the table entries carry the stencil cell in their upper 5 bits; lane c < 27 holds, for stencil cell c, hbase = sorted index
  // of its first atom minus its first slot minus hg0 (so hbase + slot = sorted index relative to the home cell) and hsh = its loop shift
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TABCAP = cm_tabcap(CM_MASK);
  constexpr int OFF_SQ = 3 * TILE_VPAD * 4 + 64 * 4 + (TILE_NT / 32) * cm_warp_bytes(CM_MASK);
  // this warp's shared-memory pointers are rebuilt from one opaque offset: carried across the cell loop they get spilled,
  // and with 4 x 48 KB of shared memory per SM the L1 that is left does not hold the stacks
  asm volatile("" : "+r"(wofs));
  const float4* const sq = (const float4*)(smem_raw + OFF_SQ);
  const uint16_t* const cslot = (const uint16_t*)(smem_raw + wofs);
  uint32_t* const mkT = (uint32_t*)(smem_raw + wofs + TABCAP * 3);
  float* const hb = (float*)(smem_raw + wofs + TABCAP * 3 + MASK_WORDS * 34 * 4);
  const float mid = a.mid, hw = a.hw;
  const float2 nmid2 = make_float2(-mid, -mid);
  float qx[NCH], qy[NCH], qz[NCH];
  int hd[HALF ? NCH : 1], hshp[HALF ? NCH : 1];  // half lists: sorted index of the candidate relative to the home cell's first atom, loop shift
  unsigned badm = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) {
    const int f = k * 32 + lane;
    float4 q = make_float4(CAND_FAR, CAND_FAR, CAND_FAR, 0.f);
    const unsigned cs = f < ncand ? (unsigned)cslot[f] : 0u;
    if (f < ncand) q = sq[HALF ? (cs & 2047u) : cs];
    qx[k] = q.x; qy[k] = q.y; qz[k] = q.z;
    if (q.w != 0.f) badm |= 1u << k;
    if constexpr (HALF) {
      hd[k] = __shfl_sync(FULL, hbase, (int)(cs >> 11)) + (int)(cs & 2047u);
      hshp[k] = __shfl_sync(FULL, hsh, (int)(cs >> 11));
    }
  }
  bool rare = badm != 0;

  for (int g0 = 0; g0 < nh; g0 += 32) {
    const int ng = min(32, nh - g0);
    const int npair = (ng + 1) >> 1;
    __syncwarp();
    bool my_bad;
    {
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < ng) p = sq[hstart + g0 + lane];
      float* d = hb + (lane >> 1) * 8 + (lane & 1);
      d[0] = p.x; d[2] = p.y; d[4] = p.z;
      my_bad = (lane < ng) && p.w != 0.f;
    }
    __syncwarp();

    float bmin = 3.0e38f;  // min |t| over this lane's pairs: <= hw means some pair fell in the uncertainty band
#pragma unroll 1
    for (int pr = 0; pr < npair; pr++) {
      const float4 A = *(const float4*)(hb + pr * 8);
      const float2 Z = *(const float2*)(hb + pr * 8 + 4);
      const float2 Ax = make_float2(A.x, A.y), Ay = make_float2(A.z, A.w);
#pragma unroll
      for (int k = 0; k < NCH; k++) {
        const float2 dx = add2_rn(Ax, make_float2(-qx[k], -qx[k]));
        const float2 dy = add2_rn(Ay, make_float2(-qy[k], -qy[k]));
        const float2 dz = add2_rn(Z, make_float2(-qz[k], -qz[k]));
        float2 t = fma2_rn(dx, dx, nmid2);
        t = fma2_rn(dy, dy, t);
        t = fma2_rn(dz, dz, t);
        bmin = fminf(bmin, fminf(fabsf(t.x), fabsf(t.y)));
        const unsigned b0 = __ballot_sync(FULL, t.x < -hw);
        const unsigned b1 = __ballot_sync(FULL, t.y < -hw);
        *(uint2*)(mkT + k * 34 + 2 * pr) = make_uint2(b0, b1);  // every lane stores the same words: no branch
      }
    }
    __syncwarp();
    rare = rare || bmin <= hw || my_bad;
    if constexpr (HALF) {
#pragma unroll
      for (int k = 0; k < NCH; k++) {
        const bool valid = k * 32 + lane < ncand;
        const unsigned word = c2_half_word(valid, valid ? hd[k] - g0 : -1, hshp[k], ng, lane);
        if (lane < ng) mkT[k * 34 + lane] &= word;
      }
      __syncwarp();
    }
    // ---- drop the self pair (same atom, zero shift): flat index fh + g0 + aa of home atom aa
    if (lane < ng) {
      const int fs = fh + g0 + lane;
      mkT[(fs >> 5) * 34 + lane] &= ~(1u << (fs & 31));
    }
    __syncwarp();
    // ---- masks to global (atom-major, MASK_WORDS per atom); per-atom counts = popcount over the atom's 8 words
    uint32_t* dst = a.masks + (hg0 + g0) * MASK_WORDS;
    const int nw = ng * MASK_WORDS;
    for (int w0 = 0; w0 < nw; w0 += 32) {
      const int w = w0 + lane;
      const int kk = w & 7, aa = w >> 3;
      uint32_t val = 0;
      if (w < nw && kk < NCH) val = mkT[kk * 34 + aa];
      if (w < nw) dst[w] = val;
      int c = __popc(val);
      c += __shfl_xor_sync(FULL, c, 1);
      c += __shfl_xor_sync(FULL, c, 2);
      c += __shfl_xor_sync(FULL, c, 4);
      if (w < nw && kk == 0) a.out.counts[a.rec.pidx[hg0 + g0 + aa]] = (uint32_t)c;
    }
  }
  return __any_sync(FULL, rare);
}

template <class TI, bool HALF>
__global__ void __launch_bounds__(TILE_NT, NL_C2_MINB) k_count_mask2(const MaskArgs<double, TI> a) {
  typedef double T;
  constexpr int TABCAP = cm_tabcap(CM_MASK);
  constexpr int CNT_WARP_BYTES = cm_warp_bytes(CM_MASK);
  constexpr int CAPSLOTS = cm_cap(CM_MASK);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* vstart = (int*)smem_raw;
  int* vgs = vstart + TILE_VPAD;
  int* vsh = vgs + TILE_VPAD;
  int* hcell = vsh + TILE_VPAD;  // [64] packed (lx, ly, lz) of each home cell
  unsigned char* wbase = (unsigned char*)(hcell + 64);
  float4* sq = (float4*)(wbase + (TILE_NT / 32) * CNT_WARP_BYTES);
  __shared__ int scan_sm[33];
  __shared__ int s_next;

  const Geo<T>& g = a.g;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned char* wb = wbase + wid * CNT_WARP_BYTES;
  const unsigned wofs = (unsigned)(wb - smem_raw);
  uint16_t* cslot = (uint16_t*)wb;
  uint32_t* mkT = (uint32_t*)(wb + TABCAP * 3);                   // [MASK_WORDS][34]: word k of home atom aa at k * 34 + aa
  float* hb = (float*)(wb + TABCAP * 3 + MASK_WORDS * 34 * 4);    // [16 pairs][8]: x0 x1 y0 y1 z0 z1 - -
  uint8_t* rare_list = (uint8_t*)(wb + TABCAP * 2);               // [<= 64] home cells queued for the exact pass (the unused cv table)
  int nrare = 0;

  const int b = blockIdx.x;
  const int bz = b / (a.ntx * a.nty);
  const int hx0 = (b % a.ntx) * a.tx, hy0 = ((b / a.ntx) % a.nty) * a.ty, hz0 = (a.zlayers ? a.zlayers[bz] : bz) * a.tz;
  const int hxn = min(a.tx, g.nc[0] - hx0), hyn = min(a.ty, g.nc[1] - hy0), hzn = min(a.tz, g.nc[2] - hz0);
  const int VX = hxn + 2, VY = hyn + 2, VZ = hzn + 2, NV = VX * VY * VZ;
  {  // quick reject: tiles without a single home atom (a slab shard sees the global grid, mostly empty)
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

  if (total > CAPSLOTS) {  // denser than the staging capacity: the whole tile takes the generic route
    for (int hc = wid; hc < nhome; hc += TILE_NT / 32) {
      const int lx = hcell[hc] & 255, ly = (hcell[hc] >> 8) & 255, lz = hcell[hc] >> 16;
      const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
      const int nh = vstart[vh + 1] - vstart[vh];
      generic_cell<T, TI, MODE_COUNT>(a.self, (long long)vgs[vh], nh, lane);
      if (lane == 0) a.cellflag[cell_linear(g.nc, hx0 + lx, hy0 + ly, hz0 + lz)] = 0;
    }
    return;
  }

  // ---- stage: Float32 position of every slot's periodic image relative to the tile origin (see nl_mask.cuh)
  {
    double O[3];
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
  }
  __syncthreads();

  while (true) {
    int hc = 0;
    if (lane == 0) hc = atomicAdd(&s_next, 1);
    hc = __shfl_sync(FULL, hc, 0);
    if (hc >= nhome) break;
    const int hcv = hcell[hc];
    const int lx = hcv & 255, ly = (hcv >> 8) & 255, lz = hcv >> 16;
    const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
    const int hstart = vstart[vh], nh = vstart[vh + 1] - hstart;
    if (nh == 0) continue;
    const long long hg0 = vgs[vh];
    // candidate table: flat candidate -> staged slot (lane c < 27 owns stencil cell c); fh = flat index of home atom 0
    int ncand, fh;
    int hbase = 0, hsh = 0;
    static_assert(cm_cap(CM_MASK) <= 2048, "staged slots must fit 11 bits beside the stencil-cell tag");
    {
      int st = 0, cn = 0;
      int l2 = lane;
      unsigned wo = wofs;
      asm volatile("" : "+r"(l2), "+r"(wo));
      uint16_t* const cs = (uint16_t*)(smem_raw + wo);
      if (l2 < 27) {
        const int v = ((lz + l2 / 9) * VY + (ly + (l2 / 3) % 3)) * VX + (lx + l2 % 3);
        st = vstart[v];
        cn = vstart[v + 1] - st;
        if constexpr (HALF) { hbase = vgs[v] - st - (int)hg0; hsh = vsh[v]; }
      }
      const int incl = warp_incl_scan(cn, l2);
      ncand = __shfl_sync(FULL, incl, 31);
      fh = __shfl_sync(FULL, incl, 12);
      if (ncand <= TABCAP) {
        const int pre = incl - cn;
        const int mx = __reduce_max_sync(FULL, cn);
        const unsigned tag = HALF ? (unsigned)l2 << 11 : 0u;   // half lists: stencil cell of the candidate in the upper 5 bits
        for (int j = 0; j < mx; j++)
          if (j < cn) cs[pre + j] = (uint16_t)(tag | (unsigned)(st + j));
      }
      __syncwarp();
    }
    if (lane == 0) a.cellflag[cell_linear(g.nc, hx0 + lx, hy0 + ly, hz0 + lz)] = ncand <= TABCAP ? 1 : 0;
    if (ncand > TABCAP) {  // too many candidates for the 256-bit masks: generic route (the fill pass does the same)
      generic_cell<T, TI, MODE_COUNT>(a.self, hg0, nh, lane);
      continue;
    }
    bool rare;
    switch ((ncand + 31) >> 5) {
      case 1: rare = c2_cell<TI, 1, HALF>(a, wofs, lane, hstart, nh, hg0, ncand, fh, hbase, hsh); break;
      case 2: rare = c2_cell<TI, 2, HALF>(a, wofs, lane, hstart, nh, hg0, ncand, fh, hbase, hsh); break;
      case 3: rare = c2_cell<TI, 3, HALF>(a, wofs, lane, hstart, nh, hg0, ncand, fh, hbase, hsh); break;
      case 4: rare = c2_cell<TI, 4, HALF>(a, wofs, lane, hstart, nh, hg0, ncand, fh, hbase, hsh); break;
      case 5: rare = c2_cell<TI, 5, HALF>(a, wofs, lane, hstart, nh, hg0, ncand, fh, hbase, hsh); break;
      case 6: rare = c2_cell<TI, 6, HALF>(a, wofs, lane, hstart, nh, hg0, ncand, fh, hbase, hsh); break;
      case 7: rare = c2_cell<TI, 7, HALF>(a, wofs, lane, hstart, nh, hg0, ncand, fh, hbase, hsh); break;
      default: rare = c2_cell<TI, 8, HALF>(a, wofs, lane, hstart, nh, hg0, ncand, fh, hbase, hsh); break;
    }
    if (rare) {
      if (lane == 0) rare_list[nrare] = (uint8_t)hc;
      nrare++;
    }
  }
  __syncwarp();
  for (int r = 0; r < nrare; r++) c2_cell_slow<TI>(a.self, sq, cslot, mkT, vstart, vgs, vsh, NV, VX, VY, hcell[rare_list[r]]);
}

}  // namespace nl
