// nl_fill2.cuh -- round-2 fill pass of the mask route: complete sectors in place, row boundaries parked.
//
// Why.  The CSR rows are laid out in ORIGINAL atom order (`first`, src/gpu_kernels.jl:159-180) while the atoms are
// processed in cell-sorted order, so with randomly ordered inputs every ~1.1 KB row lands at a random place of the
// output.  Measured (experiments/microbench_rows2.cu, experiments/microbench_hybrid.cu): what makes that slow is not
// the random placement but the PARTIAL 32-byte sectors at both ends of every row segment -- they are shared with the
// neighbouring rows, which other warps write at other times, so each one costs a DRAM read-modify-write.
//
// What.  Three kernels replace k_fill_mask + its i stream:
//   k_fill_park       the mask expansion (sorted order).  Every row segment is staged in shared memory with the
//                     destination's 16-byte phase and written with 16-byte stores -- but only the sectors the row covers
//                     COMPLETELY.  The <= 7 leading and <= 7 trailing elements of each stream are PARKED in a per-row
//                     record (j head | j tail | S head | S tail: 128 B; R head | R tail: 64 B), addressed by position
//                     inside the sector, each 32-byte block written whole.
//   k_fix_boundaries  original row order: the thread of the first row boundary inside a sector assembles all 32 bytes
//                     from the parked pieces of the rows that share it and writes the sector once.
//   k_expand_rows     the i stream is a pure function of `first`: a streaming kernel, 16-byte stores.
// The kernel itself is leaner than k_fill_mask (one packed table entry per candidate, staged records as two 16-byte
// halves, no per-slot row starts in shared memory, 512 threads per tile at <= 64 registers: 32 warps / SM instead of 16).
#pragma once
#include "nl_mask.cuh"

namespace nl {

#ifndef NL_F2_OPAQUE
#define NL_F2_OPAQUE 1
#endif
#ifndef NL_F2_NT
#define NL_F2_NT 448   // 14 warps x 2 CTAs / SM at <= 72 registers: measured best for k_fill3 (fill stage 4.84 ms at 384, 4.65 at 448, 5.2 at 480)
#endif
constexpr int F2_NT = NL_F2_NT;
constexpr int F2_NW = F2_NT / 32;
#ifndef NL_F2_SMEM_KB
#define NL_F2_SMEM_KB 104
#endif
constexpr int F2_SMEM_BYTES = NL_F2_SMEM_KB * 1024;
constexpr int PARK_A_BYTES = 128, PARK_R_BYTES = 64;

template <class T> __host__ __device__ constexpr int f2_slot_bytes() { return sizeof(T) == 8 ? 32 : 20; }
// staging buffers of one row chunk (<= 32 pairs), each with room for the destination's phase inside its 32-byte sector
template <class T, class TI> __host__ __device__ constexpr int f2_bufJ() { return 32 + 32 * (int)sizeof(TI); }
template <class T, class TI> __host__ __device__ constexpr int f2_bufS() { return 32 + 96 * (int)sizeof(TI); }
template <class T, class TI> __host__ __device__ constexpr int f2_bufR() { return 32 + 96 * (int)sizeof(T) + 32; }  // + slack: tail blocks are read whole
// per warp: candidate table (u16) | 4 hit lists | shift code per stencil cell | staging buffers
template <class T, class TI> __host__ __device__ constexpr int f2_warp_bytes() {
  return 2 * MASK_MAXCAND + 4 * MASK_MAXCAND + 32 + f2_bufJ<T, TI>() + f2_bufS<T, TI>() + f2_bufR<T, TI>();
}
template <class T, class TI> __host__ __device__ constexpr int f2_fixed_bytes() {
  return 3 * TILE_VPAD * 4 + 64 * 4 + 64 * 4 * (int)sizeof(T) + F2_NW * f2_warp_bytes<T, TI>();
}
template <class T, class TI> __host__ __device__ constexpr int f2_cap() {
  return ((F2_SMEM_BYTES - f2_fixed_bytes<T, TI>()) / f2_slot_bytes<T>() / 8 * 8) > 2040 ? 2040
                                                                                          : ((F2_SMEM_BYTES - f2_fixed_bytes<T, TI>()) / f2_slot_bytes<T>() / 8 * 8);
}
static_assert(f2_warp_bytes<double, int32_t>() % 32 == 0 && f2_warp_bytes<float, int64_t>() % 32 == 0 && f2_fixed_bytes<double, int32_t>() % 32 == 0, "alignment");

// Staged record of one slot: Float64 -> {x, y} | {z, idx, w}; Float32 -> {x, y, z, idx} | {w}.
template <class T>
__device__ __forceinline__ void f2_store(unsigned char* sA, unsigned char* sB, int slot, T x, T y, T z, uint32_t idx, uint32_t w) {
  if constexpr (sizeof(T) == 8) {
    ((double2*)sA)[slot] = make_double2(x, y);
    ((int4*)sB)[slot] = make_int4(__double2loint(z), __double2hiint(z), (int)idx, (int)w);
  } else {
    ((float4*)sA)[slot] = make_float4(x, y, z, __uint_as_float(idx));
    ((uint32_t*)sB)[slot] = w;
  }
}
template <class T>
__device__ __forceinline__ void f2_load(const unsigned char* sA, const unsigned char* sB, int slot, T& x, T& y, T& z, uint32_t& idx, uint32_t& w) {
  if constexpr (sizeof(T) == 8) {
    const double2 A = ((const double2*)sA)[slot];
    const int4 B = ((const int4*)sB)[slot];
    x = A.x; y = A.y; z = __hiloint2double(B.y, B.x); idx = (uint32_t)B.z; w = (uint32_t)B.w;
  } else {
    const float4 A = ((const float4*)sA)[slot];
    x = A.x; y = A.y; z = A.z; idx = __float_as_uint(A.w); w = ((const uint32_t*)sB)[slot];
  }
}

// One stream of one row chunk (nr <= 32 pairs starting at pair p0, EB bytes per pair).  The chunk is staged at
// buf + o (o = phase of its first byte inside a 32-byte sector), so buf + 32 * k is sector k of the destination:
//   head  = the hb bytes before the first sector boundary  -> buf[0, 32) IS the head block of the park record
//   tail  = the tb bytes after the last sector boundary    -> buf[o + nb - tb, +32) IS the tail block
//   what lies between is complete sectors, copied in place with 16-byte stores.
// Chunks of a long row are cut at multiples of 8 pairs, which are sector boundaries in every stream, so only a row's
// true ends are partial.
struct Seg { int o, nb, hb, tb; };
template <int EB> __device__ __forceinline__ Seg make_seg(uint32_t p0_low, int nr) {
  Seg s;
  s.o = (int)((p0_low * (uint32_t)EB) & 31u);
  s.nb = nr * EB;
  const int h = (32 - s.o) & 31;
  s.hb = h < s.nb ? h : s.nb;
  s.tb = (s.nb - s.hb) & 31;
  return s;
}

// Rows the generic per-atom route wrote in place (cells beyond the mask capacity, tiles beyond the staging capacity): copy
// their boundary elements into the park record, so that k_fix_boundaries can treat every row alike.
template <class T, class TI>
__device__ __noinline__ void park_row_from_output(const Sinks<T, TI>* out, unsigned char* parkA, unsigned char* parkR, uint32_t io) {
  const long long b = (long long)out->first[io] - 1, e = (long long)out->first[io + 1] - 1;
  if (e <= b) return;
  constexpr int NSI = 32 / (int)sizeof(TI), NSR = 32 / (int)sizeof(T);
  for (int stream = 0; stream < 3; stream++) {
    if (stream == 2 && !out->Ro) break;
    const int es = stream == 2 ? (int)sizeof(T) : (int)sizeof(TI);
    const int eb = stream == 0 ? es : 3 * es;
    const long long B0 = b * eb, B1 = e * eb;
    const long long nbl = B1 - B0;
    const int h = (int)((-B0) & 31);
    const int hb = h < nbl ? h : (int)nbl;
    const int tb = (int)((nbl - hb) & 31);
    const char* g = stream == 0 ? (const char*)out->jo : (stream == 1 ? (const char*)out->So : (const char*)out->Ro);
    unsigned char* rec = stream == 2 ? parkR + (size_t)io * PARK_R_BYTES : parkA + (size_t)io * PARK_A_BYTES + (stream == 1 ? 64 : 0);
    const int ns = stream == 2 ? NSR : NSI;
    for (int k = 0; k < ns; k++) {
      const int relh = k * es - (int)(B0 & 31);
      if (relh >= 0 && relh < hb) {
        if (es == 8) *(unsigned long long*)(rec + k * 8) = *(const unsigned long long*)(g + B0 + relh);
        else *(uint32_t*)(rec + k * 4) = *(const uint32_t*)(g + B0 + relh);
      }
      if (k * es < tb) {
        if (es == 8) *(unsigned long long*)(rec + 32 + k * 8) = *(const unsigned long long*)(g + B1 - tb + k * 8);
        else *(uint32_t*)(rec + 32 + k * 4) = *(const uint32_t*)(g + B1 - tb + k * 4);
      }
    }
  }
}
template <class T, class TI>
__device__ __noinline__ void generic_cell_park(const MaskArgs<T, TI>* ad, unsigned char* parkA, unsigned char* parkR, long long g0, int n, int lane) {
  for (int k = lane; k < n; k += 32) {
    generic_atom<T, TI, MODE_FILL>(g0 + k, ad->rec, ad->co, ad->g, ad->out);
    const uint32_t io = ad->rec.pidx[g0 + k];
    if ((long long)io < ad->out.n_rows) park_row_from_output<T, TI>(&ad->out, parkA, parkR, io);
  }
}

template <int NT>
__device__ __forceinline__ int tile_table_nt(const int nc[3], const int pbc[3], const void* co, bool co64, int hx0, int hy0, int hz0, int VX, int VY, int NV,
                                             int* vstart, int* vgs, int* vsh, int* scan_sm) {
  const int tid = threadIdx.x;
  int cnt = 0, gs = 0, sh = 0;
  if (tid < NV) {
    int cx, cy, cz, s0, s1, s2;
    bool ok = map_virtual(hx0 + tid % VX - 1, nc[0], pbc[0], cx, s0);
    ok = map_virtual(hy0 + (tid / VX) % VY - 1, nc[1], pbc[1], cy, s1) && ok;
    ok = map_virtual(hz0 + tid / (VX * VY) - 1, nc[2], pbc[2], cz, s2) && ok;
    if (ok) {
      const long long cl = (long long)cx + (long long)nc[0] * ((long long)cy + (long long)nc[1] * cz);
      const long long c0 = co64 ? (long long)((const long long*)co)[cl] : (long long)((const int*)co)[cl];
      const long long c1 = co64 ? (long long)((const long long*)co)[cl + 1] : (long long)((const int*)co)[cl + 1];
      gs = (int)(c0 - 1);
      cnt = (int)(c1 - c0);
      sh = pack_shift(s0, s1, s2);
    }
  }
  int total;
  const int excl = block_excl_scan<int, NT>(cnt, scan_sm, &total);
  if (tid < NV) { vstart[tid] = excl; vgs[tid] = gs; vsh[tid] = sh; }
  if (tid == NV) vstart[NV] = total;
  return total;
}

constexpr int SHP_ZERO = 1 | (1 << 2) | (1 << 4);  // pack_shift(0, 0, 0)

template <class T, class TI, bool PARK>
__global__ void __launch_bounds__(F2_NT, 2) k_fill_park(const MaskArgs<T, TI> a, unsigned char* __restrict__ parkA, unsigned char* __restrict__ parkR, int prefetch,
                                                        int s_zeroed) {
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
  char* bufS = bufJ + f2_bufJ<T, TI>();
  char* bufR = bufS + f2_bufS<T, TI>();

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

  if (total > CAP) {  // denser than the staging capacity: generic route, in place, boundaries parked afterwards
    for (int hc = wid; hc < nhome; hc += F2_NW) {
      const int lx = hcell[hc] & 255, ly = (hcell[hc] >> 8) & 255, lz = hcell[hc] >> 16;
      const int vh = ((lz + 1) * VY + (ly + 1)) * VX + (lx + 1);
      if (PARK) generic_cell_park<T, TI>(a.self, parkA, parkR, (long long)vgs[vh], vstart[vh + 1] - vstart[vh], lane);
      else generic_cell<T, TI, MODE_FILL>(a.self, (long long)vgs[vh], vstart[vh + 1] - vstart[vh], lane);
    }
    return;
  }
  const uint32_t* __restrict__ idsrc = a.out.pgid0 ? a.out.pgid0 : a.rec.pidx;  // what j publishes (shard mode: global index - 1)
  for (int sl = tid; sl < total; sl += F2_NT) {
    const int v = find_vcell(vstart, NV, sl);
    const long long src = (long long)vgs[v] + (sl - vstart[v]);
    f2_store<T>(sA, sB, sl, a.rec.px[src], a.rec.py[src], a.rec.pz[src], idsrc[src], a.rec.pw[src]);
  }
  if (tid < nhome) {
    const int lx = hcell[tid] & 255, ly = (hcell[tid] >> 8) & 255, lz = (hcell[tid] >> 16) & 255;
    if (a.cellflag[cell_linear(g.nc, hx0 + lx, hy0 + ly, hz0 + lz)]) hcell[tid] |= 1 << 24;
  }
  __syncthreads();
  const BaseT* __restrict__ srow = (const BaseT*)a.srow;
  const bool has_R = a.out.Ro != nullptr;

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
      if (PARK) generic_cell_park<T, TI>(a.self, parkA, parkR, hg0, nh, lane);
      else generic_cell<T, TI, MODE_FILL>(a.self, hg0, nh, lane);
      continue;
    }
    // row data of the first two passes (lane = 8 * atom + mask word), issued before the table building so that it hides them
    uint32_t nx_word = 0, nx2_word = 0, nx_io = 0, nx2_io = 0;
    BaseT nx_base = NOROW, nx2_base = NOROW;
    if (grp < nh) { nx_word = a.masks[(hg0 + grp) * MASK_WORDS + sub]; nx_base = srow[hg0 + grp]; nx_io = a.rec.pidx[hg0 + grp]; }
    if (4 + grp < nh) { nx2_word = a.masks[(hg0 + 4 + grp) * MASK_WORDS + sub]; nx2_base = srow[hg0 + 4 + grp]; nx2_io = a.rec.pidx[hg0 + 4 + grp]; }
    // candidate table: flat candidate -> staged slot | stencil cell; packed shift per stencil cell
    bool fastcell;
    {
      int st = 0, cn = 0, shp = SHP_ZERO;
      if (lane < 27) {
        const int v = ((lz + lane / 9) * VY + (ly + (lane / 3) % 3)) * VX + (lx + lane % 3);
        st = vstart[v];
        cn = vstart[v + 1] - st;
        if (cn > 0) shp = vsh[v];
        shc[lane] = (uint8_t)shp;
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

    for (int a0 = 0; a0 < nh; a0 += 4) {
      uint32_t word = nx_word;
      const BaseT my_base = nx_base;
      const uint32_t my_io = nx_io;
      nx_word = nx2_word; nx_base = nx2_base; nx_io = nx2_io;
      nx2_word = 0; nx2_base = NOROW; nx2_io = 0;
      if (a0 + 8 + grp < nh) {
        nx2_word = a.masks[(hg0 + a0 + 8 + grp) * MASK_WORDS + sub];
        nx2_base = srow[hg0 + a0 + 8 + grp];
        nx2_io = a.rec.pidx[hg0 + a0 + 8 + grp];
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
      if (!PARK && prefetch && my_nhit > 0 && sub < (has_R ? 6 : 4) && sub >= (prefetch >= 3 ? (prefetch == 3 ? 4 : 2) : 0) &&
          !(s_zeroed && (sub >> 1) == 1)) {  // 3: R only, 4: S and R (experiments); S rows are rarely written when the stream was zeroed up front
        // In place, the partial 32-byte sectors at both ends of a row segment are shared with the neighbouring rows, which
        // other warps write at other times: evicted half-written, each costs a DRAM read-modify-write.  Prefetching them
        // into L2 now makes the partial write land on a fully valid sector, which is later written back whole
        // (experiments/microbench_rows2.cu: 5.34 -> 4.63 ms for the stores of the headline list).
        const int st = sub >> 1;   // 0 j, 1 S, 2 R
        const long long eb = st == 0 ? (long long)sizeof(TI) : (st == 1 ? 3ll * sizeof(TI) : 3ll * sizeof(T));
        const char* gb = st == 0 ? (const char*)a.out.jo : (st == 1 ? (const char*)a.out.So : (const char*)a.out.Ro);
        const long long B = ((long long)my_base + ((sub & 1) ? my_nhit : 0)) * eb;
        if (B & 31) {
          const char* sec = gb + ((sub & 1) ? ((B - 1) & ~31ll) : (B & ~31ll));
          if (prefetch == 2) asm volatile("cp.async.bulk.prefetch.L2.global [%0], 32;" ::"l"(sec));  // exactly the sector, but issued lane by lane
          else asm volatile("prefetch.global.L2 [%0];" ::"l"(sec));                                  // one instruction, fetches the 128-byte line
        }
      }
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
        const int nhit = __shfl_sync(FULL, my_nhit, q * 8);
        if (nhit == 0) continue;
        const BaseT base = __shfl_sync(FULL, my_base, q * 8);
        const uint32_t io = __shfl_sync(FULL, my_io, q * 8);
        T xi, yi, zi;
        uint32_t idx_i, wi;
        f2_load<T>(sA, sB, hstart + a0 + q, xi, yi, zi, idx_i, wi);
        const bool wi_ok = !(wi & WIND_OVERFLOW);
        const uint8_t* L = lists + q * MASK_MAXCAND;

        for (int r0 = 0; r0 < nhit;) {
          const long long p0 = (long long)base + r0;
          int nr = nhit - r0;
          if (nr > 32) nr = PARK ? 32 - (int)((p0 + 32) & 7) : 32;  // PARK: cut long rows at multiples of 8 pairs (sector boundaries in every stream)
          const bool act = lane < nr;
          constexpr int EJ = (int)sizeof(TI), ES = 3 * (int)sizeof(TI), ER = 3 * (int)sizeof(T);
#if NL_F2_OPAQUE
          // keep the lane-dependent shared-memory addresses of this body from being hoisted out of the row loops: hoisted, they
          // do not fit the register budget of 2 CTAs / SM and get spilled to local memory (measured: LDL stalls in this loop).
          // A 32-bit offset from the shared-memory symbol keeps the address space known (LDS / STS, not generic accesses).
          unsigned wo = wofs;
          asm volatile("" : "+r"(wo));
          uint16_t* const tab = (uint16_t*)(smem_raw + wo);
          const uint8_t* const L = smem_raw + wo + 2 * MASK_MAXCAND + q * MASK_MAXCAND;
          const uint8_t* const shc = smem_raw + wo + 6 * MASK_MAXCAND;
          char* const bufJ = (char*)(smem_raw + wo + 6 * MASK_MAXCAND + 32);
          char* const bufS = bufJ + f2_bufJ<T, TI>();
          char* const bufR = bufS + f2_bufS<T, TI>();
#endif
          bool zs = true;
          int S0 = 0, S1 = 0, S2 = 0;
          T R0 = 0, R1 = 0, R2 = 0;
          uint32_t jv = 0;
          if (act) {
            const int f = (int)L[r0 + lane];
            const unsigned t = tab[f];
            const int slot = (int)(t & 2047u), c = (int)(t >> 11);
            T xj, yj, zj;
            uint32_t wj;
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
          }
          const bool zeroS = __all_sync(FULL, zs);   // the usual case away from the periodic boundary: S is not staged at all
          if constexpr (!PARK) {
            // ---- lean variant: everything in place; j straight from the lane, S and R transposed through shared memory
            TI* const jrow = a.out.jo + p0;
            TI* const Srow = a.out.So + 3 * p0;
            if (act) jrow[lane] = (TI)jv + 1;
            TI* const bS = (TI*)bufS;
            T* const bR = (T*)bufR;
            // S: zero away from the periodic boundary.  With the stream zeroed up front by k_expand_rows (s_zeroed) such
            // chunks skip S altogether: 27 % of the randomly placed bytes become a sequential fill
            const bool writeS = !(zeroS && s_zeroed);
            if (act) {
              if (has_R) { bR[3 * lane] = R0; bR[3 * lane + 1] = R1; bR[3 * lane + 2] = R2; }
              if (!zeroS) { bS[3 * lane] = (TI)S0; bS[3 * lane + 1] = (TI)S1; bS[3 * lane + 2] = (TI)S2; }
            }
            __syncwarp();
            const int nw = 3 * nr;
#pragma unroll
            for (int m = 0; m < 3; m++) {
              const int w = m * 32 + lane;
              if (w < nw) {
                if (writeS) Srow[w] = zeroS ? (TI)0 : bS[w];
                if (has_R) a.out.Ro[3 * p0 + w] = bR[w];
              }
            }
            __syncwarp();
          } else {
          const Seg sj = make_seg<EJ>((uint32_t)p0, nr), sS = make_seg<ES>((uint32_t)p0, nr), sR = make_seg<ER>((uint32_t)p0, nr);
          if (act) {
            *(TI*)(bufJ + sj.o + lane * EJ) = (TI)jv + 1;
            if (has_R) {
              T* d = (T*)(bufR + sR.o + lane * ER);
              d[0] = R0; d[1] = R1; d[2] = R2;
            }
            if (!zeroS) {
              TI* d = (TI*)(bufS + sS.o + lane * ES);
              d[0] = (TI)S0; d[1] = (TI)S1; d[2] = (TI)S2;
            }
          }
          __syncwarp();
          // ---- complete sectors, in place: 16 bytes per lane
          {
            const int n16 = (sj.nb - sj.hb - sj.tb) >> 4;                        // <= 8 (Int32) / 16 (Int64)
            if (lane < n16) *(int4*)((char*)a.out.jo + (p0 * EJ + sj.hb) + 16 * lane) = *(const int4*)(bufJ + sj.o + sj.hb + 16 * lane);
          }
          {
            const int n16 = (sS.nb - sS.hb - sS.tb) >> 4;                        // <= 24 (Int32) / 48 (Int64)
            char* d = (char*)a.out.So + (p0 * ES + sS.hb);
            const char* src = bufS + sS.o + sS.hb;
#pragma unroll
            for (int o = 0; o < (EJ == 4 ? 32 : 64); o += 32)
              if (o + lane < n16) *(int4*)(d + 16 * (o + lane)) = zeroS ? make_int4(0, 0, 0, 0) : *(const int4*)(src + 16 * (o + lane));
          }
          if (has_R) {
            const int n16 = (sR.nb - sR.hb - sR.tb) >> 4;                        // <= 48 (Float64) / 24 (Float32)
            char* d = (char*)a.out.Ro + (p0 * ER + sR.hb);
            const char* src = bufR + sR.o + sR.hb;
#pragma unroll
            for (int o = 0; o < (sizeof(T) == 8 ? 64 : 32); o += 32)
              if (o + lane < n16) *(int4*)(d + 16 * (o + lane)) = *(const int4*)(src + 16 * (o + lane));
          }
          // ---- row ends -> park record of row io: the staged head / tail SECTORS are the record's blocks, copied word by word
          {
            const int blk = lane >> 3;   // 0 j head, 1 j tail, 2 S head, 3 S tail
            const Seg& sg = (blk & 2) ? sS : sj;
            const char* src = ((blk & 2) ? bufS : bufJ) + ((blk & 1) ? sg.o + sg.nb - sg.tb : 0) + 4 * (lane & 7);
            const bool needed = (blk & 1) ? sg.tb > 0 : sg.hb > 0;
            if (needed) {
              uint32_t v = *(const uint32_t*)src;
              if (zeroS && (blk & 2)) v = 0;
              ((uint32_t*)(parkA + (size_t)io * PARK_A_BYTES))[lane] = v;
            }
            if (has_R && lane < 16) {
              const char* srcR = bufR + ((lane & 8) ? sR.o + sR.nb - sR.tb : 0) + 4 * (lane & 7);
              const bool neededR = (lane & 8) ? sR.tb > 0 : sR.hb > 0;
              if (neededR) ((uint32_t*)(parkR + (size_t)io * PARK_R_BYTES))[lane] = *(const uint32_t*)srcR;
            }
          }
          __syncwarp();
          }
          r0 += nr;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Boundary sectors, original row order.  Thread t owns row boundary i = t + 1 (between rows i-1 and i; i = n is the end of
// the arrays) in all streams.  A boundary strictly inside a sector makes that sector partial for the rows on both sides; the
// thread of the FIRST boundary inside a sector assembles its 32 bytes from the parked pieces of every row that touches it.
template <int ES, int NE, class V, class TI>
__device__ __forceinline__ void fix_stream(const TI* __restrict__ first, long long n, long long i, char* gbase, const unsigned char* __restrict__ park,
                                           int rec_bytes, int head_off) {
  constexpr long long EB = (long long)ES * NE;
  const long long Bp = ((long long)first[i] - 1) * EB;
  if ((Bp & 31) == 0) return;
  const long long s0 = Bp & ~31ll;
  if (((long long)first[i - 1] - 1) * EB > s0) return;  // an earlier boundary lies inside this sector: its thread does it
  constexpr int NS = 32 / ES;
  V vals[NS];
  const long long Pend = ((long long)first[n] - 1) * EB;
  long long r = i - 1;   // the row that holds the first byte of the sector: it started at or before s0, so this is its tail
  long long rB1 = Bp;
  bool before = true;
#pragma unroll
  for (int k = 0; k < NS; k++) {
    const long long adr = s0 + (long long)k * ES;
    while (adr >= rB1 && r < n - 1) { r++; rB1 = ((long long)first[r + 1] - 1) * EB; before = false; }
    V v = V(0);
    if (adr < Pend && adr < rB1) v = *(const V*)(park + (size_t)r * rec_bytes + head_off + (before ? 32 : 0) + k * ES);
    vals[k] = v;
  }
  if (s0 + 32 <= Pend) {
    int4* d = (int4*)(gbase + s0);
    d[0] = *(int4*)&vals[0];
    d[1] = *(int4*)&vals[16 / ES];
  } else {
#pragma unroll
    for (int k = 0; k < NS; k++)
      if (s0 + (long long)k * ES < Pend) *(V*)(gbase + s0 + k * ES) = vals[k];
  }
}
template <class T, class TI>
__global__ void __launch_bounds__(256) k_fix_boundaries(const TI* __restrict__ first, long long n_rows, TI* __restrict__ jo, TI* __restrict__ So,
                                                        T* __restrict__ Ro, const unsigned char* __restrict__ parkA, const unsigned char* __restrict__ parkR) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (i > n_rows) return;
  fix_stream<(int)sizeof(TI), 1, TI, TI>(first, n_rows, i, (char*)jo, parkA, PARK_A_BYTES, 0);
  fix_stream<(int)sizeof(TI), 3, TI, TI>(first, n_rows, i, (char*)So, parkA, PARK_A_BYTES, 64);
  if (Ro) fix_stream<(int)sizeof(T), 3, T, TI>(first, n_rows, i, (char*)Ro, parkR, PARK_R_BYTES, 0);
}

// ------------------------------------------------------------------------------------------------
// i stream: i[p] = (row of pair p) for p in [0, first[n_rows] - 1); rows through gmap in shard mode.  One block per
// EXP_RB consecutive ROWS, whose pairs are one contiguous range of the output: first[] of those rows sits in shared memory,
// 16-byte stores, 512 contiguous bytes per warp instruction.
#ifndef NL_EXP_NT
#define NL_EXP_NT 512   // A/B (experiments/README.md): 512 threads x 256 rows per block: fill stage -0.08 ms against 256 x 512
#endif
#ifndef NL_EXP_RB
#define NL_EXP_RB 256
#endif
constexpr int EXP_NT = NL_EXP_NT;
constexpr int EXP_RB = NL_EXP_RB;
template <class TI>
__global__ void __launch_bounds__(EXP_NT) k_expand_rows(const TI* __restrict__ first, long long n_rows, const TI* __restrict__ gmap, TI* __restrict__ io,
                                                        TI* __restrict__ Szero) {
  __shared__ long long sfirst[EXP_RB + 1];
  const long long r0 = (long long)blockIdx.x * EXP_RB;
  const int nr = (int)min((long long)EXP_RB, n_rows - r0);
  for (int k = threadIdx.x; k <= nr; k += EXP_NT) sfirst[k] = (long long)first[r0 + k] - 1;
  __syncthreads();
  const long long P0 = sfirst[0], P1 = sfirst[nr];
  constexpr int VEC = 16 / (int)sizeof(TI);
  if (Szero) {
    // S of this block's pairs := 0 (the fill pass then writes only the rows that cross a periodic boundary): a pure streaming
    // fill, 16-byte stores over the aligned interior, element stores for the <= 3 words at either end
    const long long w0 = 3 * P0, w1 = 3 * P1;                          // word range
    const long long a0 = (w0 + VEC - 1) & ~(long long)(VEC - 1), a1 = w1 & ~(long long)(VEC - 1);
    if (a0 >= a1) {
      for (long long w = w0 + threadIdx.x; w < w1; w += EXP_NT) Szero[w] = (TI)0;
    } else {
      if (threadIdx.x < a0 - w0) Szero[w0 + threadIdx.x] = (TI)0;
      if (threadIdx.x < w1 - a1) Szero[a1 + threadIdx.x] = (TI)0;
      int4* d = (int4*)(Szero + a0);
      const long long nv = (a1 - a0) / VEC;
      for (long long v = threadIdx.x; v < nv; v += EXP_NT) __stcs(d + v, make_int4(0, 0, 0, 0));
    }
  }
  // A warp covers 128 consecutive 16-byte pieces (aligned in the GLOBAL pair index) per round; lane l owns pieces l, l + 32,
  // l + 64, l + 96, so that every store instruction of the warp writes 512 contiguous bytes.  One binary search for the first
  // piece, then a linear walk along first[] (rows are tens of pairs long).  The first and last piece of a block may be
  // shared with the neighbouring blocks and are written element by element.
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int WPAIRS = 128 * VEC;   // pairs per warp round
  for (long long w0 = (P0 & ~(long long)(VEC - 1)) + (long long)wid * WPAIRS; w0 < P1; w0 += (long long)(EXP_NT / 32) * WPAIRS) {
    long long q = w0 + (long long)lane * VEC;
    const long long qs = q < P0 ? P0 : (q >= P1 ? P1 - 1 : q);
    int l = 0, h = nr;
    while (h - l > 1) { const int mid = (l + h) >> 1; if (sfirst[mid] <= qs) l = mid; else h = mid; }
    long long rend = sfirst[l + 1];
    TI cur = gmap ? gmap[r0 + l] : (TI)(r0 + l + 1);
#pragma unroll
    for (int c = 0; c < 4; c++, q += 32 * VEC) {
      if (q >= P1) break;
      TI v[VEC];
#pragma unroll
      for (int u = 0; u < VEC; u++) {
        while (q + u >= rend && l < nr - 1) {
          l++;
          rend = sfirst[l + 1];
          cur = gmap ? gmap[r0 + l] : (TI)(r0 + l + 1);
        }
        v[u] = cur;
      }
      if (q >= P0 && q + VEC <= P1) *(int4*)(io + q) = *(int4*)v;
      else {
#pragma unroll
        for (int u = 0; u < VEC; u++)
          if (q + u >= P0 && q + u < P1) io[q + u] = v[u];
      }
    }
  }
}

}  // namespace nl
