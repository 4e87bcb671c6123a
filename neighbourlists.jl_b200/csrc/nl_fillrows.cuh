// nl_fillrows.cuh -- the pair-fill pass in ORIGINAL atom order, one thread per output pair.
//
// Why original order: the CSR rows live in original atom order (the reference's `first`), but the atoms are
// processed most naturally in cell-sorted order, which scatters the ~1.1 KB rows all over the 11 GB output.
// Measured on B200 (scripts/microbench_rows.cu, 10 M rows, 261.8 M pairs, pure stores): rows visited in random
// order 6.5 ms, in sequential order 3.9 ms with one warp per row, cudaMemset of the same volume 1.6 ms.  Every
// sorted-order fill kernel tried here (tile-staged, list-driven, batched) sat at 7.7-8.9 ms for that reason.
//
// So: a block takes FR_RB consecutive ORIGINAL rows, i.e. one contiguous span of every output array.  It
// loads, per row, the atom's hit mask (sorted order, from k_count_mask), its home record and the 27 stencil
// cells of its home cell (first sorted index, population, periodic shift) into shared memory.  Then every
// thread owns one output pair p of the span: row by binary search in the block's slice of `first`, rank inside
// the row, rank-th set bit of the 256-bit mask, stencil cell by binary search, candidate's 32-byte AoS record
// by one gather, S and R by the contract.  i and j are stored with unit stride; S and R are transposed through
// shared memory and stored with unit stride as well: all four output streams are fully coalesced.
#pragma once
#include "nl_mask.cuh"

namespace nl {

constexpr int FR_NT = 256;
constexpr int FR_RB = 64;  // rows per block

template <class T, class TI> struct FillRowsArgs {
  const RecAoS<T>* ra;        // AoS records, sorted order; idx = index to publish (local, or global in shard mode)
  const uint32_t* pkey;       // 0-based linear cell of each sorted atom
  const uint32_t* sorted_of;  // original index -> sorted index
  const uint32_t* masks;      // hit masks, MASK_WORDS per sorted atom
  const uint8_t* cellflag;    // per cell: masks valid
  const TI* co;               // cell_offsets (1-based)
  Records<T> rec;             // SoA records (generic route only)
  Geo<T> g;
  Sinks<T, TI> out;
  long long n;
  const FillRowsArgs<T, TI>* self;  // this block in global memory (rare out-of-line paths)
};

// sorted_of = inverse of pidx; published index of every record (global id in shard mode).
template <class T, class TI>
__global__ void __launch_bounds__(256) k_fillrows_prologue(const uint32_t* __restrict__ pidx, const TI* __restrict__ gmap, long long n,
                                                           uint32_t* __restrict__ sorted_of, RecAoS<T>* __restrict__ ra) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t io = pidx[s];
  sorted_of[io] = (uint32_t)s;
  ra[s].idx = gmap ? (uint32_t)(gmap[io] - 1) : io;
}

template <class T, class TI>
__device__ __noinline__ void fillrows_generic_row(const FillRowsArgs<T, TI>* ad, long long s) {
  generic_atom<T, TI, MODE_FILL>(s, ad->rec, ad->co, ad->g, ad->out);
}

template <class T, class TI>
__device__ __noinline__ void fillrows_slow_SR(const FillRowsArgs<T, TI>* ad, T xi, T yi, T zi, T xj, T yj, T zj, uint32_t wi, uint32_t wj, int* S012,
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

template <class T> __device__ __forceinline__ void load_rec(const RecAoS<T>* p, T& x, T& y, T& z, uint32_t& idx, uint32_t& w);
template <> __device__ __forceinline__ void load_rec<double>(const RecAoS<double>* p, double& x, double& y, double& z, uint32_t& idx, uint32_t& w) {
  const uint4 lo = __ldg((const uint4*)p), hi = __ldg((const uint4*)p + 1);
  x = __hiloint2double((int)lo.y, (int)lo.x);
  y = __hiloint2double((int)lo.w, (int)lo.z);
  z = __hiloint2double((int)hi.y, (int)hi.x);
  idx = hi.z; w = hi.w;
}
template <> __device__ __forceinline__ void load_rec<float>(const RecAoS<float>* p, float& x, float& y, float& z, uint32_t& idx, uint32_t& w) {
  const uint4 lo = __ldg((const uint4*)p);
  x = __uint_as_float(lo.x); y = __uint_as_float(lo.y); z = __uint_as_float(lo.z);
  idx = lo.w;
  w = __ldg((const uint32_t*)p + 4);
}

// position of the (r+1)-th set bit of w (0 <= r < popc(w))
__device__ __forceinline__ int select_bit(uint32_t w, int r) {
  int pos = 0;
  int c = __popc(w & 0xffffu);
  if (r >= c) { r -= c; pos = 16; w >>= 16; }
  c = __popc(w & 0xffu);
  if (r >= c) { r -= c; pos += 8; w >>= 8; }
  c = __popc(w & 0xfu);
  if (r >= c) { r -= c; pos += 4; w >>= 4; }
  c = __popc(w & 0x3u);
  if (r >= c) { r -= c; pos += 2; w >>= 2; }
  if (r >= (int)(w & 1u)) pos += 1;
  return pos;
}

template <class T, class TI>
__global__ void __launch_bounds__(FR_NT, 3) k_fill_rows(const FillRowsArgs<T, TI> a) {
  __shared__ int s_first[FR_RB + 1];      // row starts relative to the block's first pair
  __shared__ int s_sorted[FR_RB];         // sorted index of each row's atom; -1: row written by the generic route
  __shared__ uint32_t s_key[FR_RB];
  __shared__ uint32_t s_mask[FR_RB][MASK_WORDS];
  __shared__ uint16_t s_mpre[FR_RB][MASK_WORDS];  // exclusive popcount prefix over the mask words
  __shared__ int s_cstart[FR_RB][27];     // sorted index of the first atom of each stencil cell
  __shared__ uint16_t s_fpre[FR_RB][28];  // flat candidate prefix over the stencil cells
  __shared__ uint8_t s_cshift[FR_RB][27]; // packed periodic shift of each stencil cell
  __shared__ T s_hx[FR_RB], s_hy[FR_RB], s_hz[FR_RB];
  __shared__ uint32_t s_hw[FR_RB], s_hid[FR_RB];
  __shared__ int stS[FR_NT * 3];
  __shared__ T stR[FR_NT * 3];
  __shared__ int s_anygen;  // some row of this block went through the generic route

  const Geo<T>& g = a.g;
  const int tid = threadIdx.x;
  const long long i0 = (long long)blockIdx.x * FR_RB;
  const int nrow = (int)min((long long)FR_RB, a.out.n_rows - i0);
  const long long p0 = (long long)a.out.first[i0] - 1;

  // ---- phase A: per-row tables
  if (tid == 0) s_anygen = 0;
  __syncthreads();
  if (tid <= nrow) s_first[tid] = (int)((long long)a.out.first[i0 + tid] - 1 - p0);
  if (tid < nrow) {
    const uint32_t s = a.sorted_of[i0 + tid];
    const uint32_t key = a.pkey[s];
    s_key[tid] = key;
    T x, y, z;
    uint32_t id, w;
    load_rec<T>(a.ra + s, x, y, z, id, w);
    s_hx[tid] = x; s_hy[tid] = y; s_hz[tid] = z; s_hw[tid] = w; s_hid[tid] = id;
    int srt = (int)s;
    if (!a.cellflag[key]) {  // no masks for this atom's cell: the generic route writes the whole row
      fillrows_generic_row<T, TI>(a.self, (long long)s);
      srt = -1;
      s_anygen = 1;
    }
    s_sorted[tid] = srt;
  }
  __syncthreads();
  for (int idx = tid; idx < nrow * MASK_WORDS; idx += FR_NT) {
    const int row = idx >> 3, k = idx & 7;
    const int srt = s_sorted[row];
    s_mask[row][k] = srt >= 0 ? a.masks[(long long)srt * MASK_WORDS + k] : 0u;
  }
  for (int idx = tid; idx < nrow * 27; idx += FR_NT) {
    const int row = idx / 27, c = idx - row * 27;
    const uint32_t key = s_key[row];
    const int hx = (int)(key % (uint32_t)g.nc[0]), hyz = (int)(key / (uint32_t)g.nc[0]);
    const int hy = hyz % g.nc[1], hz = hyz / g.nc[1];
    int cx, cy, cz, s0, s1, s2;
    bool ok = map_virtual(hx + c % 3 - 1, g.nc[0], g.pbc[0], cx, s0);
    ok = map_virtual(hy + (c / 3) % 3 - 1, g.nc[1], g.pbc[1], cy, s1) && ok;
    ok = map_virtual(hz + c / 9 - 1, g.nc[2], g.pbc[2], cz, s2) && ok;
    int st = 0, cn = 0;
    if (ok) {
      const long long cl = (long long)cx + (long long)g.nc[0] * ((long long)cy + (long long)g.nc[1] * cz);
      const long long c0 = (long long)a.co[cl], c1 = (long long)a.co[cl + 1];
      st = (int)(c0 - 1);
      cn = (int)(c1 - c0);
    }
    s_cstart[row][c] = st;
    s_fpre[row][c + 1] = (uint16_t)min(cn, 65535);
    s_cshift[row][c] = (uint8_t)pack_shift(s0, s1, s2);
  }
  __syncthreads();
  if (tid < nrow) {
    int run = 0;
    s_fpre[tid][0] = 0;
#pragma unroll 1
    for (int c = 1; c <= 27; c++) { run += s_fpre[tid][c]; s_fpre[tid][c] = (uint16_t)min(run, 65535); }
    run = 0;
#pragma unroll
    for (int k = 0; k < MASK_WORDS; k++) { s_mpre[tid][k] = (uint16_t)run; run += __popc(s_mask[tid][k]); }
  }
  __syncthreads();

  // ---- phase B: one thread per output pair of the block's span
  const int total = s_first[nrow];
  const bool want_R = a.out.Ro != nullptr;
  for (int pb = 0; pb < total; pb += FR_NT) {
    const int pl = pb + tid;
    const bool valid = pl < total;
    int S0 = 0, S1 = 0, S2 = 0;
    T R0 = 0, R1 = 0, R2 = 0;
    if (valid) {
      int lo = 0, hi = nrow;  // last row with s_first[row] <= pl
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_first[mid] <= pl) lo = mid; else hi = mid;
      }
      const int row = lo;
      if (s_sorted[row] >= 0) {
        const int rank = pl - s_first[row];
        int k = 0;
#pragma unroll
        for (int q = 1; q < MASK_WORDS; q++) k += (s_mpre[row][q] <= rank) ? 1 : 0;  // prefix is non-decreasing
        const int bit = select_bit(s_mask[row][k], rank - s_mpre[row][k]);
        const int f = 32 * k + bit;
        int c = 0;
        {
          int clo = 0, chi = 27;  // last cell with s_fpre[row][cell] <= f
          while (chi - clo > 1) {
            const int mid = (clo + chi) >> 1;
            if (s_fpre[row][mid] <= f) clo = mid; else chi = mid;
          }
          c = clo;
        }
        const long long gj = (long long)s_cstart[row][c] + (f - s_fpre[row][c]);
        T xj, yj, zj;
        uint32_t jid, wj;
        load_rec<T>(a.ra + gj, xj, yj, zj, jid, wj);
        const int shp = s_cshift[row][c];
        S0 = (shp & 3) - 1; S1 = ((shp >> 2) & 3) - 1; S2 = ((shp >> 4) & 3) - 1;
        const T xi = s_hx[row], yi = s_hy[row], zi = s_hz[row];
        const uint32_t wi = s_hw[row];
        if (wi == wj && !(wi & WIND_OVERFLOW)) {
          if (want_R) {
            T c0, c1, c2;
            mtv(g.cell, (T)S0, (T)S1, (T)S2, c0, c1, c2);
            R0 = add_rn(sub_rn(xj, xi), c0);
            R1 = add_rn(sub_rn(yj, yi), c1);
            R2 = add_rn(sub_rn(zj, zi), c2);
          }
        } else {
          int S3[3] = {S0, S1, S2};
          T R3[3];
          fillrows_slow_SR<T, TI>(a.self, xi, yi, zi, xj, yj, zj, wi, wj, S3, R3);
          S0 = S3[0]; S1 = S3[1]; S2 = S3[2];
          R0 = R3[0]; R1 = R3[1]; R2 = R3[2];
        }
        a.out.io[p0 + pl] = (TI)s_hid[row] + 1;
        a.out.jo[p0 + pl] = (TI)jid + 1;
      }
    }
    stS[3 * tid] = S0; stS[3 * tid + 1] = S1; stS[3 * tid + 2] = S2;
    if (want_R) { stR[3 * tid] = R0; stR[3 * tid + 1] = R1; stR[3 * tid + 2] = R2; }
    __syncthreads();
    // unit-stride stores of the transposed S / R of this chunk; pairs of generic rows are skipped (already written)
    const int nw = 3 * min(FR_NT, total - pb);
    TI* const So = a.out.So + 3 * (p0 + pb);
    T* const Ro = want_R ? a.out.Ro + 3 * (p0 + pb) : nullptr;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int w = m * FR_NT + tid;
      if (w < nw) {
        if (s_anygen) {  // rare: do not overwrite rows the generic route has already written
          const int pw = pb + w / 3;
          int lo = 0, hi = nrow;
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_first[mid] <= pw) lo = mid; else hi = mid;
          }
          if (s_sorted[lo] < 0) continue;
        }
        So[w] = (TI)stS[w];
        if (want_R) Ro[w] = stR[w];
      }
    }
    __syncthreads();
  }
}

}  // namespace nl
