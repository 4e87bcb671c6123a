// nl_scan_sort.cuh -- device-wide exclusive scan and a stable LSD radix sort of 32-bit keys with a
// 32-bit payload (the atom permutation).  Hand-written for sm_100a; no CUB/Thrust.
//
// Replaces AcceleratedKernels.sortperm! (src/cell_list.jl:711-718) and AcceleratedKernels.accumulate!
// (src/gpu_kernels.jl:231,279).  The sort is STABLE, so the permutation equals the CPU path's
// sortperm (src/cell_list.jl:706-708) bit for bit.
#pragma once
#include "nl_common.cuh"

namespace nl {

// ------------------------------------------------------------------ block helpers
template <class A> __device__ __forceinline__ A warp_incl_scan(A v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    A t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Exclusive scan of one value per thread across a block of NT threads (NT multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the block sum.  `sm` needs 33 entries of A.
template <class A, int NT> __device__ __forceinline__ A block_excl_scan(A v, A* sm, A* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  A inc = warp_incl_scan(v, lane);
  if (lane == 31) sm[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    A w = lane < NT / 32 ? sm[lane] : A(0);
    A winc = warp_incl_scan(w, lane);
    sm[lane] = winc - w;
    if (lane == 31) sm[32] = winc;
  }
  __syncthreads();
  A r = sm[wid] + inc - v;
  *total = sm[32];
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------ device-wide exclusive scan
// out[i] = base + sum_{k<i} in[k] for i in [0, n]; out has n+1 entries when write_total, and
// *total_out (if non-null) receives the grand total (without base).  Three launches:
// tile sums -> scan of tile sums (one block) -> tile scans.  in/out may not alias.
constexpr int SCAN_NT = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_NT * SCAN_ITEMS;

template <class In, class A> __global__ void __launch_bounds__(SCAN_NT) k_scan_tile_sums(const In* __restrict__ in, long long n, A* __restrict__ tsum) {
  __shared__ A sm[33];
  long long base = (long long)blockIdx.x * SCAN_TILE;
  A v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    long long i = base + k * SCAN_NT + threadIdx.x;
    if (i < n) v += (A)in[i];
  }
  A tot;
  block_excl_scan<A, SCAN_NT>(v, sm, &tot);
  if (threadIdx.x == 0) tsum[blockIdx.x] = tot;
}

template <class A> __global__ void __launch_bounds__(1024) k_scan_tsums(A* __restrict__ tsum, long long nt, A* __restrict__ total_out) {
  __shared__ A sm[33];
  A carry = 0;
  for (long long b = 0; b < nt; b += 1024) {
    long long i = b + threadIdx.x;
    A v = i < nt ? tsum[i] : A(0);
    A tot;
    A ex = block_excl_scan<A, 1024>(v, sm, &tot);
    if (i < nt) tsum[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <class In, class A, class Out>
__global__ void __launch_bounds__(SCAN_NT) k_scan_apply(const In* __restrict__ in, long long n, const A* __restrict__ tsum, A base,
                                                        Out* __restrict__ out, int write_total) {
  __shared__ A sm[33];
  long long t0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  A v[SCAN_ITEMS];
  A s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    long long i = t0 + k;
    v[k] = i < n ? (A)in[i] : A(0);
    s += v[k];
  }
  A tot;
  A ex = block_excl_scan<A, SCAN_NT>(s, sm, &tot) + tsum[blockIdx.x] + base;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    long long i = t0 + k;
    if (i < n) out[i] = (Out)ex;
    ex += v[k];
    if (write_total && i == n - 1) out[n] = (Out)ex;
  }
}

inline long long scan_tiles(long long n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

template <class In, class A, class Out>
inline void exclusive_scan(const In* in, long long n, Out* out, A base, bool write_total, A* tsum_scratch, A* total_out, cudaStream_t st) {
  if (n <= 0) return;
  long long nt = scan_tiles(n);
  k_scan_tile_sums<In, A><<<(unsigned)nt, SCAN_NT, 0, st>>>(in, n, tsum_scratch);
  k_scan_tsums<A><<<1, 1024, 0, st>>>(tsum_scratch, nt, total_out);
  k_scan_apply<In, A, Out><<<(unsigned)nt, SCAN_NT, 0, st>>>(in, n, tsum_scratch, base, out, write_total ? 1 : 0);
  note_launch(3);
}

// ------------------------------------------------------------------ LSD radix sort, 8-bit digits
constexpr int RS_NT = 256;                  // threads per block (8 warps)
constexpr int RS_ITEMS = 16;                // keys per thread
constexpr int RS_TILE = RS_NT * RS_ITEMS;   // 4096 keys per block
constexpr int RS_WSEG = RS_TILE / (RS_NT / 32);  // 512 consecutive keys per warp

inline long long rs_tiles(long long n) { return (n + RS_TILE - 1) / RS_TILE; }

// Per-tile digit histogram, written digit-major: hist[d * ntiles + tile].
__global__ void __launch_bounds__(RS_NT) k_rs_hist(const uint32_t* __restrict__ keys, long long n, int shift, uint32_t* __restrict__ hist, long long ntiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    long long i = base + k * RS_NT + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(long long)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter of one tile.  `offs` is the exclusive scan of `hist` (same layout).
// vals_in == nullptr means the payload is the key's own index (first pass).
__global__ void __launch_bounds__(RS_NT) k_rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                      uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, long long n,
                                                      int shift, const uint32_t* __restrict__ offs, long long ntiles) {
  __shared__ uint32_t wh[RS_NT / 32][256];  // per-warp digit counts, later per-warp exclusive bases
  __shared__ uint32_t bstart[256];          // tile-local start of each digit
  __shared__ uint32_t gbase[256];           // global start of this tile's run of each digit
  __shared__ uint32_t skey[RS_TILE];
  __shared__ uint32_t sval[RS_TILE];
  __shared__ uint32_t scan_sm[33];

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long tile0 = (long long)blockIdx.x * RS_TILE;
  const unsigned lt = (1u << lane) - 1u;

#pragma unroll
  for (int w = 0; w < RS_NT / 32; w++) wh[w][threadIdx.x] = 0;
  __syncthreads();

  uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    long long i = tile0 + (long long)wid * RS_WSEG + r * 32 + lane;
    bool ok = i < n;
    key[r] = ok ? keys_in[i] : 0u;
    val[r] = ok ? (vals_in ? vals_in[i] : (uint32_t)i) : 0u;
    uint32_t d = ok ? ((key[r] >> shift) & 255u) : 256u;  // 256 = "no key": its own match group
    unsigned m = __match_any_sync(FULL, d);
    int leader = __ffs(m) - 1;
    uint32_t old = 0;
    if (lane == leader && ok) {
      old = wh[wid][d];
      wh[wid][d] = old + __popc(m);
    }
    old = __shfl_sync(FULL, old, leader);
    rank[r] = old + __popc(m & lt);
    __syncwarp();
  }
  __syncthreads();

  // thread d: turn the per-warp counts of digit d into per-warp exclusive bases; tile count of d
  uint32_t cnt = 0;
  {
    const int d = threadIdx.x;
#pragma unroll
    for (int w = 0; w < RS_NT / 32; w++) {
      uint32_t c = wh[w][d];
      wh[w][d] = cnt;
      cnt += c;
    }
    gbase[d] = offs[(long long)d * ntiles + blockIdx.x];
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan<uint32_t, RS_NT>(cnt, scan_sm, &tot);
  bstart[threadIdx.x] = ex;
  __syncthreads();

#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    long long i = tile0 + (long long)wid * RS_WSEG + r * 32 + lane;
    if (i < n) {
      uint32_t d = (key[r] >> shift) & 255u;
      uint32_t p = bstart[d] + wh[wid][d] + rank[r];
      skey[p] = key[r];
      sval[p] = val[r];
    }
  }
  __syncthreads();

  const uint32_t tile_n = (uint32_t)min((long long)RS_TILE, n - tile0);
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    uint32_t p = k * RS_NT + threadIdx.x;
    if (p < tile_n) {
      uint32_t kk = skey[p];
      uint32_t d = (kk >> shift) & 255u;
      uint32_t dst = gbase[d] + (p - bstart[d]);
      keys_out[dst] = kk;
      vals_out[dst] = sval[p];
    }
  }
}

// Bytes of scratch for radix_sort_pairs: digit-major tile histograms + their scan + tile sums.
inline size_t rs_scratch_bytes(long long n) {
  long long nt = rs_tiles(n);
  size_t h = (size_t)256 * nt * sizeof(uint32_t);
  size_t ts = (size_t)(scan_tiles(256 * nt) + 1) * sizeof(uint32_t);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return al(h) * 2 + al(ts);
}

// Sorts (keyA, implicit iota) by the low `bits` bits; ping-pongs between A and B buffers.
// Returns 0 if the result is in the A buffers, 1 if in B.
inline int radix_sort_pairs(uint32_t* keyA, uint32_t* valA, uint32_t* keyB, uint32_t* valB, long long n, int bits, void* scratch,
                            cudaStream_t st) {
  long long nt = rs_tiles(n);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  char* p = (char*)scratch;
  uint32_t* hist = (uint32_t*)p;
  p += al((size_t)256 * nt * sizeof(uint32_t));
  uint32_t* offs = (uint32_t*)p;
  p += al((size_t)256 * nt * sizeof(uint32_t));
  uint32_t* tsum = (uint32_t*)p;
  int passes = (bits + 7) / 8;
  if (passes < 1) passes = 1;
  int cur = 0;
  for (int ps = 0; ps < passes; ps++) {
    const uint32_t* kin = cur ? keyB : keyA;
    const uint32_t* vin = ps == 0 ? nullptr : (cur ? valB : valA);
    uint32_t* kout = cur ? keyA : keyB;
    uint32_t* vout = cur ? valA : valB;
    k_rs_hist<<<(unsigned)nt, RS_NT, 0, st>>>(kin, n, ps * 8, hist, nt);
    exclusive_scan<uint32_t, uint32_t, uint32_t>(hist, 256 * nt, offs, 0u, false, tsum, nullptr, st);
    k_rs_scatter<<<(unsigned)nt, RS_NT, 0, st>>>(kin, vin, kout, vout, n, ps * 8, offs, nt);
    note_launch(2);
    cur ^= 1;
  }
  return cur;
}

}  // namespace nl
