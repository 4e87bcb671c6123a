// experiments/microbench_rows2.cu -- store-pattern study for the CSR fill pass (round 2).
//
// Question: what does it cost to write 10 M CSR rows (Poisson(26.18) entries; j: 4 B, S: 12 B, R: 24 B per entry;
// i is a pure function of first[] and gets its own streaming kernel) when the rows are visited in sequential order
// and in a random permutation, for different ways of issuing the stores?
//
//   STYLE 0  scalar   : lane-per-element 4/8-byte stores straight from registers (round-1 pattern)
//   STYLE 1  vec16    : row staged in shared memory with the destination's 16-byte phase, LDS.128 + STG.128
//   STYLE 2  bulk     : same staging, one cp.async.bulk (TMA, UBLKCP.G.S) per stream segment
//   STYLE 3  bulkmask : bulk for the interior and cp.async.bulk...cp_mask (16 B, byte mask) for the head / tail pieces
//   ALIGN 0  everything is written in place (interior at 16-byte granularity, head / tail pieces separately)
//   ALIGN 1  only the 32-byte sectors a row covers COMPLETELY are written (lower bound: no partial sectors at all)
//   ALIGN 2  complete sectors in place + the head / tail elements PARKED in a per-row record (128 B + 64 B, always
//            written whole); k_fixup then assembles every boundary sector from the parked pieces of the rows that
//            share it and writes it once, in original row order
//
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_rows2 microbench_rows2.cu
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAXC = 64;                                    // entries per staged piece
constexpr int BUF_J = 16 + 4 * MAXC, BUF_S = 16 + 12 * MAXC, BUF_R = 16 + 24 * MAXC;
constexpr int BUF_BYTES = BUF_J + BUF_S + BUF_R;           // 2608
constexpr int NBUF = 2;
constexpr int WARPS = 8;

__host__ __device__ inline int vj(long long p) { return (int)((unsigned long long)p * 2654435761ull >> 7); }
__host__ __device__ inline int vS(long long x) { return (int)(x ^ 0x5bd1e995ll); }
__host__ __device__ inline double vR(long long x) { return (double)x * 0.5; }

struct Out { int* io; int* jo; int* So; double* Ro; uint32_t* parkA; double* parkR; };

__device__ __forceinline__ void bulk_g2s_store(void* dst, const void* src_smem, int bytes) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(src_smem);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_masked16(void* dst, const void* src_smem, unsigned mask) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(src_smem);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;" ::"l"(dst), "r"(s), "h"((unsigned short)mask) : "memory");
}

// One stream segment.  gbase: array base (256-byte aligned), B0 / B1: byte range of the piece inside the array,
// sm: staged bytes; sm[(B0 & 15) + (B - B0)] holds byte B.  ES = element size.
template <int STYLE, int ALIGN, int ES>
__device__ __forceinline__ void put_stream(char* gbase, long long B0, long long B1, const char* sm, int lane, bool head_here, bool tail_here) {
  constexpr int G = ALIGN == 0 ? 16 : 32;
  const int ph = (int)(B0 & 15);
  long long I0 = (B0 + G - 1) & ~(long long)(G - 1), I1 = B1 & ~(long long)(G - 1);
  // pieces cut at multiples of 8 entries are sector aligned: no head / tail there
  const long long he = head_here ? (I0 < B1 ? I0 : B1) : B0;        // head = [B0, he)
  long long ts = tail_here ? (I1 > he ? I1 : he) : B1;              // tail = [ts, B1)
  if (!head_here) I0 = B0;
  if (!tail_here) I1 = B1;
  if (I0 < I1) {
    const int nb = (int)(I1 - I0);
    const char* s0 = sm + ph + (I0 - B0);
    if (STYLE == 1) {
      for (int o = lane * 16; o < nb; o += 512) *(int4*)(gbase + I0 + o) = *(const int4*)(s0 + o);
    } else if (STYLE >= 2) {
      if (lane == 0) bulk_g2s_store(gbase + I0, s0, nb);
    }
  }
  if (ALIGN == 0) {
    if (STYLE == 3) {
      if (lane == 0) {
        if (he > B0) {  // head chunk: the 16 bytes ending at roundup16(B0); may end early when the whole piece is inside it
          const long long c0 = B0 & ~15ll;
          const int lo = (int)(B0 - c0), hi = (int)(he - c0);
          bulk_store_masked16(gbase + c0, sm, ((1u << hi) - 1u) & ~((1u << lo) - 1u));
        }
        if (B1 > ts) {
          const long long c0 = ts & ~15ll;  // ts is 16-aligned unless it equals he
          const int lo = (int)(ts - c0), hi = (int)(B1 - c0);
          bulk_store_masked16(gbase + c0, sm + ph + (c0 - B0), ((1u << hi) - 1u) & ~((1u << lo) - 1u));
        }
      }
    } else {
      const int hn = (int)(he - B0) / ES, tn = (int)(B1 - ts) / ES;
      if (ES == 4) {
        if (lane < hn) *(int*)(gbase + B0 + 4 * lane) = *(const int*)(sm + ph + 4 * lane);
        if (lane >= 8 && lane - 8 < tn) *(int*)(gbase + ts + 4 * (lane - 8)) = *(const int*)(sm + ph + (ts - B0) + 4 * (lane - 8));
      } else {
        if (lane < hn) *(double*)(gbase + B0 + 8 * lane) = *(const double*)(sm + ph + 8 * lane);
        if (lane >= 8 && lane - 8 < tn) *(double*)(gbase + ts + 8 * (lane - 8)) = *(const double*)(sm + ph + (ts - B0) + 8 * (lane - 8));
      }
    }
  }
}

// Parked pieces of one row, addressed by POSITION INSIDE THE SECTOR: word k of a head / tail block is the element at
// byte (sector base + k * ES).  A record: [j head 8 | j tail 8 | S head 8 | S tail 8] words; R record: [head 4 | tail 4] doubles.
template <int ES, class V>
__device__ __forceinline__ V park_pick(long long B0, long long B1, const char* sm, int k, bool tail, bool head_here, bool tail_here) {
  const long long I0 = (B0 + 31) & ~31ll, I1 = B1 & ~31ll;
  const long long he = I0 < B1 ? I0 : B1;
  const long long ts = I1 > he ? I1 : he;
  long long a;
  bool ok;
  if (!tail) { a = (B0 & ~31ll) + (long long)k * ES; ok = head_here && a >= B0 && a < he; }
  else { a = (ts & ~31ll) + (long long)k * ES; ok = tail_here && a >= ts && a < B1; }
  V v = V();
  if (ok) v = *(const V*)(sm + (B0 & 15) + (a - B0));
  return v;
}

template <int STYLE, int ALIGN>
__global__ void __launch_bounds__(WARPS * 32) k_rows(const int* __restrict__ first, const int* __restrict__ order, int n, Out o) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  char* wbuf = (char*)smem + wid * (NBUF * BUF_BYTES);
  int bufsel = 0;
  const long long nw = (long long)gridDim.x * WARPS;
  for (long long w = (long long)blockIdx.x * WARPS + wid; w < n; w += nw) {
    const int row = order ? order[w] : (int)w;
    const long long b = first[row], e = first[row + 1];
    if (STYLE == 0) {
      const int cnt = (int)(e - b);
      for (int r = lane; r < cnt; r += 32) o.jo[b + r] = vj(b + r);
      for (int x = lane; x < 3 * cnt; x += 32) { o.So[3 * b + x] = vS(3 * b + x); o.Ro[3 * b + x] = vR(3 * b + x); }
      continue;
    }
    for (long long p0 = b; p0 < e;) {
      long long p1 = p0 + MAXC < e ? ((p0 + MAXC) & ~7ll) : e;   // cut at multiples of 8 entries (sector aligned in all streams)
      const int cnt = (int)(p1 - p0);
      char* buf = wbuf + bufsel * BUF_BYTES;
      bufsel ^= 1;
      if (STYLE >= 2) {  // the bulk copies that read this buffer two pieces ago must have finished READING it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      }
      char* bj = buf;
      char* bS = buf + BUF_J;
      char* bR = buf + BUF_J + BUF_S;
      const int phj = (int)((p0 * 4) & 15), phS = (int)((p0 * 12) & 15), phR = (int)((p0 * 24) & 15);
      for (int r = lane; r < cnt; r += 32) *(int*)(bj + phj + 4 * r) = vj(p0 + r);
      for (int x = lane; x < 3 * cnt; x += 32) {
        *(int*)(bS + phS + 4 * x) = vS(3 * p0 + x);
        *(double*)(bR + phR + 8 * x) = vR(3 * p0 + x);
      }
      if (STYLE >= 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      const bool hh = p0 == b, th = p1 == e;
      put_stream<STYLE, ALIGN, 4>((char*)o.jo, p0 * 4, p1 * 4, bj, lane, hh, th);
      put_stream<STYLE, ALIGN, 4>((char*)o.So, p0 * 12, p1 * 12, bS, lane, hh, th);
      put_stream<STYLE, ALIGN, 8>((char*)o.Ro, p0 * 24, p1 * 24, bR, lane, hh, th);
      if (ALIGN == 2 && (hh || th)) {
        const int k = lane & 7, blk = lane >> 3;  // blk: 0 j head, 1 j tail, 2 S head, 3 S tail
        const uint32_t v = blk < 2 ? park_pick<4, uint32_t>(p0 * 4, p1 * 4, bj, k, blk & 1, hh, th) : park_pick<4, uint32_t>(p0 * 12, p1 * 12, bS, k, blk & 1, hh, th);
        const bool mine = (blk & 1) ? th : hh;
        if (hh && th) o.parkA[(long long)row * 32 + lane] = v;               // whole record: one 128-byte line
        else if (mine) o.parkA[(long long)row * 32 + lane] = v;
        if (lane < 8) {
          const double r = park_pick<8, double>(p0 * 24, p1 * 24, bR, lane & 3, lane >> 2, hh, th);
          const bool mr = (lane >> 2) ? th : hh;
          if (mr) o.parkR[(long long)row * 8 + lane] = r;
        }
      }
      if (STYLE >= 2) { if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
      else __syncwarp();
      p0 = p1;
    }
  }
  if (STYLE >= 2) { if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
}


// STYLE 0 with the partial sectors skipped (MODE 1) or their elements written ELEMENT-WISE into the park record (MODE 2: the
// park blocks are then only partially written themselves), lean instruction stream: isolates the memory-side cost.
template <int MODE>
__global__ void __launch_bounds__(256) k_rows_scalar_sect(const int* __restrict__ first, const int* __restrict__ order, int n, Out o) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  for (long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); w < n; w += nw) {
    const int row = order ? order[w] : (int)w;
    const long long b = first[row], e = first[row + 1];
    const int cnt = (int)(e - b);
    {  // j: 8 elements per sector
      const long long lo = (b + 7) & ~7ll, hi = e & ~7ll;
      for (int r = lane; r < cnt; r += 32) {
        const long long p = b + r;
        if (p >= lo && p < hi) o.jo[p] = vj(p);
        else if (MODE == 2) o.parkA[(long long)row * 32 + (p < lo ? 0 : 8) + (p & 7)] = (uint32_t)vj(p);
      }
    }
    {
      const long long lo = (3 * b + 7) & ~7ll, hi = (3 * e) & ~7ll, lr = (3 * b + 3) & ~3ll, hr = (3 * e) & ~3ll;
      for (int x = lane; x < 3 * cnt; x += 32) {
        const long long q = 3 * b + x;
        if (q >= lo && q < hi) o.So[q] = vS(q);
        else if (MODE == 2) o.parkA[(long long)row * 32 + 16 + (q < lo ? 0 : 8) + (q & 7)] = (uint32_t)vS(q);
        if (q >= lr && q < hr) o.Ro[q] = vR(q);
        else if (MODE == 2) o.parkR[(long long)row * 8 + (q < lr ? 0 : 4) + (q & 3)] = vR(q);
      }
    }
  }
}


// Lane = PAIR, strided element stores (no transposition): j 1 x STG.32, S 3 x STG.32 (stride 12 B), R 3 x STG.64 (stride 24 B).
// MODE 0: everything in place; MODE 1: complete sectors only.  What a flat (pair-parallel) fill without staging would issue.
template <int MODE>
__global__ void __launch_bounds__(256) k_rows_strided(const int* __restrict__ first, const int* __restrict__ order, int n, Out o) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  for (long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); w < n; w += nw) {
    const int row = order ? order[w] : (int)w;
    const long long b = first[row], e = first[row + 1];
    const int cnt = (int)(e - b);
    const long long lo = (b + 7) & ~7ll, hi = e & ~7ll, ls = (3 * b + 7) & ~7ll, hs = (3 * e) & ~7ll, lr = (3 * b + 3) & ~3ll, hr = (3 * e) & ~3ll;
    for (int r = lane; r < cnt; r += 32) {
      const long long p = b + r;
      if (MODE == 0 || (p >= lo && p < hi)) o.jo[p] = vj(p);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const long long q = 3 * p + k;
        if (MODE == 0 || (q >= ls && q < hs)) o.So[q] = vS(q);
        if (MODE == 0 || (q >= lr && q < hr)) o.Ro[q] = vR(q);
      }
    }
  }
}


// STYLE 0 (everything in place) + explicit L2 prefetch of the row's partial boundary sectors before they are written:
// the partial write then lands on a sector that is fully valid in L2 and is written back whole -- the read-modify-write
// the memory system would do at eviction becomes an ordinary, pipelined read.  AHEAD = prefetch for the row of the NEXT
// iteration (more lead time) instead of the current one.
template <bool AHEAD, int HOW>
__global__ void __launch_bounds__(256) k_rows_prefetch(const int* __restrict__ first, const int* __restrict__ order, int n, Out o) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  auto pf = [&](long long b, long long e) {
    if (e <= b) return;
    // lane k < 6: stream (k >> 1) in {j, S, R}, end (k & 1)
    if (lane < 6) {
      const int st = lane >> 1;
      const long long B0 = st == 0 ? b * 4 : (st == 1 ? b * 12 : b * 24), B1 = st == 0 ? e * 4 : (st == 1 ? e * 12 : e * 24);
      const char* base = st == 0 ? (const char*)o.jo : (st == 1 ? (const char*)o.So : (const char*)o.Ro);
      const long long a = (lane & 1) ? ((B1 - 1) & ~31ll) : (B0 & ~31ll);
      const bool partial = (lane & 1) ? (B1 & 31) != 0 : (B0 & 31) != 0;
      if (partial) {
        if (HOW == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + a));
        else if (HOW == 1) { unsigned tmp; asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(tmp) : "l"(base + a)); }   // one 32-byte sector
        else if (HOW == 2) { unsigned tmp; asm volatile("ld.global.cg.L2::64B.b32 %0, [%1];" : "=r"(tmp) : "l"(base + a)); }
        else if (HOW == 3) { unsigned tmp; asm volatile("ld.global.cg.L2::128B.b32 %0, [%1];" : "=r"(tmp) : "l"(base + a)); }
        else if (HOW == 4) { asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + a)); }
        else if (HOW == 5) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], 32;" ::"l"(base + a)); }
        else if (HOW == 6) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], 64;" ::"l"(base + (a & ~63ll))); }
      }
    }
  };
  long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (AHEAD && w < n) { const int row = order ? order[w] : (int)w; pf(first[row], first[row + 1]); }
  for (; w < n; w += nw) {
    const int row = order ? order[w] : (int)w;
    const long long b = first[row], e = first[row + 1];
    if (AHEAD) { if (w + nw < n) { const int r2 = order ? order[w + nw] : (int)(w + nw); pf(first[r2], first[r2 + 1]); } }
    else pf(b, e);
    const int cnt = (int)(e - b);
    for (int r = lane; r < cnt; r += 32) o.jo[b + r] = vj(b + r);
    for (int x = lane; x < 3 * cnt; x += 32) { o.So[3 * b + x] = vS(3 * b + x); o.Ro[3 * b + x] = vR(3 * b + x); }
  }
}

// Fix-up: thread i owns row boundary i (between rows i-1 and i; i = n is the end of the array) in all three streams.
// A boundary strictly inside a sector makes that sector partial for both rows; the thread of the FIRST boundary inside
// the sector assembles all 32 bytes from the parked pieces of every row that touches it and writes the sector once.
template <int ES, int NE, class V>
__device__ __forceinline__ void fix_stream(const int* __restrict__ first, int n, int i, char* gbase, const char* park, int rec_bytes, int head_off, int tail_off) {
  const long long EB = (long long)ES * NE;
  const long long Bp = (long long)first[i] * EB;
  if ((Bp & 31) == 0) return;
  const long long s0 = Bp & ~31ll;
  if (i > 0 && (long long)first[i - 1] * EB > s0) return;   // an earlier boundary lies inside this sector: its thread does it
  constexpr int NS = 32 / ES;
  V vals[NS];
  const long long Pend = (long long)first[n] * EB;
  int r = i - 1;                       // row that holds the first byte of the sector (it started at or before s0)
  long long rB1 = Bp;                  // end of row r
  bool started_before = true;
#pragma unroll
  for (int k = 0; k < NS; k++) {
    const long long a = s0 + (long long)k * ES;
    while (a >= rB1 && r < n - 1) { r++; rB1 = (long long)first[r + 1] * EB; started_before = false; }
    V v = V();
    if (a < Pend && a < rB1) v = *(const V*)(park + (long long)r * rec_bytes + (started_before ? tail_off : head_off) + k * ES);
    vals[k] = v;
  }
  if (s0 + 32 <= Pend) {
    int4* d = (int4*)(gbase + s0);
    d[0] = *(int4*)&vals[0];
    d[1] = *(int4*)&vals[16 / ES];
  } else {
    for (int k = 0; k < NS; k++) if (s0 + (long long)k * ES < Pend) *(V*)(gbase + s0 + k * ES) = vals[k];
  }
}
__global__ void __launch_bounds__(256) k_fixup(const int* __restrict__ first, int n, Out o) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 1 || t > n) return;
  const int i = (int)t;
  fix_stream<4, 1, uint32_t>(first, n, i, (char*)o.jo, (const char*)o.parkA, 128, 0, 32);
  fix_stream<4, 3, uint32_t>(first, n, i, (char*)o.So, (const char*)o.parkA, 128, 64, 96);
  fix_stream<8, 3, double>(first, n, i, (char*)o.Ro, (const char*)o.parkR, 64, 0, 32);
}

// i stream: expand first[] into row ids; thread per 4 output elements, binary search in a shared-memory slice of first[].
constexpr int EXP_ELEMS = 4096;
__global__ void __launch_bounds__(256) k_expand_i(const int* __restrict__ first, int n, long long P, int* __restrict__ io) {
  __shared__ int sfirst[2048 + 1];
  __shared__ int s_r0, s_r1;
  const long long P0 = (long long)blockIdx.x * EXP_ELEMS, P1 = min(P, P0 + EXP_ELEMS);
  if (threadIdx.x < 2) {
    const long long target = threadIdx.x == 0 ? P0 : P1 - 1;  // row containing element target: last r with first[r] <= target
    int lo = 0, hi = n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (first[mid] <= target) lo = mid; else hi = mid; }
    if (threadIdx.x == 0) s_r0 = lo; else s_r1 = lo;
  }
  __syncthreads();
  const int r0 = s_r0, r1 = s_r1, nr = r1 - r0 + 1;
  const bool in_smem = nr <= 2048;
  if (in_smem) for (int k = threadIdx.x; k <= nr; k += 256) sfirst[k] = first[r0 + k];
  __syncthreads();
  for (long long q = P0 + 4 * threadIdx.x; q < P1; q += 1024) {
    int v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long t = q + u;
      int lo = 0, hi = nr;
      if (in_smem) { while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sfirst[mid] <= t) lo = mid; else hi = mid; } }
      else { while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (first[r0 + mid] <= t) lo = mid; else hi = mid; } }
      v[u] = r0 + lo + 1;
    }
    if (q + 4 <= P1) *(int4*)(io + q) = make_int4(v[0], v[1], v[2], v[3]);
    else for (int u = 0; u < 4 && q + u < P1; u++) io[q + u] = v[u];
  }
}

template <class F> float time_best(F f, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < reps; it++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  CK(cudaGetLastError());
  return best;
}

template <int STYLE, int ALIGN> void launch(const int* first, const int* order, int n, Out o, int blocks) {
  static bool once = false;
  const int sm = WARPS * NBUF * BUF_BYTES;
  if (!once) { CK(cudaFuncSetAttribute(k_rows<STYLE, ALIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)); once = true; }
  k_rows<STYLE, ALIGN><<<blocks, WARPS * 32, sm>>>(first, order, n, o);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 10000000;
  std::mt19937 rng(1);
  std::poisson_distribution<int> pd(26.18);
  std::vector<int> first(n + 1), perm(n);
  first[0] = 0;
  for (int i = 0; i < n; i++) first[i + 1] = first[i] + pd(rng);
  for (int i = 0; i < n; i++) perm[i] = i;
  std::shuffle(perm.begin(), perm.end(), rng);
  const long long P = first[n];
  int *d_first, *d_perm;
  Out o;
  CK(cudaMalloc(&d_first, (size_t)(n + 1) * 4)); CK(cudaMalloc(&d_perm, (size_t)n * 4));
  CK(cudaMalloc(&o.io, P * 4 + 64)); CK(cudaMalloc(&o.jo, P * 4 + 64)); CK(cudaMalloc(&o.So, P * 12 + 64)); CK(cudaMalloc(&o.Ro, P * 24 + 64));
  CK(cudaMalloc(&o.parkA, (size_t)n * 128)); CK(cudaMalloc(&o.parkR, (size_t)n * 64));
  CK(cudaMemcpy(d_first, first.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_perm, perm.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  const double bytes_jSR = (double)P * 40, bytes_all = (double)P * 44;
  printf("n = %d rows, P = %lld entries; j,S,R = %.2f GB, with i = %.2f GB\n", n, P, bytes_jSR / 1e9, bytes_all / 1e9);
  {
    float t = time_best([&] { cudaMemsetAsync(o.Ro, 1, P * 24); cudaMemsetAsync(o.So, 1, P * 12); cudaMemsetAsync(o.jo, 1, P * 4); });
    printf("%-64s %7.3f ms  %6.0f GB/s\n", "cudaMemset of j,S,R (40 B/entry)", t, bytes_jSR / t / 1e6);
    t = time_best([&] { k_expand_i<<<(unsigned)((P + EXP_ELEMS - 1) / EXP_ELEMS), 256>>>(d_first, n, P, o.io); });
    printf("%-64s %7.3f ms  %6.0f GB/s\n", "k_expand_i (i stream from first[])", t, (double)P * 4 / t / 1e6);
    t = time_best([&] { k_fixup<<<(n + 256) / 256, 256>>>(d_first, n, o); });
    printf("%-64s %7.3f ms\n", "k_fixup (boundary sectors from parked pieces)", t);
  }
  const char* sname[4] = {"scalar", "vec16", "bulk", "bulk+cp_mask"};
  const char* aname[3] = {"all in place", "complete sectors only", "complete sectors + park"};
  for (int blocks : {148 * 8}) {
    for (int mode = 0; mode < 2; mode++) {
      const int* ord = mode ? d_perm : nullptr;
#define RUN(S, A)                                                                                                 \
  {                                                                                                               \
    float t = time_best([&] { launch<S, A>(d_first, ord, n, o, blocks); });                                       \
    char nm[128];                                                                                                 \
    snprintf(nm, sizeof nm, "%s rows, %s, %s, %d CTAs", mode ? "RANDOM" : "sequential", sname[S], aname[A], blocks); \
    printf("%-64s %7.3f ms  %6.0f GB/s\n", nm, t, bytes_jSR / t / 1e6);                                           \
  }
      RUN(0, 0) RUN(1, 0) RUN(1, 1) RUN(1, 2) RUN(2, 0) RUN(2, 1) RUN(2, 2) RUN(3, 0)
    }
  }

  for (int mode = 0; mode < 2; mode++) {
    const int* ord = mode ? d_perm : nullptr;
    float t = time_best([&] { k_rows_scalar_sect<1><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar lean, complete sectors only" : "sequential rows, scalar lean, complete sectors only", t);
    t = time_best([&] { k_rows_scalar_sect<2><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar lean, complete sectors + element-wise park" : "sequential rows, scalar lean, complete sectors + element-wise park", t);
  }

  for (int mode = 0; mode < 2; mode++) {
    const int* ord = mode ? d_perm : nullptr;
    float t = time_best([&] { k_rows_strided<0><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, lane = pair, strided stores, all in place" : "sequential rows, lane = pair, strided stores, all in place", t);
    t = time_best([&] { k_rows_strided<1><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, lane = pair, strided stores, complete sectors only" : "sequential rows, lane = pair, strided stores, complete sectors only", t);
  }

  for (int mode = 0; mode < 2; mode++) {
    const int* ord = mode ? d_perm : nullptr;
    float t = time_best([&] { k_rows_prefetch<false, 0><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar in place + L2 prefetch of boundary sectors" : "sequential rows, scalar in place + L2 prefetch of boundary sectors", t);
    t = time_best([&] { k_rows_prefetch<true, 0><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar in place + L2 prefetch one row AHEAD" : "sequential rows, scalar in place + L2 prefetch one row AHEAD", t);
    t = time_best([&] { k_rows_prefetch<false, 1><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar in place + ld.cg of boundary sectors" : "sequential rows, scalar in place + ld.cg of boundary sectors", t);
    t = time_best([&] { k_rows_prefetch<false, 2><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar in place + ld.cg.L2::64B of boundary sectors" : "sequential rows, + ld.cg.L2::64B", t);
    t = time_best([&] { k_rows_prefetch<false, 3><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar in place + ld.cg.L2::128B of boundary sectors" : "sequential rows, + ld.cg.L2::128B", t);
    t = time_best([&] { k_rows_prefetch<false, 5><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar in place + cp.async.bulk.prefetch.L2 32 B" : "sequential rows, + bulk prefetch 32 B", t);
    t = time_best([&] { k_rows_prefetch<false, 6><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar in place + cp.async.bulk.prefetch.L2 64 B" : "sequential rows, + bulk prefetch 64 B", t);
    t = time_best([&] { k_rows_prefetch<false, 4><<<148 * 8, 256>>>(d_first, ord, n, o); });
    printf("%-64s %7.3f ms\n", mode ? "RANDOM rows, scalar in place + prefetch.L2::evict_last" : "sequential rows, + prefetch.L2::evict_last", t);
  }
  // correctness of the park + fix-up scheme (both store styles) and of the cp_mask variant
  for (int variant = 0; variant < 3; variant++) {
    CK(cudaMemset(o.io, 0xff, P * 4)); CK(cudaMemset(o.jo, 0xff, P * 4)); CK(cudaMemset(o.So, 0xff, P * 12)); CK(cudaMemset(o.Ro, 0xff, P * 24));
    if (variant == 0) launch<1, 2>(d_first, d_perm, n, o, 148 * 8);
    if (variant == 1) launch<2, 2>(d_first, d_perm, n, o, 148 * 8);
    if (variant == 2) launch<3, 0>(d_first, d_perm, n, o, 148 * 8);
    if (variant < 2) k_fixup<<<(n + 256) / 256, 256>>>(d_first, n, o);
    k_expand_i<<<(unsigned)((P + EXP_ELEMS - 1) / EXP_ELEMS), 256>>>(d_first, n, P, o.io);
    CK(cudaDeviceSynchronize());
    const long long chk = std::min<long long>(P, 3000000);
    long long bad = 0;
    for (int part = 0; part < 2; part++) {
      const long long q0 = part ? P - chk : 0;
      std::vector<int> hi(chk), hj(chk), hS(3 * chk); std::vector<double> hR(3 * chk);
      CK(cudaMemcpy(hi.data(), o.io + q0, chk * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hj.data(), o.jo + q0, chk * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hS.data(), o.So + 3 * q0, chk * 12, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hR.data(), o.Ro + 3 * q0, chk * 24, cudaMemcpyDeviceToHost));
      int row = (int)(std::upper_bound(first.begin(), first.end(), (int)q0) - first.begin()) - 1;
      for (long long q = 0; q < chk; q++) {
        while (first[row + 1] <= q0 + q) row++;
        if (hi[q] != row + 1 || hj[q] != vj(q0 + q)) bad++;
        for (int k = 0; k < 3; k++) if (hS[3 * q + k] != vS(3 * (q0 + q) + k) || hR[3 * q + k] != vR(3 * (q0 + q) + k)) bad++;
      }
    }
    printf("verify %s: %lld mismatches in 2 x %lld entries\n", variant == 0 ? "vec16 + park + fixup" : variant == 1 ? "bulk + park + fixup" : "bulk + cp_mask", bad, chk);
  }
  return 0;
}
