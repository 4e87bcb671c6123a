// Microbenchmark: cost of writing CSR rows (i,j: 4 B; S: 12 B; R: 24 B per entry) when rows are visited in
// sequential order vs in a random permutation.  Pure stores, no reads besides first[] / perm[].
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int POL> __device__ __forceinline__ void st(int* p, int v) { if (POL == 0) *p = v; else if (POL == 1) __stcs(p, v); else if (POL == 2) __stwt(p, v); else __stcg(p, v); }
template <int POL> __device__ __forceinline__ void st(double* p, double v) { if (POL == 0) *p = v; else if (POL == 1) __stcs(p, v); else if (POL == 2) __stwt(p, v); else __stcg(p, v); }

template <bool WITH_R, int POL>
__global__ void write_rows_pol(const int* __restrict__ first, const int* __restrict__ order, int n, int* io, int* jo, int* So, double* Ro) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += nw) {
    const int row = order ? order[w] : (int)w;
    const long long b = first[row], e = first[row + 1];
    const int cnt = (int)(e - b);
    for (int r = lane; r < cnt; r += 32) { st<POL>(io + b + r, row); st<POL>(jo + b + r, r); }
    for (int x = lane; x < 3 * cnt; x += 32) { st<POL>(So + 3 * b + x, x); if (WITH_R) st<POL>(Ro + 3 * b + x, (double)x); }
  }
}

template <bool WITH_R>
__global__ void write_rows(const int* __restrict__ first, const int* __restrict__ order, int n, int* io, int* jo, int* So, double* Ro) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += nw) {
    const int row = order ? order[w] : (int)w;
    const long long b = first[row], e = first[row + 1];
    const int cnt = (int)(e - b);
    for (int r = lane; r < cnt; r += 32) { io[b + r] = row; jo[b + r] = r; }
    for (int x = lane; x < 3 * cnt; x += 32) { So[3 * b + x] = x; if (WITH_R) Ro[3 * b + x] = (double)x; }
  }
}

// Variant A (ALIGNED): every row padded to a multiple of 8 entries, so each stream segment starts and ends on a 32-byte sector.
// Variant B (INTERIOR): the real, unaligned rows, but only the sectors a row covers COMPLETELY are written (the partial head /
// tail sectors, shared with the neighbouring rows, are skipped).  Both isolate the cost of partial-sector writes.
template <int VARIANT>
__global__ void write_rows_sect(const int* __restrict__ first, const int* __restrict__ order, int n, int* io, int* jo, int* So, double* Ro) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n; w += nw) {
    const int row = order ? order[w] : (int)w;
    long long b = first[row], e = first[row + 1];
    if (VARIANT == 0) {  // padded rows: first[] was built with multiples of 8
      const int cnt = (int)(e - b);
      for (int r = lane; r < cnt; r += 32) { io[b + r] = row; jo[b + r] = r; }
      for (int x = lane; x < 3 * cnt; x += 32) { So[3 * b + x] = x; Ro[3 * b + x] = (double)x; }
    } else {
      // element ranges rounded inwards to sector boundaries, per stream (4-byte elements: 8 per sector; doubles: 4 per sector)
      { const long long lo = (b + 7) & ~7ll, hi = e & ~7ll; for (long long r = lo + lane; r < hi; r += 32) { io[r] = row; jo[r] = (int)r; } }
      { const long long lo = (3 * b + 7) & ~7ll, hi = (3 * e) & ~7ll; for (long long x = lo + lane; x < hi; x += 32) So[x] = (int)x; }
      { const long long lo = (3 * b + 3) & ~3ll, hi = (3 * e) & ~3ll; for (long long x = lo + lane; x < hi; x += 32) Ro[x] = (double)x; }
    }
  }
}

int main() {
  const int n = 10000000;
  std::mt19937 rng(1);
  std::poisson_distribution<int> pd(26.18);
  std::vector<int> first(n + 1), perm(n);
  first[0] = 0;
  for (int i = 0; i < n; i++) first[i + 1] = first[i] + pd(rng);
  for (int i = 0; i < n; i++) perm[i] = i;
  std::shuffle(perm.begin(), perm.end(), rng);
  const long long P = first[n];
  int *d_first, *d_perm, *io, *jo, *So; double* Ro;
  CK(cudaMalloc(&d_first, (n + 1) * 4)); CK(cudaMalloc(&d_perm, n * 4));
  CK(cudaMalloc(&io, P * 4)); CK(cudaMalloc(&jo, P * 4)); CK(cudaMalloc(&So, P * 12)); CK(cudaMalloc(&Ro, P * 24));
  CK(cudaMemcpy(d_first, first.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_perm, perm.data(), n * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int with_r = 1; with_r >= 0; with_r--)
    for (int mode = 0; mode < 2; mode++) {
      float best = 1e9;
      for (int it = 0; it < 5; it++) {
        cudaEventRecord(e0);
        if (with_r) write_rows<true><<<148 * 8, 256>>>(d_first, mode ? d_perm : nullptr, n, io, jo, So, Ro);
        else write_rows<false><<<148 * 8, 256>>>(d_first, mode ? d_perm : nullptr, n, io, jo, So, Ro);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
      }
      double bytes = (double)P * (with_r ? 44 : 20);
      printf("rows %s, %s: %.3f ms  (%.0f GB/s of payload, P=%lld)\n", mode ? "RANDOM order" : "sequential", with_r ? "i,j,S,R" : "i,j,S", best,
             bytes / best / 1e6, P);
    }
  for (int pol = 1; pol <= 3; pol++) {
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
      cudaEventRecord(e0);
      if (pol == 1) write_rows_pol<true, 1><<<148 * 8, 256>>>(d_first, d_perm, n, io, jo, So, Ro);
      if (pol == 2) write_rows_pol<true, 2><<<148 * 8, 256>>>(d_first, d_perm, n, io, jo, So, Ro);
      if (pol == 3) write_rows_pol<true, 3><<<148 * 8, 256>>>(d_first, d_perm, n, io, jo, So, Ro);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
    printf("rows RANDOM order, i,j,S,R, store policy %s: %.3f ms (%.0f GB/s)\n", pol == 1 ? "st.cs (streaming)" : pol == 2 ? "st.wt (write-through)" : "st.cg", best, (double)P * 44 / best / 1e6);
  }
  // more warps in flight: 148*32 blocks
  {
    float best = 1e9;
    for (int it = 0; it < 5; it++) { cudaEventRecord(e0); write_rows<true><<<148 * 32, 256>>>(d_first, d_perm, n, io, jo, So, Ro); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
    printf("rows RANDOM order, i,j,S,R, grid 148*32: %.3f ms\n", best);
  }
  // reference: plain streaming memset-like write of the same volume
  {
    float best = 1e9;
    for (int it = 0; it < 5; it++) { cudaEventRecord(e0); cudaMemsetAsync(Ro, 1, P * 24); cudaMemsetAsync(So, 1, P * 12); cudaMemsetAsync(io, 1, P * 4); cudaMemsetAsync(jo, 1, P * 4);
      cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
    printf("cudaMemset of the same 44 B/pair: %.3f ms (%.0f GB/s)\n", best, (double)P * 44 / best / 1e6);
  }
  // partial-sector experiments
  {
    std::vector<int> first8(n + 1);
    first8[0] = 0;
    for (int i = 0; i < n; i++) first8[i + 1] = first8[i] + ((first[i + 1] - first[i] + 7) / 8) * 8;
    const long long P8 = first8[n];
    int *d_first8, *io8, *jo8, *So8; double* Ro8;
    CK(cudaMalloc(&d_first8, (n + 1) * 4)); CK(cudaMemcpy(d_first8, first8.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&io8, P8 * 4)); CK(cudaMalloc(&jo8, P8 * 4)); CK(cudaMalloc(&So8, P8 * 12)); CK(cudaMalloc(&Ro8, P8 * 24));
    for (int mode = 0; mode < 2; mode++) {
      float best = 1e9;
      for (int it = 0; it < 5; it++) { cudaEventRecord(e0); write_rows_sect<0><<<148 * 8, 256>>>(d_first8, mode ? d_perm : nullptr, n, io8, jo8, So8, Ro8); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
      printf("ALIGNED rows (padded to 8 entries, P8=%lld), %s: %.3f ms (%.0f GB/s)\n", P8, mode ? "RANDOM order" : "sequential", best, (double)P8 * 44 / best / 1e6);
    }
    for (int mode = 0; mode < 2; mode++) {
      float best = 1e9;
      for (int it = 0; it < 5; it++) { cudaEventRecord(e0); write_rows_sect<1><<<148 * 8, 256>>>(d_first, mode ? d_perm : nullptr, n, io, jo, So, Ro); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); }
      printf("INTERIOR sectors only (unaligned rows, partial head/tail sectors skipped), %s: %.3f ms\n", mode ? "RANDOM order" : "sequential", best);
    }
  }
  return 0;
}