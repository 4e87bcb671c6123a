// experiments/microbench_hybrid.cu -- store-side floor of the "compact + emit" fill (round 2).
//
// Main pass (rows visited in SORTED-atom order, i.e. a random permutation of the original rows):
//   * (j, packed S) as ONE 8-byte record per pair into a compact array laid out in PROCESSING order (sequential writes)
//   * R (24 B per pair) in place, at the row's position in the original-order CSR (random 624-byte rows, transposed stores)
// Emit pass (original order, perfectly sequential writes): thread per 4 pairs, finds the row of each pair by binary search
//   in a shared-memory slice of first[], reads the compact records of the row and writes i, j, S with 16-byte stores.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_hybrid microbench_hybrid.cu
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__host__ __device__ inline int vj(long long p) { return (int)((unsigned long long)p * 2654435761ull >> 7); }
__host__ __device__ inline unsigned vcode(long long p) { return (unsigned)((512 + (p & 1) - ((p >> 1) & 1)) | (512u << 10) | ((512 + ((p >> 2) & 1)) << 20)); }

template <bool WITH_R, bool WITH_C>
__global__ void __launch_bounds__(256) k_main(const int* __restrict__ first, const int* __restrict__ order, const int* __restrict__ first2, int n,
                                              int2* __restrict__ compact, double* __restrict__ Ro) {
  __shared__ double stR[8][96];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long nw = (long long)gridDim.x * 8;
  for (long long w = (long long)blockIdx.x * 8 + wid; w < n; w += nw) {
    const int row = order ? order[w] : (int)w;
    const long long b = first[row], cb = first2[w];
    const int cnt = first[row + 1] - (int)b;
    for (int r0 = 0; r0 < cnt; r0 += 32) {
      const int r = r0 + lane, nr = min(32, cnt - r0);
      if (r < cnt) {
        const long long p = b + r;
        if (WITH_C) compact[cb + r] = make_int2(vj(p), (int)vcode(p));
        if (WITH_R) { stR[wid][3 * lane] = (double)p; stR[wid][3 * lane + 1] = (double)p + 0.25; stR[wid][3 * lane + 2] = (double)p + 0.5; }
      }
      if (WITH_R) {
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const int x = m * 32 + lane;
          if (x < 3 * nr) Ro[3 * (b + r0) + x] = stR[wid][x];
        }
        __syncwarp();
      }
    }
  }
}

constexpr int EM_PAIRS = 4096;  // pairs per block
constexpr int EM_ROWS = 1024;   // rows whose first[] / cstart[] slices fit in shared memory
__global__ void __launch_bounds__(256) k_emit(const int* __restrict__ first, const int* __restrict__ cstart, int n, long long P,
                                              const int2* __restrict__ compact, int* __restrict__ io, int* __restrict__ jo, int* __restrict__ So) {
  __shared__ int sfirst[EM_ROWS + 1];
  __shared__ int scs[EM_ROWS];
  __shared__ int s_r0, s_r1;
  const long long P0 = (long long)blockIdx.x * EM_PAIRS, P1 = min(P, P0 + EM_PAIRS);
  if (threadIdx.x < 2) {
    const long long target = threadIdx.x == 0 ? P0 : P1 - 1;
    int lo = 0, hi = n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (first[mid] <= target) lo = mid; else hi = mid; }
    if (threadIdx.x == 0) s_r0 = lo; else s_r1 = lo;
  }
  __syncthreads();
  const int r0 = s_r0, nr = s_r1 - r0 + 1;
  const bool in_smem = nr <= EM_ROWS;
  if (in_smem) {
    for (int k = threadIdx.x; k <= nr; k += 256) sfirst[k] = first[r0 + k];
    for (int k = threadIdx.x; k < nr; k += 256) scs[k] = cstart[r0 + k];
  }
  __syncthreads();
  for (long long q = P0 + 4 * threadIdx.x; q < P1; q += 1024) {
    int lo = 0, hi = nr;
    if (in_smem) { while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sfirst[mid] <= q) lo = mid; else hi = mid; } }
    else { while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (first[r0 + mid] <= q) lo = mid; else hi = mid; } }
    int vi[4], vjj[4], vs[12];
    int rend = in_smem ? sfirst[lo + 1] : first[r0 + lo + 1];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long t = q + u;
      while (t >= rend && lo < nr - 1) { lo++; rend = in_smem ? sfirst[lo + 1] : first[r0 + lo + 1]; }
      const int rs = in_smem ? sfirst[lo] : first[r0 + lo];
      const int cs = in_smem ? scs[lo] : cstart[r0 + lo];
      int2 c = make_int2(0, 0);
      if (t < P1) c = compact[(long long)cs + (t - rs)];
      vi[u] = r0 + lo + 1;
      vjj[u] = c.x;
      const unsigned code = (unsigned)c.y;
      vs[3 * u] = (int)(code & 1023u) - 512; vs[3 * u + 1] = (int)((code >> 10) & 1023u) - 512; vs[3 * u + 2] = (int)((code >> 20) & 1023u) - 512;
    }
    if (q + 4 <= P1) {
      *(int4*)(io + q) = make_int4(vi[0], vi[1], vi[2], vi[3]);
      *(int4*)(jo + q) = make_int4(vjj[0], vjj[1], vjj[2], vjj[3]);
      int4* d = (int4*)(So + 3 * q);
      d[0] = make_int4(vs[0], vs[1], vs[2], vs[3]); d[1] = make_int4(vs[4], vs[5], vs[6], vs[7]); d[2] = make_int4(vs[8], vs[9], vs[10], vs[11]);
    } else {
      for (int u = 0; u < 4 && q + u < P1; u++) { io[q + u] = vi[u]; jo[q + u] = vjj[u]; for (int k = 0; k < 3; k++) So[3 * (q + u) + k] = vs[3 * u + k]; }
    }
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

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 10000000;
  std::mt19937 rng(1);
  std::poisson_distribution<int> pd(26.18);
  std::vector<int> first(n + 1), perm(n), first2(n + 1), cstart(n);
  first[0] = 0;
  for (int i = 0; i < n; i++) first[i + 1] = first[i] + pd(rng);
  for (int i = 0; i < n; i++) perm[i] = i;
  std::shuffle(perm.begin(), perm.end(), rng);
  first2[0] = 0;
  for (int w = 0; w < n; w++) { cstart[perm[w]] = first2[w]; first2[w + 1] = first2[w] + (first[perm[w] + 1] - first[perm[w]]); }
  std::vector<int> first2_seq(first.begin(), first.end()), cstart_seq(first.begin(), first.end() - 1);
  const long long P = first[n];
  int *d_first, *d_perm, *d_first2, *d_cstart, *d_first2s, *d_cstarts, *io, *jo, *So; int2* compact; double* Ro;
  CK(cudaMalloc(&d_first, (size_t)(n + 1) * 4)); CK(cudaMalloc(&d_perm, (size_t)n * 4)); CK(cudaMalloc(&d_first2, (size_t)(n + 1) * 4)); CK(cudaMalloc(&d_cstart, (size_t)n * 4));
  CK(cudaMalloc(&d_first2s, (size_t)(n + 1) * 4)); CK(cudaMalloc(&d_cstarts, (size_t)n * 4));
  CK(cudaMalloc(&io, P * 4 + 64)); CK(cudaMalloc(&jo, P * 4 + 64)); CK(cudaMalloc(&So, P * 12 + 64)); CK(cudaMalloc(&Ro, P * 24 + 64)); CK(cudaMalloc(&compact, P * 8 + 64));
  CK(cudaMemcpy(d_first, first.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_perm, perm.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_first2, first2.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_cstart, cstart.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_first2s, first2_seq.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_cstarts, cstart_seq.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  printf("n = %d rows, P = %lld pairs\n", n, P);
  for (int blocks : {148 * 8, 148 * 16}) {
    for (int mode = 0; mode < 2; mode++) {
      const int* ord = mode ? d_perm : nullptr;
      const int* f2 = mode ? d_first2 : d_first2s;
      float t = time_best([&] { k_main<true, true><<<blocks, 256>>>(d_first, ord, f2, n, compact, Ro); });
      printf("%-10s main: compact (8 B) + R in place (24 B), %5d CTAs   %7.3f ms  %6.0f GB/s\n", mode ? "RANDOM" : "sequential", blocks, t, (double)P * 32 / t / 1e6);
      t = time_best([&] { k_main<true, false><<<blocks, 256>>>(d_first, ord, f2, n, compact, Ro); });
      printf("%-10s main: R in place only,                      %5d CTAs   %7.3f ms  %6.0f GB/s\n", mode ? "RANDOM" : "sequential", blocks, t, (double)P * 24 / t / 1e6);
      t = time_best([&] { k_main<false, true><<<blocks, 256>>>(d_first, ord, f2, n, compact, Ro); });
      printf("%-10s main: compact only,                         %5d CTAs   %7.3f ms  %6.0f GB/s\n", mode ? "RANDOM" : "sequential", blocks, t, (double)P * 8 / t / 1e6);
    }
  }
  for (int mode = 0; mode < 2; mode++) {
    const int* cs = mode ? d_cstart : d_cstarts;
    float t = time_best([&] { k_emit<<<(unsigned)((P + EM_PAIRS - 1) / EM_PAIRS), 256>>>(d_first, cs, n, P, compact, io, jo, So); });
    printf("emit (compact rows at %s places): read 8 B, write 20 B per pair   %7.3f ms  %6.0f GB/s\n", mode ? "RANDOM" : "sequential", t, (double)P * 28 / t / 1e6);
  }
  // verify emit against the random-order main pass
  k_main<true, true><<<148 * 8, 256>>>(d_first, d_perm, d_first2, n, compact, Ro);
  k_emit<<<(unsigned)((P + EM_PAIRS - 1) / EM_PAIRS), 256>>>(d_first, d_cstart, n, P, compact, io, jo, So);
  CK(cudaDeviceSynchronize());
  const long long chk = std::min<long long>(P, 3000000);
  long long bad = 0;
  for (int part = 0; part < 2; part++) {
    const long long q0 = part ? P - chk : 0;
    std::vector<int> hi(chk), hj(chk), hS(3 * chk); std::vector<double> hR(3 * chk);
    CK(cudaMemcpy(hi.data(), io + q0, chk * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hj.data(), jo + q0, chk * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hS.data(), So + 3 * q0, chk * 12, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hR.data(), Ro + 3 * q0, chk * 24, cudaMemcpyDeviceToHost));
    int row = (int)(std::upper_bound(first.begin(), first.end(), (int)q0) - first.begin()) - 1;
    for (long long q = 0; q < chk; q++) {
      const long long p = q0 + q;
      while (first[row + 1] <= p) row++;
      const unsigned code = vcode(p);
      if (hi[q] != row + 1 || hj[q] != vj(p)) bad++;
      if (hS[3 * q] != (int)(code & 1023u) - 512 || hS[3 * q + 1] != (int)((code >> 10) & 1023u) - 512 || hS[3 * q + 2] != (int)((code >> 20) & 1023u) - 512) bad++;
      if (hR[3 * q] != (double)p || hR[3 * q + 1] != (double)p + 0.25 || hR[3 * q + 2] != (double)p + 0.5) bad++;
    }
  }
  printf("verify main + emit: %lld mismatches in 2 x %lld pairs\n", bad, chk);
  return 0;
}
