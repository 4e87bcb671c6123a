// nl_tohost.cuh -- Array(PairList): the materialised list into HOST memory (what a host consumer of
// neighbour_list gets; the reference's CPU path returns host Vectors, src/cell_list.jl:897-916, and its GPU tests
// bring the device list back with Array(...) before comparing, test/test_utils.jl:127-131).
//
// The 20 B/pair of (i, j, S) cost 100 ms of PCIe at the headline size while the list itself takes 7.5 ms to build, so the
// transfer is compressed: over the bus go `first`, `j` and ONE BYTE per pair for S (every component in {-1, 0, 1}, which is
// what a stencil of half-width 1 over wrapped atoms produces); `i` is rebuilt from `first` and S from the byte codes by host
// threads of the library, with non-temporal stores, WHILE the copies are in flight.  Lists with a shift component outside
// {-1, 0, 1} (stray atoms far outside the cell, wide stencils) send S as it is.
#pragma once
#include <atomic>
#include <thread>
#include <vector>
#include "nl_common.cuh"

namespace nl {

constexpr int TH_PAIRS = 8;  // pairs per thread: 96 / 192 bytes of S in, 8 code bytes out

// code = (Sx + 1) + 3 (Sy + 1) + 9 (Sz + 1) in 0..26; *escape |= 1 when a component is outside {-1, 0, 1}.
// Block = 256 threads x 8 pairs; S is read unit-stride as 16-byte words through shared memory.
template <class TI>
__global__ void __launch_bounds__(256) k_pack_shifts(const TI* __restrict__ S, long long P, uint8_t* __restrict__ codes, unsigned* __restrict__ escape,
                                                     int aligned) {
  constexpr int NP = 256 * TH_PAIRS;
  __shared__ TI s[NP * 3];
  const long long p0 = (long long)blockIdx.x * NP;
  const int np = (int)min((long long)NP, P - p0);
  const TI* src = S + 3 * p0;
  if (np == NP && aligned) {
    const int4* s4 = (const int4*)src;  // S is 16-byte aligned and 3 * p0 * sizeof(TI) is a multiple of 16
    int4* d4 = (int4*)s;
    constexpr int NV = NP * 3 * (int)sizeof(TI) / 16;
#pragma unroll 4
    for (int k = threadIdx.x; k < NV; k += 256) d4[k] = __ldcs(s4 + k);
  } else {
    for (int k = threadIdx.x; k < 3 * np; k += 256) s[k] = src[k];
  }
  __syncthreads();
  bool bad = false;
  uint32_t w[2] = {0, 0};
#pragma unroll
  for (int q = 0; q < TH_PAIRS; q++) {
    const int p = threadIdx.x * TH_PAIRS + q;
    uint32_t c = 13;
    if (p < np) {
      const long long a = (long long)s[3 * p], b = (long long)s[3 * p + 1], d = (long long)s[3 * p + 2];
      bad |= (a < -1 || a > 1 || b < -1 || b > 1 || d < -1 || d > 1);
      c = (uint32_t)((a + 1) + 3 * (b + 1) + 9 * (d + 1)) & 255u;
    }
    w[q >> 2] |= c << (8 * (q & 3));
  }
  if (threadIdx.x * TH_PAIRS + TH_PAIRS <= np && ((uintptr_t)codes & 7) == 0) {
    *(uint2*)(codes + p0 + threadIdx.x * TH_PAIRS) = make_uint2(w[0], w[1]);  // p0 multiple of 2048: 8-byte aligned when codes is
  } else {
    for (int q = 0; q < TH_PAIRS; q++) {
      const int p = threadIdx.x * TH_PAIRS + q;
      if (p < np) codes[p0 + p] = (uint8_t)(w[q >> 2] >> (8 * (q & 3)));
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(escape, 1u);
}

// ------------------------------------------------------------------------------------------------ host side
// The decoders live in nl_hostcodec.cpp (plain C++, AVX2 variants picked at run time).
}  // namespace nl
namespace nl_host {
void expand_rows(int int64, const void* first, const void* row_map, long long n_rows, long long p_lo, long long p_hi, void* i_out);
void unpack_shifts(int int64, const uint8_t* codes, long long p_lo, long long p_hi, void* S_out);
}  // namespace nl_host
namespace nl {
// i[p] = r + 1 (or row_map[r]) for first[r] - 1 <= p < first[r + 1] - 1 (first is 1-based), for p in [p_lo, p_hi)
template <class TI>
inline void host_expand_rows(const TI* first, const TI* row_map, long long n_rows, long long p_lo, long long p_hi, TI* i_out) {
  nl_host::expand_rows(sizeof(TI) == 8, first, row_map, n_rows, p_lo, p_hi, i_out);
}
// S[p] = decode(codes[p]) for p in [p_lo, p_hi)
template <class TI>
inline void host_unpack_shifts(const uint8_t* codes, long long p_lo, long long p_hi, TI* S_out) {
  nl_host::unpack_shifts(sizeof(TI) == 8, codes, p_lo, p_hi, S_out);
}

}  // namespace nl
