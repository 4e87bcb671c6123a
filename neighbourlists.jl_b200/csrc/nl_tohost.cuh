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
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
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
// i[p] = r + 1 for first[r] - 1 <= p < first[r + 1] - 1 (first is 1-based), for p in [p_lo, p_hi).
template <class TI>
void host_expand_rows(const TI* first, long long n_rows, long long p_lo, long long p_hi, TI* i_out) {
  if (p_hi <= p_lo) return;
  // row of p_lo: last r with first[r] - 1 <= p_lo
  long long lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)first[mid] - 1 <= p_lo) lo = mid; else hi = mid;
  }
  long long r = lo;
  long long p = p_lo;
  long long e = (long long)first[r + 1] - 1;  // end of row r
  while (e <= p && r + 1 < n_rows) { r++; e = (long long)first[r + 1] - 1; }
#if defined(__SSE2__)
  if (sizeof(TI) == 4) {
    int* out = (int*)i_out;
    while (p < p_hi && (((uintptr_t)(out + p)) & 15)) {  // head up to a 16-byte boundary
      while (e <= p) { r++; e = (long long)first[r + 1] - 1; }
      out[p++] = (int)(r + 1);
    }
    while (p + 4 <= p_hi) {
      while (e <= p) { r++; e = (long long)first[r + 1] - 1; }
      if (p + 4 <= e) {
        const __m128i v = _mm_set1_epi32((int)(r + 1));
        const long long stop = (e < p_hi ? e : p_hi) - 3;
        for (; p < stop; p += 4) _mm_stream_si128((__m128i*)(out + p), v);
      } else {
        int t[4];
        for (int k = 0; k < 4; k++) {
          while (e <= p + k) { r++; e = (long long)first[r + 1] - 1; }
          t[k] = (int)(r + 1);
        }
        _mm_stream_si128((__m128i*)(out + p), _mm_set_epi32(t[3], t[2], t[1], t[0]));
        p += 4;
      }
    }
  }
#endif
  for (; p < p_hi; p++) {
    while (e <= p) { r++; e = (long long)first[r + 1] - 1; }
    i_out[p] = (TI)(r + 1);
  }
#if defined(__SSE2__)
  _mm_sfence();
#endif
}

// S[p] = decode(codes[p]) for p in [p_lo, p_hi)
template <class TI>
void host_unpack_shifts(const uint8_t* codes, long long p_lo, long long p_hi, TI* S_out) {
  long long p = p_lo;
#if defined(__SSE2__)
  if (sizeof(TI) == 4 && (((uintptr_t)S_out) & 15) == 0) {
    alignas(16) static const struct Lut {
      int v[256][4];
      Lut() {
        for (int c = 0; c < 256; c++) {
          const int k = c < 27 ? c : 13;
          v[c][0] = k % 3 - 1; v[c][1] = (k / 3) % 3 - 1; v[c][2] = k / 9 - 1; v[c][3] = 0;
        }
      }
    } lut;
    int* out = (int*)S_out;
    for (; p < p_hi && (p & 3); p++) {
      const int* t = lut.v[codes[p]];
      out[3 * p] = t[0]; out[3 * p + 1] = t[1]; out[3 * p + 2] = t[2];
    }
    for (; p + 4 <= p_hi; p += 4) {  // 4 pairs = 48 bytes = three aligned 16-byte words
      uint32_t c4;
      memcpy(&c4, codes + p, 4);
      if (c4 == 0x0d0d0d0du) {  // the common case: no shift
        const __m128i z = _mm_setzero_si128();
        __m128i* d = (__m128i*)(out + 3 * p);
        _mm_stream_si128(d, z); _mm_stream_si128(d + 1, z); _mm_stream_si128(d + 2, z);
        continue;
      }
      const int* a = lut.v[c4 & 255], *b = lut.v[(c4 >> 8) & 255], *c = lut.v[(c4 >> 16) & 255], *e = lut.v[c4 >> 24];
      __m128i* d = (__m128i*)(out + 3 * p);
      _mm_stream_si128(d, _mm_set_epi32(b[0], a[2], a[1], a[0]));
      _mm_stream_si128(d + 1, _mm_set_epi32(c[1], c[0], b[2], b[1]));
      _mm_stream_si128(d + 2, _mm_set_epi32(e[2], e[1], e[0], c[2]));
    }
  }
#endif
  for (; p < p_hi; p++) {
    const int k = codes[p] < 27 ? codes[p] : 13;
    S_out[3 * p] = (TI)(k % 3 - 1);
    S_out[3 * p + 1] = (TI)((k / 3) % 3 - 1);
    S_out[3 * p + 2] = (TI)(k / 9 - 1);
  }
#if defined(__SSE2__)
  _mm_sfence();
#endif
}

}  // namespace nl
