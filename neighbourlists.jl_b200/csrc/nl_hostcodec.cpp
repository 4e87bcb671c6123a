// nl_hostcodec.cpp -- host side of the nl_pairs_to_host transfer format (include/nlcuda.h): i rebuilt from `first`, S from
// one-byte shift codes.  Plain C++ (g++), no CUDA: compiled as its own translation unit so that the AVX2 variants can be built
// with per-function target attributes and picked at run time.  Non-temporal stores throughout: the arrays are written once
// and are far larger than the caches.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#define NL_X86 1
#else
#define NL_X86 0
#endif

namespace nl_host {

// ------------------------------------------------------------------------------------------------ generic (any TI)
// value written for the pairs of row r: r + 1, or map[r] (shard lists: the row's GLOBAL atom index)
template <class TI> static inline TI row_value(const TI* map, long long r) { return map ? map[r] : (TI)(r + 1); }

template <class TI>
static void expand_scalar(const TI* first, const TI* map, long long n_rows, long long p, long long p_hi, long long& r, long long& e, TI* out) {
  for (; p < p_hi; p++) {
    while (e <= p) { r++; e = (long long)first[r + 1] - 1; }
    out[p] = row_value<TI>(map, r);
  }
}

// row of pair p: last r with first[r] - 1 <= p (first is 1-based); skips empty rows so that e = end of row r > p
template <class TI>
static void locate(const TI* first, long long n_rows, long long p, long long& r, long long& e) {
  long long lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)first[mid] - 1 <= p) lo = mid; else hi = mid;
  }
  r = lo;
  e = (long long)first[r + 1] - 1;
  while (e <= p && r + 1 < n_rows) { r++; e = (long long)first[r + 1] - 1; }
}

#if NL_X86
// 32-bit indices, groups of 4 (SSE2) / 8 (AVX2) pairs on aligned addresses.  Whole groups inside a row are one broadcast store; a
// group in which exactly ONE row ends is emitted without a per-element loop (lanes at or after the row end get r + 2, the others
// r + 1); groups with two or more row ends (rows shorter than the group, empty rows) take the element-wise path.
__attribute__((target("avx2"))) static void expand_avx2(const int32_t* first, const int32_t* map, long long n_rows, long long p, long long p_hi,
                                                         long long& r_io, long long& e_io, int32_t* out) {
  const __m256i iota = _mm256_set_epi32(7, 6, 5, 4, 3, 2, 1, 0);
  long long r = r_io, e = e_io;
  while (p + 8 <= p_hi) {
    while (e <= p) { r++; e = (long long)first[r + 1] - 1; }
    long long d = e - p;                                               // > 0: pairs of row r left from p on
    if (d >= 8) {                                                      // the whole group inside row r: run to the last full group of the row
      const __m256i v = _mm256_set1_epi32((int)row_value<int32_t>(map, r));
      const long long stop = (e < p_hi ? e : p_hi) - 7;
      for (; p < stop; p += 8) _mm256_stream_si256((__m256i*)(out + p), v);
      continue;
    }
    // row r ends inside the group; if the next row ends inside it as well, fall back to the element-wise path
    const long long e2 = (r + 1 < n_rows) ? (long long)first[r + 2] - 1 : p + 8;
    if (e2 < p + 8) {
      alignas(32) int32_t t[8];
      for (int k = 0; k < 8; k++) {
        while (e <= p + k) { r++; e = (long long)first[r + 1] - 1; }
        t[k] = row_value<int32_t>(map, r);
      }
      _mm256_stream_si256((__m256i*)(out + p), _mm256_load_si256((const __m256i*)t));
      p += 8;
      continue;
    }
    const __m256i past = _mm256_cmpgt_epi32(iota, _mm256_set1_epi32((int)d - 1));   // lanes k >= d belong to row r + 1
    _mm256_stream_si256((__m256i*)(out + p), _mm256_blendv_epi8(_mm256_set1_epi32((int)row_value<int32_t>(map, r)),
                                                                 _mm256_set1_epi32((int)row_value<int32_t>(map, r + 1)), past));
    r++; e = e2;
    p += 8;
  }
  r_io = r; e_io = e;
  // the caller finishes [p, p_hi) -- fewer than 8 pairs -- element by element
}

static bool have_avx2() {
  static const int v = (__builtin_cpu_supports("avx2") && !getenv("NL_HOST_NO_AVX2")) ? 1 : 0;  // the switch is for testing the SSE2 path
  return v != 0;
}
#endif

template <class TI>
static void expand_rows_t(const TI* first, const TI* map, long long n_rows, long long p_lo, long long p_hi, TI* out) {
  if (p_hi <= p_lo) return;
  long long r, e;
  locate<TI>(first, n_rows, p_lo, r, e);
  long long p = p_lo;
#if NL_X86
  if (sizeof(TI) == 4) {
    const bool avx2 = have_avx2();
    const long long G = avx2 ? 8 : 4;
    // head up to the first aligned group
    long long head = p;
    while (head < p_hi && (((uintptr_t)(out + head)) & (uintptr_t)(4 * G - 1))) head++;
    expand_scalar<TI>(first, map, n_rows, p, head, r, e, out);
    p = head;
    const long long body = p + ((p_hi - p) / G) * G;
    if (body > p) {
      if (avx2) expand_avx2((const int32_t*)first, (const int32_t*)map, n_rows, p, body, r, e, (int32_t*)out);
      else {
        // SSE2: same scheme with groups of 4 (kept simple: run loop + boundary groups)
        const int32_t* f32 = (const int32_t*)first;
        const int32_t* m32 = (const int32_t*)map;
        int32_t* o32 = (int32_t*)out;
        const __m128i iota = _mm_set_epi32(3, 2, 1, 0);
        while (p + 4 <= body) {
          while (e <= p) { r++; e = (long long)f32[r + 1] - 1; }
          const long long d = e - p;
          if (d >= 4) {
            const __m128i v = _mm_set1_epi32((int)row_value<int32_t>(m32, r));
            const long long stop = (e < body ? e : body) - 3;
            for (; p < stop; p += 4) _mm_stream_si128((__m128i*)(o32 + p), v);
            continue;
          }
          const long long e2 = (r + 1 < n_rows) ? (long long)f32[r + 2] - 1 : p + 4;
          if (e2 < p + 4) {
            alignas(16) int32_t t[4];
            for (int k = 0; k < 4; k++) {
              while (e <= p + k) { r++; e = (long long)f32[r + 1] - 1; }
              t[k] = row_value<int32_t>(m32, r);
            }
            _mm_stream_si128((__m128i*)(o32 + p), _mm_load_si128((const __m128i*)t));
            p += 4;
            continue;
          }
          const __m128i past = _mm_cmpgt_epi32(iota, _mm_set1_epi32((int)d - 1));   // SSE2 has no blend: (a & ~m) | (b & m)
          const __m128i va = _mm_set1_epi32((int)row_value<int32_t>(m32, r)), vb = _mm_set1_epi32((int)row_value<int32_t>(m32, r + 1));
          _mm_stream_si128((__m128i*)(o32 + p), _mm_or_si128(_mm_andnot_si128(past, va), _mm_and_si128(past, vb)));
          r++; e = e2;
          p += 4;
        }
      }
      p = body;
      // the vector paths leave (r, e) at the row of the last pair they wrote; catch up lazily in the scalar tail
    }
  }
#endif
  expand_scalar<TI>(first, map, n_rows, p, p_hi, r, e, out);
#if NL_X86
  _mm_sfence();
#endif
}

// ------------------------------------------------------------------------------------------------ S from one-byte codes
struct Lut {
  alignas(16) int v[256][4];
  Lut() {
    for (int c = 0; c < 256; c++) {
      const int k = c < 27 ? c : 13;
      v[c][0] = k % 3 - 1; v[c][1] = (k / 3) % 3 - 1; v[c][2] = k / 9 - 1; v[c][3] = 0;
    }
  }
};
static const Lut lut;

template <class TI>
static void unpack_shifts_t(const uint8_t* codes, long long p_lo, long long p_hi, TI* S_out) {
  long long p = p_lo;
#if NL_X86
  if (sizeof(TI) == 4 && (((uintptr_t)S_out) & 15) == 0) {
    int* out = (int*)S_out;
    for (; p < p_hi && (p & 3); p++) {
      const int* t = lut.v[codes[p]];
      out[3 * p] = t[0]; out[3 * p + 1] = t[1]; out[3 * p + 2] = t[2];
    }
    const __m128i z = _mm_setzero_si128();
    for (; p + 4 <= p_hi; p += 4) {  // 4 pairs = 48 bytes = three aligned 16-byte words
      uint32_t c4;
      memcpy(&c4, codes + p, 4);
      __m128i* d = (__m128i*)(out + 3 * p);
      if (c4 == 0x0d0d0d0du) {  // the common case: no shift
        _mm_stream_si128(d, z); _mm_stream_si128(d + 1, z); _mm_stream_si128(d + 2, z);
        continue;
      }
      const int *a = lut.v[c4 & 255], *b = lut.v[(c4 >> 8) & 255], *c = lut.v[(c4 >> 16) & 255], *e = lut.v[c4 >> 24];
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
#if NL_X86
  _mm_sfence();
#endif
}

// ------------------------------------------------------------------------------------------------ entry points for nlcuda.cu
void expand_rows(int int64, const void* first, const void* row_map, long long n_rows, long long p_lo, long long p_hi, void* i_out) {
  if (int64) expand_rows_t<int64_t>((const int64_t*)first, (const int64_t*)row_map, n_rows, p_lo, p_hi, (int64_t*)i_out);
  else expand_rows_t<int32_t>((const int32_t*)first, (const int32_t*)row_map, n_rows, p_lo, p_hi, (int32_t*)i_out);
}
void unpack_shifts(int int64, const uint8_t* codes, long long p_lo, long long p_hi, void* S_out) {
  if (int64) unpack_shifts_t<int64_t>(codes, p_lo, p_hi, (int64_t*)S_out);
  else unpack_shifts_t<int32_t>(codes, p_lo, p_hi, (int32_t*)S_out);
}

}  // namespace nl_host
