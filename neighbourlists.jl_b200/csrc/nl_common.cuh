// nl_common.cuh -- shared device helpers: the arithmetic contract and the cell-index algebra.
//
// The arithmetic contract restates, operation for operation, what the reference
// evaluates per candidate pair (src/gpu_kernels.jl:84-92) and per atom (src/cell_list.jl:66-74):
// separate IEEE multiplies and adds in T, left-associated, never fused.  All contract arithmetic
// goes through the *_rn intrinsics below, which the compiler never contracts into FMA, so parity
// does not depend on -fmad.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

namespace nl {

constexpr unsigned FULL = 0xffffffffu;

// Process-wide count of kernels this library has launched (nl_launch_count(), used by bench.py).
inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}
inline void note_launch(int n = 1) { launch_counter().fetch_add(n, std::memory_order_relaxed); }

// Last cudaError_t seen by this thread's library calls (nl_last_cuda_error()).
inline int& last_cuda_slot() {
  thread_local int e = 0;
  return e;
}

__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ long long floor_ll(double v) { return __double2ll_rd(v); }
__device__ __forceinline__ long long floor_ll(float v) { return __float2ll_rd(v); }

// Packed Float32 pairs (Blackwell FADD2 / FMUL2 / FFMA2), in PTX with explicit .rn.
// CAUTION (verified in SASS, CUDA 12.9): ptxas contracts a packed multiply feeding a packed add into FFMA2
// even with explicit .rn and -fmad=false.  Contract arithmetic may therefore use mul2_rn and add2_rn, but
// never add2_rn on the result of mul2_rn: sum packed products with scalar __fadd_rn.
__device__ __forceinline__ unsigned long long f2_bits(float2 v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 bits_f2(unsigned long long r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(r);
}
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(r);
}
__device__ __forceinline__ float2 fma2_rn(float2 a, float2 b, float2 c) {  // deliberately fused (pre-filter only)
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(r);
}

// Geometry in the compute type T (device-side copy of nl_params).
template <class T> struct Geo {
  T cell[9];  // column-major, rows = lattice vectors
  T inv[9];
  T cutoff_sq;  // cutoff * cutoff rounded in T (src/gpu_kernels.jl:317)
  int nc[3];
  int nxyz[3];
  int pbc[3];
  int nct;  // prod(nc)
};

// (M' * v)_k = (M[1,k] v1 + M[2,k] v2) + M[3,k] v3   -- StaticArrays' unrolled mat-vec
template <class T> __device__ __forceinline__ void mtv(const T* m, T v0, T v1, T v2, T& o0, T& o1, T& o2) {
  o0 = add_rn(add_rn(mul_rn(m[0], v0), mul_rn(m[1], v1)), mul_rn(m[2], v2));
  o1 = add_rn(add_rn(mul_rn(m[3], v0), mul_rn(m[4], v1)), mul_rn(m[5], v2));
  o2 = add_rn(add_rn(mul_rn(m[6], v0), mul_rn(m[7], v1)), mul_rn(m[8], v2));
}

// floor division / modulus for any sign (closed form of wrap_and_shift's loops,
// src/cell_list.jl:105-120, in 0-based indices: i0 = wrapped + shift * n).
__device__ __forceinline__ void wrap0(long long i0, int n, int& wrapped, long long& shift) {
  long long q = i0 / n, r = i0 % n;
  if (r < 0) { r += n; q -= 1; }
  wrapped = (int)r;
  shift = q;
}

// position_to_cell_index (src/cell_list.jl:66-74) followed by bin_wrap_and_shift (:141-147):
// 0-based wrapped/clamped cell c[3] and winding w[3] (0 on open axes).
template <class T>
__device__ __forceinline__ void cell_of(const Geo<T>& g, T x, T y, T z, int c[3], long long w[3]) {
  T f[3];
  mtv(g.inv, x, y, z, f[0], f[1], f[2]);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    // floor(frac * n + 1) is the 1-based raw index; subtract 1 in integers -> 0-based raw index
    long long raw = floor_ll(add_rn(mul_rn(f[k], (T)g.nc[k]), (T)1)) - 1;
    if (g.pbc[k]) {
      wrap0(raw, g.nc[k], c[k], w[k]);
    } else {
      c[k] = raw < 0 ? 0 : (raw >= g.nc[k] ? g.nc[k] - 1 : (int)raw);
      w[k] = 0;
    }
  }
}

// Winding numbers are carried per atom as 3 x 10 bits (bias 512) in one word; anything outside
// [-511, 511] sets WIND_OVERFLOW and is recomputed from the position where it is needed.
constexpr uint32_t WIND_OVERFLOW = 0x80000000u;
constexpr uint32_t WIND_ZERO = 512u | (512u << 10) | (512u << 20);
__device__ __forceinline__ uint32_t pack_wind(const long long w[3]) {
  if (w[0] < -511 || w[0] > 511 || w[1] < -511 || w[1] > 511 || w[2] < -511 || w[2] > 511) return WIND_OVERFLOW;
  return (uint32_t)(w[0] + 512) | ((uint32_t)(w[1] + 512) << 10) | ((uint32_t)(w[2] + 512) << 20);
}
__device__ __forceinline__ void unpack_wind(uint32_t p, long long w[3]) {
  w[0] = (long long)(p & 1023u) - 512;
  w[1] = (long long)((p >> 10) & 1023u) - 512;
  w[2] = (long long)((p >> 20) & 1023u) - 512;
}

// One candidate pair under the contract: S given, R = (xj - xi) + cell' * S, r2 = dot(R, R).
template <class T>
__device__ __forceinline__ T pair_r2(const Geo<T>& g, T xi, T yi, T zi, T xj, T yj, T zj, const long long S[3], T R[3]) {
  T cs[3];
  mtv(g.cell, (T)S[0], (T)S[1], (T)S[2], cs[0], cs[1], cs[2]);
  R[0] = add_rn(sub_rn(xj, xi), cs[0]);
  R[1] = add_rn(sub_rn(yj, yi), cs[1]);
  R[2] = add_rn(sub_rn(zj, zi), cs[2]);
  return add_rn(add_rn(mul_rn(R[0], R[0]), mul_rn(R[1], R[1])), mul_rn(R[2], R[2]));
}

}  // namespace nl
