// nl_shard.cuh -- multi-GPU slabs: the device side of the redistribution and of the halo exchange, and the NCCL
// calls, behind nl_shard_prepare / nl_shard_exchange (include/nlcuda.h).
//
// One process per GPU.  The box is cut into slabs of whole cell PLANES along the axis with the most cells (ties: z, the
// slowest key axis), keeping the reference's cell grid and linearisation (src/cell_list.jl:83-86, widths :94-95), so that the
// unchanged single-GPU stages run on each rank's local set with the GLOBAL geometry.
//
//   prepare   bin the local atoms to planes, per-rank plane histogram, ncclAllGather of the histograms, ONE host read:
//             every rank then knows every (source, destination) count, so no further size exchange is ever needed
//   exchange  owners by plane -> stable partition by destination (one pass of the radix sort) -> all-to-all-v
//             (ncclSend / ncclRecv group) straight into the caller's local arrays -> halo selection (one more partition)
//             -> halo exchange with ranks r-1 / r+1 into the tail of the same arrays
//
// NCCL is bound at run time (dlopen): the single-GPU library has no NCCL dependency, and inside a process that already
// loaded NCCL (PyTorch, NCCL.jl) the SAME instance is used, so a communicator made by the host framework works too.
#pragma once
#include <dlfcn.h>

#include <vector>

#include "nl_common.cuh"
#include "nl_scan_sort.cuh"

namespace nl {

// ---- NCCL, bound at run time.  Types restated from nccl.h (stable since NCCL 2.0).
struct NcclUid { char internal[128]; };
enum { NCCL_INT8 = 0, NCCL_UINT64 = 5 };
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  bool ok = false;
};
inline NcclApi& nccl() {
  static NcclApi a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);           // the instance the host framework already loaded
    if (!a.h) a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (a.h) {
      a.GetUniqueId = (int (*)(NcclUid*))dlsym(a.h, "ncclGetUniqueId");
      a.CommInitRank = (int (*)(void**, int, NcclUid, int))dlsym(a.h, "ncclCommInitRank");
      a.CommDestroy = (int (*)(void*))dlsym(a.h, "ncclCommDestroy");
      a.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(a.h, "ncclAllGather");
      a.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(a.h, "ncclSend");
      a.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(a.h, "ncclRecv");
      a.GroupStart = (int (*)())dlsym(a.h, "ncclGroupStart");
      a.GroupEnd = (int (*)())dlsym(a.h, "ncclGroupEnd");
      a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.Send && a.Recv && a.GroupStart && a.GroupEnd;
    }
  }
  return a;
}

// ---- kernels
// plane[i] = 0-based cell index of atom i along the slab axis (position_to_cell_index + bin_wrap_or_trunc, src/cell_list.jl:66-74,
// 126-132: exactly the binning of the build stage), plus this rank's plane histogram.
constexpr int SHARD_HIST_SMEM = 4096;
template <class T>
__global__ void __launch_bounds__(256) k_shard_planes(const T* __restrict__ X, long long n, Geo<T> g, int axis, int nplanes, int32_t* __restrict__ planes,
                                                      unsigned long long* __restrict__ hist) {
  __shared__ unsigned int sh[SHARD_HIST_SMEM];
  const bool in_smem = nplanes <= SHARD_HIST_SMEM;
  if (in_smem)
    for (int k = threadIdx.x; k < nplanes; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c[3];
    long long w[3];
    cell_of(g, X[3 * i], X[3 * i + 1], X[3 * i + 2], c, w);
    const int p = c[axis];
    planes[i] = p;
    if (in_smem) atomicAdd(&sh[p], 1u); else atomicAdd(&hist[p], 1ull);
  }
  __syncthreads();
  if (in_smem)
    for (int k = threadIdx.x; k < nplanes; k += blockDim.x)
      if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// key[i] = rank that owns plane[i]: the r with bounds[r] <= plane < bounds[r+1]
__global__ void __launch_bounds__(256) k_shard_owner(const int32_t* __restrict__ planes, long long n, const long long* __restrict__ bounds, int nranks,
                                                     uint32_t* __restrict__ keys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long p = planes[i];
  int lo = 0, hi = nranks;  // last r with bounds[r] <= p
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (bounds[mid] <= p) lo = mid; else hi = mid;
  }
  keys[i] = (uint32_t)lo;
}

// key[i] = 0: atom goes into the halo sent DOWN (bottom planes), 1: into the halo sent UP (top planes), 2: neither.
// Slabs are at least 2 * halo + 1 planes wide, so no atom is in both.
__global__ void __launch_bounds__(256) k_shard_halo_class(const int32_t* __restrict__ planes, long long n, long long lo, long long hi, int halo, int has_dn,
                                                          int has_up, uint32_t* __restrict__ keys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long p = planes[i];
  keys[i] = (has_dn && p < lo + halo) ? 0u : ((has_up && p >= hi - halo) ? 1u : 2u);
}

// out[k] = in[order[k0 + k]] for k < cnt: positions (3 T per atom), global indices (TI) and planes.
template <class T, class TI>
__global__ void __launch_bounds__(256) k_shard_gather(const uint32_t* __restrict__ order, long long k0, long long cnt, const T* __restrict__ X,
                                                      const TI* __restrict__ gidx, const int32_t* __restrict__ planes, T* __restrict__ Xo,
                                                      TI* __restrict__ go, int32_t* __restrict__ po) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const long long s = order[k0 + k];
  Xo[3 * k] = X[3 * s]; Xo[3 * k + 1] = X[3 * s + 1]; Xo[3 * k + 2] = X[3 * s + 2];
  go[k] = gidx[s];
  if (po) po[k] = planes[s];
}

// ---- workspace of the shard stage, for `n` = max(local atoms, owned atoms)
struct ShardWs {
  int32_t *planes, *planes_owned, *sendp;
  unsigned long long *hist_local, *hist_all;
  long long* bounds;
  uint32_t *keyA, *keyB, *valA, *valB;
  void* rs_scratch;
  void *sendX, *sendg;
  void *haloX, *halog;            // halo send buffers of the peer path (the NCCL path reuses sendX / sendg)
  unsigned long long* bar;        // 8 (nranks + 1) bytes: the all-gather used as a barrier
  size_t total;
};
inline ShardWs shard_ws(void* ws, long long n, int nplanes, int nranks, size_t fbytes, size_t ibytes) {
  ShardWs w;
  char* p = (char*)ws;
  size_t o = 0;
  auto take = [&](size_t b) { char* r = p ? p + o : nullptr; o += (b + 255) & ~(size_t)255; return (void*)r; };
  const size_t n1 = (size_t)(n > 0 ? n : 1);
  w.planes = (int32_t*)take(n1 * 4);
  w.planes_owned = (int32_t*)take(n1 * 4);
  w.sendp = (int32_t*)take(n1 * 4);
  w.hist_local = (unsigned long long*)take((size_t)nplanes * 8);
  w.hist_all = (unsigned long long*)take((size_t)nplanes * 8 * nranks);
  w.bounds = (long long*)take((size_t)(nranks + 1) * 8);
  w.keyA = (uint32_t*)take(n1 * 4);
  w.keyB = (uint32_t*)take(n1 * 4);
  w.valA = (uint32_t*)take(n1 * 4);
  w.valB = (uint32_t*)take(n1 * 4);
  w.rs_scratch = take(rs_scratch_bytes((long long)n1));
  w.sendX = take(n1 * 3 * fbytes);
  w.sendg = take(n1 * ibytes);
  w.haloX = take(n1 * 3 * fbytes);
  w.halog = take(n1 * ibytes);
  w.bar = (unsigned long long*)take((size_t)(nranks + 1) * 8);
  w.total = o;
  return w;
}

}  // namespace nl
