// nlcuda.cu -- the C ABI of libnlcuda.so (declared in include/nlcuda.h).
// Host-side stage orchestration only; kernels live in the .cuh files next to this one.
#include "../../include/nlcuda.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "nl_build.cuh"
#include "nl_common.cuh"
#include "nl_scan_sort.cuh"
#include "nl_traverse.cuh"
#include "nl_tiled.cuh"
#include "nl_mask.cuh"
#include "nl_fillrows.cuh"
#include "nl_fill2.cuh"
#include "nl_fill3.cuh"
#include "nl_count2.cuh"
#include "nl_shard.cuh"
#include "nl_access.cuh"
#include "nl_tohost.cuh"

namespace {

using namespace nl;


static_assert(sizeof(nl_params) == 192, "nl_params layout is part of the ABI");
static_assert(sizeof(nl_shard_info) == 1632 + 8 + 8 * NL_MAX_RANKS + 16, "nl_shard_info layout is part of the ABI");
static_assert(sizeof(nl_shard_peers) == 8 + 8 + 8 + 8 + 16 * NL_MAX_RANKS, "nl_shard_peers layout is part of the ABI");

inline int cuda_fail(cudaError_t e) {
  last_cuda_slot() = (int)e;
  return NL_ERR_CUDA;
}
#define NL_CUDA(expr)                          \
  do {                                         \
    cudaError_t e__ = (expr);                  \
    if (e__ != cudaSuccess) return cuda_fail(e__); \
  } while (0)
#define NL_LAUNCH_CHECK() NL_CUDA(cudaGetLastError())
#define NL_LAUNCHED(n) nl::note_launch(n)

inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }
inline size_t fsize(const nl_params* p) { return p->float_type == NL_F64 ? 8 : 4; }

int check_params(const nl_params* p, int64_t N) {
  if (!p) return NL_ERR_BAD_ARG;
  if (p->float_type != NL_F32 && p->float_type != NL_F64) return NL_ERR_BAD_ARG;
  if (p->int_type != NL_I32 && p->int_type != NL_I64) return NL_ERR_BAD_ARG;
  if (N < 0) return NL_ERR_BAD_ARG;
  long long nct = 1;
  for (int k = 0; k < 3; k++) {
    if (p->ncells[k] < 1 || p->nxyz[k] < 1) return NL_ERR_BAD_ARG;
    nct *= p->ncells[k];
    if (nct >= 2147483647ll) return NL_ERR_UNSUPPORTED;
  }
  if (N >= 2147483647ll) return NL_ERR_UNSUPPORTED;
  if (!(p->cutoff > 0.0)) return NL_ERR_BAD_ARG;
  if ((p->reserved[0] & ~NL_FLAG_HALF) || p->reserved[1] || p->reserved[2] || p->reserved[3] || p->reserved[4]) return NL_ERR_BAD_ARG;
  return NL_OK;
}

template <class T> Geo<T> make_geo(const nl_params* p) {
  Geo<T> g;
  for (int k = 0; k < 9; k++) { g.cell[k] = (T)p->cell[k]; g.inv[k] = (T)p->inv_cell[k]; }
  volatile T c = (T)p->cutoff;  // volatile: one rounded multiply in T, as clist.cutoff^2
  volatile T c2 = c * c;
  g.cutoff_sq = c2;
  long long nct = 1;
  for (int k = 0; k < 3; k++) { g.nc[k] = p->ncells[k]; g.nxyz[k] = p->nxyz[k]; g.pbc[k] = p->pbc[k] ? 1 : 0; nct *= p->ncells[k]; }
  g.nct = (int)nct;
  return g;
}

int key_bits(long long nct) {
  int b = 1;
  while (b < 32 && (1ll << b) < nct) b++;
  return b;
}

// ---------------------------------------------------------------- workspace layouts
struct BuildWs {
  uint32_t *keyA, *keyB, *valA, *valB;
  void* rs_scratch;
  uint32_t* cnt;                 // bucket build: one counter per cell
  unsigned long long* cnt_tsum;  //               scan scratch
  size_t total;
};
// The bucket (counting-sort) build serves grids that are not much larger than the atom count and cells that are not crowded
// (the in-cell ordering reads a cell's whole segment per atom); everything else takes the radix sort.  NL_BUILD=radix forces it.
inline bool bucket_build_ok(long long nct, int64_t N) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NL_BUILD"); v = (e && e[0] == 'r') ? 0 : 1; }
  return v == 1 && N > 0 && nct <= 4 * (long long)N + 4096 && (long long)N <= 24 * nct;
}
BuildWs build_ws(void* ws, int64_t N, long long nct) {
  BuildWs w;
  char* p = (char*)ws;
  size_t o = 0;
  auto take = [&](size_t b) { char* r = p ? p + o : nullptr; o += al256(b); return (void*)r; };
  size_t nb = (size_t)(N > 0 ? N : 1) * 4;
  w.keyA = (uint32_t*)take(nb);
  w.keyB = (uint32_t*)take(nb);
  w.valA = (uint32_t*)take(nb);
  w.valB = (uint32_t*)take(nb);
  w.rs_scratch = take(rs_scratch_bytes(N > 0 ? N : 1));
  const long long ncnt = bucket_build_ok(nct, N) ? nct : 0;
  w.cnt = (uint32_t*)take((size_t)(ncnt + 1) * 4);
  w.cnt_tsum = (unsigned long long*)take((size_t)(scan_tiles(ncnt + 1) + 1) * 8);
  w.total = o;
  return w;
}
inline long long params_nct(const nl_params* p) { return (long long)p->ncells[0] * p->ncells[1] * p->ncells[2]; }

struct PairWs {
  void* hdr;
  void *px, *py, *pz;
  uint32_t *pidx, *pw, *counts, *pgid0, *pkey, *sorted_of;
  void* ra;
  void* srow;  // fill pass: 0-based row start of every SORTED atom (all ones: no row), TI-wide
  int* zlayers; // tile layers along z that a windowed (slab shard) call launches
  uint8_t* planes;  // device copy of the caller's plane_active promise (checked against cell_offsets)
  unsigned long long* tsum;
  unsigned long long* total;
  void* tiled;  // tiled-kernel scratch (tile table, hit masks)
  unsigned char* parkA;  // fill pass: parked row ends, 128 B per row (j head | j tail | S head | S tail)
  unsigned char* parkR;  //            64 B per row (R head | R tail)
  void* stamp;           // what nl_count_pairs left in this workspace (WsStamp); nl_fill_pairs* check it
  size_t total_bytes;
};
inline int fill_variant();
PairWs pair_ws(void* ws, const nl_params* prm, int64_t N) {
  PairWs w;
  char* p = (char*)ws;
  size_t o = 0;
  auto take = [&](size_t b) { char* r = p ? p + o : nullptr; o += al256(b); return (void*)r; };
  size_t n1 = (size_t)(N > 0 ? N : 1);
  w.hdr = take(4096);  // device copies of kernel argument blocks (rare out-of-line paths read them from here)
  w.px = take(n1 * fsize(prm));
  w.py = take(n1 * fsize(prm));
  w.pz = take(n1 * fsize(prm));
  w.pidx = (uint32_t*)take(n1 * 4);
  w.pw = (uint32_t*)take(n1 * 4);
  w.counts = (uint32_t*)take(n1 * 4);
  w.pgid0 = (uint32_t*)take(n1 * 4);
  w.pkey = (uint32_t*)take(n1 * 4);
  w.sorted_of = (uint32_t*)take(n1 * 4);
  w.ra = take(n1 * 32);
  w.srow = take(n1 * 8);
  w.zlayers = (int*)take(((size_t)prm->ncells[2] + 1) * 4);
  w.planes = (uint8_t*)take((size_t)prm->ncells[2]);
  w.tsum = (unsigned long long*)take((size_t)(scan_tiles((long long)n1) + 1) * 8);
  w.total = (unsigned long long*)take(256);
  w.tiled = take(tiled_scratch_bytes(prm, N));
  const size_t npark = fill_variant() == 2 ? n1 : 1;  // only the NL_FILL=park experiment uses the park records (192 B per atom)
  w.parkA = (unsigned char*)take(npark * PARK_A_BYTES);
  w.parkR = (unsigned char*)take(npark * PARK_R_BYTES);
  w.stamp = take(256);
  w.total_bytes = o;
  return w;
}

// ---------------------------------------------------------------- stage implementations
template <class T, class TI>
int build_cells_impl(const nl_params* p, const void* X, int64_t N, void* Xs, void* perm, void* cell_id, void* cell_offsets, void* ws,
                     cudaStream_t st) {
  Geo<T> g = make_geo<T>(p);
  const long long nct = g.nct;
  BuildWs w = build_ws(ws, N, nct);
  const uint32_t* skeys = w.keyA;
  if (bucket_build_ok(nct, N)) {
    const unsigned nb = (unsigned)((N + 255) / 256);
    NL_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)nct * 4, st));
    k_bin_count<T><<<nb, 256, 0, st>>>((const T*)X, N, g, w.keyA, w.valA, w.cnt);
    typedef typename std::conditional<sizeof(TI) == 4, uint32_t, unsigned long long>::type Acc;
    exclusive_scan<uint32_t, Acc, TI>(w.cnt, nct, (TI*)cell_offsets, (Acc)1, true, (Acc*)w.cnt_tsum, (Acc*)nullptr, st);
    k_bucket_scatter<TI><<<nb, 256, 0, st>>>(w.keyA, w.valA, (const TI*)cell_offsets, N, w.valB);
    k_finalize_buckets<T, TI><<<nb, 256, 0, st>>>(w.valB, w.keyA, (const TI*)cell_offsets, (const T*)X, N, (T*)Xs, (TI*)perm, (TI*)cell_id);
    NL_LAUNCHED(3);
    NL_LAUNCH_CHECK();
    return NL_OK;
  }
  if (N > 0) {
    const unsigned nb = (unsigned)((N + 255) / 256);
    k_bin<T><<<nb, 256, 0, st>>>((const T*)X, N, g, w.keyA);
    NL_LAUNCHED(1);
    NL_LAUNCH_CHECK();
    int where = radix_sort_pairs(w.keyA, w.valA, w.keyB, w.valB, N, key_bits(nct), w.rs_scratch, st);
    NL_LAUNCH_CHECK();
    skeys = where ? w.keyB : w.keyA;
    const uint32_t* svals = where ? w.valB : w.valA;
    k_finalize_sorted<T, TI><<<nb, 256, 0, st>>>(skeys, svals, (const T*)X, N, (T*)Xs, (TI*)perm, (TI*)cell_id);
    NL_LAUNCHED(1);
    NL_LAUNCH_CHECK();
  }
  k_cell_offsets<TI><<<(unsigned)((nct + 1 + 255) / 256), 256, 0, st>>>(skeys, N, nct, (TI*)cell_offsets);
  NL_LAUNCHED(1);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

template <class T, class TI>
int cell_ids_impl(const nl_params* p, const void* X, int64_t N, void* out, cudaStream_t st) {
  Geo<T> g = make_geo<T>(p);
  k_cell_ids<T, TI><<<(unsigned)((N + 255) / 256), 256, 0, st>>>((const T*)X, N, g, (TI*)out);
  NL_LAUNCHED(1);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

inline bool fill_tiled_requested();

template <class T, class TI>
int prep_impl(const nl_params* p, const void* Xs, int64_t N, const void* perm, PairWs& w, Geo<T>& g, cudaStream_t st) {
  if (N > 0) {
    k_prep_records<T, TI><<<(unsigned)((N + 255) / 256), 256, 0, st>>>((const T*)Xs, (const TI*)perm, N, g, (T*)w.px, (T*)w.py, (T*)w.pz,
                                                                      w.pidx, w.pw, fill_tiled_requested() ? nullptr : (RecAoS<T>*)w.ra, w.pkey);
    NL_LAUNCHED(1);
    NL_LAUNCH_CHECK();
  }
  return NL_OK;
}

template <class T> Records<T> records_of(const PairWs& w) {
  Records<T> r;
  r.px = (const T*)w.px; r.py = (const T*)w.py; r.pz = (const T*)w.pz; r.pidx = w.pidx; r.pw = w.pw;
  return r;
}

// Which traversal serves this problem.  A pure function of (params, N): nl_count_pairs and
// nl_fill_pairs must reach the same verdict.
enum { PATH_GENERIC = 0, PATH_TILED = 1, PATH_MASK = 2 };
template <class T> struct Plan {
  int path;
  TileShape ts_exact, ts_count, ts_fill, ts_fill2;
  int fill2_ok;         // the 512-thread parked-boundary fill (nl_fill2.cuh) has a tile for this density
  MaskThresholds th;
  int lazy_ok;          // the packed counting kernel can serve the lazy sinks (count; LJ for Float32)
  TileShape ts_lazy;
  MaskThresholds th_lazy;
  int ljf_ok;           // ... and the fused force sink (Float32; its own tile: the slots also carry accumulators)
  TileShape ts_ljf;
};
template <class T, class TI> Plan<T> make_plan(const nl_params* p, const Geo<T>& g, int64_t N) {
  Plan<T> pl;
  pl.path = PATH_GENERIC;
  pl.th = MaskThresholds{0.f, 0.f, 0.f, 0};
  pl.th_lazy = pl.th;
  pl.lazy_ok = 0;
  pl.ljf_ok = 0;
  pl.fill2_ok = 0;
  if (N <= 0) return pl;
  if (!tiled_applicable<T>(p, g, N, tile_cap<T>(), pl.ts_exact)) return pl;
  pl.path = PATH_TILED;
  const double dens = (double)N / (double)g.nct;
  if (27.0 * dens <= 400.0 && pick_tile<T>(g, N, cm_cap(CM_COUNT), pl.ts_lazy)) {  // candidate tables hold 512 entries
    pl.lazy_ok = 1;
    if (sizeof(T) == 8) {
      pl.th_lazy = mask_thresholds(p->cell, p->ncells, pl.ts_lazy, (double)g.cutoff_sq);
      pl.lazy_ok = pl.th_lazy.ok;
    }
  }
  if (sizeof(T) == 4 && 27.0 * dens <= 400.0 && pick_tile<T>(g, N, cm_cap(CM_LJF), pl.ts_ljf)) pl.ljf_ok = 1;
  if (27.0 * dens > 200.0) return pl;  // candidate lists would overflow the 256-bit masks too often
  if (!pick_tile<T>(g, N, CNT_CAP2, pl.ts_count) || !pick_tile<T>(g, N, fill_cap<T, TI>(), pl.ts_fill)) return pl;
  if (sizeof(T) == 8) {
    pl.th = mask_thresholds(p->cell, p->ncells, pl.ts_count, (double)g.cutoff_sq);
    if (!pl.th.ok) return pl;
  }
  pl.fill2_ok = pick_tile<T>(g, N, f2_cap<T, TI>(), pl.ts_fill2) ? 1 : 0;
  pl.path = PATH_MASK;
  return pl;
}

// Two fill kernels exist: the sorted-order, tile-staged k_fill_mask (default: 8.3 ms at the headline
// size, bound by randomly placed row writes) and the original-order, thread-per-pair k_fill_rows (10.8 ms, bound by
// random record gathers).  NL_FILL_ROWS=1 selects the latter for A/B measurements.
inline bool fill_tiled_requested() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NL_FILL_ROWS"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// NL_FILL=legacy selects round 1's k_fill_mask (every element written in place by the mask expansion) for A/B measurements;
// it is also what runs when an output pointer is not 32-byte aligned.
// NL_FILL=park selects the experiment variant of the round-2 kernel that writes only complete sectors in place and parks the row
// ends for k_fix_boundaries (nl_fill2.cuh; measured slower than the default: the parking costs more instructions than the
// read-modify-writes it removes).
inline int fill_variant() {  // 0 legacy k_fill_mask, 1 k_fill3 (default), 2 parked k_fill_park<true>, 3 lean in-place k_fill_park<false>
  static int v = -1;
  if (v < 0) { const char* e = getenv("NL_FILL"); v = !e ? 1 : (e[0] == 'l' ? 0 : (e[0] == 'p' ? 2 : (e[0] == '2' ? 3 : 1))); }
  return v;
}

// What nl_count_pairs leaves in the workspace for nl_fill_pairs* to check (include/nlcuda.h: NL_ERR_WORKSPACE).
struct WsStamp {
  unsigned long long magic;
  long long N, nct;
  unsigned long long total;
  int float_type, int_type, flags, windowed;
};
constexpr unsigned long long WS_MAGIC = 0x4e4c57535f523032ull;  // "NLWS_R02"

inline int fill_prefetch() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NL_FILL_PREFETCH"); v = !e ? 1 : (e[0] >= '0' && e[0] <= '4' ? e[0] - '0' : 1); }
  return v;
}

inline int count_variant() {  // 1: k_count_mask2 (default), 0: k_count_mask
  static int v = -1;
  if (v < 0) { const char* e = getenv("NL_COUNT"); v = (e && e[0] == 'l') ? 0 : 1; }
  return v;
}
inline bool fill_szero() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NL_FILL_SZERO"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// Opt-in to > 48 KB of dynamic shared memory.  The attribute is per DEVICE, so the "done" flags are per device too
// (one process may drive several GPUs); a benign race at worst sets it twice.
struct SmemOnce { bool done[64] = {}; };
template <class F> int set_smem_once(F* fn, int bytes, SmemOnce& once) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e);
  if (dev < 0 || dev >= 64 || !once.done[dev]) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e);
    if (dev >= 0 && dev < 64) once.done[dev] = true;
  }
  return NL_OK;
}

struct TiledScratch { uint32_t* masks; uint8_t* cellflag; };
TiledScratch tiled_scratch(void* base, int64_t N) {
  TiledScratch t;
  t.masks = (uint32_t*)((char*)base + 256);
  t.cellflag = (uint8_t*)((char*)base + 256 + al256((size_t)(N > 0 ? N : 1) * 32));
  return t;
}

// Slab shards (nl_*_window): only the tile layers along z that contain an active cell plane are launched.  The rest of a
// shard's view of the GLOBAL grid is empty; launching it cost 1.1 ms per list at 8 ranks (scripts/exp_slab_tiles.py).
template <class T, class TI>
int apply_plane_window(MaskArgs<T, TI>& a, const uint8_t* plane_active, int* zl_dev, unsigned& nblk, cudaStream_t st) {
  nblk = (unsigned)((long long)a.ntx * a.nty * a.ntz);
  if (!plane_active) return NL_OK;
  std::vector<int> layers;
  for (int k = 0; k < a.ntz; k++) {
    bool on = false;
    for (int z = k * a.tz; z < std::min((k + 1) * a.tz, a.g.nc[2]) && !on; z++) on = plane_active[z] != 0;
    if (on) layers.push_back(k);
  }
  if ((int)layers.size() == a.ntz) return NL_OK;
  if (layers.empty()) layers.push_back(0);
  NL_CUDA(cudaMemcpyAsync(zl_dev, layers.data(), layers.size() * sizeof(int), cudaMemcpyHostToDevice, st));  // pageable source: staged before return
  a.zlayers = zl_dev;
  nblk = (unsigned)((long long)a.ntx * a.nty * (long long)layers.size());
  return NL_OK;
}

// MODE_COUNT with want_mask (materialisation) or without (lazy count), MODE_FILL, MODE_LJ.
template <class T, class TI, int MODE>
int traverse(const nl_params* p, int64_t N, const void* co, const PairWs& w, const Geo<T>& g, const Sinks<T, TI>& sk, bool want_mask,
             cudaStream_t st, unsigned long long total_pairs = 0) {
  if (N <= 0) return NL_OK;
  Records<T> rec = records_of<T>(w);
  const Plan<T> pl = make_plan<T, TI>(p, g, N);
  constexpr bool LJ_FAST = MODE == MODE_LJ && sizeof(T) == 4;
  constexpr bool LJF_FAST = MODE == MODE_LJF && sizeof(T) == 4;
  if ((pl.lazy_ok && ((MODE == MODE_COUNT && !want_mask) || LJ_FAST)) || (pl.ljf_ok && LJF_FAST)) {
    // lazy sinks on the packed counting kernel: no masks, candidate lists up to 512
    constexpr int CM = MODE == MODE_LJ ? CM_LJ : (MODE == MODE_LJF ? CM_LJF : CM_COUNT);
    MaskArgs<T, TI> a;
    mask_args<T, TI>(a, N, (const TI*)co, rec, g, sk, LJF_FAST ? pl.ts_ljf : pl.ts_lazy, nullptr);
    a.mid = pl.th_lazy.mid; a.hw = pl.th_lazy.hw; a.dguard = pl.th_lazy.dguard;
    a.self = (const MaskArgs<T, TI>*)((char*)w.hdr + 3072);
    NL_CUDA(cudaMemcpyAsync((void*)a.self, &a, sizeof(a), cudaMemcpyHostToDevice, st));
    const unsigned nblk = (unsigned)((long long)a.ntx * a.nty * a.ntz);
    if constexpr (CM == CM_COUNT || LJ_FAST || LJF_FAST) {
      static SmemOnce done;
      int rc = set_smem_once(k_count_mask<T, TI, CM>, cm_smem_bytes(CM), done);
      if (rc) return rc;
      k_count_mask<T, TI, CM><<<nblk, TILE_NT, cm_smem_bytes(CM), st>>>(a);
    }
    NL_LAUNCHED(1);
  } else if (pl.path == PATH_MASK && (MODE == MODE_COUNT || MODE == MODE_FILL)) {
    TiledScratch tsx = tiled_scratch(w.tiled, N);
    MaskArgs<T, TI> a;
    const bool out_aligned = MODE == MODE_FILL && (((uintptr_t)sk.io | (uintptr_t)sk.jo | (uintptr_t)sk.So | (uintptr_t)sk.Ro) & 31) == 0;
    const int fvar = fill_variant();
    const bool use_park = MODE == MODE_FILL && pl.fill2_ok && out_aligned && fvar != 0 && fill_tiled_requested();
    mask_args<T, TI>(a, N, (const TI*)co, rec, g, sk, MODE == MODE_FILL ? (use_park ? pl.ts_fill2 : pl.ts_fill) : pl.ts_count, tsx.masks);
    a.cellflag = tsx.cellflag;
    a.mid = pl.th.mid; a.hw = pl.th.hw; a.dguard = pl.th.dguard;
    unsigned nblk = 0;
    {
      int rcw = apply_plane_window<T, TI>(a, sk.plane_active, w.zlayers, nblk, st);
      if (rcw) return rcw;
    }
    static_assert(sizeof(MaskArgs<T, TI>) <= 1024, "argument block");
    a.self = (const MaskArgs<T, TI>*)((char*)w.hdr + (MODE == MODE_FILL ? 2048 : (want_mask ? 0 : 3072)));
    if (MODE == MODE_FILL && !fill_tiled_requested()) {
      // original-order, thread-per-pair fill (nl_fillrows.cuh)
      k_fillrows_prologue<T, TI><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(w.pidx, sk.gmap, N, w.sorted_of, (RecAoS<T>*)w.ra);
      NL_LAUNCHED(1);
      FillRowsArgs<T, TI> fa;
      fa.ra = (const RecAoS<T>*)w.ra; fa.pkey = w.pkey; fa.sorted_of = w.sorted_of; fa.masks = tsx.masks; fa.cellflag = tsx.cellflag;
      fa.co = (const TI*)co; fa.rec = rec; fa.g = g; fa.out = sk; fa.n = N;
      static_assert(sizeof(FillRowsArgs<T, TI>) <= 1024, "argument block");
      fa.self = (const FillRowsArgs<T, TI>*)((char*)w.hdr + 1024);
      NL_CUDA(cudaMemcpyAsync((void*)fa.self, &fa, sizeof(fa), cudaMemcpyHostToDevice, st));
      if (sk.n_rows > 0) {
        k_fill_rows<T, TI><<<(unsigned)((sk.n_rows + FR_RB - 1) / FR_RB), FR_NT, 0, st>>>(fa);
        NL_LAUNCHED(1);
      }
      NL_LAUNCH_CHECK();
      return NL_OK;
    } else if (MODE == MODE_FILL && use_park) {
      // round-2 fill: complete sectors in place + parked row ends (k_fill_park), boundary sectors in original row order
      // (k_fix_boundaries), i stream from first[] alone (k_expand_rows)
      static SmemOnce done, done_p;
      int rc = fvar == 2 ? set_smem_once(k_fill_park<T, TI, true>, F2_SMEM_BYTES, done_p) : set_smem_once(k_fill_park<T, TI, false>, F2_SMEM_BYTES, done);
      if (rc) return rc;
      a.srow = w.srow;
      k_row_starts<T, TI><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(w.pidx, sk.first, N, sk.n_rows, (typename FillBase<TI>::type*)w.srow);
      NL_CUDA(cudaMemcpyAsync((void*)a.self, &a, sizeof(a), cudaMemcpyHostToDevice, st));
      // the i stream is a pure function of first[]; the same streaming kernel zeroes S up front (NL_FILL_SZERO=0: off), so that
      // the expansion only writes the S rows that cross a periodic boundary
      const bool szero = fvar != 2 && fill_szero();
      const bool expand = total_pairs > 0 && sk.n_rows > 0;
      if (expand && szero)
        k_expand_rows<TI><<<(unsigned)((sk.n_rows + EXP_RB - 1) / EXP_RB), EXP_NT, 0, st>>>(sk.first, sk.n_rows, sk.gmap, sk.io, sk.So);
      if (fvar == 2) {
        k_fill_park<T, TI, true><<<nblk, F2_NT, F2_SMEM_BYTES, st>>>(a, w.parkA, w.parkR, 0, 0);
        k_fix_boundaries<T, TI><<<(unsigned)((sk.n_rows + 255) / 256), 256, 0, st>>>(sk.first, sk.n_rows, sk.jo, sk.So, sk.Ro, w.parkA, w.parkR);
        NL_LAUNCHED(1);
      } else if (expand && szero && fvar == 1) {
        // default: plain / general row bodies over the pre-zeroed S stream (nl_fill3.cuh); NL_FILL=2 keeps k_fill_park for A/B
        static SmemOnce done3, done3n;
        if (sk.Ro) {
          rc = set_smem_once(k_fill3<T, TI, true>, F2_SMEM_BYTES, done3);
          if (rc) return rc;
          k_fill3<T, TI, true><<<nblk, F2_NT, F2_SMEM_BYTES, st>>>(a, fill_prefetch());
        } else {
          rc = set_smem_once(k_fill3<T, TI, false>, F2_SMEM_BYTES, done3n);
          if (rc) return rc;
          k_fill3<T, TI, false><<<nblk, F2_NT, F2_SMEM_BYTES, st>>>(a, fill_prefetch());
        }
      } else {
        k_fill_park<T, TI, false><<<nblk, F2_NT, F2_SMEM_BYTES, st>>>(a, nullptr, nullptr, fill_prefetch(), expand && szero ? 1 : 0);
      }
      if (expand && !szero)
        k_expand_rows<TI><<<(unsigned)((sk.n_rows + EXP_RB - 1) / EXP_RB), EXP_NT, 0, st>>>(sk.first, sk.n_rows, sk.gmap, sk.io, nullptr);
      NL_LAUNCHED(2);
    } else if (MODE == MODE_FILL) {
      static SmemOnce done;
      int rc = set_smem_once(k_fill_mask<T, TI>, FILL_SMEM_BYTES, done);
      if (rc) return rc;
      // row starts in SORTED order: the N random gathers of first[] run here at full memory-level parallelism instead of
      // inside the (low-occupancy) fill kernel's staging loop
      a.srow = w.srow;
      k_row_starts<T, TI><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(w.pidx, sk.first, N, sk.n_rows, (typename FillBase<TI>::type*)w.srow);
      NL_LAUNCHED(1);
      NL_CUDA(cudaMemcpyAsync((void*)a.self, &a, sizeof(a), cudaMemcpyHostToDevice, st));
      k_fill_mask<T, TI><<<nblk, TILE_NT, FILL_SMEM_BYTES, st>>>(a);
    } else {
      // MODE_COUNT: with masks for the fill pass, or (lazy count on a problem the lazy plan rejected) without
      NL_CUDA(cudaMemcpyAsync((void*)a.self, &a, sizeof(a), cudaMemcpyHostToDevice, st));
      bool launched = false;
      if constexpr (sizeof(T) == 8) {
        // Float64 lists: candidates register-resident, home-atom pairs in the outer loop (nl_count2.cuh);
        // NL_COUNT=legacy keeps round 1's chunk-major kernel for A/B measurements
        if (count_variant() == 1) {
          static SmemOnce done2, done2h;
          int rc = sk.half ? set_smem_once(k_count_mask2<TI, true>, cm_smem_bytes(CM_MASK), done2h)
                           : set_smem_once(k_count_mask2<TI, false>, cm_smem_bytes(CM_MASK), done2);
          if (rc) return rc;
          if (sk.half) k_count_mask2<TI, true><<<nblk, TILE_NT, cm_smem_bytes(CM_MASK), st>>>(a);   // the half rule filters the hit words
          else k_count_mask2<TI, false><<<nblk, TILE_NT, cm_smem_bytes(CM_MASK), st>>>(a);
          launched = true;
        }
      }
      if (!launched) {
        static SmemOnce done;
        int rc = set_smem_once(k_count_mask<T, TI, CM_MASK>, cm_smem_bytes(CM_MASK), done);
        if (rc) return rc;
        k_count_mask<T, TI, CM_MASK><<<nblk, TILE_NT, cm_smem_bytes(CM_MASK), st>>>(a);
      }
    }
    NL_LAUNCHED(1);
  } else if (pl.path != PATH_GENERIC) {
    int rc = tiled_traverse<T, TI, MODE>(p, N, (const TI*)co, rec, g, sk, pl.ts_exact, w.tiled, st);
    if (rc != NL_OK) return rc;
  } else {
    k_traverse_generic<T, TI, MODE><<<(unsigned)((N + 127) / 128), 128, 0, st>>>(rec, (const TI*)co, N, g, sk);
    NL_LAUNCHED(1);
  }
  NL_LAUNCH_CHECK();
  return NL_OK;
}

// Windowed calls: the caller promises that every z plane of cells it did not mark active is empty.  A broken promise would
// leave atoms without counts (their tile layers are never launched), so it is checked against cell_offsets: flag = 1.
template <class TI>
__global__ void k_check_planes(const TI* __restrict__ co, long long nxy, int nz, const uint8_t* __restrict__ active, unsigned long long* __restrict__ flag) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z < nz && !active[z] && co[(long long)(z + 1) * nxy] != co[(long long)z * nxy]) *flag = 1ull;
}

template <class T, class TI>
int count_pairs_impl(const nl_params* p, const void* Xs, int64_t N, const void* perm, const void* co, void* first, int64_t* total_host,
                     void* ws, const uint8_t* plane_active, cudaStream_t st) {
  Geo<T> g = make_geo<T>(p);
  PairWs w = pair_ws(ws, p, N);
  unsigned long long tot2[2] = {0, 0};   // pair total | "plane_active promise broken" flag
  if (N > 0) {
    int rc = prep_impl<T, TI>(p, Xs, N, perm, w, g, st);
    if (rc) return rc;
    Sinks<T, TI> sk = {};
    sk.counts = w.counts;
    sk.half = p->reserved[0] & NL_FLAG_HALF;
    sk.plane_active = plane_active;
    NL_CUDA(cudaMemsetAsync(w.total, 0, 16, st));
    if (plane_active) {
      NL_CUDA(cudaMemsetAsync(w.counts, 0, (size_t)N * 4, st));  // atoms outside the launched layers (there must be none) count zero
      NL_CUDA(cudaMemcpyAsync(w.planes, plane_active, (size_t)g.nc[2], cudaMemcpyHostToDevice, st));
      k_check_planes<TI><<<(unsigned)((g.nc[2] + 255) / 256), 256, 0, st>>>((const TI*)co, (long long)g.nc[0] * g.nc[1], g.nc[2], w.planes, w.total + 1);
      NL_LAUNCHED(1);
    }
    rc = traverse<T, TI, MODE_COUNT>(p, N, co, w, g, sk, true, st);
    if (rc) return rc;
    exclusive_scan<uint32_t, unsigned long long, TI>(w.counts, N, (TI*)first, 1ull, true, w.tsum, w.total, st);
    NL_LAUNCH_CHECK();
    NL_CUDA(cudaMemcpyAsync(tot2, w.total, sizeof(tot2), cudaMemcpyDeviceToHost, st));
  } else {
    TI one = 1;
    NL_CUDA(cudaMemcpyAsync(first, &one, sizeof(TI), cudaMemcpyHostToDevice, st));
  }
  NL_CUDA(cudaStreamSynchronize(st));
  const unsigned long long total = tot2[0];
  *total_host = (int64_t)total;
  if (tot2[1]) return NL_ERR_BAD_ARG;  // atoms in a cell plane the caller declared empty
  if (sizeof(TI) == 4 && total + 1 > 2147483647ull) return NL_ERR_OVERFLOW;
  if (N > 0) {
    WsStamp s = {WS_MAGIC, (long long)N, (long long)g.nct, total, p->float_type, p->int_type, p->reserved[0], plane_active ? 1 : 0};
    NL_CUDA(cudaMemcpyAsync(w.stamp, &s, sizeof(s), cudaMemcpyHostToDevice, st));  // pageable source: staged before the call returns
  }
  return NL_OK;
}

template <class T, class TI>
int fill_pairs_impl(const nl_params* p, int64_t N, const void* co, const void* first, void* io, void* jo, void* So, void* Ro, void* ws,
                    int64_t n_rows, const void* gmap, const uint8_t* plane_active, cudaStream_t st) {
  Geo<T> g = make_geo<T>(p);
  PairWs w = pair_ws(ws, p, N);
  Sinks<T, TI> sk = {};
  sk.first = (const TI*)first;
  sk.io = (TI*)io; sk.jo = (TI*)jo; sk.So = (TI*)So; sk.Ro = (T*)Ro;
  sk.n_rows = n_rows; sk.gmap = (const TI*)gmap;
  sk.half = p->reserved[0] & NL_FLAG_HALF;
  sk.plane_active = plane_active;
  // the workspace must be the one nl_count_pairs filled for this very problem (masks, counts, records live in it)
  WsStamp s = {};
  NL_CUDA(cudaMemcpyAsync(&s, w.stamp, sizeof(s), cudaMemcpyDeviceToHost, st));
  NL_CUDA(cudaStreamSynchronize(st));
  if (s.magic != WS_MAGIC || s.N != (long long)N || s.nct != (long long)g.nct || s.float_type != p->float_type || s.int_type != p->int_type ||
      s.flags != p->reserved[0] || s.windowed != (plane_active ? 1 : 0))
    return NL_ERR_WORKSPACE;
  if (gmap && N > 0) {
    k_make_pgid<TI><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(w.pidx, (const TI*)gmap, N, w.pgid0);
    NL_LAUNCHED(1);
    NL_LAUNCH_CHECK();
    sk.pgid0 = w.pgid0;
  }
  return traverse<T, TI, MODE_FILL>(p, N, co, w, g, sk, true, st, s.total);
}

template <class TI> __global__ void k_counts_to_ti(const uint32_t* __restrict__ c, long long n, TI* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (TI)c[i];
}

template <class T, class TI>
int lazy_count_impl(const nl_params* p, const void* Xs, int64_t N, const void* perm, const void* co, void* counts_out, void* ws,
                    cudaStream_t st) {
  if (N <= 0) return NL_OK;
  Geo<T> g = make_geo<T>(p);
  PairWs w = pair_ws(ws, p, N);
  int rc = prep_impl<T, TI>(p, Xs, N, perm, w, g, st);
  if (rc) return rc;
  Sinks<T, TI> sk = {};
  sk.counts = w.counts;
  rc = traverse<T, TI, MODE_COUNT>(p, N, co, w, g, sk, false, st);
  if (rc) return rc;
  k_counts_to_ti<TI><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(w.counts, N, (TI*)counts_out);
  NL_LAUNCHED(1);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

template <class T, class TI>
int lazy_lj_impl(const nl_params* p, const void* Xs, int64_t N, const void* perm, const void* co, double eps, double sigma, double* e_out,
                 void* ws, cudaStream_t st) {
  NL_CUDA(cudaMemsetAsync(e_out, 0, sizeof(double), st));
  if (N <= 0) return NL_OK;
  Geo<T> g = make_geo<T>(p);
  PairWs w = pair_ws(ws, p, N);
  int rc = prep_impl<T, TI>(p, Xs, N, perm, w, g, st);
  if (rc) return rc;
  Sinks<T, TI> sk = {};
  sk.energy = e_out;
  sk.lj_eps = eps;
  sk.lj_sigma2 = sigma * sigma;
  return traverse<T, TI, MODE_LJ>(p, N, co, w, g, sk, false, st);
}

// ---------------------------------------------------------------- accessors / adapters (nl_access.cuh)
template <class T, class TI>
int pairs_R_impl(const nl_params* p, const void* X, const void* i, const void* j, const void* S, int64_t p_lo, int64_t p_hi, void* R,
                 cudaStream_t st) {
  const long long np = p_hi - p_lo;
  if (np <= 0) return NL_OK;
  Geo<T> g = make_geo<T>(p);
  k_pairs_R<T, TI><<<(unsigned)((np + 255) / 256), 256, 0, st>>>((const T*)X, (const TI*)i, (const TI*)j, (const TI*)S, p_lo, np, g, (T*)R);
  NL_LAUNCHED(1);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

template <class T, class TI>
int rows_padded_impl(const nl_params* p, const void* X, const void* first, const void* j, const void* S, const void* rows, int64_t n_sel,
                     int32_t width, void* n_out, void* j_out, void* S_out, void* R_out, cudaStream_t st) {
  Geo<T> g = make_geo<T>(p);
  k_rows_padded<T, TI><<<(unsigned)((n_sel + 7) / 8), 256, 0, st>>>((const T*)X, (const TI*)first, (const TI*)j, (const TI*)S, (const TI*)rows,
                                                                     n_sel, width, g, (TI*)n_out, (TI*)j_out, (TI*)S_out, (T*)R_out);
  NL_LAUNCHED(1);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

template <class T, class TI>
int lazy_neighbours_impl(const nl_params* p, const void* Xo, const void* Xs, const void* perm, const void* co, const void* atoms, int64_t n_sel,
                         int32_t width, void* n_out, void* j_out, void* S_out, void* R_out, cudaStream_t st) {
  Geo<T> g = make_geo<T>(p);
  k_lazy_neighbours<T, TI><<<(unsigned)((n_sel + 7) / 8), 256, 0, st>>>((const T*)Xo, (const T*)Xs, (const TI*)perm, (const TI*)co, g,
                                                                         (const TI*)atoms, n_sel, width, (TI*)n_out, (TI*)j_out, (TI*)S_out,
                                                                         (T*)R_out);
  NL_LAUNCHED(1);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

template <class T> int bbox_impl(const void* X, int64_t N, void* out, void* ws, cudaStream_t st) {
  const int nb = (int)std::min<long long>(RED_BLOCKS, (N + 255) / 256);
  k_bbox_partial<T><<<nb, 256, 0, st>>>((const T*)X, N, (T*)ws);
  k_bbox_final<T><<<1, 256, 0, st>>>((const T*)ws, nb, (T*)out);
  NL_LAUNCHED(2);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

template <class T> int maxdisp_impl(const void* X, const void* Y, int64_t N, void* out, void* ws, cudaStream_t st) {
  const int nb = (int)std::max<long long>(1, std::min<long long>(RED_BLOCKS, (N + 255) / 256));
  k_maxdisp_partial<T><<<nb, 256, 0, st>>>((const T*)X, (const T*)Y, N, (T*)ws);
  k_max_final<T><<<1, 256, 0, st>>>((const T*)ws, nb, (T*)out);
  NL_LAUNCHED(2);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

template <class T, class TI>
int lazy_ljf_impl(const nl_params* p, const void* Xs, int64_t N, const void* perm, const void* co, double eps, double sigma, void* fe_out,
                  void* ws, cudaStream_t st) {
  if (N <= 0) return NL_OK;
  NL_CUDA(cudaMemsetAsync(fe_out, 0, (size_t)N * 4 * sizeof(T), st));
  Geo<T> g = make_geo<T>(p);
  PairWs w = pair_ws(ws, p, N);
  int rc = prep_impl<T, TI>(p, Xs, N, perm, w, g, st);
  if (rc) return rc;
  Sinks<T, TI> sk = {};
  sk.fe = (T*)fe_out;
  sk.lj_eps = eps;
  sk.lj_sigma2 = sigma * sigma;
  return traverse<T, TI, MODE_LJF>(p, N, co, w, g, sk, false, st);
}

// ---------------------------------------------------------------- multi-GPU slabs (nl_shard.cuh)
#define NL_NCCL(expr)                         \
  do {                                        \
    if ((expr) != 0) return NL_ERR_NCCL;      \
  } while (0)

inline int slab_axis(const nl_params* p) {  // most planes; ties go to the SLOWEST key axis (z), so that a slab is one range of keys
  int axis = 2;
  if (p->ncells[1] > p->ncells[axis]) axis = 1;
  if (p->ncells[0] > p->ncells[axis]) axis = 0;
  return axis;
}

template <class T, class TI>
int shard_prepare_impl(const nl_params* p, const void* X, int64_t n, void* comm, int rank, int nranks, nl_shard_info* info, void* ws,
                       cudaStream_t st) {
  Geo<T> g = make_geo<T>(p);
  const int axis = slab_axis(p), nplanes = p->ncells[axis], halo = p->nxyz[axis];
  ShardWs w = shard_ws(ws, n, nplanes, nranks, sizeof(T), sizeof(TI));
  NL_CUDA(cudaMemsetAsync(w.hist_local, 0, (size_t)nplanes * 8, st));
  if (n > 0) {
    const unsigned nb = (unsigned)std::min<long long>((n + 255) / 256, 148 * 8);
    k_shard_planes<T><<<nb, 256, 0, st>>>((const T*)X, n, g, axis, nplanes, w.planes, w.hist_local);
    NL_LAUNCHED(1);
    NL_LAUNCH_CHECK();
  }
  if (nranks > 1) {
    if (!nccl().ok || !comm) return NL_ERR_NCCL;
    NL_NCCL(nccl().AllGather(w.hist_local, w.hist_all, (size_t)nplanes, NCCL_UINT64, comm, st));
  } else {
    NL_CUDA(cudaMemcpyAsync(w.hist_all, w.hist_local, (size_t)nplanes * 8, cudaMemcpyDeviceToDevice, st));
  }
  std::vector<unsigned long long> h((size_t)nplanes * nranks);
  NL_CUDA(cudaMemcpyAsync(h.data(), w.hist_all, h.size() * 8, cudaMemcpyDeviceToHost, st));
  NL_CUDA(cudaStreamSynchronize(st));
  std::vector<int64_t> tot(nplanes, 0);
  for (int r = 0; r < nranks; r++)
    for (int q = 0; q < nplanes; q++) tot[q] += (int64_t)h[(size_t)r * nplanes + q];
  nl_shard_info& f = *info;
  f = nl_shard_info{};
  f.axis = axis; f.halo = halo; f.periodic = p->pbc[axis] ? 1 : 0; f.nranks = nranks; f.rank = rank; f.nplanes = nplanes;
  f.n_local = n;
  int rc = nl_shard_plan(tot.data(), nplanes, nranks, halo, f.bounds);
  if (rc) return rc;
  auto sum = [&](const unsigned long long* row, int64_t a, int64_t b) { int64_t s = 0; for (int64_t q = a; q < b; q++) s += (int64_t)row[q]; return s; };
  auto sumt = [&](int64_t a, int64_t b) { int64_t s = 0; for (int64_t q = a; q < b; q++) s += tot[q]; return s; };
  for (int d = 0; d < nranks; d++) f.send_count[d] = sum(&h[(size_t)rank * nplanes], f.bounds[d], f.bounds[d + 1]);
  for (int r = 0; r < nranks; r++) f.recv_count[r] = sum(&h[(size_t)r * nplanes], f.bounds[rank], f.bounds[rank + 1]);
  f.n_owned = sumt(f.bounds[rank], f.bounds[rank + 1]);
  f.up_peer = rank + 1; f.dn_peer = rank - 1;
  if (f.periodic) { f.up_peer = (rank + 1) % nranks; f.dn_peer = (rank + nranks - 1) % nranks; }
  f.has_up = nranks > 1 && f.up_peer >= 0 && f.up_peer < nranks;
  f.has_dn = nranks > 1 && f.dn_peer >= 0 && f.dn_peer < nranks;
  const int64_t lo = f.bounds[rank], hi = f.bounds[rank + 1];
  if (f.has_dn) { f.n_send_dn = sumt(lo, lo + halo); f.n_halo_dn = sumt(f.bounds[f.dn_peer + 1] - halo, f.bounds[f.dn_peer + 1]); }
  if (f.has_up) { f.n_send_up = sumt(hi - halo, hi); f.n_halo_up = sumt(f.bounds[f.up_peer], f.bounds[f.up_peer] + halo); }
  // what the peer path needs to know about the other ranks (all from the gathered histograms: the same on every rank)
  for (int r = 0; r < nranks; r++) {
    const unsigned long long* row = &h[(size_t)r * nplanes];
    f.n_max_all = std::max(f.n_max_all, std::max(sum(row, 0, nplanes), sumt(f.bounds[r], f.bounds[r + 1])));
    f.src_offset[r] = sum(row, 0, f.bounds[rank]);  // slabs are contiguous plane ranges in rank order
  }
  // the halo from below is the dn peer's UP block, which follows its DOWN block; the halo from above is the up peer's DOWN block
  if (f.has_dn) {
    const int q = f.dn_peer;
    const bool q_has_dn = f.periodic || q > 0;
    f.halo_src_offset_dn = q_has_dn ? sumt(f.bounds[q], f.bounds[q] + halo) : 0;
  }
  f.halo_src_offset_up = 0;
  return NL_OK;
}

// NL_SHARD_PROFILE=1: device time of the phases of nl_shard_exchange, printed by rank 0 (synchronises; measurement only).
struct PhaseTimer {
  bool on;
  cudaStream_t st;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  PhaseTimer(bool enable, cudaStream_t s) : on(enable), st(s) {}
  void mark(const char* name) {
    if (!on) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
    cudaEventRecord(e, st);
    marks.emplace_back(name, e);
  }
  void report(const char* what) {
    if (!on || marks.size() < 2) return;
    cudaStreamSynchronize(st);
    fprintf(stderr, "[%s ms]", what);
    for (size_t k = 1; k < marks.size(); k++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, marks[k - 1].second, marks[k].second);
      fprintf(stderr, " %s=%.3f", marks[k].first, ms);
    }
    fprintf(stderr, "\n");
    for (auto& m : marks) cudaEventDestroy(m.second);
  }
};
inline bool shard_profile() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NL_SHARD_PROFILE"); v = (e && e[0] && e[0] != '0') ? 1 : 0; }
  return v == 1;
}

template <class T, class TI>
int shard_exchange_impl(const nl_params* p, const nl_shard_info* info, const void* X, const void* gidx, int64_t n, void* comm,
                        const nl_shard_peers* peers, void* X_all, void* g_all, uint8_t* plane_active_out, void* ws, cudaStream_t st) {
  const nl_shard_info& f = *info;
  Geo<T> g = make_geo<T>(p);
  const int G = f.nranks, me = f.rank;
  // peer path: every rank lays its workspace out for the SAME capacity, so that a peer's send buffers can be addressed
  const int64_t n_max = peers ? peers->cap : std::max<int64_t>(n, f.n_owned);
  ShardWs w = shard_ws(ws, n_max, f.nplanes, G, sizeof(T), sizeof(TI));
  auto peer_view = [&](int r) { return shard_ws(peers->peer_ws[r], n_max, f.nplanes, G, sizeof(T), sizeof(TI)); };
  auto barrier = [&]() { return nccl().AllGather(w.bar, w.bar + 1, 1, NCCL_UINT64, comm, st); };
  T* Xa = (T*)X_all;
  TI* ga = (TI*)g_all;
  if (plane_active_out) {
    const int nz = p->ncells[2];
    if (f.axis != 2 || G == 1) {
      for (int z = 0; z < nz; z++) plane_active_out[z] = 1;
    } else {
      for (int z = 0; z < nz; z++) plane_active_out[z] = 0;
      for (int64_t z = f.bounds[me] - f.halo; z < f.bounds[me + 1] + f.halo; z++) {
        if (z >= 0 && z < nz) plane_active_out[z] = 1;
        else if (f.periodic) plane_active_out[((z % nz) + nz) % nz] = 1;
      }
    }
  }
  if (G == 1) {
    if (n > 0) {
      NL_CUDA(cudaMemcpyAsync(Xa, X, (size_t)n * 3 * sizeof(T), cudaMemcpyDeviceToDevice, st));
      NL_CUDA(cudaMemcpyAsync(ga, gidx, (size_t)n * sizeof(TI), cudaMemcpyDeviceToDevice, st));
    }
    return NL_OK;
  }
  if (!nccl().ok || !comm) return NL_ERR_NCCL;
  PhaseTimer pt(shard_profile() && me == 0, st);
  pt.mark("start");
  // ---- owners, stable partition by destination
  const uint32_t* order = nullptr;
  if (n > 0) {
    NL_CUDA(cudaMemsetAsync(w.hist_local, 0, (size_t)f.nplanes * 8, st));
    k_shard_planes<T><<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, st>>>((const T*)X, n, g, f.axis, f.nplanes, w.planes, w.hist_local);
    NL_CUDA(cudaMemcpyAsync(w.bounds, f.bounds, (size_t)(G + 1) * 8, cudaMemcpyHostToDevice, st));
    k_shard_owner<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w.planes, n, w.bounds, G, w.keyA);
    const int where = radix_sort_pairs(w.keyA, w.valA, w.keyB, w.valB, n, key_bits(G), w.rs_scratch, st);
    order = where ? w.valB : w.valA;
    k_shard_gather<T, TI><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(order, 0, n, (const T*)X, (const TI*)gidx, w.planes, (T*)w.sendX, (TI*)w.sendg, w.sendp);
    NL_LAUNCHED(3);
    NL_LAUNCH_CHECK();
  }
  pt.mark("partition");
  // ---- all-to-all-v straight into the local arrays: [atoms that stay | from rank 0 | from rank 1 | ...]
  std::vector<int64_t> soff(G + 1, 0), roff(G, 0);
  for (int d = 0; d < G; d++) soff[d + 1] = soff[d] + f.send_count[d];
  {
    int64_t o = f.send_count[me];
    for (int s = 0; s < G; s++) if (s != me) { roff[s] = o; o += f.recv_count[s]; }
    if (o != f.n_owned || soff[G] != n) return NL_ERR_BAD_ARG;  // info does not belong to these atoms
  }
  if (f.send_count[me] > 0) {
    const int64_t k0 = soff[me], c = f.send_count[me];
    NL_CUDA(cudaMemcpyAsync(Xa, (const T*)w.sendX + 3 * k0, (size_t)c * 3 * sizeof(T), cudaMemcpyDeviceToDevice, st));
    NL_CUDA(cudaMemcpyAsync(ga, (const TI*)w.sendg + k0, (size_t)c * sizeof(TI), cudaMemcpyDeviceToDevice, st));
    NL_CUDA(cudaMemcpyAsync(w.planes_owned, w.sendp + k0, (size_t)c * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (peers) {
    // every send buffer is ready once this (stream-ordered) barrier has completed; then PULL my blocks out of the peers' buffers
    NL_NCCL(barrier());
    for (int k = 1; k < G; k++) {
      const int r = (me + k) % G;  // staggered: at any time every rank reads from a different peer
      const int64_t c = f.recv_count[r];
      if (c <= 0) continue;
      const ShardWs pw = peer_view(r);
      const int64_t s0 = f.src_offset[r], k0 = roff[r];
      NL_CUDA(cudaMemcpyAsync(Xa + 3 * k0, (const T*)pw.sendX + 3 * s0, (size_t)c * 3 * sizeof(T), cudaMemcpyDefault, st));
      NL_CUDA(cudaMemcpyAsync(ga + k0, (const TI*)pw.sendg + s0, (size_t)c * sizeof(TI), cudaMemcpyDefault, st));
      NL_CUDA(cudaMemcpyAsync(w.planes_owned + k0, pw.sendp + s0, (size_t)c * 4, cudaMemcpyDefault, st));
    }
  } else {
    bool any = false;
    for (int r = 0; r < G; r++) any = any || (r != me && (f.send_count[r] > 0 || f.recv_count[r] > 0));
    if (any) {
      NL_NCCL(nccl().GroupStart());
      for (int r = 0; r < G; r++) {
        if (r == me) continue;
        if (f.send_count[r] > 0) {
          const int64_t k0 = soff[r], c = f.send_count[r];
          NL_NCCL(nccl().Send((const T*)w.sendX + 3 * k0, (size_t)c * 3 * sizeof(T), NCCL_INT8, r, comm, st));
          NL_NCCL(nccl().Send((const TI*)w.sendg + k0, (size_t)c * sizeof(TI), NCCL_INT8, r, comm, st));
          NL_NCCL(nccl().Send(w.sendp + k0, (size_t)c * 4, NCCL_INT8, r, comm, st));
        }
        if (f.recv_count[r] > 0) {
          const int64_t k0 = roff[r], c = f.recv_count[r];
          NL_NCCL(nccl().Recv(Xa + 3 * k0, (size_t)c * 3 * sizeof(T), NCCL_INT8, r, comm, st));
          NL_NCCL(nccl().Recv(ga + k0, (size_t)c * sizeof(TI), NCCL_INT8, r, comm, st));
          NL_NCCL(nccl().Recv(w.planes_owned + k0, (size_t)c * 4, NCCL_INT8, r, comm, st));
        }
      }
      NL_NCCL(nccl().GroupEnd());
    }
  }
  pt.mark("all-to-all");
  // ---- halos: bottom planes to the rank below, top planes to the rank above
  const int64_t no = f.n_owned, ns = f.n_send_dn + f.n_send_up;
  if (no > 0 && ns > 0) {
    k_shard_halo_class<<<(unsigned)((no + 255) / 256), 256, 0, st>>>(w.planes_owned, no, f.bounds[me], f.bounds[me + 1], f.halo, f.has_dn, f.has_up, w.keyA);
    const int where = radix_sort_pairs(w.keyA, w.valA, w.keyB, w.valB, no, 2, w.rs_scratch, st);
    const uint32_t* order2 = where ? w.valB : w.valA;
    // peer path: a separate halo buffer -- the peers may still be pulling their all-to-all blocks out of sendX / sendg
    k_shard_gather<T, TI><<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(order2, 0, ns, Xa, ga, nullptr, (T*)(peers ? w.haloX : w.sendX),
                                                                        (TI*)(peers ? w.halog : w.sendg), nullptr);
    NL_LAUNCHED(2);
    NL_LAUNCH_CHECK();
  }
  pt.mark("halo-select");
  if (peers) {
    // second barrier: every halo buffer is ready (and, stream-ordered, every rank has finished pulling from sendX / sendg, so the
    // next list may overwrite them; its own first barrier protects the halo buffers in the same way)
    NL_NCCL(barrier());
    if (f.has_dn && f.n_halo_dn > 0) {
      const ShardWs pw = peer_view(f.dn_peer);
      NL_CUDA(cudaMemcpyAsync(Xa + 3 * no, (const T*)pw.haloX + 3 * f.halo_src_offset_dn, (size_t)f.n_halo_dn * 3 * sizeof(T), cudaMemcpyDefault, st));
      NL_CUDA(cudaMemcpyAsync(ga + no, (const TI*)pw.halog + f.halo_src_offset_dn, (size_t)f.n_halo_dn * sizeof(TI), cudaMemcpyDefault, st));
    }
    if (f.has_up && f.n_halo_up > 0) {
      const ShardWs pw = peer_view(f.up_peer);
      NL_CUDA(cudaMemcpyAsync(Xa + 3 * (no + f.n_halo_dn), (const T*)pw.haloX + 3 * f.halo_src_offset_up, (size_t)f.n_halo_up * 3 * sizeof(T), cudaMemcpyDefault, st));
      NL_CUDA(cudaMemcpyAsync(ga + no + f.n_halo_dn, (const TI*)pw.halog + f.halo_src_offset_up, (size_t)f.n_halo_up * sizeof(TI), cudaMemcpyDefault, st));
    }
  } else if (ns > 0 || f.n_halo_dn > 0 || f.n_halo_up > 0) {
    // message order matters when both neighbours are the same rank (2 ranks, periodic): everyone sends [up, down] and
    // receives [from below, from above], so the k-th send to a peer meets its k-th receive
    NL_NCCL(nccl().GroupStart());
    if (f.has_up && f.n_send_up > 0) {
      NL_NCCL(nccl().Send((const T*)w.sendX + 3 * f.n_send_dn, (size_t)f.n_send_up * 3 * sizeof(T), NCCL_INT8, f.up_peer, comm, st));
      NL_NCCL(nccl().Send((const TI*)w.sendg + f.n_send_dn, (size_t)f.n_send_up * sizeof(TI), NCCL_INT8, f.up_peer, comm, st));
    }
    if (f.has_dn && f.n_send_dn > 0) {
      NL_NCCL(nccl().Send((const T*)w.sendX, (size_t)f.n_send_dn * 3 * sizeof(T), NCCL_INT8, f.dn_peer, comm, st));
      NL_NCCL(nccl().Send((const TI*)w.sendg, (size_t)f.n_send_dn * sizeof(TI), NCCL_INT8, f.dn_peer, comm, st));
    }
    if (f.has_dn && f.n_halo_dn > 0) {
      NL_NCCL(nccl().Recv(Xa + 3 * no, (size_t)f.n_halo_dn * 3 * sizeof(T), NCCL_INT8, f.dn_peer, comm, st));
      NL_NCCL(nccl().Recv(ga + no, (size_t)f.n_halo_dn * sizeof(TI), NCCL_INT8, f.dn_peer, comm, st));
    }
    if (f.has_up && f.n_halo_up > 0) {
      NL_NCCL(nccl().Recv(Xa + 3 * (no + f.n_halo_dn), (size_t)f.n_halo_up * 3 * sizeof(T), NCCL_INT8, f.up_peer, comm, st));
      NL_NCCL(nccl().Recv(ga + no + f.n_halo_dn, (size_t)f.n_halo_up * sizeof(TI), NCCL_INT8, f.up_peer, comm, st));
    }
    NL_NCCL(nccl().GroupEnd());
  }
  pt.mark("halo-exchange");
  pt.report("nl_shard_exchange");
  return NL_OK;
}

#define NL_DISPATCH(p, FN, ...)                                                              \
  ((p)->float_type == NL_F64                                                                 \
       ? ((p)->int_type == NL_I64 ? FN<double, int64_t>(__VA_ARGS__) : FN<double, int32_t>(__VA_ARGS__)) \
       : ((p)->int_type == NL_I64 ? FN<float, int64_t>(__VA_ARGS__) : FN<float, int32_t>(__VA_ARGS__)))

int check_ws(const void* ws, size_t have, size_t need) {
  if (need == 0) return NL_OK;
  if (!ws || ((uintptr_t)ws & 255) != 0 || have < need) return NL_ERR_WORKSPACE;
  return NL_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// nl_pairs_to_host (+ _begin / _finish): see nl_tohost.cuh.  One job = one transfer.  begin copies `first` on the job's private
// stream and starts the host threads; finish enqueues the device side (pack kernel, code chunks, j) on the caller's stream, feeds
// the threads as the chunks arrive and waits for everything.  The threads take work from two queues -- slices of i (bound by
// instruction rate) and slices of the arrived S chunks (bound by memory bandwidth) -- half of them preferring one, half the
// other, so that both kinds of work are in flight at any time.
}  // namespace
struct nl_to_host_job {
  int dev = 0;
  int int64 = 0;
  int64_t n_rows = 0, P = 0, i_from = 0;
  const void* first_h = nullptr;
  const void* rowmap_h = nullptr;   // shard lists: global index per row (host copy), else null
  void* i_h = nullptr;
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_first = nullptr;
  int T = 0;
  int n_islices = 0;                // i slices of [0, i_from)
  int64_t islice = 0;
  std::vector<std::thread> workers;
  std::atomic<int> ready{0};        // 1: first is on the host and belongs to P; -1: abort
  std::atomic<int> phase2{0};       // 1: unpack the code chunks as they arrive; 2: nothing to unpack (S copied as it is); -1: abort
  std::atomic<int> chunks_ready{0};
  std::atomic<int> next_i{0}, next_s{0};
  std::atomic<int> error{0};
  // set by finish before phase2 is published
  const uint8_t* codes_h = nullptr;
  void* S_h = nullptr;
  int nchunks = 0, sslices = 0;     // S slices per chunk
  int64_t chunk = 0;
};
namespace {

template <class TI>
void to_host_worker(nl_to_host_job* J, int t) {
  if (t == 0) {
    int ok = cudaSetDevice(J->dev) == cudaSuccess && cudaEventSynchronize(J->ev_first) == cudaSuccess;
    if (ok && ((int64_t)((const TI*)J->first_h)[J->n_rows] - 1 != J->P || ((const TI*)J->first_h)[0] != 1)) {
      J->error.store(NL_ERR_BAD_ARG);  // first and P do not belong together
      ok = 0;
    } else if (!ok) {
      J->error.store(NL_ERR_CUDA);
    }
    J->ready.store(ok ? 1 : -1, std::memory_order_release);
  }
  int v;
  while ((v = J->ready.load(std::memory_order_acquire)) == 0) std::this_thread::yield();
  if (v < 0) return;
  const int64_t P = J->P;
  const bool prefer_s = (t & 1) != 0;
  auto take_i = [&]() {
    if (J->next_i.load(std::memory_order_relaxed) >= J->n_islices) return false;
    const int k = J->next_i.fetch_add(1, std::memory_order_relaxed);
    if (k >= J->n_islices) return false;
    const int64_t a = std::min<int64_t>(J->i_from, J->islice * k), b = std::min<int64_t>(J->i_from, J->islice * (k + 1));
    host_expand_rows<TI>((const TI*)J->first_h, (const TI*)J->rowmap_h, (long long)J->n_rows, a, b, (TI*)J->i_h);
    return true;
  };
  auto take_s = [&]() {
    if (J->phase2.load(std::memory_order_acquire) != 1) return false;
    const int avail = J->chunks_ready.load(std::memory_order_acquire) * J->sslices;
    int s = J->next_s.load(std::memory_order_relaxed);
    while (s < avail && !J->next_s.compare_exchange_weak(s, s + 1, std::memory_order_relaxed)) {}
    if (s >= avail) return false;
    const int k = s / J->sslices, u = s % J->sslices;
    const int64_t c0 = (int64_t)k * J->chunk, c1 = std::min<int64_t>(P, c0 + J->chunk);
    const int64_t q = ((c1 - c0 + J->sslices - 1) / J->sslices + 3) & ~(int64_t)3;
    host_unpack_shifts<TI>(J->codes_h, std::min(c1, c0 + q * u), std::min(c1, c0 + q * (u + 1)), (TI*)J->S_h);
    return true;
  };
  while (true) {
    if (prefer_s ? (take_s() || take_i()) : (take_i() || take_s())) continue;
    const int p2 = J->phase2.load(std::memory_order_acquire);
    if (p2 < 0 || J->chunks_ready.load(std::memory_order_acquire) < 0) return;
    const bool i_done = J->next_i.load(std::memory_order_relaxed) >= J->n_islices;
    const bool s_done = p2 == 2 || (p2 == 1 && J->next_s.load(std::memory_order_relaxed) >= J->nchunks * J->sslices);
    if (i_done && s_done) return;
    std::this_thread::yield();
  }
}

void to_host_job_free(nl_to_host_job* J) {
  for (auto& w : J->workers) if (w.joinable()) w.join();
  if (J->ev_first) cudaEventDestroy(J->ev_first);
  if (J->aux) cudaStreamDestroy(J->aux);
  delete J;
}

// `first` must be final in stream order of `st` at the time of the call (after nl_count_pairs it is: that call synchronises)
int to_host_begin(int int_type, const void* first, int64_t n_rows, int64_t P, int64_t i_from, const void* row_index, void* row_index_host,
                  void* first_host, void* i_host, int nthreads, cudaStream_t st, nl_to_host_job** job_out) {
  *job_out = nullptr;
  if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency() / 2);  // memory-bound work: one thread per core pair
  nl_to_host_job* J = new (std::nothrow) nl_to_host_job();
  if (!J) return NL_ERR_BAD_ARG;
  J->int64 = int_type == NL_I64;
  J->n_rows = n_rows; J->P = P; J->i_from = i_from; J->first_h = first_host; J->i_h = i_host;
  J->rowmap_h = row_index ? row_index_host : nullptr;
  J->T = P > 0 ? std::max(1, std::min<int>(nthreads, 256)) : 0;
  J->n_islices = i_from > 0 ? (int)std::min<int64_t>(4 * J->T, (i_from + 65535) / 65536) : 0;
  J->islice = J->n_islices ? ((i_from + J->n_islices - 1) / J->n_islices + 63) & ~(int64_t)63 : 0;
  const size_t w = J->int64 ? 8 : 4;
  cudaEvent_t ev_in = nullptr;
  cudaError_t ce = cudaGetDevice(&J->dev);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&J->aux, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&J->ev_first, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventRecord(ev_in, st);           // whatever produced `first` on the caller's stream comes first
  if (ce == cudaSuccess) ce = cudaStreamWaitEvent(J->aux, ev_in, 0);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(first_host, first, (size_t)(n_rows + 1) * w, cudaMemcpyDeviceToHost, J->aux);
  if (ce == cudaSuccess && row_index && n_rows > 0) ce = cudaMemcpyAsync(row_index_host, row_index, (size_t)n_rows * w, cudaMemcpyDeviceToHost, J->aux);
  if (ce == cudaSuccess) ce = cudaEventRecord(J->ev_first, J->aux);
  if (ev_in) cudaEventDestroy(ev_in);
  if (ce != cudaSuccess) {
    to_host_job_free(J);
    return cuda_fail(ce);
  }
  J->workers.reserve(J->T);
  for (int t = 0; t < J->T; t++) {
    if (J->int64) J->workers.emplace_back(to_host_worker<int64_t>, J, t);
    else J->workers.emplace_back(to_host_worker<int32_t>, J, t);
  }
  *job_out = J;
  return NL_OK;
}

template <class TI>
int to_host_finish_impl(nl_to_host_job* J, const void* i_d, const void* j_d, const void* S_d, void* j_h, void* S_h, void* dscratch, void* hscratch,
                        cudaStream_t st) {
  const int64_t P = J->P;
  int rc = NL_OK;
  auto stop = [&](int code) {
    if (rc == NL_OK) rc = code;
    J->phase2.store(-1, std::memory_order_release);
    J->chunks_ready.store(-1, std::memory_order_release);
  };
  std::vector<cudaEvent_t> ev;
  if (P > 0) {
    uint8_t* codes_d = (uint8_t*)dscratch;
    uint8_t* codes_h = (uint8_t*)hscratch;
    unsigned* flag_d = (unsigned*)(codes_d + al256((size_t)P));
    volatile unsigned* flag_h = (volatile unsigned*)(codes_h + al256((size_t)P));
    const int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(16, (P + (32ll << 20) - 1) / (32ll << 20)));
    const int64_t chunk = ((P + nchunks - 1) / nchunks + 63) & ~(int64_t)63;  // multiple of 64 pairs: slices of S stay 16-byte aligned
    ev.assign(nchunks + 1, nullptr);
    cudaError_t ce = cudaSuccess;
    for (auto& e : ev)
      if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(flag_d, 0, 4, st);
    if (ce == cudaSuccess) {
      const unsigned nb = (unsigned)((P + 256 * TH_PAIRS - 1) / (256 * TH_PAIRS));
      k_pack_shifts<TI><<<nb, 256, 0, st>>>((const TI*)S_d, (long long)P, codes_d, flag_d, (((uintptr_t)S_d) & 15) == 0 ? 1 : 0);
      NL_LAUNCHED(1);
      ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync((void*)flag_h, flag_d, 4, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaEventRecord(ev[0], st);
    for (int k = 0; k < nchunks && ce == cudaSuccess; k++) {
      const int64_t c0 = (int64_t)k * chunk, c1 = std::min<int64_t>(P, c0 + chunk);
      if (c1 > c0) ce = cudaMemcpyAsync(codes_h + c0, codes_d + c0, (size_t)(c1 - c0), cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess) ce = cudaEventRecord(ev[k + 1], st);
    }
    if (ce == cudaSuccess && J->i_from < P)  // this part of i crosses the bus, [0, i_from) is rebuilt from first
      ce = cudaMemcpyAsync((TI*)J->i_h + J->i_from, (const TI*)i_d + J->i_from, (size_t)(P - J->i_from) * sizeof(TI), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(j_h, j_d, (size_t)P * sizeof(TI), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaEventSynchronize(ev[0]);
    if (ce != cudaSuccess) stop(cuda_fail(ce));
    else if (*flag_h) {  // a shift component outside {-1, 0, 1}: S goes over the bus as it is
      J->phase2.store(2, std::memory_order_release);
      ce = cudaMemcpyAsync(S_h, S_d, (size_t)P * 3 * sizeof(TI), cudaMemcpyDeviceToHost, st);
      if (ce != cudaSuccess) stop(cuda_fail(ce));
    } else {
      J->codes_h = codes_h; J->S_h = S_h; J->nchunks = nchunks; J->chunk = chunk;
      J->sslices = std::max(1, std::min(J->T, 64));
      J->phase2.store(1, std::memory_order_release);
      for (int k = 0; k < nchunks && rc == NL_OK; k++) {
        ce = cudaEventSynchronize(ev[k + 1]);
        if (ce != cudaSuccess) stop(cuda_fail(ce));
        else J->chunks_ready.store(k + 1, std::memory_order_release);
      }
    }
  }
  cudaError_t ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) stop(cuda_fail(ce));
  ce = cudaStreamSynchronize(J->aux);
  if (ce != cudaSuccess) stop(cuda_fail(ce));
  if (J->phase2.load() == 0) J->phase2.store(2, std::memory_order_release);  // P == 0: nothing was started
  for (auto& w : J->workers) w.join();
  for (auto e : ev) if (e) cudaEventDestroy(e);
  if (rc == NL_OK && J->error.load()) rc = J->error.load();
  to_host_job_free(J);
  return rc;
}

int to_host_finish(nl_to_host_job* job, const void* i, const void* j, const void* S, void* j_host, void* S_host, void* dev_scratch,
                   void* host_scratch, size_t scratch_bytes, cudaStream_t st) {
  const int64_t P = job->P;
  int bad = NL_OK;
  if (P > 0 && (!j || !S || !j_host || !S_host || (job->i_from < P && !i))) bad = NL_ERR_BAD_ARG;
  else if (P > 0 && (!dev_scratch || !host_scratch || scratch_bytes < al256((size_t)P) + 256 ||
                     ((((uintptr_t)dev_scratch) | ((uintptr_t)host_scratch)) & 15)))
    bad = NL_ERR_WORKSPACE;
  if (bad) {  // the job is consumed either way: stop its workers and release it
    job->phase2.store(-1, std::memory_order_release);
    job->chunks_ready.store(-1, std::memory_order_release);
    cudaStreamSynchronize(job->aux);
    to_host_job_free(job);
    return bad;
  }
  return job->int64 ? to_host_finish_impl<int64_t>(job, i, j, S, j_host, S_host, dev_scratch, host_scratch, st)
                    : to_host_finish_impl<int32_t>(job, i, j, S, j_host, S_host, dev_scratch, host_scratch, st);
}

}  // namespace

// ================================================================ exported C ABI
extern "C" {

int nl_version(void) { return NL_VERSION; }

const char* nl_strerror(int code) {
  switch (code) {
    case NL_OK: return "ok";
    case NL_ERR_BAD_ARG: return "nlcuda: bad argument (null pointer, negative N, bad type tag, ncells < 1, nxyz < 1 or cutoff <= 0)";
    case NL_ERR_WORKSPACE: return "nlcuda: workspace missing, misaligned (256 B) or smaller than nl_workspace_bytes()";
    case NL_ERR_CUDA: return "nlcuda: CUDA error (see nl_last_cuda_error())";
    case NL_ERR_OVERFLOW: return "nlcuda: number of pairs overflows int_type; use a 64-bit int_type";
    case NL_ERR_UNSUPPORTED: return "nlcuda: N or prod(ncells) >= 2^31 - 1 is not supported (use a larger cutoff or a smaller simulation cell)";
    case NL_ERR_NCCL: return "nlcuda: NCCL is not loadable in this process (libnccl.so.2) or an NCCL call failed";
    default: return "nlcuda: unknown error code";
  }
}

int nl_last_cuda_error(void) { return nl::last_cuda_slot(); }

long long nl_launch_count(void) { return nl::launch_counter().load(); }

size_t nl_workspace_bytes(const nl_params* params, int64_t N, int stage) {
  if (check_params(params, N) != NL_OK) return 0;
  if (stage == NL_STAGE_BUILD) return build_ws(nullptr, N, params_nct(params)).total;
  if (stage == NL_STAGE_PAIRS) return pair_ws(nullptr, params, N).total_bytes;
  return 0;
}

int nl_build_cells(const nl_params* params, const void* X, int64_t N, void* X_sorted, void* perm, void* cell_id, void* cell_offsets,
                   void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (!cell_offsets || (N > 0 && (!X || !X_sorted || !perm || !cell_id))) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, build_ws(nullptr, N, params_nct(params)).total);
  if (rc) return rc;
  return NL_DISPATCH(params, build_cells_impl, params, X, N, X_sorted, perm, cell_id, cell_offsets, ws, (cudaStream_t)stream);
}

int nl_count_pairs(const nl_params* params, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets, void* first,
                   int64_t* total_pairs_host, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (!first || !total_pairs_host || !cell_offsets || (N > 0 && (!X_sorted || !perm))) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, pair_ws(nullptr, params, N).total_bytes);
  if (rc) return rc;
  return NL_DISPATCH(params, count_pairs_impl, params, X_sorted, N, perm, cell_offsets, first, total_pairs_host, ws, nullptr, (cudaStream_t)stream);
}

int nl_fill_pairs(const nl_params* params, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets, const void* first,
                  void* i_out, void* j_out, void* S_out, void* R_out, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (N == 0) return NL_OK;
  if (!first || !cell_offsets || !X_sorted || !perm || !i_out || !j_out || !S_out) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, pair_ws(nullptr, params, N).total_bytes);
  if (rc) return rc;
  return NL_DISPATCH(params, fill_pairs_impl, params, N, cell_offsets, first, i_out, j_out, S_out, R_out, ws, N, nullptr, nullptr, (cudaStream_t)stream);
}

int nl_fill_pairs_rows(const nl_params* params, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets, const void* first,
                       int64_t n_rows, const void* index_map, void* i_out, void* j_out, void* S_out, void* R_out, void* ws, size_t ws_bytes,
                       void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (N == 0 || n_rows == 0) return NL_OK;
  if (n_rows < 0 || n_rows > N) return NL_ERR_BAD_ARG;
  if (!first || !cell_offsets || !X_sorted || !perm || !i_out || !j_out || !S_out) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, pair_ws(nullptr, params, N).total_bytes);
  if (rc) return rc;
  return NL_DISPATCH(params, fill_pairs_impl, params, N, cell_offsets, first, i_out, j_out, S_out, R_out, ws, n_rows, index_map, nullptr,
                     (cudaStream_t)stream);
}

int nl_count_pairs_window(const nl_params* params, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets, void* first,
                          int64_t* total_pairs_host, const uint8_t* plane_active, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (!first || !total_pairs_host || !cell_offsets || (N > 0 && (!X_sorted || !perm))) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, pair_ws(nullptr, params, N).total_bytes);
  if (rc) return rc;
  return NL_DISPATCH(params, count_pairs_impl, params, X_sorted, N, perm, cell_offsets, first, total_pairs_host, ws, plane_active,
                     (cudaStream_t)stream);
}

int nl_fill_pairs_window(const nl_params* params, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets, const void* first,
                         int64_t n_rows, const void* index_map, const uint8_t* plane_active, void* i_out, void* j_out, void* S_out, void* R_out,
                         void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (N == 0 || n_rows == 0) return NL_OK;
  if (n_rows < 0 || n_rows > N) return NL_ERR_BAD_ARG;
  if (!first || !cell_offsets || !X_sorted || !perm || !i_out || !j_out || !S_out) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, pair_ws(nullptr, params, N).total_bytes);
  if (rc) return rc;
  return NL_DISPATCH(params, fill_pairs_impl, params, N, cell_offsets, first, i_out, j_out, S_out, R_out, ws, n_rows, index_map, plane_active,
                     (cudaStream_t)stream);
}

int nl_shard_plan(const int64_t* plane_hist, int32_t nplanes, int32_t nranks, int32_t halo, int64_t* bounds_out) {
  if (!plane_hist || !bounds_out || nplanes < 1 || nranks < 1 || halo < 0) return NL_ERR_BAD_ARG;
  const long long minw = nranks > 1 ? 2ll * halo + 1 : 1;
  if ((long long)nranks * minw > nplanes) return NL_ERR_BAD_ARG;
  long long total = 0;
  for (int p = 0; p < nplanes; p++) total += plane_hist[p];
  bounds_out[0] = 0;
  long long cum = 0;  // atoms in planes [0, b)
  int b = 0;
  for (int r = 1; r < nranks; r++) {
    // smallest b with cum(b) >= total * r / nranks   (exact rational comparison: cum * nranks >= total * r)
    while (b < nplanes && (__int128)cum * nranks < (__int128)total * r) cum += plane_hist[b++];
    long long bb = b;
    if (bb < bounds_out[r - 1] + minw) bb = bounds_out[r - 1] + minw;              // this slab wide enough
    if (bb > nplanes - (long long)(nranks - r) * minw) bb = nplanes - (long long)(nranks - r) * minw;  // room for the rest
    while (b < bb) cum += plane_hist[b++];
    while (b > bb) cum -= plane_hist[--b];
    bounds_out[r] = bb;
  }
  bounds_out[nranks] = nplanes;
  return NL_OK;
}

size_t nl_shard_workspace_bytes(const nl_params* params, int64_t n_max, int32_t nranks) {
  if (check_params(params, n_max) != NL_OK || nranks < 1 || nranks > NL_MAX_RANKS) return 0;
  return shard_ws(nullptr, n_max, params->ncells[slab_axis(params)], nranks, fsize(params), params->int_type == NL_I64 ? 8 : 4).total;
}

int nl_shard_prepare(const nl_params* params, const void* X, int64_t n, void* comm, int32_t rank, int32_t nranks, nl_shard_info* info_out,
                     void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, n);
  if (rc) return rc;
  if (!info_out || nranks < 1 || nranks > NL_MAX_RANKS || rank < 0 || rank >= nranks || (n > 0 && !X)) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, nl_shard_workspace_bytes(params, n, nranks));
  if (rc) return rc;
  return NL_DISPATCH(params, shard_prepare_impl, params, X, n, comm, rank, nranks, info_out, ws, (cudaStream_t)stream);
}

int nl_shard_exchange(const nl_params* params, const nl_shard_info* info, const void* X, const void* gidx, int64_t n, void* comm, void* X_all,
                      void* gidx_all, uint8_t* plane_active_out, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, n);
  if (rc) return rc;
  if (!info || info->nranks < 1 || info->nranks > NL_MAX_RANKS || info->n_local != n || (n > 0 && (!X || !gidx))) return NL_ERR_BAD_ARG;
  if (info->n_owned + info->n_halo_dn + info->n_halo_up > 0 && (!X_all || !gidx_all)) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, nl_shard_workspace_bytes(params, std::max<int64_t>(n, info->n_owned), info->nranks));
  if (rc) return rc;
  return NL_DISPATCH(params, shard_exchange_impl, params, info, X, gidx, n, comm, nullptr, X_all, gidx_all, plane_active_out, ws, (cudaStream_t)stream);
}

int nl_shard_exchange_peer(const nl_params* params, const nl_shard_info* info, const void* X, const void* gidx, int64_t n, void* comm,
                           const nl_shard_peers* peers, void* X_all, void* gidx_all, uint8_t* plane_active_out, void* ws, size_t ws_bytes,
                           void* stream) {
  int rc = check_params(params, n);
  if (rc) return rc;
  if (!info || info->nranks < 1 || info->nranks > NL_MAX_RANKS || info->n_local != n || (n > 0 && (!X || !gidx))) return NL_ERR_BAD_ARG;
  if (info->n_owned + info->n_halo_dn + info->n_halo_up > 0 && (!X_all || !gidx_all)) return NL_ERR_BAD_ARG;
  if (!peers || peers->nranks != info->nranks || peers->rank != info->rank) return NL_ERR_BAD_ARG;
  if (ws != peers->ws || ws_bytes != peers->ws_bytes || info->n_max_all > peers->cap) return NL_ERR_WORKSPACE;  // the same verdict on every rank
  rc = check_ws(ws, ws_bytes, nl_shard_workspace_bytes(params, peers->cap, info->nranks));
  if (rc) return rc;
  return NL_DISPATCH(params, shard_exchange_impl, params, info, X, gidx, n, comm, info->nranks > 1 ? peers : nullptr, X_all, gidx_all,
                     plane_active_out, ws, (cudaStream_t)stream);
}

int nl_shard_connect(const nl_params* params, int64_t cap, void* comm, int32_t rank, int32_t nranks, void* ws, size_t ws_bytes,
                     nl_shard_peers* peers_out, void* stream) {
  int rc = check_params(params, cap);
  if (rc) return rc;
  if (!peers_out || nranks < 1 || nranks > NL_MAX_RANKS || rank < 0 || rank >= nranks || cap < 1) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, nl_shard_workspace_bytes(params, cap, nranks));
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  nl_shard_peers& P = *peers_out;
  P = nl_shard_peers{};
  P.nranks = nranks; P.rank = rank; P.cap = cap; P.ws_bytes = ws_bytes; P.ws = ws;
  if (nranks == 1) return NL_OK;
  if (!nccl().ok || !comm) return NL_ERR_NCCL;
  // the allocation `ws` lives in (a caching allocator hands out interior pointers; IPC handles name whole allocations)
  typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
  static GetRange get_range = nullptr;
  if (!get_range) {
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) get_range = (GetRange)dlsym(h, "cuMemGetAddressRange_v2");
    if (!get_range) return NL_ERR_UNSUPPORTED;
  }
  unsigned long long base = 0;
  size_t alloc_bytes = 0;
  if (get_range(&base, &alloc_bytes, (unsigned long long)(uintptr_t)ws) != 0) return NL_ERR_BAD_ARG;
  struct Desc { cudaIpcMemHandle_t h; unsigned long long offset, bytes, cap; unsigned long long pad; };
  static_assert(sizeof(Desc) == 96, "descriptor size");
  Desc mine{};
  NL_CUDA(cudaIpcGetMemHandle(&mine.h, (void*)(uintptr_t)base));
  mine.offset = (unsigned long long)(uintptr_t)ws - base; mine.bytes = ws_bytes; mine.cap = (unsigned long long)cap;
  std::vector<Desc> all(nranks);
  char* dsc = (char*)ws;  // the workspace is free at connect time: descriptor at 0, the gathered ones from 4096 on
  NL_CUDA(cudaMemcpyAsync(dsc, &mine, sizeof(Desc), cudaMemcpyHostToDevice, st));
  NL_NCCL(nccl().AllGather(dsc, dsc + 4096, sizeof(Desc), NCCL_INT8, comm, st));
  NL_CUDA(cudaMemcpyAsync(all.data(), dsc + 4096, sizeof(Desc) * nranks, cudaMemcpyDeviceToHost, st));
  NL_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r < nranks; r++)
    if (all[r].cap != (unsigned long long)cap || all[r].bytes != ws_bytes) return NL_ERR_BAD_ARG;  // the layouts must agree (seen by every rank alike)
  for (int r = 0; r < nranks; r++) {
    if (r == rank) continue;
    void* m = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&m, all[r].h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; q++) if (P.peer_base[q]) cudaIpcCloseMemHandle(P.peer_base[q]);
      P = nl_shard_peers{};
      return cuda_fail(e);
    }
    P.peer_base[r] = m;
    P.peer_ws[r] = (char*)m + all[r].offset;
  }
  // nobody may start writing into its workspace (the first exchange) before every rank has read the descriptors out of it
  NL_NCCL(nccl().AllGather(dsc + 2048, dsc + 4096, 8, NCCL_INT8, comm, st));
  NL_CUDA(cudaStreamSynchronize(st));
  return NL_OK;
}

int nl_shard_disconnect(nl_shard_peers* peers) {
  if (!peers) return NL_ERR_BAD_ARG;
  int rc = NL_OK;
  for (int r = 0; r < NL_MAX_RANKS; r++)
    if (peers->peer_base[r]) {
      if (cudaIpcCloseMemHandle(peers->peer_base[r]) != cudaSuccess) rc = NL_ERR_CUDA;
      peers->peer_base[r] = nullptr;
      peers->peer_ws[r] = nullptr;
    }
  return rc;
}

int nl_nccl_unique_id(void* id128_out) {
  if (!id128_out) return NL_ERR_BAD_ARG;
  if (!nccl().ok) return NL_ERR_NCCL;
  return nccl().GetUniqueId((NcclUid*)id128_out) == 0 ? NL_OK : NL_ERR_NCCL;
}

int nl_nccl_comm_init(void** comm_out, int32_t nranks, const void* id128, int32_t rank) {
  if (!comm_out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return NL_ERR_BAD_ARG;
  if (!nccl().ok) return NL_ERR_NCCL;
  NcclUid id;
  memcpy(&id, id128, sizeof(id));
  return nccl().CommInitRank(comm_out, nranks, id, rank) == 0 ? NL_OK : NL_ERR_NCCL;
}

int nl_nccl_comm_destroy(void* comm) {
  if (!comm) return NL_OK;
  if (!nccl().ok) return NL_ERR_NCCL;
  return nccl().CommDestroy(comm) == 0 ? NL_OK : NL_ERR_NCCL;
}

int nl_cell_ids(const nl_params* params, const void* X, int64_t N, void* cell_id_out, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (N == 0) return NL_OK;
  if (!X || !cell_id_out) return NL_ERR_BAD_ARG;
  return NL_DISPATCH(params, cell_ids_impl, params, X, N, cell_id_out, (cudaStream_t)stream);
}

int nl_lazy_count(const nl_params* params, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets, void* counts_out,
                  void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (N == 0) return NL_OK;
  if (!counts_out || !cell_offsets || !X_sorted || !perm) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, pair_ws(nullptr, params, N).total_bytes);
  if (rc) return rc;
  return NL_DISPATCH(params, lazy_count_impl, params, X_sorted, N, perm, cell_offsets, counts_out, ws, (cudaStream_t)stream);
}

int nl_lazy_lj_energy(const nl_params* params, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets, double eps,
                      double sigma, double* energy_out, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (!energy_out || (N > 0 && (!cell_offsets || !X_sorted || !perm))) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, pair_ws(nullptr, params, N).total_bytes);
  if (rc) return rc;
  return NL_DISPATCH(params, lazy_lj_impl, params, X_sorted, N, perm, cell_offsets, eps, sigma, energy_out, ws, (cudaStream_t)stream);
}

int nl_lazy_lj_forces(const nl_params* params, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets, double eps,
                      double sigma, void* fe_out, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (N == 0) return NL_OK;
  if (!fe_out || !cell_offsets || !X_sorted || !perm || ((uintptr_t)fe_out & 15) != 0) return NL_ERR_BAD_ARG;
  rc = check_ws(ws, ws_bytes, pair_ws(nullptr, params, N).total_bytes);
  if (rc) return rc;
  return NL_DISPATCH(params, lazy_ljf_impl, params, X_sorted, N, perm, cell_offsets, eps, sigma, fe_out, ws, (cudaStream_t)stream);
}

int nl_pairs_R(const nl_params* params, const void* X, int64_t N, const void* i, const void* j, const void* S, int64_t p_lo, int64_t p_hi,
               void* R_out, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (p_lo < 0 || p_hi < p_lo) return NL_ERR_BAD_ARG;
  if (p_hi == p_lo) return NL_OK;
  if (!X || !i || !j || !S || !R_out) return NL_ERR_BAD_ARG;
  return NL_DISPATCH(params, pairs_R_impl, params, X, i, j, S, p_lo, p_hi, R_out, (cudaStream_t)stream);
}

int nl_max_neighbours(const nl_params* params, const void* first, int64_t N, int64_t* max_out, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (N == 0 || !first || !max_out) return NL_ERR_BAD_ARG;  // reference: maximum over an empty collection throws
  cudaStream_t st = (cudaStream_t)stream;
  NL_CUDA(cudaMemsetAsync(max_out, 0, sizeof(int64_t), st));
  const unsigned nb = (unsigned)std::min<long long>(RED_BLOCKS, (N + 255) / 256);
  if (params->int_type == NL_I64) k_max_neighbours<int64_t><<<nb, 256, 0, st>>>((const int64_t*)first, N, (unsigned long long*)max_out);
  else k_max_neighbours<int32_t><<<nb, 256, 0, st>>>((const int32_t*)first, N, (unsigned long long*)max_out);
  NL_LAUNCHED(1);
  NL_LAUNCH_CHECK();
  return NL_OK;
}

int nl_rows_padded(const nl_params* params, const void* X, int64_t N, const void* first, const void* j, const void* S, const void* rows,
                   int64_t n_sel, int32_t width, void* n_out, void* j_out, void* S_out, void* R_out, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (n_sel < 0 || width < 0) return NL_ERR_BAD_ARG;
  if (n_sel == 0) return NL_OK;
  if (!X || !first || !rows || !n_out || (width > 0 && (!j || !S || !j_out))) return NL_ERR_BAD_ARG;
  return NL_DISPATCH(params, rows_padded_impl, params, X, first, j, S, rows, n_sel, width, n_out, j_out, S_out, R_out, (cudaStream_t)stream);
}

int nl_lazy_neighbours(const nl_params* params, const void* X_orig, const void* X_sorted, int64_t N, const void* perm, const void* cell_offsets,
                       const void* atoms, int64_t n_sel, int32_t width, void* n_out, void* j_out, void* S_out, void* R_out, void* stream) {
  int rc = check_params(params, N);
  if (rc) return rc;
  if (n_sel < 0 || width < 0) return NL_ERR_BAD_ARG;
  if (n_sel == 0) return NL_OK;
  if (N == 0 || !X_orig || !X_sorted || !perm || !cell_offsets || !atoms || !n_out || (width > 0 && !j_out)) return NL_ERR_BAD_ARG;
  return NL_DISPATCH(params, lazy_neighbours_impl, params, X_orig, X_sorted, perm, cell_offsets, atoms, n_sel, width, n_out, j_out, S_out, R_out,
                     (cudaStream_t)stream);
}

int nl_bounding_box(int32_t float_type, const void* X, int64_t N, void* minmax_out, void* ws, size_t ws_bytes, void* stream) {
  if (float_type != NL_F32 && float_type != NL_F64) return NL_ERR_BAD_ARG;
  if (N < 1 || !X || !minmax_out) return NL_ERR_BAD_ARG;  // reference: maximum over an empty collection throws
  int rc = check_ws(ws, ws_bytes, NL_REDUCE_WS_BYTES);
  if (rc) return rc;
  return float_type == NL_F64 ? bbox_impl<double>(X, N, minmax_out, ws, (cudaStream_t)stream)
                              : bbox_impl<float>(X, N, minmax_out, ws, (cudaStream_t)stream);
}

int nl_max_displacement2(int32_t float_type, const void* X, const void* X_ref, int64_t N, void* d2_out, void* ws, size_t ws_bytes,
                         void* stream) {
  if (float_type != NL_F32 && float_type != NL_F64) return NL_ERR_BAD_ARG;
  if (N < 0 || !d2_out || (N > 0 && (!X || !X_ref))) return NL_ERR_BAD_ARG;
  int rc = check_ws(ws, ws_bytes, NL_REDUCE_WS_BYTES);
  if (rc) return rc;
  return float_type == NL_F64 ? maxdisp_impl<double>(X, X_ref, N, d2_out, ws, (cudaStream_t)stream)
                              : maxdisp_impl<float>(X, X_ref, N, d2_out, ws, (cudaStream_t)stream);
}

size_t nl_to_host_scratch_bytes(int64_t P) { return al256((size_t)(P > 0 ? P : 0)) + 256; }

int nl_pairs_to_host(const nl_params* params, const void* first, int64_t n_rows, const void* i, int64_t i_copy_from, const void* row_index,
                     void* row_index_host, const void* j, const void* S, int64_t P, void* first_host, void* i_host, void* j_host, void* S_host,
                     void* dev_scratch, void* host_scratch, size_t scratch_bytes, int32_t nthreads, void* stream) {
  if (!params || (params->int_type != NL_I32 && params->int_type != NL_I64)) return NL_ERR_BAD_ARG;
  if (n_rows < 0 || P < 0 || !first || !first_host) return NL_ERR_BAD_ARG;
  if (P > 0 && (!j || !S || !i_host || !j_host || !S_host)) return NL_ERR_BAD_ARG;
  if (i_copy_from < 0 || i_copy_from > P || (i_copy_from < P && !i)) return NL_ERR_BAD_ARG;
  if (row_index && !row_index_host) return NL_ERR_BAD_ARG;
  if (i_copy_from < P) i_copy_from &= ~(int64_t)63;  // the two parts of i meet on a 256-byte boundary
  if (P > 0) {
    if (!dev_scratch || !host_scratch || scratch_bytes < nl_to_host_scratch_bytes(P)) return NL_ERR_WORKSPACE;
    if ((((uintptr_t)dev_scratch) | ((uintptr_t)host_scratch)) & 15) return NL_ERR_WORKSPACE;
  }
  nl_to_host_job* job = nullptr;
  int rc = to_host_begin(params->int_type, first, n_rows, P, i_copy_from, row_index, row_index_host, first_host, i_host, nthreads, (cudaStream_t)stream, &job);
  if (rc) return rc;
  return to_host_finish(job, i, j, S, j_host, S_host, dev_scratch, host_scratch, scratch_bytes, (cudaStream_t)stream);
}

int nl_pairs_to_host_begin(const nl_params* params, const void* first, int64_t n_rows, int64_t P, void* first_host, void* i_host,
                           int32_t nthreads, void* stream, nl_to_host_job** job_out) {
  if (!params || (params->int_type != NL_I32 && params->int_type != NL_I64)) return NL_ERR_BAD_ARG;
  if (!job_out || n_rows < 0 || P < 0 || !first || !first_host || (P > 0 && !i_host)) return NL_ERR_BAD_ARG;
  return to_host_begin(params->int_type, first, n_rows, P, P, nullptr, nullptr, first_host, i_host, nthreads, (cudaStream_t)stream, job_out);
}

int nl_pairs_to_host_finish(nl_to_host_job* job, const void* j, const void* S, void* j_host, void* S_host, void* dev_scratch,
                            void* host_scratch, size_t scratch_bytes, void* stream) {
  if (!job) return NL_ERR_BAD_ARG;
  return to_host_finish(job, nullptr, j, S, j_host, S_host, dev_scratch, host_scratch, scratch_bytes, (cudaStream_t)stream);
}

int nl_host_expand_rows(int32_t int_type, const void* first, const void* row_index, int64_t n_rows, int64_t p_lo, int64_t p_hi, void* i_out) {
  if ((int_type != NL_I32 && int_type != NL_I64) || n_rows < 0 || p_lo < 0 || p_hi < p_lo) return NL_ERR_BAD_ARG;
  if (p_hi == p_lo) return NL_OK;
  if (!first || !i_out || n_rows == 0) return NL_ERR_BAD_ARG;
  if (int_type == NL_I64) {
    if (p_hi > ((const int64_t*)first)[n_rows] - 1) return NL_ERR_BAD_ARG;
    host_expand_rows<int64_t>((const int64_t*)first, (const int64_t*)row_index, n_rows, p_lo, p_hi, (int64_t*)i_out);
  } else {
    if (p_hi > (int64_t)((const int32_t*)first)[n_rows] - 1) return NL_ERR_BAD_ARG;
    host_expand_rows<int32_t>((const int32_t*)first, (const int32_t*)row_index, n_rows, p_lo, p_hi, (int32_t*)i_out);
  }
  return NL_OK;
}

int nl_host_unpack_shifts(int32_t int_type, const uint8_t* codes, int64_t p_lo, int64_t p_hi, void* S_out) {
  if ((int_type != NL_I32 && int_type != NL_I64) || p_lo < 0 || p_hi < p_lo) return NL_ERR_BAD_ARG;
  if (p_hi == p_lo) return NL_OK;
  if (!codes || !S_out) return NL_ERR_BAD_ARG;
  if (int_type == NL_I64) host_unpack_shifts<int64_t>(codes, p_lo, p_hi, (int64_t*)S_out);
  else host_unpack_shifts<int32_t>(codes, p_lo, p_hi, (int32_t*)S_out);
  return NL_OK;
}

}  // extern "C"
