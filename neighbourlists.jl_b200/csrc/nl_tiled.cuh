// nl_tiled.cuh -- tiled, shared-memory-staged traversal for the common geometry (placeholder:
// not yet enabled; every problem goes through k_traverse_generic).
#pragma once
#include "../../include/nlcuda.h"
#include "nl_traverse.cuh"

namespace nl {

inline size_t tiled_scratch_bytes(const nl_params*, int64_t) { return 256; }
template <class T> inline bool tiled_applicable(const nl_params*, const Geo<T>&) { return false; }
template <class T, class TI, int MODE>
inline int tiled_traverse(const nl_params*, int64_t, const TI*, const Records<T>&, const Geo<T>&, const Sinks<T, TI>&, void*, cudaStream_t) {
  return NL_ERR_UNSUPPORTED;
}

}  // namespace nl
