/* host_list.c -- the C ABI of libnlcuda.so from plain C99: positions in host memory in, the complete neighbour list
 * (i, j, S, first) in host memory out.  Everything the library needs crosses the boundary as plain pointers and sizes
 * (include/nlcuda.h); device memory comes from the CUDA runtime, which is the only other dependency.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/host_list.c -o host_list \
 *       -L neighbourlists.jl_b200 -lnlcuda -L /usr/local/cuda/lib64 -lcudart -lm
 *
 * The geometry (inv_cell, ncells, nxyz) is computed by the CALLER, exactly as the reference's analyze_cell does
 * (src/cell_list.jl:152-170) -- here for a cubic box, where it is trivial.                                           */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nlcuda.h"

#define CHECK(call)                                                           \
  do {                                                                        \
    int rc__ = (call);                                                        \
    if (rc__ != NL_OK) {                                                      \
      fprintf(stderr, "%s -> %s\n", #call, nl_strerror(rc__));                \
      return 1;                                                               \
    }                                                                         \
  } while (0)

int main(int argc, char** argv) {
  const int64_t N = argc > 1 ? atoll(argv[1]) : 100000;
  const double rho = 0.05, rc = 5.0, L = cbrt((double)N / rho);

  /* nl_params of a cubic periodic box (rows of `cell` are the lattice vectors; column-major 3 x 3) */
  nl_params p;
  memset(&p, 0, sizeof p);
  p.float_type = NL_F64;
  p.int_type = NL_I32;
  for (int k = 0; k < 3; k++) {
    p.cell[k + 3 * k] = L;
    p.inv_cell[k + 3 * k] = 1.0 / L;
    p.ncells[k] = (int32_t)fmax(floor(L / rc), 1.0);
    p.nxyz[k] = (int32_t)ceil(rc * p.ncells[k] / L);
    p.pbc[k] = 1;
  }
  p.cutoff = rc;
  int64_t nct = (int64_t)p.ncells[0] * p.ncells[1] * p.ncells[2];

  /* positions */
  double* X = (double*)malloc((size_t)N * 3 * sizeof(double));
  srand(10);
  for (int64_t k = 0; k < 3 * N; k++) X[k] = L * (rand() / (RAND_MAX + 1.0));

  /* device buffers: the caller owns everything */
  void *dX, *dXs, *dperm, *dcid, *dco, *dfirst, *wsb, *wsp;
  size_t nb = nl_workspace_bytes(&p, N, NL_STAGE_BUILD), np = nl_workspace_bytes(&p, N, NL_STAGE_PAIRS);
  cudaMalloc(&dX, (size_t)N * 24); cudaMalloc(&dXs, (size_t)N * 24);
  cudaMalloc(&dperm, (size_t)N * 4); cudaMalloc(&dcid, (size_t)N * 4);
  cudaMalloc(&dco, (size_t)(nct + 1) * 4); cudaMalloc(&dfirst, (size_t)(N + 1) * 4);
  cudaMalloc(&wsb, nb); cudaMalloc(&wsp, np);
  cudaMemcpy(dX, X, (size_t)N * 24, cudaMemcpyHostToDevice);

  /* build_cell_list + the counting pass (the one host synchronisation: P is data dependent) */
  int64_t P = 0;
  CHECK(nl_build_cells(&p, dX, N, dXs, dperm, dcid, dco, wsb, nb, NULL));
  CHECK(nl_count_pairs(&p, dXs, N, dperm, dco, dfirst, &P, wsp, np, NULL));

  /* outputs: device arrays for the fill pass, pinned host arrays for the list */
  void *di, *dj, *dS, *dscr;
  int32_t *hfirst, *hi, *hj, *hS;
  void* hscr;
  size_t ns = nl_to_host_scratch_bytes(P);
  cudaMalloc(&di, (size_t)P * 4); cudaMalloc(&dj, (size_t)P * 4); cudaMalloc(&dS, (size_t)P * 12); cudaMalloc(&dscr, ns);
  cudaMallocHost((void**)&hfirst, (size_t)(N + 1) * 4); cudaMallocHost((void**)&hi, (size_t)P * 4);
  cudaMallocHost((void**)&hj, (size_t)P * 4); cudaMallocHost((void**)&hS, (size_t)P * 12); cudaMallocHost(&hscr, ns);

  /* the transfer starts while the fill pass runs: `first` is copied and i rebuilt by host threads of the library */
  nl_to_host_job* job = NULL;
  CHECK(nl_pairs_to_host_begin(&p, dfirst, N, P, hfirst, hi, 0, NULL, &job));
  CHECK(nl_fill_pairs(&p, dXs, N, dperm, dco, dfirst, di, dj, dS, NULL, wsp, np, NULL));
  CHECK(nl_pairs_to_host_finish(job, dj, dS, hj, hS, dscr, hscr, ns, NULL));

  printf("%lld atoms, %lld pairs (%.1f per atom); first pair: i=%d j=%d S=(%d,%d,%d)\n", (long long)N, (long long)P, (double)P / (double)N,
         hi[0], hj[0], hS[0], hS[1], hS[2]);
  return 0;
}
