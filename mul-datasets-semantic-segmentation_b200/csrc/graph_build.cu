// graph_build.cu — CSR / CSC / quad-padded CSR index lists of a column-one-hot 0/1 bipartite graph, built ON THE DEVICE.
//
// Reference context: the SEG stage's bi_graphs are 0/1 matrices [C_ds, C_uni] with at most one 1 per column (UOT /
// pretrain graphs, lib/models/ltbgnn_direct_learn.py:426-439; ClassRemap.getRemapMatrix of a single-label remap,
// lib/class_remap.py:176-183).  The kernels of proj.cu / mds_bwd.cu walk them as index lists.  Building those lists on
// the host needs a device -> host copy of the matrix whenever the caller hands over a new tensor (a trainer that
// rebuilds its graphs every iteration, e.g. the EMA graphs of lib/models/ltbgnn_sfg.py, synchronises the stream once per
// dataset per step).  Here the caller DECLARES the kind, the lists are built by one CTA per graph with no host
// involvement, and a matrix that is not column-one-hot 0/1 raises MDSEG_ERR_GRAPH_KIND in the error flag.
//
// Order of the lists (identical to the host builder of ops.BipartiteGraphs, so results are bit-identical):
//   csr_col   rows ascending, columns ascending inside a row          [C_uni] (nnz <= C_uni used)
//   csc_row   columns ascending                                       [C_uni]
//   csr4_col  as csr_col, every row padded to whole quads with its last column   [C_uni + 3 * C_ds]
#include "common.cuh"

namespace mdseg {
namespace {

constexpr int kThreads = 1024;

__global__ void __launch_bounds__(kThreads)
graph_build_onehot_kernel(const float* __restrict__ G, int C_ds, int C_uni, int* __restrict__ csr_ptr,
                          int* __restrict__ csr_col, int* __restrict__ csc_ptr, int* __restrict__ csc_row,
                          int* __restrict__ csr4_ptr, int* __restrict__ csr4_col, int* __restrict__ row_of /*[C_uni] scratch*/,
                          int* err_flag) {
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  // 1. the row of every column (-1: empty column); anything but a single 1 breaks the declared kind
  for (int u = threadIdx.x; u < C_uni; u += blockDim.x) {
    int r = -1, bad = 0;
    for (int n = 0; n < C_ds; ++n) {
      const float v = G[(int64_t)n * C_uni + u];
      if (v != 0.f) {
        if (r >= 0 || v != 1.f) bad = 1;
        r = n;
      }
    }
    row_of[u] = r;
    if (bad) s_bad = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_bad && err_flag) atomicOr(err_flag, MDSEG_ERR_GRAPH_KIND);
  // 2. one thread per row: its columns in ascending order (C_ds * C_uni <= a few 10^5 reads of a cached vector)
  for (int n = threadIdx.x; n < C_ds; n += blockDim.x) {
    int cnt = 0;
    for (int u = 0; u < C_uni; ++u) cnt += row_of[u] == n ? 1 : 0;
    csr_ptr[n + 1] = cnt;  // counts first, prefix sums below
    csr4_ptr[n + 1] = (cnt + 3) >> 2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // C_ds <= a few hundred: a serial scan is a few microseconds
    csr_ptr[0] = 0; csr4_ptr[0] = 0;
    for (int n = 0; n < C_ds; ++n) { csr_ptr[n + 1] += csr_ptr[n]; csr4_ptr[n + 1] += csr4_ptr[n]; }
    int acc = 0;
    csc_ptr[0] = 0;
    for (int u = 0; u < C_uni; ++u) {
      if (row_of[u] >= 0) csc_row[acc++] = row_of[u];
      csc_ptr[u + 1] = acc;
    }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < C_ds; n += blockDim.x) {
    int o = csr_ptr[n], o4 = 4 * csr4_ptr[n], last = -1;
    for (int u = 0; u < C_uni; ++u)
      if (row_of[u] == n) { csr_col[o++] = u; csr4_col[o4++] = u; last = u; }
    for (const int e4 = 4 * csr4_ptr[n + 1]; o4 < e4; ++o4) csr4_col[o4] = last;
  }
}

}  // namespace
}  // namespace mdseg

extern "C" size_t mdseg_graph_build_onehot_ints(int C_ds, int C_uni) {
  // csr_ptr, csc_ptr, csr4_ptr, csr_col, csc_row, csr4_col, row_of scratch — in this order, each rounded up to 4 ints
  auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
  return r4(C_ds + 1) + r4(C_uni + 1) + r4(C_ds + 1) + r4(C_uni) + r4(C_uni) + r4((size_t)C_uni + 3 * (size_t)C_ds) + r4(C_uni);
}

extern "C" int mdseg_graph_build_onehot(const float* G, int C_ds, int C_uni, int* buf, int32_t* err_flag, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(G && buf && C_ds > 0 && C_uni > 0, "mdseg_graph_build_onehot: bad arguments");
  auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
  int* csr_ptr = buf;
  int* csc_ptr = csr_ptr + r4(C_ds + 1);
  int* csr4_ptr = csc_ptr + r4(C_uni + 1);
  int* csr_col = csr4_ptr + r4(C_ds + 1);
  int* csc_row = csr_col + r4(C_uni);
  int* csr4_col = csc_row + r4(C_uni);
  int* row_of = csr4_col + r4((size_t)C_uni + 3 * (size_t)C_ds);
  graph_build_onehot_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(G, C_ds, C_uni, csr_ptr, csr_col, csc_ptr, csc_row,
                                                                      csr4_ptr, csr4_col, row_of, err_flag);
  MDSEG_LAUNCH_OK();
  return 0;
}
