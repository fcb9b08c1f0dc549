// linalg.cu -- CSR mat-vec and diagonal extraction on the device-resident AIJ
// arrays.  Not part of the assembly hot path: they serve the standalone
// Newton-Krylov harness (firedrake_ts_b200/ts.py), i.e. "the step after the
// path" of SURVEY.md section 8(f) rank 4.
#include "fdts_common.cuh"

namespace {
__global__ void spmv_warp_per_row(int64_t nrows, const int64_t *__restrict__ rowptr,
                                  const int32_t *__restrict__ colidx, const double *__restrict__ vals,
                                  const double *__restrict__ x, double *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  double acc = 0.0;
  for (int64_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) acc += vals[k] * __ldg(x + colidx[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc;
}

__global__ void diag_kernel(int64_t nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                            const double *__restrict__ vals, int64_t nown, int64_t nnodes, int ncomp, int layout,
                            double *__restrict__ d) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  int64_t col = row;
  if (layout == FDTS_BLOCKED) {
    const int64_t m = row / nown;
    col = m * nnodes + (row - m * nown);
  }
  double v = 0.0;
  for (int64_t k = rowptr[row]; k < rowptr[row + 1]; ++k)
    if (colidx[k] == col) v = vals[k];
  d[row] = v;
}
}  // namespace

extern "C" int fdts_csr_spmv(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals,
                             const double *x, double *y, void *stream) {
  if (nrows <= 0) return 0;
  const int threads = 256;
  const int64_t blocks = (nrows * 32 + threads - 1) / threads;
  spmv_warp_per_row<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(nrows, rowptr, colidx, vals, x, y);
  FDTS_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fdts_csr_diagonal(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals,
                                 int64_t nown_nodes, int64_t nnodes, int32_t ncomp, int32_t layout, double *d,
                                 void *stream) {
  if (nrows <= 0) return 0;
  diag_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nrows, rowptr, colidx, vals,
                                                                                 nown_nodes, nnodes, ncomp, layout, d);
  FDTS_CHECK(cudaGetLastError());
  return 0;
}
