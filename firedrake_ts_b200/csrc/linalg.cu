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

// CSR-stream mat-vec: a CTA owns NT consecutive rows; the non-zeros of those rows are one contiguous
// range of vals / colidx, which is read with fully coalesced loads (U independent loads in flight per
// thread), multiplied with the gathered x and parked in shared memory; then thread t sums the segment
// of row t.  HBM traffic is the 12 bytes per non-zero of the CSR arrays plus the x gather (L2 hits on
// tile-numbered spaces), without the partially used sectors and the 5 shuffles per row of the
// warp-per-row kernel.
template <int NT, int U>
__global__ void __launch_bounds__(NT) spmv_stream(int64_t nrows, const int64_t *__restrict__ rowptr,
                                                  const int32_t *__restrict__ colidx, const double *__restrict__ vals,
                                                  const double *__restrict__ x, double *__restrict__ y) {
  constexpr int CH = NT * U;
  __shared__ double prod[CH];
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * NT;
  const int64_t r1 = min(nrows, r0 + NT);
  const int64_t k0 = rowptr[r0], k1 = rowptr[r1];
  const int64_t r = r0 + tid;
  int64_t ra = 0, rb = 0;
  if (r < r1) {
    ra = rowptr[r];
    rb = rowptr[r + 1];
  }
  double acc = 0.0;
  for (int64_t kb = k0; kb < k1; kb += CH) {
    double v[U];
    int32_t c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t k = kb + u * NT + tid;
      c[u] = k < k1 ? __ldcs(colidx + k) : -1;
      v[u] = k < k1 ? __ldcs(vals + k) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) prod[u * NT + tid] = c[u] >= 0 ? v[u] * __ldg(x + c[u]) : 0.0;
    __syncthreads();
    const int64_t lo = max(ra, kb), hi = min(rb, kb + CH);
    for (int64_t k = lo; k < hi; ++k) acc += prod[k - kb];
    __syncthreads();
  }
  if (r < r1) y[r] = acc;
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
  constexpr int NT = 128;
  spmv_stream<NT, 8><<<(unsigned)((nrows + NT - 1) / NT), NT, 0, (cudaStream_t)stream>>>(nrows, rowptr, colidx, vals, x, y);
  FDTS_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fdts_csr_spmv_warp(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals,
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
