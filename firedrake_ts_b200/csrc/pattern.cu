// pattern.cu -- one-off device construction of the AIJ sparsity and index maps.
//
// Replaces what the reference obtains on first touch of ctx._jac in
// _TSContext.set_ijacobian (firedrake_ts/solving_utils.py:120-122): Firedrake's
// MatrixAssembler.allocate builds the op2.Sparsity and preallocates the PETSc
// AIJ Mat [upstream].  Here the node-level CSR pattern (owned node rows x local
// node columns, sorted columns -- what PETSc AIJ stores), the row->cell
// adjacency and the row-relative cell->CSR-slot map are built on the GPU from
// the cell->node map alone; integer work, bit-exact against the oracle's
// oracle_node_pattern (tests/test_pattern_gpu.py).
#include <cub/cub.cuh>

#include "fdts_common.cuh"

namespace {

__global__ void count_incidence(const int32_t *__restrict__ cell_node_t, int64_t nc, int nd, int64_t nown,
                                int32_t *__restrict__ cnt) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nd) return;
  const int32_t n = cell_node_t[t];
  if (n < nown) atomicAdd(cnt + n, 1);
}

__global__ void fill_incidence(const int32_t *__restrict__ cell_node_t, int64_t nc, int nd, int64_t nown,
                               const int64_t *__restrict__ adj_ptr, int32_t *__restrict__ cursor,
                               int32_t *__restrict__ adj_cell) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nd) return;
  const int32_t n = cell_node_t[t];
  if (n >= nown) return;
  const int i = (int)(t / nc);
  const int64_t c = t - (int64_t)i * nc;
  const int p = atomicAdd(cursor + n, 1);
  adj_cell[adj_ptr[n] + p] = (int32_t)(c * 32 + i);
}

// make the adjacency order deterministic (ascending cell): insertion sort of short lists
__global__ void sort_incidence(const int64_t *__restrict__ adj_ptr, int64_t nown, int32_t *__restrict__ adj_cell) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nown) return;
  const int64_t b = adj_ptr[n], e = adj_ptr[n + 1];
  for (int64_t i = b + 1; i < e; ++i) {
    const int32_t v = adj_cell[i];
    int64_t j = i - 1;
    while (j >= b && adj_cell[j] > v) {
      adj_cell[j + 1] = adj_cell[j];
      --j;
    }
    adj_cell[j + 1] = v;
  }
}

// one warp per owned node row: gather candidate columns, bitonic sort, unique.
template <bool WRITE>
__global__ void row_pattern(const int32_t *__restrict__ cell_node_t, int64_t nc, int nd, int64_t nown,
                            const int64_t *__restrict__ adj_ptr, const int32_t *__restrict__ adj_cell, int kpad,
                            int32_t *__restrict__ rowlen, const int64_t *__restrict__ nrowptr,
                            int32_t *__restrict__ ncolidx) {
  extern __shared__ int32_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= nown) return;
  int32_t *s = sm + (size_t)warp * kpad;
  const int64_t b = adj_ptr[row];
  const int deg = (int)(adj_ptr[row + 1] - b);
  const int k = deg * nd;
  for (int t = lane; t < kpad; t += 32) {
    int32_t v = INT32_MAX;
    if (t < k) {
      const int a = t / nd, j = t - a * nd;
      const int64_t cell = adj_cell[b + a] >> 5;
      v = cell_node_t[(int64_t)j * nc + cell];
    }
    s[t] = v;
  }
  __syncwarp();
  for (int kk = 2; kk <= kpad; kk <<= 1)
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < kpad; t += 32) {
        const int p = t ^ j;
        if (p > t) {
          const int32_t x = s[t], y = s[p];
          const bool up = (t & kk) == 0;
          if ((x > y) == up) {
            s[t] = y;
            s[p] = x;
          }
        }
      }
      __syncwarp();
    }
  int count = 0;
  const int64_t obase = WRITE ? nrowptr[row] : 0;
  for (int t0 = 0; t0 < kpad; t0 += 32) {
    const int t = t0 + lane;
    const int32_t v = s[t];
    const bool first = (v != INT32_MAX) && (t == 0 || s[t - 1] != v);
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (WRITE && first) ncolidx[obase + count + __popc(m & ((1u << lane) - 1u))] = v;
    count += __popc(m);
  }
  if (!WRITE && lane == 0) rowlen[row] = count;
}

struct Widen {
  __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};
using WideIt = cub::TransformInputIterator<int64_t, Widen, const int32_t *>;

// row-relative slot map: position of node j of cell c inside node row i of cell c
__global__ void slot_positions(const int32_t *__restrict__ cell_node_t, int64_t nc, int nd, int64_t nown,
                               const int64_t *__restrict__ nrowptr, const int32_t *__restrict__ ncolidx,
                               uint8_t *__restrict__ pos8_t, int *__restrict__ err) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nd) return;
  const int i = (int)(t / nc);
  const int64_t c = t - (int64_t)i * nc;
  const int32_t rn = cell_node_t[t];
  if (rn >= nown) {
    for (int j = 0; j < nd; ++j) pos8_t[(int64_t)(i * nd + j) * nc + c] = 255;
    return;
  }
  const int64_t b = nrowptr[rn];
  const int len = (int)(nrowptr[rn + 1] - b);
  for (int j = 0; j < nd; ++j) {
    const int32_t col = cell_node_t[(int64_t)j * nc + c];
    int lo = 0, hi = len - 1, found = -1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const int32_t v = ncolidx[b + mid];
      if (v == col) {
        found = mid;
        break;
      }
      if (v < col) lo = mid + 1; else hi = mid - 1;
    }
    if (found < 0 || found > 254) atomicExch(err, 1);
    pos8_t[(int64_t)(i * nd + j) * nc + c] = (uint8_t)found;
  }
}

// per cell: bit i set iff local node i is an owned, non-Dirichlet row (one coalesced 32-bit load in the
// scatter kernels instead of nd dependent byte loads of isbc)
__global__ void cell_row_masks(const int32_t *__restrict__ cell_node_t, int64_t nc, int nd, int64_t nown,
                               const uint8_t *__restrict__ isbc, uint32_t *__restrict__ rowmask) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  uint32_t m = 0;
  for (int i = 0; i < nd; ++i) {
    const int32_t n = cell_node_t[(int64_t)i * nc + c];
    if (n < nown && !isbc[n]) m |= 1u << i;
  }
  rowmask[c] = m;
}

__global__ void mark_bc(const int32_t *__restrict__ bc_nodes, int64_t nbc, uint8_t *__restrict__ isbc) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nbc) isbc[bc_nodes[t]] = 1;
}

// dof-level slot of the diagonal entry of every owned Dirichlet dof
__global__ void bc_diag_slots(const int32_t *__restrict__ bc_nodes, int64_t nbc, int ncomp, int layout, int64_t nown,
                              int64_t node_nnz, const int64_t *__restrict__ nrowptr,
                              const int32_t *__restrict__ ncolidx, int64_t *__restrict__ slots) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nbc) return;
  const int32_t n = bc_nodes[t];
  if (n >= nown) {
    for (int m = 0; m < ncomp; ++m) slots[t * ncomp + m] = -1;
    return;
  }
  const int64_t b = nrowptr[n];
  const int64_t len = nrowptr[n + 1] - b;
  int64_t pos = 0;
  for (int64_t k = 0; k < len; ++k)
    if (ncolidx[b + k] == n) pos = k;
  for (int m = 0; m < ncomp; ++m) {
    int64_t s;
    if (layout == FDTS_INTERLEAVED)
      s = b * ncomp * ncomp + (int64_t)m * len * ncomp + pos * ncomp + m;
    else
      s = (int64_t)m * ncomp * node_nnz + (int64_t)ncomp * b + (int64_t)m * len + pos;
    slots[t * ncomp + m] = s;
  }
}

}  // namespace

#define GRID(n, b) (unsigned)(((n) + (b)-1) / (b))

int fdts_build_pattern(fdts_engine *e, cudaStream_t s) {
  const int64_t nc = e->cfg.num_cells, nown = e->cfg.num_owned_nodes, nn = e->cfg.num_nodes;
  const int nd = e->cfg.nd;
  if (nc >= (1ll << 26)) {
    fdts_set_error("more than 2^26 local cells per rank are not supported by the adjacency encoding");
    return 4;
  }
  int32_t *cnt = nullptr, *rowlen = nullptr;
  int *derr = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  FDTS_CHECK(cudaMalloc(&cnt, sizeof(int32_t) * (nown + 1)));
  FDTS_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (nown + 1), s));
  FDTS_CHECK(cudaMalloc(&derr, sizeof(int)));
  FDTS_CHECK(cudaMemsetAsync(derr, 0, sizeof(int), s));
  if (nc * nd > 0) count_incidence<<<GRID(nc * nd, 256), 256, 0, s>>>(e->cell_node, nc, nd, nown, cnt);
  FDTS_CHECK(cudaMalloc(&e->adj_ptr, sizeof(int64_t) * (nown + 4)));
  // exclusive scan int32 -> int64 (cnt has nown+1 entries, the last one zero)
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, WideIt(cnt, Widen()), e->adj_ptr, nown + 1, s);
  size_t tb2 = 0;
  int32_t *dmax = nullptr;
  FDTS_CHECK(cudaMalloc(&dmax, sizeof(int32_t) * 2));
  cub::DeviceReduce::Max(nullptr, tb2, cnt, dmax, nown + 1, s);
  tmp_bytes = tmp_bytes > tb2 ? tmp_bytes : tb2;
  FDTS_CHECK(cudaMalloc(&tmp, tmp_bytes + 16));
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, WideIt(cnt, Widen()), e->adj_ptr, nown + 1, s);
  cub::DeviceReduce::Max(tmp, tmp_bytes, cnt, dmax, nown + 1, s);
  int64_t total_adj = 0;
  int32_t maxdeg = 0;
  FDTS_CHECK(cudaMemcpyAsync(&total_adj, e->adj_ptr + nown, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  FDTS_CHECK(cudaMemcpyAsync(&maxdeg, dmax, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FDTS_CHECK(cudaStreamSynchronize(s));
  e->max_degree = maxdeg;
  FDTS_CHECK(cudaMalloc(&e->adj_cell, sizeof(int32_t) * (total_adj > 0 ? total_adj : 1)));
  FDTS_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (nown + 1), s));
  if (nc * nd > 0)
    fill_incidence<<<GRID(nc * nd, 256), 256, 0, s>>>(e->cell_node, nc, nd, nown, e->adj_ptr, cnt, e->adj_cell);
  if (nown > 0) sort_incidence<<<GRID(nown, 128), 128, 0, s>>>(e->adj_ptr, nown, e->adj_cell);

  // row lengths, then columns
  int kpad = 32;
  while (kpad < maxdeg * nd) kpad <<= 1;
  const int warps = kpad <= 2048 ? 4 : 1;
  const size_t smem = sizeof(int32_t) * (size_t)kpad * warps;
  if (smem > 200 * 1024) {
    fdts_set_error("node degree too large for the sparsity builder");
    return 4;
  }
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(row_pattern<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(row_pattern<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  FDTS_CHECK(cudaMalloc(&rowlen, sizeof(int32_t) * (nown + 1)));
  FDTS_CHECK(cudaMemsetAsync(rowlen, 0, sizeof(int32_t) * (nown + 1), s));
  if (nown > 0)
    row_pattern<false><<<GRID(nown, warps), warps * 32, smem, s>>>(e->cell_node, nc, nd, nown, e->adj_ptr,
                                                                     e->adj_cell, kpad, rowlen, nullptr, nullptr);
  FDTS_CHECK(cudaMalloc(&e->nrowptr, sizeof(int64_t) * (nown + 1)));
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, WideIt(rowlen, Widen()), e->nrowptr, nown + 1, s);
  cub::DeviceReduce::Max(tmp, tmp_bytes, rowlen, dmax, nown + 1, s);
  int32_t maxrow = 0;
  FDTS_CHECK(cudaMemcpyAsync(&e->node_nnz, e->nrowptr + nown, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  FDTS_CHECK(cudaMemcpyAsync(&maxrow, dmax, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FDTS_CHECK(cudaStreamSynchronize(s));
  e->max_row_nodes = maxrow;
  if (maxrow > 255) {
    fdts_set_error("node rows longer than 255 entries are not supported by the uint8 slot map");
    return 4;
  }
  FDTS_CHECK(cudaMalloc(&e->ncolidx, sizeof(int32_t) * (e->node_nnz > 0 ? e->node_nnz : 1)));
  if (nown > 0)
    row_pattern<true><<<GRID(nown, warps), warps * 32, smem, s>>>(e->cell_node, nc, nd, nown, e->adj_ptr,
                                                                    e->adj_cell, kpad, nullptr, e->nrowptr,
                                                                    e->ncolidx);
  // row-relative slot map
  FDTS_CHECK(cudaMalloc(&e->pos8, (size_t)(nc * nd * nd > 0 ? nc * nd * nd : 1)));
  if (nc * nd > 0)
    slot_positions<<<GRID(nc * nd, 256), 256, 0, s>>>(e->cell_node, nc, nd, nown, e->nrowptr, e->ncolidx, e->pos8,
                                                       derr);
  // boundary conditions
  FDTS_CHECK(cudaMalloc(&e->isbc, (size_t)(nn + 1)));
  FDTS_CHECK(cudaMemsetAsync(e->isbc, 0, (size_t)(nn + 1), s));
  const int64_t nbc = e->cfg.num_bc_nodes;
  if (nbc > 0) {
    mark_bc<<<GRID(nbc, 256), 256, 0, s>>>(e->bc_nodes, nbc, e->isbc);
    FDTS_CHECK(cudaMalloc(&e->bc_diag_slot, sizeof(int64_t) * nbc * e->cfg.ncomp));
    bc_diag_slots<<<GRID(nbc, 256), 256, 0, s>>>(e->bc_nodes, nbc, e->cfg.ncomp, e->cfg.layout, nown, e->node_nnz,
                                                  e->nrowptr, e->ncolidx, e->bc_diag_slot);
  }
  FDTS_CHECK(cudaMalloc(&e->cell_rowmask, sizeof(uint32_t) * (size_t)(nc > 0 ? nc : 1)));
  if (nc > 0) cell_row_masks<<<GRID(nc, 256), 256, 0, s>>>(e->cell_node, nc, nd, nown, e->isbc, e->cell_rowmask);
  int herr = 0;
  FDTS_CHECK(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, s));
  FDTS_CHECK(cudaStreamSynchronize(s));
  FDTS_CHECK(cudaGetLastError());
  cudaFree(cnt);
  cudaFree(rowlen);
  cudaFree(derr);
  cudaFree(dmax);
  cudaFree(tmp);
  if (herr) {
    fdts_set_error("internal error: cell entry outside the sparsity pattern");
    return 5;
  }
  return 0;
}
