// gather_plan.cu -- one-off construction of the owner-computes gather plan (K6).
//
// The deterministic scatter variant of SURVEY.md section 8(a) ("K6 deterministic
// segmented scatter variant (sorted contribution lists)").  Owned node rows are
// cut into contiguous row blocks (one CTA each).  For every block the plan holds
//   * the sorted list of cells touching its rows (the CTA recomputes their
//     element matrices into shared memory -- overlap-1 redundancy, exactly like
//     the exec-halo of the MPI decomposition, only at CTA scale);
//   * for every CSR slot of the block, the list of staged element-matrix entries
//     that sum to it, as 16-bit shared-memory addresses in ascending cell order
//     (fixed order => bitwise reproducible sums), stored SELL-32-sigma style: inside
//     windows of 2048 slots the slots are sorted by contribution count, cut into
//     slices of 32, and each slice stores its addresses column-major, padded to the
//     slice's longest list with the address of a staged zero -- so the summation
//     loop of a warp is divergence-free and its index loads are coalesced;
//   * for every row, the list of staged element-vector entries (residual).
// Dirichlet rows/columns get empty lists (PyOP2's negative-lgmap "drop").
// All of it is integer work on the GPU, derived from the cell->node map, the
// node-level CSR and the row->cell adjacency built in pattern.cu.
#include <cub/cub.cuh>

#include "fdts_common.cuh"

namespace {

constexpr int PLAN_THREADS = 256;

__global__ void block_candidates(const int64_t *__restrict__ bstart, int64_t nb, const int64_t *__restrict__ adj_ptr,
                                 int64_t *__restrict__ ncand) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb) ncand[b] = adj_ptr[bstart[b + 1]] - adj_ptr[bstart[b]];
}

// one CTA per row block: sorted unique cells adjacent to the block's rows
template <bool WRITE>
__global__ void __launch_bounds__(PLAN_THREADS) block_cells(const int64_t *__restrict__ bstart,
                                                            const int64_t *__restrict__ adj_ptr,
                                                            const int32_t *__restrict__ adj_cell, int kpad, int cmax,
                                                            int32_t *__restrict__ bcount,
                                                            int32_t *__restrict__ bcells) {
  extern __shared__ int32_t cand[];
  __shared__ int warp_tot[PLAN_THREADS / 32];
  const int64_t b = blockIdx.x;
  const int64_t a0 = adj_ptr[bstart[b]];
  const int n = (int)(adj_ptr[bstart[b + 1]] - a0);
  for (int t = threadIdx.x; t < kpad; t += PLAN_THREADS) cand[t] = t < n ? (adj_cell[a0 + t] >> 5) : INT32_MAX;
  __syncthreads();
  for (int kk = 2; kk <= kpad; kk <<= 1)
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kpad; t += PLAN_THREADS) {
        const int p = t ^ j;
        if (p > t) {
          const int32_t x = cand[t], y = cand[p];
          const bool up = (t & kk) == 0;
          if ((x > y) == up) {
            cand[t] = y;
            cand[p] = x;
          }
        }
      }
      __syncthreads();
    }
  // ordered compaction of first occurrences, PLAN_THREADS elements per round
  int base = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < kpad; t0 += PLAN_THREADS) {
    const int t = t0 + threadIdx.x;
    const int32_t v = cand[t];
    const bool first = v != INT32_MAX && (t == 0 || cand[t - 1] != v);
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (WRITE && first) bcells[b * cmax + off + __popc(m & ((1u << lane) - 1u))] = v;
    for (int w = 0; w < PLAN_THREADS / 32; ++w) base += warp_tot[w];
    __syncthreads();
  }
  if (!WRITE && threadIdx.x == 0) bcount[b] = base;
}

__device__ __forceinline__ int sym_index(int i, int j, int nd) {
  if (i > j) {
    const int t = i;
    i = j;
    j = t;
  }
  return i * nd - (i * (i - 1)) / 2 + (j - i);
}

__device__ __forceinline__ int find_cell(const int32_t *__restrict__ list, int n, int32_t cell) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int32_t v = list[mid];
    if (v == cell) return mid;
    if (v < cell) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// MODE 0: count contributions per slot (cnt8).  MODE 1: write the staged addresses.
template <int MODE>
__global__ void slot_contributions(const int64_t *__restrict__ bstart, const int64_t *__restrict__ adj_ptr,
                                   const int32_t *__restrict__ adj_cell, const int32_t *__restrict__ cell_node_t,
                                   int64_t nc, int nd, int nep, const uint8_t *__restrict__ pos8_t,
                                   const uint8_t *__restrict__ isbc, const int64_t *__restrict__ nrp,
                                   uint8_t *__restrict__ cnt8, const uint32_t *__restrict__ sellpos,
                                   const int64_t *__restrict__ bbase, const int32_t *__restrict__ bcells,
                                   const int32_t *__restrict__ bcount, int cmax, uint16_t *__restrict__ cidx,
                                   int *__restrict__ err) {
  const int64_t b = blockIdx.x;
  const int64_t r0 = bstart[b], r1 = bstart[b + 1];
  for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    if (isbc[r]) continue;
    const int64_t rb = nrp[r];
    for (int64_t a = adj_ptr[r]; a < adj_ptr[r + 1]; ++a) {
      const int32_t code = adj_cell[a];
      const int64_t cell = code >> 5;
      const int i = code & 31;
      int cs = 0;
      if (MODE == 1) {
        cs = find_cell(bcells + b * cmax, bcount[b], (int32_t)cell);
        if (cs < 0) {
          atomicExch(err, 2);
          continue;
        }
      }
      for (int j = 0; j < nd; ++j) {
        if (isbc[cell_node_t[(int64_t)j * nc + cell]]) continue;
        const int64_t slot = rb + pos8_t[(int64_t)(i * nd + j) * nc + cell];
        if (MODE == 0) {
          if (cnt8[slot] == 255) atomicExch(err, 3);
          cnt8[slot]++;
        } else {
          const int k = cnt8[slot]++;
          cidx[bbase[b] + sellpos[slot] + (int64_t)k * 32] = (uint16_t)(cs * nep + sym_index(i, j, nd));
        }
      }
    }
  }
}

constexpr int SELL_WINDOW = 2048;

__global__ void block_slices(const int64_t *__restrict__ bstart, int64_t nb, const int64_t *__restrict__ nrp,
                             int64_t *__restrict__ nsl_padded) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int64_t n = nrp[bstart[b + 1]] - nrp[bstart[b]];
  const int64_t full = n / SELL_WINDOW, rem = n - full * SELL_WINDOW;
  const int64_t nsl = full * (SELL_WINDOW / 32) + (rem + 31) / 32;
  nsl_padded[b] = (nsl + 1 + 3) & ~3ll;  // +1 end marker, 16-byte granules
}

// one CTA per row block: SELL-32-sigma layout of its slots (sigma = SELL_WINDOW)
__global__ void __launch_bounds__(PLAN_THREADS) sell_layout(const int64_t *__restrict__ bstart,
                                                            const int64_t *__restrict__ nrp,
                                                            const uint8_t *__restrict__ cnt8,
                                                            const int64_t *__restrict__ slbase,
                                                            uint16_t *__restrict__ sdst, uint32_t *__restrict__ soff,
                                                            uint32_t *__restrict__ sellpos,
                                                            int64_t *__restrict__ btotal) {
  __shared__ uint32_t keys[SELL_WINDOW];
  __shared__ uint32_t woff[SELL_WINDOW / 32 + 1];
  __shared__ uint32_t running_s;
  const int64_t b = blockIdx.x;
  const int64_t s0 = nrp[bstart[b]], s1 = nrp[bstart[b + 1]];
  const int64_t nslots = s1 - s0;
  uint32_t *boff = soff + slbase[b];
  if (threadIdx.x == 0) running_s = 0;
  int64_t slice0 = 0;
  for (int64_t wb = 0; wb < nslots; wb += SELL_WINDOW) {
    const int nw = (int)((nslots - wb) < SELL_WINDOW ? (nslots - wb) : SELL_WINDOW);
    for (int t = threadIdx.x; t < SELL_WINDOW; t += PLAN_THREADS)
      keys[t] = t < nw ? (((uint32_t)(255 - cnt8[s0 + wb + t]) << 16) | (uint32_t)t) : 0xFFFFFFFFu;
    __syncthreads();
    for (int kk = 2; kk <= SELL_WINDOW; kk <<= 1)
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < SELL_WINDOW; t += PLAN_THREADS) {
          const int q = t ^ j;
          if (q > t) {
            const uint32_t x = keys[t], y = keys[q];
            const bool up = (t & kk) == 0;
            if ((x > y) == up) {
              keys[t] = y;
              keys[q] = x;
            }
          }
        }
        __syncthreads();
      }
    const int nsl = (nw + 31) / 32;
    if (threadIdx.x == 0) {
      uint32_t run = running_s;
      for (int j = 0; j < nsl; ++j) {
        woff[j] = run;
        run += 32u * (255u - (keys[32 * j] >> 16));
      }
      woff[nsl] = run;
      running_s = run;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nsl; j += PLAN_THREADS) boff[slice0 + j] = woff[j];
    for (int q = threadIdx.x; q < nw; q += PLAN_THREADS) {
      const uint32_t t = keys[q] & 0xFFFFu;
      sdst[s0 + wb + q] = (uint16_t)t;
      sellpos[s0 + wb + t] = woff[q >> 5] + (uint32_t)(q & 31);
    }
    slice0 += nsl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    boff[slice0] = running_s;
    btotal[b] = running_s;  // a multiple of 32 entries = 64 bytes
  }
}

__global__ void fill_u16(uint16_t *__restrict__ p, int64_t n, uint16_t v) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = v;
}

// residual lists: per row the staged element-vector entries, cell ascending
template <int MODE>
__global__ void row_contributions(const int64_t *__restrict__ bstart, const int64_t *__restrict__ adj_ptr,
                                  const int32_t *__restrict__ adj_cell, const uint8_t *__restrict__ isbc, int ndp,
                                  const int32_t *__restrict__ bcells, const int32_t *__restrict__ bcount, int cmax,
                                  uint16_t *__restrict__ ridx, int *__restrict__ err) {
  const int64_t b = blockIdx.x;
  const int64_t r0 = bstart[b], r1 = bstart[b + 1];
  for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    for (int64_t a = adj_ptr[r]; a < adj_ptr[r + 1]; ++a) {
      const int32_t code = adj_cell[a];
      const int cs = find_cell(bcells + b * cmax, bcount[b], code >> 5);
      if (cs < 0) {
        atomicExch(err, 5);
        continue;
      }
      ridx[a] = (uint16_t)(cs * ndp + (code & 31));
    }
    (void)isbc;
  }
}

__global__ void block_extents(const int64_t *__restrict__ bstart, int64_t nb, const int64_t *__restrict__ nrp,
                              const int64_t *__restrict__ adj_ptr, const int64_t *__restrict__ bbase,
                              const int64_t *__restrict__ slbase, int64_t *__restrict__ out4) {
  // out4: max contributions, max slots, max rows, max adjacency entries, max padded slice entries per block
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int64_t r0 = bstart[b], r1 = bstart[b + 1];
  atomicMax((unsigned long long *)&out4[0], (unsigned long long)(bbase[b + 1] - bbase[b]));
  atomicMax((unsigned long long *)&out4[1], (unsigned long long)(nrp[r1] - nrp[r0]));
  atomicMax((unsigned long long *)&out4[2], (unsigned long long)(r1 - r0));
  atomicMax((unsigned long long *)&out4[3], (unsigned long long)(adj_ptr[r1] - adj_ptr[r0]));
  atomicMax((unsigned long long *)&out4[4], (unsigned long long)(slbase[b + 1] - slbase[b]));
}

// one CTA per row block: block-local vertex table (coordinates copied) and per-cell local vertex ids
template <bool WRITE>
__global__ void __launch_bounds__(PLAN_THREADS) block_vertices(const int32_t *__restrict__ bcells,
                                                               const int32_t *__restrict__ bcount, int cmax,
                                                               const int32_t *__restrict__ cell_coord_t, int64_t nc,
                                                               int dim, const double *__restrict__ coords, int kpad,
                                                               int vmax, int32_t *__restrict__ vcount,
                                                               double *__restrict__ bvcoords,
                                                               uint32_t *__restrict__ bcc, int *__restrict__ err) {
  extern __shared__ int32_t cand[];  // kpad candidates, then kpad unique
  int32_t *uniq = cand + kpad;
  __shared__ int warp_tot[PLAN_THREADS / 32];
  const int64_t b = blockIdx.x;
  const int ncell = bcount[b];
  const int nv1 = dim + 1;
  const int n = ncell * nv1;
  for (int t = threadIdx.x; t < kpad; t += PLAN_THREADS) {
    int32_t v = INT32_MAX;
    if (t < n) {
      const int k = t / nv1, a = t - k * nv1;
      v = cell_coord_t[(int64_t)a * nc + bcells[b * cmax + k]];
    }
    cand[t] = v;
  }
  __syncthreads();
  for (int kk = 2; kk <= kpad; kk <<= 1)
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kpad; t += PLAN_THREADS) {
        const int p = t ^ j;
        if (p > t) {
          const int32_t x = cand[t], y = cand[p];
          const bool up = (t & kk) == 0;
          if ((x > y) == up) {
            cand[t] = y;
            cand[p] = x;
          }
        }
      }
      __syncthreads();
    }
  int base = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < kpad; t0 += PLAN_THREADS) {
    const int t = t0 + threadIdx.x;
    const int32_t v = cand[t];
    const bool first = v != INT32_MAX && (t == 0 || cand[t - 1] != v);
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (first) uniq[off + __popc(m & ((1u << lane) - 1u))] = v;
    for (int w = 0; w < PLAN_THREADS / 32; ++w) base += warp_tot[w];
    __syncthreads();
  }
  const int nv = base;
  if (!WRITE) {
    if (threadIdx.x == 0) vcount[b] = nv;
    return;
  }
  if (nv > 255) {
    if (threadIdx.x == 0) atomicExch(err, 6);
    return;
  }
  for (int t = threadIdx.x; t < nv * dim; t += PLAN_THREADS) {
    const int lv = t / dim, d = t - lv * dim;
    bvcoords[((int64_t)b * vmax + lv) * dim + d] = coords[(int64_t)uniq[lv] * dim + d];
  }
  for (int k = threadIdx.x; k < ncell; k += PLAN_THREADS) {
    uint32_t packed = 0;
    for (int a = 0; a < nv1; ++a) {
      const int32_t gv = cell_coord_t[(int64_t)a * nc + bcells[b * cmax + k]];
      const int lv = find_cell(uniq, nv, gv);
      packed |= (uint32_t)(lv & 255) << (8 * a);
    }
    bcc[b * (int64_t)((cmax + 3) & ~3) + k] = packed;
  }
}

__global__ void block_descriptors(const int64_t *__restrict__ bstart, int64_t nb, const int64_t *__restrict__ nrp,
                                  const int64_t *__restrict__ bbase, const int64_t *__restrict__ slbase,
                                  const int32_t *__restrict__ bcount, const int32_t *__restrict__ vcount,
                                  BlockDesc *__restrict__ d) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  BlockDesc x;
  x.s0 = nrp[bstart[b]];
  x.cbase = bbase[b];
  x.nslots = (int32_t)(nrp[bstart[b + 1]] - x.s0);
  x.ncontrib = (int32_t)(bbase[b + 1] - bbase[b]);
  x.ncell = bcount[b];
  x.nvert = vcount[b];
  x.slbase = slbase[b];
  x.nslpad = (int32_t)(slbase[b + 1] - slbase[b]);
  x.pad = 0;
  d[b] = x;
}

static inline bool hext_overflow(int32_t v) { return v < 0; }

struct Widen8 {
  __host__ __device__ int64_t operator()(uint8_t v) const { return (int64_t)v; }
};

}  // namespace

#define GRID(n, b) (unsigned)(((n) + (b)-1) / (b))

void fdts_free_plan(fdts_engine *e) {
  GatherPlan &p = e->plan;
  void *ptrs[] = {p.bstart, p.bcells, p.bcount, p.bbase, p.sdst, p.soff, p.slbase, p.cidx, p.ridx, p.bvcoords, p.bcc, p.bdesc};
  for (void *q : ptrs)
    if (q) cudaFree(q);
  p = GatherPlan();
}

// starts: host array of nblocks+1 owned-node row offsets (contiguous row blocks)
int fdts_build_plan(fdts_engine *e, const int64_t *starts, int64_t nb) {
  fdts_free_plan(e);
  GatherPlan &p = e->plan;
  cudaStream_t s = e->own_stream;
  const int64_t nown = e->cfg.num_owned_nodes, nc = e->cfg.num_cells;
  const int nd = e->cfg.nd;
  if (nb <= 0 || starts[0] != 0 || starts[nb] != nown) {
    fdts_set_error("row blocks must cover the owned rows [0, num_owned_nodes)");
    return 1;
  }
  if (e->cfg.ncomp != 1) {
    fdts_set_error("gather plan: scalar spaces only");
    return 3;
  }
  p.nblocks = nb;
  p.ne = nd * (nd + 1) / 2;
  p.nep = p.ne | 1;  // odd stride: conflict-free staging stores
  p.ndp = nd | 1;
  FDTS_CHECK(cudaMalloc(&p.bstart, sizeof(int64_t) * (nb + 1)));
  FDTS_CHECK(cudaMemcpyAsync(p.bstart, starts, sizeof(int64_t) * (nb + 1), cudaMemcpyHostToDevice, s));
  // candidate counts -> padded sort size
  int64_t *ncand = nullptr, *btotal = nullptr;
  int *derr = nullptr;
  FDTS_CHECK(cudaMalloc(&ncand, sizeof(int64_t) * nb));
  FDTS_CHECK(cudaMalloc(&btotal, sizeof(int64_t) * (nb + 1)));
  FDTS_CHECK(cudaMalloc(&derr, sizeof(int)));
  FDTS_CHECK(cudaMemsetAsync(derr, 0, sizeof(int), s));
  block_candidates<<<GRID(nb, 256), 256, 0, s>>>(p.bstart, nb, e->adj_ptr, ncand);
  void *tmp = nullptr;
  size_t tb = 0, tb2 = 0;
  int64_t *dmax = nullptr;
  FDTS_CHECK(cudaMalloc(&dmax, sizeof(int64_t) * 2));
  cub::DeviceReduce::Max(nullptr, tb, ncand, dmax, nb, s);
  cub::DeviceScan::ExclusiveSum(nullptr, tb2, btotal, btotal, nb + 1, s);
  tb = tb > tb2 ? tb : tb2;
  FDTS_CHECK(cudaMalloc(&tmp, tb + 16));
  cub::DeviceReduce::Max(tmp, tb, ncand, dmax, nb, s);
  int64_t maxcand = 0;
  FDTS_CHECK(cudaMemcpyAsync(&maxcand, dmax, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  FDTS_CHECK(cudaStreamSynchronize(s));
  int kpad = PLAN_THREADS;
  while (kpad < maxcand) kpad <<= 1;
  int rc = 0;
  do {
    if (kpad > 32768) {
      fdts_set_error("gather plan: a row block touches too many cells (use smaller / more compact row blocks)");
      rc = 4;
      break;
    }
    const size_t sm = sizeof(int32_t) * (size_t)kpad;
    if (sm > 48 * 1024) {
      cudaFuncSetAttribute(block_cells<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      cudaFuncSetAttribute(block_cells<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    }
    if (cudaMalloc(&p.bcount, sizeof(int32_t) * nb) != cudaSuccess) { rc = 2; break; }
    block_cells<false><<<(unsigned)nb, PLAN_THREADS, sm, s>>>(p.bstart, e->adj_ptr, e->adj_cell, kpad, 0, p.bcount,
                                                            nullptr);
    int32_t *cm = nullptr;
    if (cudaMalloc(&cm, sizeof(int32_t) * 2) != cudaSuccess) { rc = 2; break; }
    size_t tb3 = 0;
    cub::DeviceReduce::Max(nullptr, tb3, p.bcount, cm, nb, s);
    void *tmp3 = nullptr;
    if (cudaMalloc(&tmp3, tb3 + 16) != cudaSuccess) { rc = 2; break; }
    cub::DeviceReduce::Max(tmp3, tb3, p.bcount, cm, nb, s);
    int32_t cmax = 0;
    cudaMemcpyAsync(&cmax, cm, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    cudaFree(cm);
    cudaFree(tmp3);
    p.cmax = cmax;
    if ((int64_t)cmax * p.nep > 65535) {
      fdts_set_error("gather plan: staged entries per row block exceed the 16-bit address space");
      rc = 4;
      break;
    }
    if (cudaMalloc(&p.bcells, sizeof(int32_t) * (size_t)nb * cmax) != cudaSuccess) {
      fdts_set_error("gather plan: out of memory (block cell lists)");
      rc = 2;
      break;
    }
    block_cells<true><<<(unsigned)nb, PLAN_THREADS, sm, s>>>(p.bstart, e->adj_ptr, e->adj_cell, kpad, cmax, nullptr,
                                                           p.bcells);
    // slot contribution counts -> SELL layout
    uint8_t *cnt8 = nullptr;
    uint32_t *sellpos = nullptr;
    const int64_t nnz = e->node_nnz;
    if (cudaMalloc(&cnt8, (size_t)nnz + 1) != cudaSuccess || cudaMalloc(&p.sdst, sizeof(uint16_t) * (nnz + 32)) != cudaSuccess ||
        cudaMalloc(&sellpos, sizeof(uint32_t) * (nnz + 1)) != cudaSuccess ||
        cudaMalloc(&p.bbase, sizeof(int64_t) * (nb + 1)) != cudaSuccess ||
        cudaMalloc(&p.slbase, sizeof(int64_t) * (nb + 1)) != cudaSuccess) {
      fdts_set_error("gather plan: out of memory (slot tables)");
      rc = 2;
      break;
    }
    cudaMemsetAsync(cnt8, 0, (size_t)nnz + 1, s);
    cudaMemsetAsync(p.sdst, 0, sizeof(uint16_t) * (nnz + 32), s);
    slot_contributions<0><<<(unsigned)nb, 128, 0, s>>>(p.bstart, e->adj_ptr, e->adj_cell, e->cell_node, nc, nd, p.nep,
                                                      e->pos8, e->isbc, e->nrowptr, cnt8, nullptr, nullptr, nullptr,
                                                      nullptr, cmax, nullptr, derr);
    cudaMemsetAsync(btotal, 0, sizeof(int64_t) * (nb + 1), s);
    block_slices<<<GRID(nb, 256), 256, 0, s>>>(p.bstart, nb, e->nrowptr, btotal);
    cub::DeviceScan::ExclusiveSum(tmp, tb, btotal, p.slbase, nb + 1, s);
    int64_t total_slices = 0;
    cudaMemcpyAsync(&total_slices, p.slbase + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (cudaMalloc(&p.soff, sizeof(uint32_t) * (size_t)(total_slices + 16)) != cudaSuccess) {
      fdts_set_error("gather plan: out of memory (slice offsets)");
      rc = 2;
      cudaFree(cnt8);
      cudaFree(sellpos);
      break;
    }
    cudaMemsetAsync(p.soff, 0, sizeof(uint32_t) * (size_t)(total_slices + 16), s);
    cudaMemsetAsync(btotal, 0, sizeof(int64_t) * (nb + 1), s);
    sell_layout<<<(unsigned)nb, PLAN_THREADS, 0, s>>>(p.bstart, e->nrowptr, cnt8, p.slbase, p.sdst, p.soff, sellpos,
                                                    btotal);
    cub::DeviceScan::ExclusiveSum(tmp, tb, btotal, p.bbase, nb + 1, s);
    int64_t total = 0;
    cudaMemcpyAsync(&total, p.bbase + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    p.total_contrib = total;
    {
      int64_t *ext = nullptr, hext[5] = {0, 0, 0, 0, 0};
      if (cudaMalloc(&ext, sizeof(int64_t) * 5) != cudaSuccess) { rc = 2; break; }
      cudaMemsetAsync(ext, 0, sizeof(int64_t) * 5, s);
      block_extents<<<GRID(nb, 256), 256, 0, s>>>(p.bstart, nb, e->nrowptr, e->adj_ptr, p.bbase, p.slbase, ext);
      cudaMemcpyAsync(hext, ext, sizeof(int64_t) * 5, cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
      cudaFree(ext);
      p.max_block_contrib = (int32_t)hext[0];
      p.max_block_slots = (int32_t)hext[1];
      p.max_block_rows = (int32_t)hext[2];
      p.max_block_adj = (int32_t)hext[3];
      p.max_block_slices = (int32_t)hext[4];
    }
    p.zaddr = cmax * p.nep;  // staged zero that padded SELL entries point at
    if (p.zaddr > 65535 || hext_overflow(p.max_block_contrib)) {
      fdts_set_error("gather plan: a row block exceeds the 16-bit staging address space");
      rc = 4;
      cudaFree(cnt8);
      cudaFree(sellpos);
      break;
    }
    if (cudaMalloc(&p.cidx, sizeof(uint16_t) * (size_t)(total + 64)) != cudaSuccess ||
        cudaMalloc(&p.ridx, sizeof(uint16_t) * (size_t)(nc * nd + 32)) != cudaSuccess) {
      fdts_set_error("gather plan: out of memory (contribution lists)");
      rc = 2;
      cudaFree(cnt8);
      cudaFree(sellpos);
      break;
    }
    fill_u16<<<GRID(total + 64, 256), 256, 0, s>>>(p.cidx, total + 64, (uint16_t)p.zaddr);
    cudaMemsetAsync(cnt8, 0, (size_t)nnz + 1, s);
    slot_contributions<1><<<(unsigned)nb, 128, 0, s>>>(p.bstart, e->adj_ptr, e->adj_cell, e->cell_node, nc, nd, p.nep,
                                                      e->pos8, e->isbc, e->nrowptr, cnt8, sellpos, p.bbase, p.bcells,
                                                      p.bcount, cmax, p.cidx, derr);
    row_contributions<1><<<(unsigned)nb, 128, 0, s>>>(p.bstart, e->adj_ptr, e->adj_cell, e->isbc, p.ndp, p.bcells,
                                                     p.bcount, cmax, p.ridx, derr);
    cudaStreamSynchronize(s);
    cudaFree(sellpos);
    {  // bulk-copyable phase-A inputs: block vertex tables, local vertex ids, descriptors
      const int dim = e->cfg.dim;
      int kpv = PLAN_THREADS;
      while (kpv < cmax * (dim + 1)) kpv <<= 1;
      const size_t smv = sizeof(int32_t) * 2 * (size_t)kpv;
      if (smv > 48 * 1024) {
        cudaFuncSetAttribute(block_vertices<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv);
        cudaFuncSetAttribute(block_vertices<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv);
      }
      int32_t *vcount = nullptr, *vm = nullptr;
      if (cudaMalloc(&vcount, sizeof(int32_t) * nb) != cudaSuccess || cudaMalloc(&vm, sizeof(int32_t) * 2) != cudaSuccess) {
        rc = 2;
        cudaFree(cnt8);
        break;
      }
      block_vertices<false><<<(unsigned)nb, PLAN_THREADS, smv, s>>>(p.bcells, p.bcount, cmax, e->cell_coord, nc, dim,
                                                                  e->coords, kpv, 0, vcount, nullptr, nullptr, derr);
      size_t tb4 = 0;
      cub::DeviceReduce::Max(nullptr, tb4, vcount, vm, nb, s);
      void *tmp4 = nullptr;
      cudaMalloc(&tmp4, tb4 + 16);
      cub::DeviceReduce::Max(tmp4, tb4, vcount, vm, nb, s);
      int32_t vmax = 0;
      cudaMemcpyAsync(&vmax, vm, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
      cudaFree(tmp4);
      cudaFree(vm);
      vmax = (vmax + 1) & ~1;  // keep vmax*dim*8 a multiple of 16 bytes
      p.vmax = vmax;
      if (vmax > 254 || cudaMalloc(&p.bvcoords, sizeof(double) * (size_t)nb * vmax * dim + 64) != cudaSuccess ||
          cudaMalloc(&p.bcc, sizeof(uint32_t) * ((size_t)nb * ((cmax + 3) & ~3) + 16)) != cudaSuccess ||
          cudaMalloc(&p.bdesc, sizeof(BlockDesc) * (size_t)(nb + 4)) != cudaSuccess) {
        fdts_set_error("gather plan: block vertex tables do not fit");
        rc = 2;
        cudaFree(vcount);
        cudaFree(cnt8);
        break;
      }
      cudaMemsetAsync(p.bvcoords, 0, sizeof(double) * (size_t)nb * vmax * dim + 64, s);
      cudaMemsetAsync(p.bdesc, 0, sizeof(BlockDesc) * (size_t)(nb + 4), s);
      block_vertices<true><<<(unsigned)nb, PLAN_THREADS, smv, s>>>(p.bcells, p.bcount, cmax, e->cell_coord, nc, dim,
                                                                 e->coords, kpv, vmax, nullptr, p.bvcoords, p.bcc, derr);
      block_descriptors<<<GRID(nb, 256), 256, 0, s>>>(p.bstart, nb, e->nrowptr, p.bbase, p.slbase, p.bcount, vcount,
                                                      p.bdesc);
      cudaStreamSynchronize(s);
      cudaFree(vcount);
    }
    int herr = 0;
    cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaError_t ce = cudaStreamSynchronize(s);
    cudaFree(cnt8);
    if (ce != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      fdts_set_error(std::string("gather plan kernels failed: ") + cudaGetErrorString(ce));
      rc = 2;
      break;
    }
    if (herr) {
      fdts_set_error("gather plan: internal inconsistency (code " + std::to_string(herr) + ")");
      rc = 5;
      break;
    }
    p.ready = true;
  } while (0);
  cudaFree(ncand);
  cudaFree(btotal);
  cudaFree(derr);
  cudaFree(dmax);
  cudaFree(tmp);
  if (rc) fdts_free_plan(e);
  return rc;
}

// =====================================================================================
// Plan version 2: sector-grouped SELL with split slices and a pattern dictionary.
// =====================================================================================
#include <algorithm>
#include <unordered_map>
#include <vector>

namespace {

constexpr int SPLIT_CAP = 8;        // slots with more contributions are summed by 4 lanes each
constexpr int MAX_SECTORS = 8192;   // sectors per row block handled by the in-smem sort
constexpr int MAX_SLICES = 2048;    // slices per row block
constexpr uint32_t SPLIT_FLAG = 0x80000000u;

// SELL storage of one slice of width w: ceil(w/2) rows of 32 lanes x 2 columns (one 32-bit load per
// lane fetches two consecutive columns): entry (column k, lane l) sits at (k>>1)*64 + 2*l + (k&1).
__device__ __forceinline__ uint32_t slice_entries(int w) { return 64u * (uint32_t)((w + 1) >> 1); }
__device__ __forceinline__ int64_t sell_entry(int k, int lane) { return (int64_t)(k >> 1) * 64 + 2 * lane + (k & 1); }

// one CTA per row block.  G = slots per sector (4: 32-byte sectors, 2: 16-byte half sectors, 1: single slots).
// Sectors are sorted (split first, wide first); a normal slice holds 32/G sectors (one slot per lane),
// a split slice 8/G sectors (four lanes per slot).  WRITE=false: slice and entry counts only.
template <bool WRITE>
__global__ void __launch_bounds__(PLAN_THREADS) sell2_layout(const int64_t *__restrict__ bstart,
                                                             const int64_t *__restrict__ nrp,
                                                             const uint8_t *__restrict__ cnt8, int G,
                                                             const int64_t *__restrict__ meta_off,
                                                             uint2 *__restrict__ cmeta, int32_t *__restrict__ bncls,
                                                             uint16_t *__restrict__ gdst,
                                                             uint32_t *__restrict__ sellpos,
                                                             int64_t *__restrict__ btotal, int64_t *__restrict__ bnsl,
                                                             int *__restrict__ err) {
  __shared__ uint32_t keys[MAX_SECTORS];
  __shared__ uint32_t woff[MAX_SLICES + 2];
  __shared__ int nsplit_s;
  const int lg = G == 4 ? 2 : (G == 2 ? 1 : 0);
  const int per_norm = 32 >> lg, per_split = 8 >> lg;  // sectors per normal / split slice
  const int64_t b = blockIdx.x;
  const int64_t s0 = nrp[bstart[b]], s1 = nrp[bstart[b + 1]];
  const int64_t sec0 = s0 >> lg;
  const int64_t nsec64 = s1 > s0 ? ((s1 + G - 1) >> lg) - sec0 : 0;
  if (nsec64 > MAX_SECTORS) {
    if (threadIdx.x == 0) {
      atomicExch(err, 7);
      btotal[b] = 0;
      bnsl[b] = 0;
    }
    return;
  }
  const int nsec = (int)nsec64;
  if (threadIdx.x == 0) nsplit_s = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < MAX_SECTORS; t += PLAN_THREADS) {
    uint32_t key = 0xFFFFFFFFu;
    if (t < nsec) {
      int m = 0;
      for (int q = 0; q < G; ++q) {
        const int64_t slot = (int64_t)G * (sec0 + t) + q;
        if (slot >= s0 && slot < s1) m = max(m, (int)cnt8[slot]);
      }
      const int split = m > SPLIT_CAP ? 1 : 0;
      const int width = split ? (m + 3) >> 2 : m;
      if (split) atomicAdd(&nsplit_s, 1);
      key = ((uint32_t)(0x1ff - ((split << 8) | width)) << 16) | (uint32_t)t;  // split first, wide first
    }
    keys[t] = key;
  }
  __syncthreads();
  for (int kk = 2; kk <= MAX_SECTORS; kk <<= 1)
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < MAX_SECTORS; t += PLAN_THREADS) {
        const int q = t ^ j;
        if (q > t) {
          const uint32_t x = keys[t], y = keys[q];
          const bool up = (t & kk) == 0;
          if ((x > y) == up) {
            keys[t] = y;
            keys[q] = x;
          }
        }
      }
      __syncthreads();
    }
  const int nsplit = nsplit_s;
  const int nslA = (nsplit + per_split - 1) / per_split, nslB = (nsec - nsplit + per_norm - 1) / per_norm;
  const int nsl = nslA + nslB;
  if (nsl > MAX_SLICES) {
    if (threadIdx.x == 0) {
      atomicExch(err, 7);
      btotal[b] = 0;
      bnsl[b] = 0;
    }
    return;
  }
  auto first_sector = [&](int j) { return j < nslA ? per_split * j : nsplit + per_norm * (j - nslA); };
  auto width_of = [&](uint32_t key) { return (int)((0x1ff - (key >> 16)) & 0xff); };
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    int ncls = 0, cfirst = 0;
    uint32_t ckey = 0xffffffffu, coff = 0;
    auto flush = [&](int jend) {
      if (ckey == 0xffffffffu) return;
      if (WRITE && ncls < PLAN2_MAX_CLASSES)
        cmeta[b * PLAN2_MAX_CLASSES + ncls] =
            make_uint2((uint32_t)cfirst | ((uint32_t)(jend - cfirst) << 16), (coff >> 6) | ((ckey & 0xffu) << 16) | ((ckey >> 8) << 24));
      ++ncls;
    };
    for (int j = 0; j < nsl; ++j) {
      woff[j] = run;
      const int wj = width_of(keys[first_sector(j)]);
      const uint32_t key = (uint32_t)wj | (j < nslA ? 0x100u : 0u);
      if (key != ckey) {
        flush(j);
        ckey = key;
        cfirst = j;
        coff = run;
      }
      run += slice_entries(wj);
    }
    flush(nsl);
    woff[nsl] = run;
    btotal[b] = run;
    bnsl[b] = WRITE ? nsl : ((nsl + 3) & ~3);
    if ((run >> 6) > 65535u) atomicExch(err, 8);
    if (ncls > PLAN2_MAX_CLASSES) atomicExch(err, 9);
    if (WRITE) bncls[b] = ncls;
  }
  if (!WRITE) return;
  __syncthreads();
  const int64_t mo = meta_off[b];
  for (int j = threadIdx.x; j < ((nsl + 3) & ~3); j += PLAN_THREADS) {
    for (int g = 0; g < per_norm; ++g) {
      uint16_t v = 0xffff;
      if (j < nslA) {
        if (g < per_split && per_split * j + g < nsplit) v = (uint16_t)(keys[per_split * j + g] & 0xffffu);
      } else if (j < nsl) {
        const int q = nsplit + per_norm * (j - nslA) + g;
        if (q < nsec) v = (uint16_t)(keys[q] & 0xffffu);
      }
      gdst[(mo + j) * per_norm + g] = v;
    }
  }
  for (int q = threadIdx.x; q < nsec; q += PLAN_THREADS) {
    const int t = (int)(keys[q] & 0xffffu);
    int j, lanebase;
    uint32_t flag = 0;
    if (q < nsplit) {
      j = q / per_split;
      lanebase = (q % per_split) * 4 * G;
      flag = SPLIT_FLAG;
    } else {
      const int qq = q - nsplit;
      j = nslA + qq / per_norm;
      lanebase = (qq % per_norm) * G;
    }
    for (int sub = 0; sub < G; ++sub) {
      const int64_t slot = (int64_t)G * (sec0 + t) + sub;
      // sellpos: slice offset (units of 64 entries) << 8 | first lane of the slot | split flag
      if (slot >= s0 && slot < s1) sellpos[slot] = ((woff[j] >> 6) << 8) | (uint32_t)(lanebase + (flag ? sub * 4 : sub)) | flag;
    }
  }
}

// padding entries point at one of 16 staged zeros, bank = lane & 15 (conflict free among themselves)
__global__ void fill_padding(uint16_t *__restrict__ p, int64_t n, int zaddr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = (uint16_t)(zaddr + (int)((t >> 1) & 15));
}

// writes the staged addresses of every contribution into its SELL position (one thread per row,
// contributions in ascending cell order => reproducible sums)
__global__ void slot_contributions2(const int64_t *__restrict__ bstart, const int64_t *__restrict__ adj_ptr,
                                    const int32_t *__restrict__ adj_cell, const int32_t *__restrict__ cell_node_t,
                                    int64_t nc, int nd, int nep, const uint8_t *__restrict__ pos8_t,
                                    const uint8_t *__restrict__ isbc, const int64_t *__restrict__ nrp,
                                    uint8_t *__restrict__ cnt8, const uint32_t *__restrict__ sellpos,
                                    const int64_t *__restrict__ idx_off, const int32_t *__restrict__ bcells,
                                    const int32_t *__restrict__ bcount, int cmax, uint16_t *__restrict__ cidx,
                                    int *__restrict__ err) {
  const int64_t b = blockIdx.x;
  const int64_t r0 = bstart[b], r1 = bstart[b + 1];
  for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    if (isbc[r]) continue;
    const int64_t rb = nrp[r];
    for (int64_t a = adj_ptr[r]; a < adj_ptr[r + 1]; ++a) {
      const int32_t code = adj_cell[a];
      const int64_t cell = code >> 5;
      const int i = code & 31;
      const int cs = find_cell(bcells + b * cmax, bcount[b], (int32_t)cell);
      if (cs < 0) {
        atomicExch(err, 2);
        continue;
      }
      for (int j = 0; j < nd; ++j) {
        if (isbc[cell_node_t[(int64_t)j * nc + cell]]) continue;
        const int64_t slot = rb + pos8_t[(int64_t)(i * nd + j) * nc + cell];
        const int k = cnt8[slot]++;
        const uint32_t sp = sellpos[slot];
        const int64_t base = (int64_t)((sp & ~SPLIT_FLAG) >> 8) * 64;
        const int lane0 = (int)(sp & 0xffu);
        const int64_t at = (sp & SPLIT_FLAG) ? base + sell_entry(k >> 2, lane0 + (k & 3)) : base + sell_entry(k, lane0);
        cidx[idx_off[b] + at] = (uint16_t)(cs * nep + sym_index(i, j, nd));
      }
    }
  }
}

__global__ void block_descriptors2(const int64_t *__restrict__ bstart, int64_t nb, const int64_t *__restrict__ nrp,
                                   const int64_t *__restrict__ idx_off, const int64_t *__restrict__ meta_off,
                                   const int64_t *__restrict__ bnsl_true, const int32_t *__restrict__ bncls,
                                   const int32_t *__restrict__ bcount, const int32_t *__restrict__ vcount, int cmaxp,
                                   int vmax, int dim, BlockDesc2 *__restrict__ d) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  BlockDesc2 x;
  x.s0 = nrp[bstart[b]];
  x.idx_off = idx_off[b];
  x.meta_off = meta_off[b];
  x.cc_off = b * (int64_t)cmaxp;
  x.vc_off = b * (int64_t)vmax * dim;
  x.nslots = (int32_t)(nrp[bstart[b + 1]] - x.s0);
  x.nidx = (int32_t)(idx_off[b + 1] - idx_off[b]);
  x.ncell = bcount[b];
  x.nvert = vcount[b];
  x.nsl = (int32_t)bnsl_true[b];
  x.ncls = bncls[b];
  x.cls_off = b * (int64_t)PLAN2_MAX_CLASSES;
  x.pattern = (int32_t)b;
  x.pad1 = 0;
  d[b] = x;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// 128-bit content hash of a block's pattern (position-keyed, order-independent sum => parallel)
__global__ void __launch_bounds__(PLAN_THREADS) block_hash(const BlockDesc2 *__restrict__ d,
                                                           const uint16_t *__restrict__ cidx,
                                                           const uint2 *__restrict__ cmeta,
                                                           const uint16_t *__restrict__ gdst, int G,
                                                           const uint32_t *__restrict__ bcc, int per_norm,
                                                           uint64_t *__restrict__ hash2) {
  __shared__ unsigned long long h0, h1;
  const BlockDesc2 x = d[blockIdx.x];
  if (threadIdx.x == 0) {
    h0 = mix64(((uint64_t)x.nslots << 32) ^ (uint64_t)x.nidx) + mix64(((uint64_t)x.ncell << 32) ^ (uint64_t)x.nvert ^ 0x51ull << 60);
    h1 = mix64(((uint64_t)x.nsl << 32) ^ (uint64_t)(x.s0 & (G - 1)) ^ ((uint64_t)x.ncls << 8) ^ 0x7ull << 58);
  }
  __syncthreads();
  uint64_t a0 = 0, a1 = 0;
  auto add = [&](uint64_t stream, uint64_t pos, uint64_t val) {
    const uint64_t k = (stream << 56) ^ (pos << 20) ^ val;
    a0 += mix64(k);
    a1 += mix64(k ^ 0xa5a5a5a5deadbeefull);
  };
  for (int t = threadIdx.x; t < x.nidx; t += PLAN_THREADS) add(1, t, cidx[x.idx_off + t]);
  const int nslp = (x.nsl + 3) & ~3;
  for (int t = threadIdx.x; t < x.ncls; t += PLAN_THREADS) {
    const uint2 c = cmeta[x.cls_off + t];
    add(2, t, ((uint64_t)c.y << 32) | c.x);
  }
  for (int t = threadIdx.x; t < nslp * per_norm; t += PLAN_THREADS) add(3, t, gdst[x.meta_off * per_norm + t]);
  for (int t = threadIdx.x; t < x.ncell; t += PLAN_THREADS) add(4, t, bcc[x.cc_off + t]);
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&h0, (unsigned long long)a0);
    atomicAdd(&h1, (unsigned long long)a1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    hash2[2 * blockIdx.x] = h0;
    hash2[2 * blockIdx.x + 1] = h1;
  }
}

__global__ void apply_representatives(BlockDesc2 *__restrict__ d, const int64_t *__restrict__ rep, int64_t nb) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const int64_t r = rep[b];
  if (r == b) return;
  d[b].idx_off = d[r].idx_off;
  d[b].meta_off = d[r].meta_off;
  d[b].cc_off = d[r].cc_off;
  d[b].cls_off = d[r].cls_off;
  d[b].pattern = (int32_t)r;
}

struct PairHash {
  size_t operator()(const std::pair<uint64_t, uint64_t> &p) const { return (size_t)(p.first ^ (p.second * 0x9e3779b97f4a7c15ull)); }
};

}  // namespace

// Bank-conflict-aware layout of the unique patterns (host side).  Two degrees of freedom that do not
// change any sum except for its order:
//   * the order in which a lane adds its contributions (a permutation of its column entries);
//   * which store groups (sectors) of one slice class share a half-warp (slices of a class are
//     interchangeable; a group moves together with its sector id).
// Goal: the 16 lanes of a half-warp hit 16 different 8-byte banks in every column.  Greedy packing of
// half-warps for narrow classes (few permutations per lane => the choice of neighbours matters most),
// then a few refinement sweeps over the lanes of every slice; padding lanes finally take a staged zero
// in the least loaded bank.
namespace {
struct BankLoad {
  struct Ent {
    uint16_t addr, n;
  };
  int w = 0;
  std::vector<std::vector<Ent>> cells;  // [column][half][bank]
  void reset(int width) {
    w = width;
    cells.assign((size_t)w * 2 * 16, {});
  }
  std::vector<Ent> &at(int k, int half, int bank) { return cells[((size_t)k * 2 + half) * 16 + bank]; }
  void add(int k, int half, uint16_t x) {
    auto &v = at(k, half, x & 15);
    for (auto &e : v)
      if (e.addr == x) {
        ++e.n;
        return;
      }
    v.push_back({x, 1});
  }
  void remove(int k, int half, uint16_t x) {
    auto &v = at(k, half, x & 15);
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i].addr == x) {
        if (--v[i].n == 0) v.erase(v.begin() + (long)i);
        return;
      }
  }
  int cost(int k, int half, uint16_t x) {
    const auto &v = at(k, half, x & 15);
    for (const auto &e : v)
      if (e.addr == x) return 0;  // same address: broadcast
    return 2 * (int)v.size() + 1;
  }
};

struct LaneOrder {
  uint64_t rng = 0x243f6a8885a308d3ull;
  std::vector<int> best, cand, ident;
  uint32_t next() {
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    return (uint32_t)(rng >> 33);
  }
  // best order of `vals` (entries >= zaddr are padding) against `load`; returns its cost
  int choose(BankLoad &load, int half, const std::vector<uint16_t> &vals, int zaddr, std::vector<uint16_t> &out,
             bool quick = false) {
    const int w = (int)vals.size();
    ident.resize(w);
    for (int k = 0; k < w; ++k) ident[k] = k;
    int bestc = -1;
    auto consider = [&](const std::vector<int> &pm) {
      int c = 0;
      for (int k = 0; k < w; ++k)
        if (vals[pm[k]] < zaddr) c += load.cost(k, half, vals[pm[k]]);
      if (bestc < 0 || c < bestc) {
        bestc = c;
        best = pm;
      }
    };
    consider(ident);  // keep the current order on ties
    if (w <= 5) {
      cand = ident;
      do consider(cand); while (bestc > 0 && std::next_permutation(cand.begin(), cand.end()));
    } else {
      for (int r = 0; r < w && bestc > 0; ++r) {
        cand.resize(w);
        for (int k = 0; k < w; ++k) cand[k] = (k + r) % w;
        consider(cand);
        std::reverse(cand.begin(), cand.end());
        consider(cand);
      }
      for (int t = 0; t < 128 && bestc > 0 && !quick; ++t) {
        cand = ident;
        for (int k = w - 1; k > 0; --k) std::swap(cand[k], cand[next() % (uint32_t)(k + 1)]);
        consider(cand);
      }
    }
    out.resize(w);
    for (int k = 0; k < w; ++k) out[k] = vals[best[k]];
    return bestc;
  }
};
}  // namespace

// ---- the same lane-order refinement on the GPU, for plans with many unique patterns (meshes without repeated
// blocks: every block keeps its own tables, and walking them on the host would take minutes).  One warp per slice;
// the two half-warps are independent (a 64-bit shared-memory load is served half-warp by half-warp).  The 16
// lanes of a half take turns: for the lane whose turn it is, all 16 lanes evaluate candidate orders of its w
// addresses (all w! permutations for w <= 5, the 2 w rotations / reversed rotations above) against the bank
// occupancy of the other 15 lanes, the best one (ties: the current order) is kept.  Padding entries finally
// take the staged zero of the least occupied bank.  Measured at C4 on the dictionary path: lane orders alone
// bring the gather from 7.45 to 5.94 ms, regrouping the store groups as well (host only) to 5.87 ms.
constexpr int OPT_WARPS = 8, OPT_MAXW = 8;

__device__ __forceinline__ void perm_from_index(int p, int w, int (&pm)[OPT_MAXW]) {
  // p-th permutation of 0..w-1 in the factorial number system (p = 0: identity)
  int avail[OPT_MAXW];
#pragma unroll
  for (int k = 0; k < OPT_MAXW; ++k) avail[k] = k;
  int f = 1;
  for (int k = 2; k < w; ++k) f *= k;  // (w-1)!
  for (int k = 0; k < w; ++k) {
    const int q = p / f;
    p -= q * f;
    pm[k] = avail[q];
    for (int t = q; t < w - 1 - k; ++t) avail[t] = avail[t + 1];
    if (w - 1 - k > 0) f /= (w - 1 - k);
  }
}

__global__ void __launch_bounds__(OPT_WARPS * 32) optimize_lane_orders(const BlockDesc2 *__restrict__ bdesc,
                                                                       const uint2 *__restrict__ cmeta,
                                                                       uint16_t *__restrict__ cidx, int zaddr, int sweeps) {
  __shared__ uint16_t sA[OPT_WARPS][32][OPT_MAXW];
  __shared__ uint16_t sC[OPT_WARPS][2][OPT_MAXW][OPT_MAXW];
  const int64_t b = blockIdx.x;
  const BlockDesc2 d = bdesc[b];
  if (d.pattern != (int32_t)b || d.nidx == 0) return;  // shared tables are optimised once, through their owner
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
  uint16_t(*A)[OPT_MAXW] = sA[warp];
  uint16_t(*Cm)[OPT_MAXW] = sC[warp][half];
  for (int c = 0; c < d.ncls; ++c) {
    const uint2 ce = cmeta[d.cls_off + c];
    const int count = (int)(ce.x >> 16), w = (int)((ce.y >> 16) & 0xffu);
    if (w < 2 || w > OPT_MAXW || count < 1) continue;
    const int pairs = (w + 1) >> 1;
    uint16_t *cbase = cidx + d.idx_off + (size_t)(ce.y & 0xffffu) * 64;
    int nperm = 2 * w;
    if (w <= 5) {
      nperm = 1;
      for (int k = 2; k <= w; ++k) nperm *= k;
    }
    for (int sl = warp; sl < count; sl += OPT_WARPS) {
      auto at = [&](int k) -> uint16_t & { return cbase[((size_t)sl * pairs + (size_t)(k >> 1)) * 64 + 2 * lane + (k & 1)]; };
      for (int k = 0; k < w; ++k) A[lane][k] = at(k);
      __syncwarp();
      for (int sweep = 0; sweep < sweeps; ++sweep) {
        bool changed = false;
        for (int j = 0; j < 16; ++j) {
          const int L = half * 16 + j;  // the lane of my half whose turn it is
          // conflicts of value m in column k with the other lanes of the half
          for (int q = hl; q < w * w; q += 16) {
            const int k = q / w, m = q - k * w;
            const uint16_t v = A[L][m];
            int cnt = 0;
            if (v < zaddr) {
              for (int l = half * 16; l < half * 16 + 16; ++l) {
                const uint16_t x = A[l][k];
                cnt += (l != L && x < zaddr && ((x ^ v) & 15) == 0 && x != v) ? 1 : 0;
              }
            }
            Cm[k][m] = (uint16_t)cnt;
          }
          __syncwarp();
          uint32_t bestkey = 0xffffffffu;
          for (int pi = hl; pi < nperm; pi += 16) {
            int pm[OPT_MAXW];
            if (w <= 5) {
              perm_from_index(pi, w, pm);
            } else {
              const int r = pi >> 1;
              for (int k = 0; k < w; ++k) pm[k] = (pi & 1) ? (w - 1 - ((k + r) % w)) : ((k + r) % w);
            }
            uint32_t cst = 0;
            for (int k = 0; k < w; ++k) cst += Cm[k][pm[k]];
            const uint32_t key = (cst << 12) | (uint32_t)pi;  // ties: the lowest index, i.e. the current order first
            bestkey = key < bestkey ? key : bestkey;
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, bestkey, o);
            bestkey = other < bestkey ? other : bestkey;
          }
          const int bp = (int)(bestkey & 0xfffu);
          if (bp != 0 && hl == 0) {  // one lane of the half rewrites row L
            int pm[OPT_MAXW];
            if (w <= 5) {
              perm_from_index(bp, w, pm);
            } else {
              const int r = bp >> 1;
              for (int k = 0; k < w; ++k) pm[k] = (bp & 1) ? (w - 1 - ((k + r) % w)) : ((k + r) % w);
            }
            uint16_t old[OPT_MAXW];
            for (int k = 0; k < w; ++k) old[k] = A[L][k];
            for (int k = 0; k < w; ++k) A[L][k] = old[pm[k]];
          }
          changed |= bp != 0;
          __syncwarp();
        }
        if (!__any_sync(0xffffffffu, changed)) break;
      }
      // padding: the staged zero of the least occupied bank of the column (per half)
      for (int k = 0; k < w; ++k) {
        const uint16_t v = A[lane][k];
        const bool real = v < zaddr;
        int bestb = 0, bestn = 99;
        for (int bk = 0; bk < 16; ++bk) {
          const unsigned m = __ballot_sync(0xffffffffu, real && (v & 15) == bk);
          const int n = __popc(half ? (m >> 16) : (m & 0xffffu));
          if (n < bestn) {
            bestn = n;
            bestb = bk;
          }
        }
        at(k) = real ? v : (uint16_t)(zaddr + bestb);
      }
      __syncwarp();
    }
  }
}

static void optimize_pattern(const BlockDesc2 &d, const uint2 *cls, uint16_t *idx, uint16_t *gdst, int G, int zaddr,
                             bool regroup = true) {
  const int per_norm = 32 / G;
  LaneOrder lo;
  BankLoad load;
  std::vector<uint16_t> tmp;
  for (int c = 0; c < d.ncls; ++c) {
    const int first = (int)(cls[c].x & 0xffffu), count = (int)(cls[c].x >> 16);
    const int w = (int)((cls[c].y >> 16) & 0xffu);
    const bool split = (cls[c].y >> 24) != 0;
    if (w < 1 || count < 1) continue;
    const size_t pairs = (size_t)((w + 1) >> 1);
    uint16_t *cbase = idx + (size_t)(cls[c].y & 0xffffu) * 64;
    auto at = [&](int sl, int k, int lane) -> uint16_t & { return cbase[((size_t)sl * pairs + (size_t)(k >> 1)) * 64 + 2 * lane + (k & 1)]; };
    // ---- regroup the store groups of narrow, non-split classes
    if (regroup && !split && w <= 8 && count >= 2) {
      struct Unit {
        uint16_t gd;
        std::vector<std::vector<uint16_t>> lanes;  // [G][w]
      };
      std::vector<Unit> pool, pads, order;
      for (int sl = 0; sl < count; ++sl)
        for (int g = 0; g < per_norm; ++g) {
          Unit u;
          u.gd = gdst[(size_t)(first + sl) * per_norm + g];
          u.lanes.assign(G, std::vector<uint16_t>(w));
          for (int q = 0; q < G; ++q)
            for (int k = 0; k < w; ++k) u.lanes[q][k] = at(sl, k, g * G + q);
          (u.gd == 0xffff ? pads : pool).push_back(std::move(u));
        }
      const int gph = per_norm / 2;  // groups per half-warp
      std::vector<char> used(pool.size(), 0);
      size_t remaining = pool.size(), head = 0;
      while (remaining > 0) {
        load.reset(w);
        for (int slot = 0; slot < gph && remaining > 0; ++slot) {
          while (used[head]) ++head;
          size_t pick = head;
          if (slot > 0) {
            int bestc = -1;
            size_t seen = 0;
            for (size_t u = head; u < pool.size() && seen < 256; ++u) {
              if (used[u]) continue;
              ++seen;
              int cst = 0;
              for (int q = 0; q < G; ++q) cst += lo.choose(load, 0, pool[u].lanes[q], zaddr, tmp, true);
              if (bestc < 0 || cst < bestc) {
                bestc = cst;
                pick = u;
              }
              if (bestc == 0) break;
            }
          }
          Unit u = pool[pick];
          used[pick] = 1;
          --remaining;
          for (int q = 0; q < G; ++q) {
            lo.choose(load, 0, u.lanes[q], zaddr, tmp);
            u.lanes[q] = tmp;
            for (int k = 0; k < w; ++k)
              if (tmp[k] < zaddr) load.add(k, 0, tmp[k]);
          }
          order.push_back(std::move(u));
        }
      }
      for (auto &u : pads) order.push_back(std::move(u));
      size_t n = 0;
      for (int sl = 0; sl < count; ++sl)
        for (int g = 0; g < per_norm; ++g, ++n) {
          const Unit &u = order[n];
          gdst[(size_t)(first + sl) * per_norm + g] = u.gd;
          for (int q = 0; q < G; ++q)
            for (int k = 0; k < w; ++k) at(sl, k, g * G + q) = u.lanes[q][k];
        }
    }
    // ---- per-slice refinement of the lane orders, then padding
    for (int sl = 0; sl < count; ++sl) {
      load.reset(w);
      std::vector<std::vector<uint16_t>> cur(32, std::vector<uint16_t>(w));
      for (int lane = 0; lane < 32; ++lane)
        for (int k = 0; k < w; ++k) cur[lane][k] = at(sl, k, lane);
      for (int sweep = 0; sweep < 4; ++sweep) {
        bool changed = false;
        for (int lane = 0; lane < 32; ++lane) {
          const int half = lane >> 4;
          auto &vals = cur[lane];
          int nreal = 0;
          for (int k = 0; k < w; ++k) nreal += vals[k] < zaddr;
          if (nreal == 0) continue;
          if (sweep > 0)
            for (int k = 0; k < w; ++k)
              if (vals[k] < zaddr) load.remove(k, half, vals[k]);
          lo.choose(load, half, vals, zaddr, tmp);
          changed |= tmp != vals;
          vals = tmp;
          for (int k = 0; k < w; ++k)
            if (vals[k] < zaddr) load.add(k, half, vals[k]);
        }
        if (sweep > 0 && !changed) break;
      }
      for (int lane = 0; lane < 32; ++lane)
        for (int k = 0; k < w; ++k) {
          uint16_t v = cur[lane][k];
          if (v >= zaddr) {  // padding: a staged zero in the least loaded bank
            const int half = lane >> 4;
            int bb = 0;
            size_t bl = (size_t)-1;
            for (int bk = 0; bk < 16; ++bk) {
              const auto &cellv = load.at(k, half, bk);
              size_t l = cellv.size();
              for (const auto &e : cellv)
                if (e.addr == (uint16_t)(zaddr + bk)) --l;  // the zero of this bank is already there: free
              if (l < bl) {
                bl = l;
                bb = bk;
              }
            }
            v = (uint16_t)(zaddr + bb);
            load.add(k, half, v);
          }
          at(sl, k, lane) = v;
        }
    }
  }
}

void fdts_free_plan2(fdts_engine *e) {
  GatherPlan2 &p = e->plan2;
  void *ptrs[] = {p.cidx, p.cmeta, p.gdst, p.bcc, p.bvcoords, p.bdesc, p.bmetric};
  for (void *q : ptrs)
    if (q) cudaFree(q);
  p = GatherPlan2();
}

#define P2_ALLOC(ptr, bytes, what)                                              \
  if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) {                    \
    fdts_set_error(std::string("gather plan 2: out of device memory (") + what + ")"); \
    rc = 2;                                                                     \
    break;                                                                      \
  }

int fdts_build_plan2(fdts_engine *e, const int64_t *starts, int64_t nb) {
  fdts_free_plan2(e);
  GatherPlan2 &p = e->plan2;
  cudaStream_t s = e->own_stream;
  const int64_t nown = e->cfg.num_owned_nodes, nc = e->cfg.num_cells, nnz = e->node_nnz;
  const int nd = e->cfg.nd, dim = e->cfg.dim;
  if (nb <= 0 || starts[0] != 0 || starts[nb] != nown) {
    fdts_set_error("row blocks must cover the owned rows [0, num_owned_nodes)");
    return 1;
  }
  if (e->cfg.ncomp != 1) {
    fdts_set_error("gather plan: scalar spaces only");
    return 3;
  }
  p.nblocks = nb;
  p.nep = (nd * (nd + 1) / 2) | 1;
  const int G = e->opt_sector == 2 ? 2 : (e->opt_sector == 1 ? 1 : 4);
  const int per_norm = 32 / G;
  p.sector = G;
  int64_t *bstart = nullptr, *ncand = nullptr, *btotal = nullptr, *bnsl = nullptr, *idx_off = nullptr, *meta_off = nullptr,
          *dmax = nullptr, *rep_d = nullptr;
  int32_t *bcount = nullptr, *bcells = nullptr, *vcount = nullptr, *imax = nullptr, *bncls = nullptr;
  uint8_t *cnt8 = nullptr;
  uint32_t *sellpos = nullptr;
  uint64_t *hash_d = nullptr;
  int *derr = nullptr;
  void *tmp = nullptr;
  size_t tb = 0;
  int rc = 0;
  auto max_i64 = [&](const int64_t *src, int64_t n, int64_t &out) {
    size_t t2 = 0;
    cub::DeviceReduce::Max(nullptr, t2, src, dmax, n, s);
    if (t2 > tb) return 2;
    cub::DeviceReduce::Max(tmp, t2, src, dmax, n, s);
    cudaMemcpyAsync(&out, dmax, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 2;
  };
  auto max_i32 = [&](const int32_t *src, int64_t n, int32_t &out) {
    size_t t2 = 0;
    cub::DeviceReduce::Max(nullptr, t2, src, imax, n, s);
    if (t2 > tb) return 2;
    cub::DeviceReduce::Max(tmp, t2, src, imax, n, s);
    cudaMemcpyAsync(&out, imax, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 2;
  };
  auto scan = [&](const int64_t *src, int64_t *dst, int64_t n) {
    size_t t2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t2, src, dst, n, s);
    if (t2 > tb) return 2;
    cub::DeviceScan::ExclusiveSum(tmp, t2, src, dst, n, s);
    return 0;
  };
  do {
    tb = 1 << 22;  // ample for CUB reductions/scans over nb+1 entries
    {
      size_t t2 = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, t2, btotal, btotal, nb + 1, s);
      if (t2 > tb) tb = t2;
    }
    P2_ALLOC(tmp, tb + 16, "scratch");
    P2_ALLOC(bstart, sizeof(int64_t) * (nb + 1), "block starts");
    P2_ALLOC(ncand, sizeof(int64_t) * (nb + 1), "scratch");
    P2_ALLOC(btotal, sizeof(int64_t) * (nb + 1), "scratch");
    P2_ALLOC(bnsl, sizeof(int64_t) * (nb + 1), "scratch");
    P2_ALLOC(idx_off, sizeof(int64_t) * (nb + 1), "scratch");
    P2_ALLOC(meta_off, sizeof(int64_t) * (nb + 1), "scratch");
    P2_ALLOC(dmax, sizeof(int64_t) * 2, "scratch");
    P2_ALLOC(imax, sizeof(int32_t) * 2, "scratch");
    P2_ALLOC(derr, sizeof(int), "scratch");
    P2_ALLOC(bcount, sizeof(int32_t) * nb, "block cell counts");
    P2_ALLOC(vcount, sizeof(int32_t) * nb, "block vertex counts");
    P2_ALLOC(bncls, sizeof(int32_t) * nb, "block class counts");
    cudaMemsetAsync(derr, 0, sizeof(int), s);
    cudaMemcpyAsync(bstart, starts, sizeof(int64_t) * (nb + 1), cudaMemcpyHostToDevice, s);
    // ---- cells of every block
    block_candidates<<<GRID(nb, 256), 256, 0, s>>>(bstart, nb, e->adj_ptr, ncand);
    int64_t maxcand = 0;
    if ((rc = max_i64(ncand, nb, maxcand))) break;
    int kpad = PLAN_THREADS;
    while (kpad < maxcand) kpad <<= 1;
    if (kpad > 32768) {
      fdts_set_error("gather plan: a row block touches too many cells (use smaller / more compact row blocks)");
      rc = 4;
      break;
    }
    const size_t sm = sizeof(int32_t) * (size_t)kpad;
    if (sm > 48 * 1024) {
      cudaFuncSetAttribute(block_cells<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      cudaFuncSetAttribute(block_cells<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    }
    block_cells<false><<<(unsigned)nb, PLAN_THREADS, sm, s>>>(bstart, e->adj_ptr, e->adj_cell, kpad, 0, bcount, nullptr);
    int32_t cmax = 0;
    if ((rc = max_i32(bcount, nb, cmax))) break;
    p.cmax = cmax;
    p.zaddr = (cmax * p.nep + 15) & ~15;  // 16 staged zeros, one per 8-byte bank
    if (p.zaddr + 16 > 65535) {
      fdts_set_error("gather plan: staged entries per row block exceed the 16-bit address space (use smaller row blocks)");
      rc = 4;
      break;
    }
    P2_ALLOC(bcells, sizeof(int32_t) * (size_t)nb * cmax, "block cell lists");
    block_cells<true><<<(unsigned)nb, PLAN_THREADS, sm, s>>>(bstart, e->adj_ptr, e->adj_cell, kpad, cmax, nullptr, bcells);
    // ---- contribution counts per slot, sector-grouped SELL layout
    P2_ALLOC(cnt8, (size_t)nnz + 1, "slot counters");
    P2_ALLOC(sellpos, sizeof(uint32_t) * (size_t)(nnz + 1), "slot positions");
    cudaMemsetAsync(cnt8, 0, (size_t)nnz + 1, s);
    slot_contributions<0><<<(unsigned)nb, 128, 0, s>>>(bstart, e->adj_ptr, e->adj_cell, e->cell_node, nc, nd, p.nep, e->pos8,
                                                      e->isbc, e->nrowptr, cnt8, nullptr, nullptr, nullptr, nullptr, cmax,
                                                      nullptr, derr);
    cudaMemsetAsync(btotal, 0, sizeof(int64_t) * (nb + 1), s);
    cudaMemsetAsync(bnsl, 0, sizeof(int64_t) * (nb + 1), s);
    sell2_layout<false><<<(unsigned)nb, PLAN_THREADS, 0, s>>>(bstart, e->nrowptr, cnt8, G, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                            btotal, bnsl, derr);
    if ((rc = scan(btotal, idx_off, nb + 1)) || (rc = scan(bnsl, meta_off, nb + 1))) break;
    int64_t total = 0, total_sl = 0, max_idx = 0, max_sl = 0;
    cudaMemcpyAsync(&total, idx_off + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&total_sl, meta_off + nb, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    if ((rc = max_i64(btotal, nb, max_idx)) || (rc = max_i64(bnsl, nb, max_sl))) break;
    int herr = 0;
    cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (herr == 9) {
      fdts_set_error("gather plan 2: a row block has more than 32 distinct slice widths");
      rc = 4;
      break;
    }
    if (herr == 7 || herr == 8) {
      fdts_set_error("gather plan 2: a row block has too many CSR slots (at most 8192 store groups / 2048 slices; use smaller row blocks)");
      rc = 4;
      break;
    }
    p.total_idx = total;
    p.max_nidx = (int32_t)max_idx;
    p.max_nsl = (int32_t)max_sl;
    P2_ALLOC(p.cidx, sizeof(uint16_t) * (size_t)(total + 64), "contribution lists");
    P2_ALLOC(p.cmeta, sizeof(uint2) * ((size_t)nb + 1) * PLAN2_MAX_CLASSES, "slice class table");
    P2_ALLOC(p.gdst, sizeof(uint16_t) * (size_t)(total_sl + 16) * per_norm, "slice destinations");
    fill_padding<<<GRID(total + 64, 256), 256, 0, s>>>(p.cidx, total + 64, p.zaddr);
    cudaMemsetAsync(p.cmeta, 0, sizeof(uint2) * ((size_t)nb + 1) * PLAN2_MAX_CLASSES, s);
    cudaMemsetAsync(p.gdst, 0xff, sizeof(uint16_t) * (size_t)(total_sl + 16) * per_norm, s);
    // second pass writes the tables; the true (unpadded) slice counts land in ncand (reused)
    sell2_layout<true><<<(unsigned)nb, PLAN_THREADS, 0, s>>>(bstart, e->nrowptr, cnt8, G, meta_off, p.cmeta, bncls, p.gdst, sellpos,
                                                           btotal, ncand, derr);
    cudaMemsetAsync(cnt8, 0, (size_t)nnz + 1, s);
    slot_contributions2<<<(unsigned)nb, 128, 0, s>>>(bstart, e->adj_ptr, e->adj_cell, e->cell_node, nc, nd, p.nep, e->pos8,
                                                    e->isbc, e->nrowptr, cnt8, sellpos, idx_off, bcells, bcount, cmax,
                                                    p.cidx, derr);
    cudaStreamSynchronize(s);
    cudaFree(sellpos);
    sellpos = nullptr;
    cudaFree(cnt8);
    cnt8 = nullptr;
    // ---- block vertex tables
    int kpv = PLAN_THREADS;
    while (kpv < cmax * (dim + 1)) kpv <<= 1;
    const size_t smv = sizeof(int32_t) * 2 * (size_t)kpv;
    if (smv > 48 * 1024) {
      cudaFuncSetAttribute(block_vertices<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv);
      cudaFuncSetAttribute(block_vertices<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv);
    }
    block_vertices<false><<<(unsigned)nb, PLAN_THREADS, smv, s>>>(bcells, bcount, cmax, e->cell_coord, nc, dim, e->coords, kpv,
                                                                0, vcount, nullptr, nullptr, derr);
    int32_t vmax = 0;
    if ((rc = max_i32(vcount, nb, vmax))) break;
    vmax = (vmax + 1) & ~1;  // keep vmax*dim*8 a multiple of 16 bytes
    p.vmax = vmax;
    if (vmax > 254) {
      fdts_set_error("gather plan: a row block touches more than 254 coordinate nodes (use smaller row blocks)");
      rc = 4;
      break;
    }
    const int cmaxp = (cmax + 3) & ~3;
    p.cmaxp = cmaxp;
    P2_ALLOC(p.bvcoords, sizeof(double) * (size_t)nb * vmax * dim + 64, "block vertex tables");
    P2_ALLOC(p.bcc, sizeof(uint32_t) * ((size_t)nb * cmaxp + 16), "block cell vertices");
    P2_ALLOC(p.bdesc, sizeof(BlockDesc2) * (size_t)(nb + 4), "block descriptors");
    cudaMemsetAsync(p.bvcoords, 0, sizeof(double) * (size_t)nb * vmax * dim + 64, s);
    cudaMemsetAsync(p.bcc, 0, sizeof(uint32_t) * ((size_t)nb * cmaxp + 16), s);
    cudaMemsetAsync(p.bdesc, 0, sizeof(BlockDesc2) * (size_t)(nb + 4), s);
    block_vertices<true><<<(unsigned)nb, PLAN_THREADS, smv, s>>>(bcells, bcount, cmax, e->cell_coord, nc, dim, e->coords, kpv,
                                                               vmax, nullptr, p.bvcoords, p.bcc, derr);
    block_descriptors2<<<GRID(nb, 256), 256, 0, s>>>(bstart, nb, e->nrowptr, idx_off, meta_off, ncand, bncls, bcount, vcount, cmaxp,
                                                    vmax, dim, p.bdesc);
    cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaError_t ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      fdts_set_error(std::string("gather plan 2 kernels failed: ") + cudaGetErrorString(ce));
      rc = 2;
      break;
    }
    if (herr) {
      fdts_set_error("gather plan 2: internal inconsistency (code " + std::to_string(herr) + ")");
      rc = 5;
      break;
    }
    // ---- pattern dictionary: blocks with identical index lists share the first copy
    p.unique_patterns = nb;
    p.shared_idx = total;
    if (e->opt_dedup) {
      P2_ALLOC(hash_d, sizeof(uint64_t) * 2 * (size_t)nb, "pattern hashes");
      P2_ALLOC(rep_d, sizeof(int64_t) * (size_t)nb, "pattern representatives");
      block_hash<<<(unsigned)nb, PLAN_THREADS, 0, s>>>(p.bdesc, p.cidx, p.cmeta, p.gdst, G, p.bcc, per_norm, hash_d);
      std::vector<uint64_t> h(2 * (size_t)nb);
      std::vector<BlockDesc2> hd((size_t)nb);
      cudaMemcpyAsync(h.data(), hash_d, sizeof(uint64_t) * 2 * (size_t)nb, cudaMemcpyDeviceToHost, s);
      cudaMemcpyAsync(hd.data(), p.bdesc, sizeof(BlockDesc2) * (size_t)nb, cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        fdts_set_error("gather plan 2: hashing failed");
        rc = 2;
        break;
      }
      std::unordered_map<std::pair<uint64_t, uint64_t>, int64_t, PairHash> first;
      first.reserve((size_t)nb);
      std::vector<int64_t> rep((size_t)nb);
      int64_t uniq = 0, shared = 0;
      for (int64_t b = 0; b < nb; ++b) {
        auto key = std::make_pair(h[2 * b], h[2 * b + 1]);
        auto it = first.find(key);
        if (it == first.end()) {
          first.emplace(key, b);
          rep[b] = b;
          ++uniq;
          shared += hd[b].nidx;
        } else {
          rep[b] = it->second;
        }
      }
      cudaMemcpyAsync(rep_d, rep.data(), sizeof(int64_t) * (size_t)nb, cudaMemcpyHostToDevice, s);
      apply_representatives<<<GRID(nb, 256), 256, 0, s>>>(p.bdesc, rep_d, nb);
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        fdts_set_error("gather plan 2: applying the pattern dictionary failed");
        rc = 2;
        break;
      }
      p.unique_patterns = uniq;
      p.shared_idx = shared;
      if (e->opt_optimize && uniq > 256) {  // too many unique patterns for the host walk (~50 ms each): lane orders on the GPU
        optimize_lane_orders<<<(unsigned)nb, OPT_WARPS * 32, 0, s>>>(p.bdesc, p.cmeta, p.cidx, p.zaddr, 3);
        if (cudaStreamSynchronize(s) != cudaSuccess) {
          fdts_set_error("gather plan 2: lane-order optimisation failed");
          rc = 2;
          break;
        }
        p.optimized = 2;
      }
      if (e->opt_optimize && uniq <= 256) {
        std::vector<uint2> hm;
        std::vector<uint16_t> hi, hg;
        bool ok = true;
        for (int64_t b = 0; b < nb && ok; ++b) {
          if (rep[b] != b || hd[b].nsl == 0 || hd[b].nidx == 0) continue;
          hm.resize((size_t)hd[b].ncls);
          hi.resize((size_t)hd[b].nidx);
          hg.resize((size_t)hd[b].nsl * per_norm);
          ok = cudaMemcpy(hm.data(), p.cmeta + hd[b].cls_off, sizeof(uint2) * hm.size(), cudaMemcpyDeviceToHost) == cudaSuccess &&
               cudaMemcpy(hi.data(), p.cidx + hd[b].idx_off, sizeof(uint16_t) * hi.size(), cudaMemcpyDeviceToHost) == cudaSuccess &&
               cudaMemcpy(hg.data(), p.gdst + hd[b].meta_off * per_norm, sizeof(uint16_t) * hg.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
          if (!ok) break;
          optimize_pattern(hd[b], hm.data(), hi.data(), hg.data(), G, p.zaddr, e->opt_optimize != 2);
          ok = cudaMemcpy(p.cidx + hd[b].idx_off, hi.data(), sizeof(uint16_t) * hi.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
               cudaMemcpy(p.gdst + hd[b].meta_off * per_norm, hg.data(), sizeof(uint16_t) * hg.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        }
        if (!ok) {
          fdts_set_error("gather plan 2: pattern optimisation failed");
          rc = 2;
          break;
        }
        p.optimized = 1;
      }
      // ---- compaction: keep only the tables of the unique patterns (C4: 12.2 GB built, 0.6 MB referenced)
      if (uniq * 4 <= nb) {
        int64_t n_idx = 0, n_sl = 0, n_cc = 0, k = 0;
        std::vector<int64_t> nidx_of((size_t)nb, -1), nsl_of((size_t)nb, -1), ncc_of((size_t)nb, -1), ncls_of((size_t)nb, -1);
        for (int64_t b = 0; b < nb; ++b) {
          if (rep[b] != b) continue;
          nidx_of[b] = n_idx;
          nsl_of[b] = n_sl;
          ncc_of[b] = n_cc;
          ncls_of[b] = k * PLAN2_MAX_CLASSES;
          n_idx += hd[b].nidx;
          n_sl += (hd[b].nsl + 3) & ~3;
          n_cc += (hd[b].ncell + 3) & ~3;
          ++k;
        }
        uint16_t *cidx2 = nullptr, *gdst2 = nullptr;
        uint2 *cmeta2 = nullptr;
        uint32_t *bcc2 = nullptr;
        bool ok = cudaMalloc(&cidx2, sizeof(uint16_t) * (size_t)(n_idx + 64)) == cudaSuccess &&
                  cudaMalloc(&gdst2, sizeof(uint16_t) * (size_t)(n_sl + 16) * per_norm) == cudaSuccess &&
                  cudaMalloc(&cmeta2, sizeof(uint2) * (size_t)(uniq + 1) * PLAN2_MAX_CLASSES) == cudaSuccess &&
                  cudaMalloc(&bcc2, sizeof(uint32_t) * (size_t)(n_cc + 16)) == cudaSuccess;
        for (int64_t b = 0; b < nb && ok; ++b) {
          if (rep[b] != b) continue;
          const BlockDesc2 &d0 = hd[b];
          const size_t nslp = (size_t)((d0.nsl + 3) & ~3), nccp = (size_t)((d0.ncell + 3) & ~3);
          ok = (d0.nidx == 0 || cudaMemcpyAsync(cidx2 + nidx_of[b], p.cidx + d0.idx_off, sizeof(uint16_t) * (size_t)d0.nidx, cudaMemcpyDeviceToDevice, s) == cudaSuccess) &&
               (nslp == 0 || cudaMemcpyAsync(gdst2 + nsl_of[b] * per_norm, p.gdst + d0.meta_off * per_norm, sizeof(uint16_t) * nslp * per_norm, cudaMemcpyDeviceToDevice, s) == cudaSuccess) &&
               cudaMemcpyAsync(cmeta2 + ncls_of[b], p.cmeta + d0.cls_off, sizeof(uint2) * PLAN2_MAX_CLASSES, cudaMemcpyDeviceToDevice, s) == cudaSuccess &&
               (nccp == 0 || cudaMemcpyAsync(bcc2 + ncc_of[b], p.bcc + d0.cc_off, sizeof(uint32_t) * nccp, cudaMemcpyDeviceToDevice, s) == cudaSuccess);
        }
        if (ok) {
          for (int64_t b = 0; b < nb; ++b) {  // final descriptors: own s0 / vertex table, tables of the representative
            const int64_t r = rep[b];
            hd[b].idx_off = nidx_of[r];
            hd[b].meta_off = nsl_of[r];
            hd[b].cc_off = ncc_of[r];
            hd[b].cls_off = ncls_of[r];
            hd[b].pattern = (int32_t)r;
          }
          ok = cudaMemcpyAsync(p.bdesc, hd.data(), sizeof(BlockDesc2) * (size_t)nb, cudaMemcpyHostToDevice, s) == cudaSuccess &&
               cudaStreamSynchronize(s) == cudaSuccess;
        }
        if (!ok) {
          cudaFree(cidx2); cudaFree(gdst2); cudaFree(cmeta2); cudaFree(bcc2);
          fdts_set_error("gather plan 2: compacting the pattern dictionary failed");
          rc = 2;
          break;
        }
        cudaFree(p.cidx); cudaFree(p.gdst); cudaFree(p.cmeta); cudaFree(p.bcc);
        p.cidx = cidx2; p.gdst = gdst2; p.cmeta = cmeta2; p.bcc = bcc2;
        p.compacted = 1;
      }
    } else if (e->opt_optimize) {  // no dictionary: every block owns its tables; lane orders on the GPU
      optimize_lane_orders<<<(unsigned)nb, OPT_WARPS * 32, 0, s>>>(p.bdesc, p.cmeta, p.cidx, p.zaddr, 3);
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        fdts_set_error("gather plan 2: lane-order optimisation failed");
        rc = 2;
        break;
      }
      p.optimized = 2;
    }
    p.ready = true;
  } while (0);
  void *scratch[] = {tmp, bstart, ncand, btotal, bnsl, idx_off, meta_off, dmax, imax, derr, bcount, vcount, bncls, bcells, cnt8,
                     sellpos, hash_d, rep_d};
  for (void *q : scratch)
    if (q) cudaFree(q);
  if (rc) fdts_free_plan2(e);
  return rc;
}

// =====================================================================================
// Residual tile plan: per-tile pre-reduction of the element vectors before the atomics.
// =====================================================================================
namespace {

constexpr int RT_KEYS = 4096;  // (cell, local node) pairs per tile handled by the in-smem sort

__global__ void tile_cells(const int64_t *__restrict__ tstart, int64_t nt, int cmax, int32_t *__restrict__ bcount,
                           int32_t *__restrict__ bcells) {
  const int64_t t = blockIdx.x;
  const int64_t c0 = tstart[t], n = tstart[t + 1] - c0;
  if (threadIdx.x == 0) bcount[t] = (int32_t)n;
  if (bcells)
    for (int k = threadIdx.x; k < n; k += blockDim.x) bcells[t * cmax + k] = (int32_t)(c0 + k);
}

__global__ void widen_i32(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= n) out[t] = t < n ? (int64_t)in[t] : 0;
}

__global__ void iota64(int64_t *p, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = t;
}

__device__ __forceinline__ void bitonic_sort_u32(uint32_t *keys, int n) {  // n power of two, all threads of the CTA
  for (int kk = 2; kk <= n; kk <<= 1)
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int q = t ^ j;
        if (q > t) {
          const uint32_t x = keys[t], y = keys[q];
          const bool up = (t & kk) == 0;
          if ((x > y) == up) {
            keys[t] = y;
            keys[q] = x;
          }
        }
      }
      __syncthreads();
    }
}

// one CTA per tile: sorted unique nodes of its cells (relative ids, "no output" flag) and the
// tile-local node index of every (cell, local node).  WRITE=false: count only.
template <bool WRITE>
__global__ void __launch_bounds__(PLAN_THREADS) tile_nodes(const int64_t *__restrict__ tstart,
                                                           const int32_t *__restrict__ cell_node_t, int64_t nc, int nd,
                                                           int64_t nown, const uint8_t *__restrict__ isbc, int umaxp,
                                                           int cmaxp, int32_t *__restrict__ ucount,
                                                           int64_t *__restrict__ tbase, int32_t *__restrict__ tnode,
                                                           uint16_t *__restrict__ tlid, int *__restrict__ err) {
  __shared__ uint32_t keys[RT_KEYS];
  __shared__ int32_t uniq[RT_KEYS];
  __shared__ int warp_tot[PLAN_THREADS / 32];
  const int64_t t = blockIdx.x;
  const int64_t c0 = tstart[t];
  const int ncell = (int)(tstart[t + 1] - c0);
  const int n = ncell * nd;
  if (n > RT_KEYS) {
    if (threadIdx.x == 0) {
      atomicExch(err, 11);
      ucount[t] = 0;
    }
    return;
  }
  for (int q = threadIdx.x; q < RT_KEYS; q += PLAN_THREADS) {
    uint32_t v = 0xFFFFFFFFu;
    if (q < n) {
      const int i = q / ncell, k = q - i * ncell;
      v = (uint32_t)cell_node_t[(int64_t)i * nc + c0 + k];
    }
    keys[q] = v;
  }
  __syncthreads();
  bitonic_sort_u32(keys, RT_KEYS);
  int base = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < RT_KEYS; t0 += PLAN_THREADS) {
    const int q = t0 + threadIdx.x;
    const uint32_t v = keys[q];
    const bool first = v != 0xFFFFFFFFu && (q == 0 || keys[q - 1] != v);
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (first) uniq[off + __popc(m & ((1u << lane) - 1u))] = (int32_t)v;
    for (int w = 0; w < PLAN_THREADS / 32; ++w) base += warp_tot[w];
    __syncthreads();
  }
  const int nu = base;
  if (!WRITE) {
    if (threadIdx.x == 0) ucount[t] = nu;
    return;
  }
  if (nu > 65534 || nu > umaxp) {
    if (threadIdx.x == 0) atomicExch(err, 12);
    return;
  }
  const int32_t b0 = nu > 0 ? uniq[0] : 0;
  if (threadIdx.x == 0) tbase[t] = b0;
  for (int q = threadIdx.x; q < umaxp; q += PLAN_THREADS) {
    int32_t v = (int32_t)0x80000000;
    if (q < nu) {
      const int32_t node = uniq[q];
      v = node - b0;
      if (node >= nown || isbc[node]) v |= (int32_t)0x80000000;  // ghost or Dirichlet row: no output
    }
    tnode[t * (int64_t)umaxp + q] = v;
  }
  for (int q = threadIdx.x; q < n; q += PLAN_THREADS) {
    const int i = q / ncell, k = q - i * ncell;
    const int32_t node = cell_node_t[(int64_t)i * nc + c0 + k];
    const int l = find_cell(uniq, nu, node);
    tlid[(t * nd + i) * (int64_t)cmaxp + k] = (uint16_t)l;
  }
}

// MODE 0: contributions per unique node (cnt8).  MODE 1: staged addresses into their SELL positions,
// contributions of a node in ascending (cell, local index) order.
template <int MODE>
__global__ void __launch_bounds__(PLAN_THREADS) tile_contributions(const int64_t *__restrict__ tstart, int nd, int ndp,
                                                                   int cmaxp, const uint16_t *__restrict__ tlid,
                                                                   const int64_t *__restrict__ tptr,
                                                                   uint8_t *__restrict__ cnt8,
                                                                   const uint32_t *__restrict__ sellpos,
                                                                   const int64_t *__restrict__ idx_off,
                                                                   uint16_t *__restrict__ cidx, int *__restrict__ err) {
  __shared__ uint32_t keys[RT_KEYS];
  const int64_t t = blockIdx.x;
  const int ncell = (int)(tstart[t + 1] - tstart[t]);
  const int n = ncell * nd;
  if (n > RT_KEYS || ncell > 512) {
    if (threadIdx.x == 0) atomicExch(err, 11);
    return;
  }
  for (int q = threadIdx.x; q < RT_KEYS; q += PLAN_THREADS) {
    uint32_t v = 0xFFFFFFFFu;
    if (q < n) {
      const int i = q / ncell, k = q - i * ncell;
      v = ((uint32_t)tlid[(t * nd + i) * (int64_t)cmaxp + k] << 14) | ((uint32_t)k << 5) | (uint32_t)i;
    }
    keys[q] = v;
  }
  __syncthreads();
  bitonic_sort_u32(keys, RT_KEYS);
  for (int p = threadIdx.x; p < n; p += PLAN_THREADS) {
    const uint32_t key = keys[p];
    const uint32_t l = key >> 14;
    int lo = 0, hi = p;  // first position of this node's segment
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((keys[mid] >> 14) < l) lo = mid + 1; else hi = mid;
    }
    const int kth = p - lo;
    const int64_t slot = tptr[t] + l;
    if (MODE == 0) {
      if (p + 1 == n || (keys[p + 1] >> 14) != l) {
        if (kth + 1 > 255) atomicExch(err, 3);
        cnt8[slot] = (uint8_t)(kth + 1);
      }
    } else {
      const uint32_t sp = sellpos[slot];
      const int64_t base = (int64_t)((sp & ~SPLIT_FLAG) >> 8) * 64;
      const int lane0 = (int)(sp & 0xffu);
      const int64_t at = (sp & SPLIT_FLAG) ? base + sell_entry(kth >> 2, lane0 + (kth & 3)) : base + sell_entry(kth, lane0);
      const int k = (int)((key >> 5) & 511u), i = (int)(key & 31u);
      cidx[idx_off[t] + at] = (uint16_t)(k * ndp + i);
    }
  }
}

__global__ void tile_descriptors(int64_t nt, const int64_t *__restrict__ tbase, const int64_t *__restrict__ idx_off,
                                 const int64_t *__restrict__ meta_off, const int64_t *__restrict__ nsl_true,
                                 const int32_t *__restrict__ bncls, const int32_t *__restrict__ bcount,
                                 const int32_t *__restrict__ vcount, const int32_t *__restrict__ ucount, int cmaxp,
                                 int vmax, int dim, BlockDesc2 *__restrict__ d) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  BlockDesc2 x;
  x.s0 = tbase[t];
  x.idx_off = idx_off[t];
  x.meta_off = meta_off[t];
  x.cc_off = t * (int64_t)cmaxp;
  x.vc_off = t * (int64_t)vmax * dim;
  x.cls_off = t * (int64_t)PLAN2_MAX_CLASSES;
  x.nslots = ucount[t];
  x.nidx = (int32_t)(idx_off[t + 1] - idx_off[t]);
  x.ncell = bcount[t];
  x.nvert = vcount[t];
  x.nsl = (int32_t)nsl_true[t];
  x.ncls = bncls[t];
  x.pattern = (int32_t)t;
  x.pad1 = 0;
  d[t] = x;
}

__global__ void __launch_bounds__(PLAN_THREADS) tile_hash(const BlockDesc2 *__restrict__ d,
                                                          const uint16_t *__restrict__ cidx,
                                                          const uint2 *__restrict__ cmeta,
                                                          const uint16_t *__restrict__ gdst,
                                                          const uint32_t *__restrict__ bcc,
                                                          const int32_t *__restrict__ tnode,
                                                          const uint16_t *__restrict__ tlid, int umaxp, int cmaxp, int nd,
                                                          int mask, uint64_t *__restrict__ hash2) {
  __shared__ unsigned long long h0, h1;
  const int64_t t = blockIdx.x;
  const BlockDesc2 x = d[t];
  if (threadIdx.x == 0) {
    h0 = mix64(((uint64_t)x.nslots << 32) ^ (uint64_t)x.nidx) + mix64(((uint64_t)x.ncell << 32) ^ (uint64_t)x.nvert ^ 0x52ull << 60);
    h1 = mix64(((uint64_t)x.nsl << 32) ^ ((uint64_t)x.ncls << 8) ^ 0x9ull << 58);
  }
  __syncthreads();
  uint64_t a0 = 0, a1 = 0;
  auto add = [&](uint64_t stream, uint64_t pos, uint64_t val) {
    const uint64_t k = (stream << 56) ^ (pos << 24) ^ val;
    a0 += mix64(k);
    a1 += mix64(k ^ 0xa5a5a5a5deadbeefull);
  };
  if (mask & 1)
    for (int q = threadIdx.x; q < x.nidx; q += PLAN_THREADS) add(1, q, cidx[x.idx_off + q]);
  if (mask & 2)
    for (int q = threadIdx.x; q < x.ncls; q += PLAN_THREADS) {
      const uint2 c = cmeta[x.cls_off + q];
      add(2, q, ((uint64_t)c.y << 32) | c.x);
    }
  const int nslp = (x.nsl + 3) & ~3;
  if (mask & 4)
    for (int q = threadIdx.x; q < nslp * 32; q += PLAN_THREADS) add(3, q, gdst[x.meta_off * 32 + q]);
  if (mask & 8)
    for (int q = threadIdx.x; q < x.ncell; q += PLAN_THREADS) add(4, q, bcc[x.cc_off + q]);
  if (mask & 16)
    for (int q = threadIdx.x; q < x.nslots; q += PLAN_THREADS) add(5, q, (uint32_t)tnode[t * umaxp + q]);
  if (mask & 32)
  for (int q = threadIdx.x; q < x.ncell * nd; q += PLAN_THREADS) {
    const int i = q / x.ncell, k = q - i * x.ncell;
    add(6, q, tlid[(t * nd + i) * (int64_t)cmaxp + k]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&h0, (unsigned long long)a0);
    atomicAdd(&h1, (unsigned long long)a1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    hash2[2 * t] = h0;
    hash2[2 * t + 1] = h1;
  }
}

}  // namespace

void fdts_free_rplan(fdts_engine *e) {
  ResidualPlan &p = e->rplan;
  void *ptrs[] = {p.cidx, p.cmeta, p.gdst, p.bcc, p.bvcoords, p.tnode, p.tlid, p.tdesc, p.tstart_dev};
  for (void *q : ptrs)
    if (q) cudaFree(q);
  delete[] p.tstart_host;
  p = ResidualPlan();
}

int fdts_build_rplan(fdts_engine *e, const int64_t *starts, int64_t nt) {
  fdts_free_rplan(e);
  ResidualPlan &p = e->rplan;
  cudaStream_t s = e->own_stream;
  const int64_t nc = e->cfg.num_cells, nown = e->cfg.num_owned_nodes;
  const int nd = e->cfg.nd, dim = e->cfg.dim;
  if (nt <= 0 || starts[0] != 0 || starts[nt] != nc) {
    fdts_set_error("cell tiles must cover the cells [0, num_cells)");
    return 1;
  }
  if (e->cfg.ncomp != 1) {
    fdts_set_error("residual tile plan: scalar spaces only");
    return 3;
  }
  p.ntiles = nt;
  p.ndp = nd | 1;
  int64_t *tstart = nullptr, *tbase = nullptr, *tptr = nullptr, *iota = nullptr, *btotal = nullptr, *bnsl = nullptr,
          *idx_off = nullptr, *meta_off = nullptr, *nsl_true = nullptr, *dmax = nullptr, *rep_d = nullptr, *u64tmp = nullptr;
  int32_t *bcount = nullptr, *bcells = nullptr, *vcount = nullptr, *ucount = nullptr, *bncls = nullptr, *imax = nullptr;
  uint8_t *cnt8 = nullptr;
  uint32_t *sellpos = nullptr;
  uint64_t *hash_d = nullptr;
  int *derr = nullptr;
  void *tmp = nullptr;
  size_t tb = 1 << 22;
  int rc = 0;
  auto max_i64 = [&](const int64_t *src, int64_t n, int64_t &out) {
    size_t t2 = 0;
    cub::DeviceReduce::Max(nullptr, t2, src, dmax, n, s);
    if (t2 > tb) return 2;
    cub::DeviceReduce::Max(tmp, t2, src, dmax, n, s);
    cudaMemcpyAsync(&out, dmax, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 2;
  };
  auto max_i32 = [&](const int32_t *src, int64_t n, int32_t &out) {
    size_t t2 = 0;
    cub::DeviceReduce::Max(nullptr, t2, src, imax, n, s);
    if (t2 > tb) return 2;
    cub::DeviceReduce::Max(tmp, t2, src, imax, n, s);
    cudaMemcpyAsync(&out, imax, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : 2;
  };
  auto scan = [&](const int64_t *src, int64_t *dst, int64_t n) {
    size_t t2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t2, src, dst, n, s);
    if (t2 > tb) return 2;
    cub::DeviceScan::ExclusiveSum(tmp, t2, src, dst, n, s);
    return 0;
  };
  do {
    p.tstart_host = new int64_t[nt + 1];
    memcpy(p.tstart_host, starts, sizeof(int64_t) * (nt + 1));
    P2_ALLOC(tmp, tb + 16, "scratch");
    P2_ALLOC(tstart, sizeof(int64_t) * (nt + 1), "tile starts");
    P2_ALLOC(tbase, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(tptr, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(iota, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(btotal, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(bnsl, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(idx_off, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(meta_off, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(nsl_true, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(u64tmp, sizeof(int64_t) * (nt + 1), "scratch");
    P2_ALLOC(dmax, sizeof(int64_t) * 2, "scratch");
    P2_ALLOC(imax, sizeof(int32_t) * 2, "scratch");
    P2_ALLOC(derr, sizeof(int), "scratch");
    P2_ALLOC(bcount, sizeof(int32_t) * nt, "scratch");
    P2_ALLOC(vcount, sizeof(int32_t) * nt, "scratch");
    P2_ALLOC(ucount, sizeof(int32_t) * nt, "scratch");
    P2_ALLOC(bncls, sizeof(int32_t) * nt, "scratch");
    cudaMemsetAsync(derr, 0, sizeof(int), s);
    cudaMemcpyAsync(tstart, starts, sizeof(int64_t) * (nt + 1), cudaMemcpyHostToDevice, s);
    iota64<<<GRID(nt + 1, 256), 256, 0, s>>>(iota, nt + 1);
    // ---- cells and vertices of every tile
    tile_cells<<<(unsigned)nt, 128, 0, s>>>(tstart, nt, 0, bcount, nullptr);
    int32_t cmax = 0;
    if ((rc = max_i32(bcount, nt, cmax))) break;
    if ((int64_t)cmax * nd > RT_KEYS || cmax > 512 || cmax * p.ndp + 16 > 65535) {
      fdts_set_error("residual tile plan: a cell tile is too large (cells x nodes per cell must not exceed 4096)");
      rc = 4;
      break;
    }
    p.cmax = cmax;
    p.cmaxp = (cmax + 3) & ~3;
    p.zaddr = (cmax * p.ndp + 15) & ~15;
    P2_ALLOC(bcells, sizeof(int32_t) * (size_t)nt * cmax, "tile cell lists");
    tile_cells<<<(unsigned)nt, 128, 0, s>>>(tstart, nt, cmax, bcount, bcells);
    int kpv = PLAN_THREADS;
    while (kpv < cmax * (dim + 1)) kpv <<= 1;
    const size_t smv = sizeof(int32_t) * 2 * (size_t)kpv;
    if (smv > 48 * 1024) {
      cudaFuncSetAttribute(block_vertices<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv);
      cudaFuncSetAttribute(block_vertices<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv);
    }
    block_vertices<false><<<(unsigned)nt, PLAN_THREADS, smv, s>>>(bcells, bcount, cmax, e->cell_coord, nc, dim, e->coords, kpv,
                                                                0, vcount, nullptr, nullptr, derr);
    int32_t vmax = 0;
    if ((rc = max_i32(vcount, nt, vmax))) break;
    vmax = (vmax + 1) & ~1;
    p.vmax = vmax;
    if (vmax > 254) {
      fdts_set_error("residual tile plan: a cell tile touches more than 254 coordinate nodes (use smaller tiles)");
      rc = 4;
      break;
    }
    P2_ALLOC(p.bvcoords, sizeof(double) * (size_t)nt * vmax * dim + 64, "tile vertex tables");
    P2_ALLOC(p.bcc, sizeof(uint32_t) * ((size_t)nt * p.cmaxp + 16), "tile cell vertices");
    cudaMemsetAsync(p.bvcoords, 0, sizeof(double) * (size_t)nt * vmax * dim + 64, s);
    cudaMemsetAsync(p.bcc, 0, sizeof(uint32_t) * ((size_t)nt * p.cmaxp + 16), s);
    block_vertices<true><<<(unsigned)nt, PLAN_THREADS, smv, s>>>(bcells, bcount, cmax, e->cell_coord, nc, dim, e->coords, kpv,
                                                               vmax, nullptr, p.bvcoords, p.bcc, derr);
    // ---- unique nodes, local node ids
    tile_nodes<false><<<(unsigned)nt, PLAN_THREADS, 0, s>>>(tstart, e->cell_node, nc, nd, nown, e->isbc, 0, p.cmaxp, ucount,
                                                           nullptr, nullptr, nullptr, derr);
    int32_t umax = 0;
    if ((rc = max_i32(ucount, nt, umax))) break;
    p.umaxp = (umax + 3) & ~3;
    P2_ALLOC(p.tnode, sizeof(int32_t) * ((size_t)nt * p.umaxp + 16), "tile node lists");
    P2_ALLOC(p.tlid, sizeof(uint16_t) * ((size_t)nt * p.cmaxp * nd + 16), "tile local node ids");
    cudaMemsetAsync(p.tlid, 0, sizeof(uint16_t) * ((size_t)nt * p.cmaxp * nd + 16), s);
    tile_nodes<true><<<(unsigned)nt, PLAN_THREADS, 0, s>>>(tstart, e->cell_node, nc, nd, nown, e->isbc, p.umaxp, p.cmaxp,
                                                          ucount, tbase, p.tnode, p.tlid, derr);
    // ---- per-node contribution counts, SELL layout (store groups of one node)
    widen_i32<<<GRID(nt + 1, 256), 256, 0, s>>>(ucount, nt, u64tmp);
    if ((rc = scan(u64tmp, tptr, nt + 1))) break;
    int64_t total_u = 0;
    cudaMemcpyAsync(&total_u, tptr + nt, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    P2_ALLOC(cnt8, (size_t)total_u + 1, "node counters");
    P2_ALLOC(sellpos, sizeof(uint32_t) * (size_t)(total_u + 1), "node positions");
    cudaMemsetAsync(cnt8, 0, (size_t)total_u + 1, s);
    tile_contributions<0><<<(unsigned)nt, PLAN_THREADS, 0, s>>>(tstart, nd, p.ndp, p.cmaxp, p.tlid, tptr, cnt8, nullptr,
                                                               nullptr, nullptr, derr);
    cudaMemsetAsync(btotal, 0, sizeof(int64_t) * (nt + 1), s);
    cudaMemsetAsync(bnsl, 0, sizeof(int64_t) * (nt + 1), s);
    sell2_layout<false><<<(unsigned)nt, PLAN_THREADS, 0, s>>>(iota, tptr, cnt8, 1, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                            btotal, bnsl, derr);
    if ((rc = scan(btotal, idx_off, nt + 1)) || (rc = scan(bnsl, meta_off, nt + 1))) break;
    int64_t total = 0, total_sl = 0, max_idx = 0, max_sl = 0;
    cudaMemcpyAsync(&total, idx_off + nt, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    cudaMemcpyAsync(&total_sl, meta_off + nt, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    if ((rc = max_i64(btotal, nt, max_idx)) || (rc = max_i64(bnsl, nt, max_sl))) break;
    int herr = 0;
    cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    if (herr) {
      fdts_set_error("residual tile plan: tile too large or internal inconsistency (code " + std::to_string(herr) + ")");
      rc = 4;
      break;
    }
    p.max_nidx = (int32_t)max_idx;
    p.max_nsl = (int32_t)max_sl;
    P2_ALLOC(p.cidx, sizeof(uint16_t) * (size_t)(total + 64), "contribution lists");
    P2_ALLOC(p.cmeta, sizeof(uint2) * ((size_t)nt + 1) * PLAN2_MAX_CLASSES, "slice class table");
    P2_ALLOC(p.gdst, sizeof(uint16_t) * (size_t)(total_sl + 16) * 32, "slice destinations");
    P2_ALLOC(p.tdesc, sizeof(BlockDesc2) * (size_t)(nt + 4), "tile descriptors");
    fill_padding<<<GRID(total + 64, 256), 256, 0, s>>>(p.cidx, total + 64, p.zaddr);
    cudaMemsetAsync(p.cmeta, 0, sizeof(uint2) * ((size_t)nt + 1) * PLAN2_MAX_CLASSES, s);
    cudaMemsetAsync(p.gdst, 0xff, sizeof(uint16_t) * (size_t)(total_sl + 16) * 32, s);
    cudaMemsetAsync(p.tdesc, 0, sizeof(BlockDesc2) * (size_t)(nt + 4), s);
    sell2_layout<true><<<(unsigned)nt, PLAN_THREADS, 0, s>>>(iota, tptr, cnt8, 1, meta_off, p.cmeta, bncls, p.gdst, sellpos,
                                                           btotal, nsl_true, derr);
    tile_contributions<1><<<(unsigned)nt, PLAN_THREADS, 0, s>>>(tstart, nd, p.ndp, p.cmaxp, p.tlid, tptr, cnt8, sellpos, idx_off,
                                                               p.cidx, derr);
    tile_descriptors<<<GRID(nt, 256), 256, 0, s>>>(nt, tbase, idx_off, meta_off, nsl_true, bncls, bcount, vcount, ucount,
                                                  p.cmaxp, vmax, dim, p.tdesc);
    cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaError_t ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess || cudaGetLastError() != cudaSuccess || herr) {
      fdts_set_error(std::string("residual tile plan kernels failed: ") + cudaGetErrorString(ce) + " code " + std::to_string(herr));
      rc = 2;
      break;
    }
    // ---- pattern dictionary + bank-conflict-aware ordering of the unique patterns
    p.unique_patterns = nt;
    if (e->opt_dedup) {
      P2_ALLOC(hash_d, sizeof(uint64_t) * 2 * (size_t)nt, "pattern hashes");
      P2_ALLOC(rep_d, sizeof(int64_t) * (size_t)nt, "pattern representatives");
      tile_hash<<<(unsigned)nt, PLAN_THREADS, 0, s>>>(p.tdesc, p.cidx, p.cmeta, p.gdst, p.bcc, p.tnode, p.tlid, p.umaxp, p.cmaxp,
                                                    nd, 63, hash_d);
      std::vector<uint64_t> h(2 * (size_t)nt);
      std::vector<BlockDesc2> hd((size_t)nt);
      cudaMemcpyAsync(h.data(), hash_d, sizeof(uint64_t) * 2 * (size_t)nt, cudaMemcpyDeviceToHost, s);
      cudaMemcpyAsync(hd.data(), p.tdesc, sizeof(BlockDesc2) * (size_t)nt, cudaMemcpyDeviceToHost, s);
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        fdts_set_error("residual tile plan: hashing failed");
        rc = 2;
        break;
      }
      std::unordered_map<std::pair<uint64_t, uint64_t>, int64_t, PairHash> first;
      first.reserve((size_t)nt);
      std::vector<int64_t> rep((size_t)nt);
      int64_t uniq = 0;
      for (int64_t t = 0; t < nt; ++t) {
        auto key = std::make_pair(h[2 * t], h[2 * t + 1]);
        auto it = first.find(key);
        if (it == first.end()) {
          first.emplace(key, t);
          rep[t] = t;
          ++uniq;
        } else {
          rep[t] = it->second;
        }
      }
      cudaMemcpyAsync(rep_d, rep.data(), sizeof(int64_t) * (size_t)nt, cudaMemcpyHostToDevice, s);
      apply_representatives<<<GRID(nt, 256), 256, 0, s>>>(p.tdesc, rep_d, nt);
      if (cudaStreamSynchronize(s) != cudaSuccess) {
        fdts_set_error("residual tile plan: applying the pattern dictionary failed");
        rc = 2;
        break;
      }
      p.unique_patterns = uniq;
      if (e->opt_optimize && uniq <= 4096) {
        std::vector<uint2> hm;
        std::vector<uint16_t> hi, hg;
        bool ok = true;
        for (int64_t t = 0; t < nt && ok; ++t) {
          if (rep[t] != t || hd[t].nsl == 0 || hd[t].nidx == 0) continue;
          hm.resize((size_t)hd[t].ncls);
          hi.resize((size_t)hd[t].nidx);
          hg.resize((size_t)hd[t].nsl * 32);
          ok = cudaMemcpy(hm.data(), p.cmeta + hd[t].cls_off, sizeof(uint2) * hm.size(), cudaMemcpyDeviceToHost) == cudaSuccess &&
               cudaMemcpy(hi.data(), p.cidx + hd[t].idx_off, sizeof(uint16_t) * hi.size(), cudaMemcpyDeviceToHost) == cudaSuccess &&
               cudaMemcpy(hg.data(), p.gdst + hd[t].meta_off * 32, sizeof(uint16_t) * hg.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
          if (!ok) break;
          optimize_pattern(hd[t], hm.data(), hi.data(), hg.data(), 1, p.zaddr);
          ok = cudaMemcpy(p.cidx + hd[t].idx_off, hi.data(), sizeof(uint16_t) * hi.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
               cudaMemcpy(p.gdst + hd[t].meta_off * 32, hg.data(), sizeof(uint16_t) * hg.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        }
        if (!ok) {
          fdts_set_error("residual tile plan: pattern optimisation failed");
          rc = 2;
          break;
        }
      }
    }
    p.tstart_dev = tstart;
    tstart = nullptr;
    p.ready = true;
  } while (0);
  void *scratch[] = {tmp, tstart, tbase, tptr, iota, btotal, bnsl, idx_off, meta_off, nsl_true, u64tmp, dmax, imax, derr, bcount,
                     vcount, ucount, bncls, bcells, cnt8, sellpos, hash_d, rep_d};
  for (void *q : scratch)
    if (q) cudaFree(q);
  if (rc) fdts_free_rplan(e);
  return rc;
}
