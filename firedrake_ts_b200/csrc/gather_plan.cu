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
