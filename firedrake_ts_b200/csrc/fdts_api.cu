// fdts_api.cu -- the C ABI of libfdts_b200.so (see include/fdts.h for the
// reference interfaces each entry point replaces).
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "kargs.cuh"

static thread_local std::string g_err;
void fdts_set_error(const std::string &msg) { g_err = msg; }
extern "C" const char *fdts_last_error(void) { return g_err.c_str(); }
extern "C" const char *fdts_version(void) { return "fdts_b200 0.1 (sm_100a)"; }

namespace {

__global__ void transpose_i32(const int32_t *__restrict__ in, int64_t rows, int cols, int32_t *__restrict__ out) {
  // in: rows x cols (AoS), out: cols x rows (SoA)
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * cols) return;
  const int c = (int)(t / rows);
  const int64_t r = t - (int64_t)c * rows;
  out[t] = in[r * cols + c];
}

__global__ void set_values(double *__restrict__ v, const int64_t *__restrict__ slots, int64_t n, double val) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && slots[t] >= 0) v[slots[t]] = val;
}

__global__ void expand_rowptr(const int64_t *__restrict__ nrp, int64_t nown, int nc, int layout, int64_t node_nnz,
                              int64_t *__restrict__ rowptr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nown * nc) return;
  if (t == nown * nc) {
    rowptr[t] = node_nnz * nc * nc;
    return;
  }
  int64_t n;
  int m;
  if (layout == FDTS_INTERLEAVED) {
    n = t / nc;
    m = (int)(t - n * nc);
    const int64_t len = nrp[n + 1] - nrp[n];
    rowptr[t] = nrp[n] * nc * nc + (int64_t)m * len * nc;
  } else {
    m = (int)(t / nown);
    n = t - (int64_t)m * nown;
    rowptr[t] = (int64_t)m * nc * node_nnz + (int64_t)nc * nrp[n];
  }
}

__global__ void expand_colidx(const int64_t *__restrict__ nrp, const int32_t *__restrict__ nci, int64_t nown, int nc,
                              int layout, int64_t node_nnz, int64_t nnodes, int32_t *__restrict__ colidx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (n >= nown) return;
  const int64_t b = nrp[n];
  const int64_t len = nrp[n + 1] - b;
  for (int64_t k = lane; k < len; k += 32) {
    const int64_t col = nci[b + k];
    for (int m = 0; m < nc; ++m)
      for (int g = 0; g < nc; ++g) {
        if (layout == FDTS_INTERLEAVED)
          colidx[b * nc * nc + (int64_t)m * len * nc + k * nc + g] = (int32_t)(col * nc + g);
        else
          colidx[(int64_t)m * nc * node_nnz + (int64_t)nc * b + (int64_t)g * len + k] =
              (int32_t)((int64_t)g * nnodes + col);
      }
  }
}

__global__ void pack_nodes(const double *__restrict__ x, const int32_t *__restrict__ nodes, int64_t count, int nc,
                           int layout, int64_t nnodes, double *__restrict__ buf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * nc) return;
  const int m = (int)(t / count);
  const int64_t k = t - (int64_t)m * count;
  const int64_t n = nodes[k];
  buf[t] = x[layout == FDTS_INTERLEAVED ? n * nc + m : (int64_t)m * nnodes + n];
}

__global__ void unpack_nodes(double *__restrict__ x, int64_t first, int64_t count, int nc, int layout, int64_t nnodes,
                             const double *__restrict__ buf) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * nc) return;
  const int m = (int)(t / count);
  const int64_t n = first + (t - (int64_t)m * count);
  x[layout == FDTS_INTERLEAVED ? n * nc + m : (int64_t)m * nnodes + n] = buf[t];
}

template <typename T>
int upload(T **dst, const T *src, size_t n, int memspace) {
  FDTS_CHECK(cudaMalloc((void **)dst, sizeof(T) * (n > 0 ? n : 1)));
  if (n == 0) return 0;
  FDTS_CHECK(cudaMemcpy(*dst, src, sizeof(T) * n,
                        memspace == FDTS_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
  return 0;
}

}  // namespace

#define GRID(n, b) (unsigned)(((n) + (b)-1) / (b))
#define RET_IF(x)          \
  do {                     \
    int rc_ = (x);         \
    if (rc_) return rc_;   \
  } while (0)

KArgs fdts_make_kargs(fdts_engine *e, const double *x, const double *xdot, double shift, double *out, int64_t c0,
                      int64_t c1) {
  KArgs a;
  a.coords = e->coords;
  a.cell_coord_t = e->cell_coord;
  a.cell_node_t = e->cell_node;
  a.B = e->B;
  a.w = e->w;
  a.x = x;
  a.xdot = xdot;
  a.out = out;
  a.isbc = e->isbc;
  a.rowmask = e->cell_rowmask;
  a.nrowptr = e->nrowptr;
  a.pos8_t = e->pos8;
  a.ncells_total = e->cfg.num_cells;
  a.c0 = c0;
  a.c1 = c1;
  a.nown = e->cfg.num_owned_nodes;
  a.nnodes = e->cfg.num_nodes;
  a.node_nnz = e->node_nnz;
  a.nq = e->cfg.nq;
  a.shift = shift;
  a.p0 = e->cfg.params[0];
  a.p1 = e->cfg.params[1];
  a.p2 = e->cfg.params[2];
#ifdef FDTS_DEV_SWITCHES
  a.dbg = (int32_t)e->opt_debug;
#endif
  return a;
}

int fdts_launch_residual(fdts_engine *e, const double *x, const double *xdot, double *f, int64_t c0, int64_t c1,
                         cudaStream_t s) {
  KArgs a = fdts_make_kargs(e, x, xdot, 0.0, f, c0, c1);
  int rc = 3;
  switch (e->cfg.family) {
    case FDTS_HEAT: rc = fdts_generic_heat(false, e, a, s); break;
    case FDTS_DIFFUSION: rc = fdts_generic_diffusion(false, e, a, s); break;
    case FDTS_BBM: rc = fdts_generic_bbm(false, e, a, s); break;
    case FDTS_BURGERS: rc = fdts_generic_burgers(false, e, a, s); break;
    case FDTS_CAHN_HILLIARD: rc = fdts_generic_ch(false, e, a, s); break;
  }
  if (rc == 3) fdts_set_error("no residual kernel for this family/dimension/degree");
  if (rc == 2) fdts_set_error(std::string("residual kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  if (rc == 0 && c1 > c0) e->launches++;
  return rc;
}

int fdts_launch_jacobian(fdts_engine *e, const double *x, const double *xdot, double shift, double *vals, int64_t c0,
                         int64_t c1, cudaStream_t s) {
  KArgs a = fdts_make_kargs(e, x, xdot, shift, vals, c0, c1);
  int rc = 3;
  switch (e->cfg.family) {
    case FDTS_HEAT: rc = fdts_generic_heat(true, e, a, s); break;
    case FDTS_DIFFUSION: rc = fdts_generic_diffusion(true, e, a, s); break;
    case FDTS_BBM:
      rc = fdts_bbm_jacobian_dmma(e, a, s);  // tensor-core contraction where it applies (P2/P3), else generic
      if (rc == 3) rc = fdts_generic_bbm(true, e, a, s);
      break;
    case FDTS_BURGERS: rc = fdts_generic_burgers(true, e, a, s); break;
    case FDTS_CAHN_HILLIARD: rc = fdts_generic_ch(true, e, a, s); break;
  }
  if (rc == 3) fdts_set_error("no Jacobian kernel for this family/dimension/degree");
  if (rc == 2) fdts_set_error(std::string("Jacobian kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  if (rc == 0 && c1 > c0) e->launches++;
  return rc;
}

extern "C" int fdts_create(const fdts_config *cfg, fdts_handle *out) {
  if (!cfg || !out) {
    fdts_set_error("null argument");
    return 1;
  }
  if (cfg->dim < 1 || cfg->dim > 3 || cfg->nd < 2 || cfg->nd > 20 || cfg->ncomp < 1 || cfg->ncomp > 3) {
    fdts_set_error("unsupported dim/nd/ncomp");
    return 1;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fdts_set_error("no CUDA device available: this engine has no CPU fallback");
    return 2;
  }
  FDTS_CHECK(cudaSetDevice(cfg->device));
  fdts_engine *e = new fdts_engine();
  e->cfg = *cfg;
  e->device = cfg->device;
  const int d = cfg->dim, nd = cfg->nd, ms = cfg->memspace;
  const int64_t nc = cfg->num_cells;
  const bool quad = cfg->cell_type == FDTS_QUADRILATERAL;
  if (cfg->cell_type != FDTS_SIMPLEX && !(quad && d == 2 && cfg->degree == 1 && nd == 4)) {
    fdts_set_error("unsupported cell type (quadrilaterals: Q1 in two dimensions)");
    delete e;
    return 1;
  }
  const int ncv = quad ? 4 : d + 1;  // coordinate nodes per cell
  int rc = 0;
  int32_t *tmp_cc = nullptr, *tmp_cn = nullptr;
  do {
    if ((rc = upload(&e->coords, cfg->coords, (size_t)cfg->num_coord_nodes * d, ms))) break;
    if ((rc = upload(&tmp_cc, cfg->cell_coord, (size_t)nc * ncv, ms))) break;
    if ((rc = upload(&tmp_cn, cfg->cell_node, (size_t)nc * nd, ms))) break;
    if (cudaMalloc(&e->cell_coord, sizeof(int32_t) * (size_t)(nc * ncv + 1)) != cudaSuccess ||
        cudaMalloc(&e->cell_node, sizeof(int32_t) * (size_t)(nc * nd + 1)) != cudaSuccess) {
      fdts_set_error("out of device memory for the cell maps");
      rc = 2;
      break;
    }
    if (nc > 0) {
      transpose_i32<<<GRID(nc * ncv, 256), 256>>>(tmp_cc, nc, ncv, e->cell_coord);
      transpose_i32<<<GRID(nc * nd, 256), 256>>>(tmp_cn, nc, nd, e->cell_node);
    }
    if ((rc = upload(&e->B, cfg->B, (size_t)cfg->nq * (d + 1) * nd, ms))) break;
    if ((rc = upload(&e->w, cfg->w, (size_t)cfg->nq, ms))) break;
    if ((rc = upload(&e->R, cfg->R, (size_t)(d + 1) * (d + 1) * nd * nd, ms))) break;
    if ((rc = upload(&e->M0, cfg->M0, (size_t)nd, ms))) break;
    if ((rc = upload(&e->bc_nodes, cfg->bc_nodes, (size_t)cfg->num_bc_nodes, ms))) break;
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
      fdts_set_error("stream creation failed");
      rc = 2;
      break;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
      fdts_set_error("upload failed");
      rc = 2;
      break;
    }
    rc = fdts_build_pattern(e, e->own_stream);
    if (rc) break;
    {  // which compiled reference table (if any) matches the tensors we were given?
      const size_t nR = (size_t)(d + 1) * (d + 1) * nd * nd;
      std::vector<double> Rh(nR);
      if (cudaMemcpy(Rh.data(), e->R, sizeof(double) * nR, cudaMemcpyDeviceToHost) != cudaSuccess) {
        fdts_set_error("reading back the reference tensors failed");
        rc = 2;
        break;
      }
      e->variant = quad ? -1 : fdts_heat_match_tables(d, cfg->degree, Rh.data());
      // default path: closed form + atomics for the heat family when the tables match
      e->opt_scatter = ((cfg->family == FDTS_HEAT || cfg->family == FDTS_DIFFUSION) && cfg->ncomp == 1 && e->variant >= 0) ? 1 : 0;
    }
  } while (0);
  cudaFree(tmp_cc);
  cudaFree(tmp_cn);
  // host pointers are not retained
  e->cfg.coords = nullptr; e->cfg.cell_coord = nullptr; e->cfg.cell_node = nullptr;
  e->cfg.B = nullptr; e->cfg.w = nullptr; e->cfg.R = nullptr; e->cfg.M0 = nullptr; e->cfg.bc_nodes = nullptr;
  if (rc) {
    fdts_destroy(e);
    return rc;
  }
  *out = e;
  return 0;
}

extern "C" int fdts_destroy(fdts_handle e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  void *ptrs[] = {e->coords, e->cell_coord, e->cell_node, e->B, e->w, e->R, e->M0, e->isbc, e->bc_nodes,
                  e->bc_diag_slot, e->nrowptr, e->ncolidx, e->pos8, e->adj_ptr, e->adj_cell, e->h2d_x, e->h2d_xdot,
                  e->d_f, e->d_vals, e->cell_rowmask, e->pos8_cm, e->scr_x, e->scr_xdot, e->bbm_T};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  fdts_free_plan(e);
  fdts_free_plan2(e);
  fdts_free_rplan(e);
  fdts_free_halo(e);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  if (e->copy_stream) {
    cudaStreamDestroy(e->copy_stream);
    for (cudaEvent_t ev : e->chunk_event)
      if (ev) cudaEventDestroy(ev);
  }
  delete e;
  return 0;
}

extern "C" int fdts_sizes(fdts_handle e, fdts_sizes_t *out) {
  if (!e || !out) return 1;
  const int nc = e->cfg.ncomp;
  out->num_rows = e->cfg.num_owned_nodes * nc;
  out->num_cols = e->cfg.num_nodes * nc;
  out->nnz = e->node_nnz * nc * nc;
  out->node_nnz = e->node_nnz;
  out->max_row_nodes = e->max_row_nodes;
  out->reserved = e->max_degree;
  return 0;
}

extern "C" int fdts_get_node_csr(fdts_handle e, int64_t *rowptr, int32_t *colidx, void *stream) {
  if (!e) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (rowptr)
    FDTS_CHECK(cudaMemcpyAsync(rowptr, e->nrowptr, sizeof(int64_t) * (e->cfg.num_owned_nodes + 1),
                               cudaMemcpyDeviceToDevice, s));
  if (colidx && e->node_nnz)
    FDTS_CHECK(cudaMemcpyAsync(colidx, e->ncolidx, sizeof(int32_t) * e->node_nnz, cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" int fdts_get_csr(fdts_handle e, int64_t *rowptr, int32_t *colidx, void *stream) {
  if (!e) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nown = e->cfg.num_owned_nodes;
  const int nc = e->cfg.ncomp;
  if (rowptr) {
    expand_rowptr<<<GRID(nown * nc + 1, 256), 256, 0, s>>>(e->nrowptr, nown, nc, e->cfg.layout, e->node_nnz, rowptr);
    e->launches++;
  }
  if (colidx && nown > 0) {
    expand_colidx<<<GRID(nown, 4), 128, 0, s>>>(e->nrowptr, e->ncolidx, nown, nc, e->cfg.layout, e->node_nnz,
                                               e->cfg.num_nodes, colidx);
    e->launches++;
  }
  FDTS_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fdts_set_param(fdts_handle e, int32_t index, double value) {
  if (!e || index < 0 || index > 3) return 1;
  e->cfg.params[index] = value;
  return 0;
}

extern "C" int fdts_set_option(fdts_handle e, const char *key, int64_t value) {
  if (!e || !key) return 1;
  if (!strcmp(key, "scatter")) {
    if (value < 0 || value > 2) {
      fdts_set_error("scatter must be 0 (generic, atomic), 1 (closed form, atomic) or 2 (owner-computes gather)");
      return 1;
    }
    if (value > 0 && e->cfg.cell_type != FDTS_SIMPLEX) {
      fdts_set_error("closed-form paths (scatter 1, 2) exist for simplices only");
      return 1;
    }
    e->opt_scatter = value;
    return 0;
  }
  if (!strcmp(key, "gather_residual")) {
    e->opt_gather_residual = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(key, "persist")) {
    e->opt_persist = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(key, "plan_version")) {
    if (value != 1 && value != 2) {
      fdts_set_error("plan_version must be 1 (windowed SELL + residual lists) or 2 (sector SELL + pattern dictionary)");
      return 1;
    }
    e->opt_plan_version = value;
    return 0;
  }
  if (!strcmp(key, "sector")) {
    if (value != 1 && value != 2 && value != 4) {
      fdts_set_error("sector must be 4 (32-byte store groups), 2 (16-byte store groups) or 1 (single slots)");
      return 1;
    }
    e->opt_sector = value;
    e->sector_auto = false;
    return 0;
  }
  if (!strcmp(key, "optimize")) {
    e->opt_optimize = value == 2 ? 2 : (value ? 1 : 0);  // 2: lane orders only, no regrouping of store groups (experiment)
    return 0;
  }
  if (!strcmp(key, "residual_tiles")) {
    e->opt_residual_tiles = value < 0 ? 0 : (value > 3 ? 3 : value);  // 0 off, 1 staged inputs, 2 direct gathers, 3 direct gathers + reduction tables in shared memory
    return 0;
  }
  if (!strcmp(key, "dmma")) {
    e->opt_dmma = value < 0 ? 0 : (value > 2 ? 2 : value);  // 2 tensor contraction, 1 quadrature DMMA, 0 generic
    return 0;
  }
  if (!strcmp(key, "overlap")) {
    e->opt_overlap = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(key, "cache_metric")) {
    e->opt_cache_metric = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(key, "pipelined_readback")) {
    e->opt_pipelined_readback = value ? 1 : 0;
    return 0;
  }
  if (!strcmp(key, "threads")) {
    if (value != 512 && value != 1024) {
      fdts_set_error("threads must be 512 or 1024");
      return 1;
    }
    e->opt_threads = value;
    return 0;
  }
#ifdef FDTS_DEV_SWITCHES
  if (!strcmp(key, "debug")) {
    e->opt_debug = value;
    return 0;
  }
#endif
  if (!strcmp(key, "dedup")) {
    e->opt_dedup = value ? 1 : 0;
    return 0;
  }
  fdts_set_error(std::string("unknown option ") + key);
  return 1;
}

// PETSc hands the callbacks GLOBAL vectors (owned dofs only); the kernels read LOCAL vectors (owned + ghost).
// Interleaved spaces: the owned dofs are the head of the local vector.  Blocked (mixed V*V) spaces: field m
// occupies [m*num_owned, (m+1)*num_owned) of the global and [m*num_nodes, ...) of the local vector.
extern "C" int fdts_owned_to_local(fdts_handle e, const double *owned, double *local, void *stream) {
  if (!e || !owned || !local) {
    fdts_set_error("null argument");
    return 1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nown = (size_t)e->cfg.num_owned_nodes, nn = (size_t)e->cfg.num_nodes, nc = (size_t)e->cfg.ncomp;
  if (nown == 0) return 0;
  if (e->cfg.layout == FDTS_INTERLEAVED || nc == 1) {
    FDTS_CHECK(cudaMemcpyAsync(local, owned, sizeof(double) * nown * nc, cudaMemcpyDeviceToDevice, s));
  } else {
    FDTS_CHECK(cudaMemcpy2DAsync(local, sizeof(double) * nn, owned, sizeof(double) * nown, sizeof(double) * nown, nc,
                                 cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

extern "C" int fdts_set_row_blocks(fdts_handle e, const int64_t *starts_host, int64_t nblocks) {
  if (!e || !starts_host) return 1;
  if (e->cfg.cell_type != FDTS_SIMPLEX) {
    fdts_set_error("gather plans are built for simplices (the quadrilateral path uses the quadrature kernels)");
    return 1;
  }
  FDTS_CHECK(cudaSetDevice(e->device));
  if (e->opt_plan_version == 2) {
    fdts_free_plan(e);
    int rc = fdts_build_plan2(e, starts_host, nblocks);
    // meshes without repeated blocks stream their index tables from HBM on every call: with unoptimised lane orders,
    // 16-byte store groups pad the SELL lists less (4.83 G instead of 6.10 G entries at C4), which outweighs their
    // partial-sector stores (7.89 -> 7.48 ms); with the GPU lane-order optimiser 32-byte groups win (6.42 vs 6.69 ms)
    if (rc == 0 && e->sector_auto && !e->opt_optimize && e->opt_sector == 4 && nblocks >= 64 &&
        e->plan2.unique_patterns * 4 > nblocks) {
      fdts_free_plan2(e);
      e->opt_sector = 2;
      rc = fdts_build_plan2(e, starts_host, nblocks);
      e->opt_sector = 4;
    }
    if (rc == 0) {  // the kernel's shared-memory footprint (vertex-table variant, kernels_heat.cu: launch_gather2_nt)
      const GatherPlan2 &p = e->plan2;
      auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
      const int dim = e->cfg.dim;
      const size_t inA = up16(8 * (size_t)p.vmax * dim + 16) + up16(4 * (size_t)p.cmax + 16);
      const size_t need = up16(sizeof(double) * ((size_t)p.zaddr + 16)) + up16(2 * (size_t)p.max_nidx + 64) +
                          up16(8 * (size_t)PLAN2_MAX_CLASSES + 16) + up16(2 * (size_t)(32 / p.sector) * (size_t)p.max_nsl + 16) +
                          2 * inA + 32 + 4 * sizeof(BlockDesc2);
      if (need > 227 * 1024) {
        fdts_free_plan2(e);
        fdts_set_error("gather plan: a row block needs " + std::to_string(need / 1024) +
                       " KB of shared memory (limit 227 KB): use smaller row blocks");
        return 4;
      }
    }
    return rc;
  }
  fdts_free_plan2(e);
  return fdts_build_plan(e, starts_host, nblocks);
}

extern "C" int fdts_set_cell_tiles(fdts_handle e, const int64_t *starts_host, int64_t ntiles) {
  if (!e || !starts_host) return 1;
  if (e->cfg.cell_type != FDTS_SIMPLEX) {
    fdts_set_error("residual tile plans are built for simplices");
    return 1;
  }
  FDTS_CHECK(cudaSetDevice(e->device));
  return fdts_build_rplan(e, starts_host, ntiles);
}

extern "C" int fdts_get_option(fdts_handle e, const char *key, int64_t *value) {
  if (!e || !key || !value) return 1;
  if (!strcmp(key, "scatter")) {
    *value = e->opt_scatter;
    return 0;
  }
  if (!strcmp(key, "dmma")) {
    *value = e->opt_dmma;
    return 0;
  }
  const bool v2 = e->plan2.ready;
  if (!strcmp(key, "h2d_bytes")) {
    *value = e->h2d_bytes;
    return 0;
  }
  if (!strcmp(key, "d2h_bytes")) {
    *value = e->d2h_bytes;
    return 0;
  }
  if (!strcmp(key, "cache_metric")) {
    *value = e->opt_cache_metric;
    return 0;
  }
  if (!strcmp(key, "plan_metric_cached")) {
    *value = e->plan2.bmetric != nullptr;
    return 0;
  }
  if (!strcmp(key, "plan_sector")) {
    *value = v2 ? e->plan2.sector : 0;
    return 0;
  }
  if (!strcmp(key, "plan_optimized")) {  // 0 no, 1 host (store groups + lane orders), 2 GPU (lane orders)
    *value = v2 ? e->plan2.optimized : 0;
    return 0;
  }
  if (!strcmp(key, "plan_cells_max")) {
    *value = v2 ? e->plan2.cmax : e->plan.cmax;
    return 0;
  }
  if (!strcmp(key, "plan_contributions")) {
    *value = v2 ? e->plan2.total_idx : e->plan.total_contrib;
    return 0;
  }
  if (!strcmp(key, "plan_blocks")) {
    *value = v2 ? e->plan2.nblocks : e->plan.nblocks;
    return 0;
  }
  if (!strcmp(key, "plan_version")) {
    *value = v2 ? 2 : (e->plan.ready ? 1 : 0);
    return 0;
  }
  if (!strcmp(key, "residual_tiles")) {
    *value = e->rplan.ready ? e->rplan.ntiles : 0;
    return 0;
  }
  if (!strcmp(key, "residual_patterns")) {
    *value = e->rplan.ready ? e->rplan.unique_patterns : 0;
    return 0;
  }
  if (!strcmp(key, "plan_patterns")) {
    *value = v2 ? e->plan2.unique_patterns : e->plan.nblocks;
    return 0;
  }
  if (!strcmp(key, "plan_shared_contributions")) {
    *value = v2 ? e->plan2.shared_idx : e->plan.total_contrib;
    return 0;
  }
  if (!strcmp(key, "plan_compacted")) {
    *value = v2 ? e->plan2.compacted : 0;
    return 0;
  }
  if (!strcmp(key, "plan_vertices_max")) {
    *value = v2 ? e->plan2.vmax : e->plan.vmax;
    return 0;
  }
  fdts_set_error(std::string("unknown option ") + key);
  return 1;
}

// closed-form kernels: the heat family, and the diffusion family a (u_t,v) + k (grad u,grad v) - s (1,v) by
// linearity (fdts_launch_heat).  The diffusion Jacobian with k = 0 (a pure mass matrix) and partial cell
// ranges of a k != 1 Jacobian stay on the quadrature kernels.
static bool use_fast_heat(fdts_engine *e, int which = 0, bool full_range = true) {
  if (!(e->opt_scatter >= 1 && e->cfg.ncomp == 1 && e->variant >= 0)) return false;
  if (e->cfg.family == FDTS_HEAT) return true;
  if (e->cfg.family != FDTS_DIFFUSION) return false;
  if (which == 0) return true;
  const double k = e->cfg.params[2];
  return k == 1.0 || (k != 0.0 && full_range);
}

extern "C" int fdts_ifunction_range(fdts_handle e, double t, const double *x, const double *xdot, double *f,
                                    int64_t c0, int64_t c1, int32_t flags, void *stream) {
  (void)t;
  if (!e || !x || !xdot || !f) {
    fdts_set_error("null argument");
    return 1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (use_fast_heat(e) && !(e->opt_scatter == 2 && e->opt_gather_residual && e->plan.ready))
    return fdts_launch_heat(e, 0, x, xdot, 0.0, f, c0, c1, flags & 1, s);  // closed form + atomics on [c0, c1)
  if (flags & 1) FDTS_CHECK(cudaMemsetAsync(f, 0, sizeof(double) * e->cfg.num_owned_nodes * e->cfg.ncomp, s));
  RET_IF(fdts_launch_residual(e, x, xdot, f, c0, c1, s));
  return 0;
}

extern "C" int fdts_ijacobian_range(fdts_handle e, double t, const double *x, const double *xdot, double shift,
                                    double *vals, int64_t c0, int64_t c1, int32_t flags, void *stream) {
  (void)t;
  if (!e || !x || !xdot || !vals) {
    fdts_set_error("null argument");
    return 1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int nc = e->cfg.ncomp;
  if (flags & 1) FDTS_CHECK(cudaMemsetAsync(vals, 0, sizeof(double) * e->node_nnz * nc * nc, s));
  RET_IF(fdts_launch_jacobian(e, x, xdot, shift, vals, c0, c1, s));
  if ((flags & 2) && e->cfg.num_bc_nodes > 0) {
    const int64_t n = e->cfg.num_bc_nodes * nc;
    set_values<<<GRID(n, 256), 256, 0, s>>>(vals, e->bc_diag_slot, n, 1.0);
    e->launches++;
  }
  FDTS_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fdts_ifunction(fdts_handle e, double t, const double *x, const double *xdot, double *f, void *stream) {
  if (!e) return 1;
  if (use_fast_heat(e)) return fdts_launch_heat(e, 0, x, xdot, 0.0, f, 0, e->cfg.num_cells, 3, (cudaStream_t)stream);
  return fdts_ifunction_range(e, t, x, xdot, f, 0, e->cfg.num_cells, 1, stream);
}

extern "C" int fdts_ijacobian(fdts_handle e, double t, const double *x, const double *xdot, double shift, double *vals,
                              void *stream) {
  if (!e) return 1;
  if (use_fast_heat(e, 1)) return fdts_launch_heat(e, 1, x, xdot, shift, vals, 0, e->cfg.num_cells, 3, (cudaStream_t)stream);
  return fdts_ijacobian_range(e, t, x, xdot, shift, vals, 0, e->cfg.num_cells, 3, stream);
}

static int ensure_staging(fdts_engine *e) {
  const int64_t nloc = e->cfg.num_nodes * e->cfg.ncomp;
  if (!e->h2d_x) {
    FDTS_CHECK(cudaMalloc(&e->h2d_x, sizeof(double) * (nloc + 1)));
    FDTS_CHECK(cudaMalloc(&e->h2d_xdot, sizeof(double) * (nloc + 1)));
    FDTS_CHECK(cudaMalloc(&e->d_f, sizeof(double) * (e->cfg.num_owned_nodes * e->cfg.ncomp + 1)));
  }
  return 0;
}

extern "C" int fdts_ifunction_host(fdts_handle e, double t, const double *x, const double *xdot, double *f) {
  if (!e) return 1;
  RET_IF(ensure_staging(e));
  cudaStream_t s = e->own_stream;
  const int64_t nloc = e->cfg.num_nodes * e->cfg.ncomp, nown = e->cfg.num_owned_nodes * e->cfg.ncomp;
  FDTS_CHECK(cudaMemcpyAsync(e->h2d_x, x, sizeof(double) * nloc, cudaMemcpyHostToDevice, s));
  FDTS_CHECK(cudaMemcpyAsync(e->h2d_xdot, xdot, sizeof(double) * nloc, cudaMemcpyHostToDevice, s));
  RET_IF(fdts_ifunction(e, t, e->h2d_x, e->h2d_xdot, e->d_f, s));
  FDTS_CHECK(cudaMemcpyAsync(f, e->d_f, sizeof(double) * nown, cudaMemcpyDeviceToHost, s));
  FDTS_CHECK(cudaStreamSynchronize(s));
  e->h2d_bytes += 2 * (int64_t)sizeof(double) * nloc;
  e->d2h_bytes += (int64_t)sizeof(double) * nown;
  return 0;
}

extern "C" int fdts_ijacobian_host(fdts_handle e, double t, const double *x, const double *xdot, double shift,
                                   double *vals) {
  if (!e) return 1;
  RET_IF(ensure_staging(e));
  const int nc = e->cfg.ncomp;
  const int64_t nnz = e->node_nnz * nc * nc;
  if (!e->d_vals) FDTS_CHECK(cudaMalloc(&e->d_vals, sizeof(double) * (nnz + 1)));
  cudaStream_t s = e->own_stream;
  const int64_t nloc = e->cfg.num_nodes * nc;
  // the Jacobian of the linear families (heat, diffusion: shift a M + k K) does not read the state: no upload
  const bool state_free = e->cfg.family == FDTS_HEAT || e->cfg.family == FDTS_DIFFUSION;
  if (!state_free) {
    FDTS_CHECK(cudaMemcpyAsync(e->h2d_x, x, sizeof(double) * nloc, cudaMemcpyHostToDevice, s));
    FDTS_CHECK(cudaMemcpyAsync(e->h2d_xdot, xdot, sizeof(double) * nloc, cudaMemcpyHostToDevice, s));
    e->h2d_bytes += 2 * (int64_t)sizeof(double) * nloc;
  }
  // Owner-computes gather (plan 2): the CSR values of a range of row blocks are final when its launch ends, so the
  // Jacobian is assembled in 8 block ranges and the read-back of range i runs on a copy stream while range i + 1
  // is assembled (the PCIe copy of the values is 97 % of this call; this hides the kernel behind it).
  constexpr int NCH = 8;
  const bool unscaled = e->cfg.family == FDTS_HEAT || (e->cfg.family == FDTS_DIFFUSION && e->cfg.params[2] == 1.0);
  if (e->opt_pipelined_readback && e->opt_scatter == 2 && e->plan2.ready && !e->opt_overlap && nc == 1 && e->plan2.nblocks >= 64 * NCH && unscaled &&
      use_fast_heat(e, 1)) {
    const int64_t nb = e->plan2.nblocks;
    if (!e->copy_stream) {
      FDTS_CHECK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
      for (int i = 0; i < NCH; ++i) FDTS_CHECK(cudaEventCreateWithFlags(&e->chunk_event[i], cudaEventDisableTiming));
    }
    if (e->chunk_for_blocks != nb) {
      e->chunk_slot.assign(NCH + 1, 0);
      for (int i = 0; i < NCH; ++i)  // s0 is the first member of BlockDesc2
        FDTS_CHECK(cudaMemcpy(&e->chunk_slot[i], reinterpret_cast<const char *>(e->plan2.bdesc + (nb * i) / NCH), sizeof(int64_t),
                              cudaMemcpyDeviceToHost));
      e->chunk_slot[NCH] = nnz;
      e->chunk_for_blocks = nb;
    }
    int rc = 0;
    for (int i = 0; i < NCH && !rc; ++i) {
      e->gather_b0 = (nb * i) / NCH;
      e->gather_b1 = (nb * (i + 1)) / NCH;
      rc = fdts_ijacobian(e, t, e->h2d_x, e->h2d_xdot, shift, e->d_vals, s);
      if (rc) break;
      FDTS_CHECK(cudaEventRecord(e->chunk_event[i], s));
      FDTS_CHECK(cudaStreamWaitEvent(e->copy_stream, e->chunk_event[i], 0));
      const int64_t s0 = e->chunk_slot[i], s1 = e->chunk_slot[i + 1];
      if (s1 > s0)
        FDTS_CHECK(cudaMemcpyAsync(vals + s0, e->d_vals + s0, sizeof(double) * (s1 - s0), cudaMemcpyDeviceToHost, e->copy_stream));
    }
    e->gather_b0 = e->gather_b1 = 0;
    if (rc) return rc;
    FDTS_CHECK(cudaStreamSynchronize(e->copy_stream));
    FDTS_CHECK(cudaStreamSynchronize(s));
    e->d2h_bytes += (int64_t)sizeof(double) * nnz;
    return 0;
  }
  RET_IF(fdts_ijacobian(e, t, e->h2d_x, e->h2d_xdot, shift, e->d_vals, s));
  FDTS_CHECK(cudaMemcpyAsync(vals, e->d_vals, sizeof(double) * nnz, cudaMemcpyDeviceToHost, s));
  FDTS_CHECK(cudaStreamSynchronize(s));
  e->d2h_bytes += (int64_t)sizeof(double) * nnz;
  return 0;
}

extern "C" int fdts_halo_pack(fdts_handle e, const double *x, const int32_t *nodes, int64_t count, double *buf,
                              void *stream) {
  if (!e) return 1;
  if (count <= 0) return 0;
  const int nc = e->cfg.ncomp;
  pack_nodes<<<GRID(count * nc, 256), 256, 0, (cudaStream_t)stream>>>(x, nodes, count, nc, e->cfg.layout,
                                                                      e->cfg.num_nodes, buf);
  e->launches++;
  FDTS_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int fdts_halo_unpack(fdts_handle e, double *x, int64_t first, int64_t count, const double *buf,
                                void *stream) {
  if (!e) return 1;
  if (count <= 0) return 0;
  const int nc = e->cfg.ncomp;
  unpack_nodes<<<GRID(count * nc, 256), 256, 0, (cudaStream_t)stream>>>(x, first, count, nc, e->cfg.layout,
                                                                        e->cfg.num_nodes, buf);
  e->launches++;
  FDTS_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int64_t fdts_launch_count(fdts_handle e) { return e ? e->launches : 0; }
