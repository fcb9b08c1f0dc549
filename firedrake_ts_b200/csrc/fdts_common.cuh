// fdts_common.cuh -- engine state and device helpers shared by all translation units.
#pragma once
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/fdts.h"

#define FDTS_MAXD 3

// owner-computes gather plan (gather_plan.cu)
struct GatherPlan {
  bool ready = false;
  int64_t nblocks = 0;
  int64_t *bstart = nullptr;    // nblocks+1 row offsets
  int32_t cmax = 0;             // max cells per block (stride of bcells)
  int32_t *bcells = nullptr;    // nblocks x cmax sorted cell ids
  int32_t *bcount = nullptr;    // cells per block
  int64_t *bbase = nullptr;     // nblocks+1 offsets into cidx
  uint16_t *sdst = nullptr;     // per slot position in SELL order: destination offset inside its 2048-slot window
  uint32_t *soff = nullptr;     // per slice: offset of its first entry in the block's cidx segment (+ end marker)
  int64_t *slbase = nullptr;    // nblocks+1 offsets into soff (padded to 4 entries)
  uint16_t *cidx = nullptr;     // staged element-matrix addresses
  uint16_t *ridx = nullptr;     // per adjacency entry: staged element-vector address
  int64_t total_contrib = 0;
  int32_t max_block_contrib = 0;  // padded to a multiple of 8
  int32_t max_block_slots = 0;
  int32_t max_block_rows = 0;
  int32_t max_block_adj = 0;
  int32_t max_block_slices = 0;   // padded slice-offset entries per block
  int32_t zaddr = 0;              // staged address holding 0.0
  int ne = 0, nep = 0, ndp = 0; // staged entries per cell (symmetric packing), odd-padded strides
  // bulk-copyable phase-A inputs of the persistent kernel
  int32_t vmax = 0;             // max coordinate nodes per block
  double *bvcoords = nullptr;   // nblocks x vmax x dim coordinates of the block's vertices
  uint32_t *bcc = nullptr;      // nblocks x cmax: (dim+1) block-local vertex ids, one byte each
  struct BlockDesc *bdesc = nullptr;
};

struct BlockDesc {              // 48 bytes, one per row block
  int64_t s0;                   // first node-level CSR slot of the block
  int64_t cbase;                // first entry of the block in cidx
  int64_t slbase;               // first entry of the block in soff
  int32_t nslots, ncontrib;     // ncontrib: padded SELL entries (multiple of 32)
  int32_t ncell, nvert;
  int32_t nslpad, pad;          // slice-offset entries incl. end marker, padded to 4
};

// owner-computes gather plan, version 2 (gather_plan.cu: fdts_build_plan2).
// Sector-grouped SELL: the unit of sorting is a 32-byte sector (4 consecutive CSR slots), so the lanes
// that sum a sector also store it (full-sector writes, no transposition buffer, no barriers inside the
// sweep).  Slots with more than SPLIT_CAP contributions are summed by 4 lanes each ("split" slices).
// Blocks whose index lists are identical share one copy (pattern dictionary).
struct BlockDesc2 {             // 80 bytes, one per row block
  int64_t s0;                   // first node-level CSR slot of the block
  int64_t idx_off;              // first entry of the block's pattern in cidx (uint16 units, multiple of 64)
  int64_t meta_off;             // first slice of the pattern in gdst (multiple of 4 slices; 32/sector ids per slice)
  int64_t cc_off;               // first entry of the pattern in bcc (multiple of 4)
  int64_t vc_off;               // first coordinate of the block's vertex table in bvcoords
  int64_t cls_off;              // first class of the pattern in cmeta (multiple of PLAN2_MAX_CLASSES)
  int32_t nslots, nidx;         // CSR slots of the block; SELL entries (multiple of 64)
  int32_t ncell, nvert;
  int32_t nsl, ncls;            // slices; slice classes (runs of slices with equal width / split flag)
  int32_t pattern, pad1;        // pattern id (first block with identical tables)
};
constexpr int PLAN2_MAX_CLASSES = 32;
// slice class (8 bytes): lo = first slice | count << 16, hi = first offset (units of 64 entries) | width << 16 | split << 24

struct GatherPlan2 {
  bool ready = false;
  int64_t nblocks = 0;
  int32_t cmax = 0, vmax = 0, nep = 0, zaddr = 0;
  int32_t max_nidx = 0, max_nsl = 0;  // per-block maxima (max_nsl padded to 4)
  int32_t sector = 4;           // slots per store group: 4 = 32-byte sectors, 2 = 16-byte half sectors
  uint16_t *cidx = nullptr;     // staged element-matrix addresses, SELL order
  uint2 *cmeta = nullptr;       // per block: PLAN2_MAX_CLASSES slice classes
  uint16_t *gdst = nullptr;     // per slice: 32/sector block-local sector ids (0xffff = none)
  uint32_t *bcc = nullptr;      // per cell: (dim+1) block-local vertex ids, one byte each
  double *bvcoords = nullptr;   // per block: vmax x dim vertex coordinates
  BlockDesc2 *bdesc = nullptr;
  double *bmetric = nullptr;    // optional (option "cache_metric"): per block [NP + 1][cmaxp] cell metrics |detJ| G_ab and |detJ|, built on first use
  int32_t cmaxp = 0;            // cmax padded to 4 (row length of bmetric)
  int64_t total_idx = 0, unique_patterns = 0, shared_idx = 0;
  int32_t optimized = 0;        // unique patterns went through the bank-conflict-aware column ordering
  int32_t compacted = 0;        // only the tables of the unique patterns are kept (the duplicates were freed)
};

// residual tile plan (gather_plan.cu: fdts_build_rplan): cells are cut into contiguous tiles; a CTA
// stages the tile's node values and vertex coordinates with coalesced loads, computes the element
// vectors, pre-reduces them per unique node in shared memory (SELL lists, same machinery as plan 2)
// and issues ONE atomic per (tile, node) instead of one per (cell, node).  Identical tiles share
// their tables (pattern dictionary).
struct ResidualPlan {
  bool ready = false;
  int64_t ntiles = 0;
  int32_t cmax = 0, cmaxp = 0, vmax = 0, umaxp = 0, ndp = 0, zaddr = 0;
  int32_t max_nidx = 0, max_nsl = 0;
  uint16_t *cidx = nullptr;     // staged element-vector addresses, SELL order (pairs layout)
  uint2 *cmeta = nullptr;       // slice classes per tile
  uint16_t *gdst = nullptr;     // per slice: 32 tile-local node indices (0xffff = none)
  uint32_t *bcc = nullptr;      // per cell: (dim+1) tile-local vertex ids, one byte each
  double *bvcoords = nullptr;   // per tile: vmax x dim vertex coordinates
  int32_t *tnode = nullptr;     // per tile: umaxp node ids relative to the tile's first node; bit 31 = no output
  uint16_t *tlid = nullptr;     // per tile: [nd][cmaxp] tile-local node index of every cell
  BlockDesc2 *tdesc = nullptr;  // s0 = first node id of the tile, nslots = unique nodes
  int64_t *tstart_host = nullptr;  // ntiles+1 cell offsets (host copy, for cell-range calls)
  int64_t *tstart_dev = nullptr;   // the same on the device
  int64_t unique_patterns = 0;
};

struct fdts_halo;  // halo_nccl.cu

struct fdts_engine {
  fdts_halo *halo = nullptr;  // NCCL ghost exchange state (fdts_halo_setup)
  fdts_config cfg;  // scalars only are meaningful after create; pointers below are device copies
  int device = 0;
  // mesh / maps (device)
  double *coords = nullptr;
  int32_t *cell_coord = nullptr;
  int32_t *cell_node = nullptr;
  // element tables (device)
  double *B = nullptr, *w = nullptr, *R = nullptr, *M0 = nullptr;
  // boundary conditions
  uint8_t *isbc = nullptr;       // per local node
  uint32_t *cell_rowmask = nullptr;  // per cell: bit i = local node i is an owned non-Dirichlet row
  int32_t *bc_nodes = nullptr;   // owned Dirichlet nodes
  int64_t num_bc_owned = 0;
  int64_t *bc_diag_slot = nullptr;  // dof-level slot of the diagonal entry, per (bc node, comp)
  // node-level sparsity
  int64_t *nrowptr = nullptr;    // num_owned_nodes + 1
  int32_t *ncolidx = nullptr;    // node_nnz
  int64_t node_nnz = 0;
  int32_t max_row_nodes = 0;
  uint8_t *pos8 = nullptr;       // num_cells x nd x nd : position of node j inside node row i
  uint8_t *pos8_cm = nullptr;    // cell-major copy (DMMA kernel: one warp per cell), built on first use
  // row -> (cell, local index) adjacency, built for the pattern and reused by gather kernels
  int64_t *adj_ptr = nullptr;    // num_owned_nodes + 1
  int32_t *adj_cell = nullptr;   // cell * 32 + local index (nd <= 20 < 32)
  int32_t max_degree = 0;
  // options
  int64_t opt_scatter = 0;
  int64_t opt_gather_residual = 0;  // gather path also for the residual (slower than atomics today)
  int64_t opt_persist = 1;       // gather path: persistent pipelined Jacobian kernel
  GatherPlan plan;
  GatherPlan2 plan2;
  ResidualPlan rplan;
  int64_t opt_plan_version = 2;  // what fdts_set_row_blocks builds: 1 = windowed SELL (+ residual lists), 2 = sector SELL
  int64_t opt_dedup = 1;         // plan 2: share identical block patterns
  int64_t opt_sector = 4;        // plan 2: slots per store group (4 or 2)
  bool sector_auto = true;       // "sector" not set by the caller: 16-byte groups when the pattern dictionary does not pay and the lane orders are not optimised (the streamed index tables shrink by 21 %, 7.89 -> 7.48 ms at C4 without dictionary)
  int64_t opt_optimize = 1;      // plan 2: bank-conflict-aware column order for the unique patterns
  int64_t opt_residual_tiles = 1; // residual: tile kernel (pre-reduced atomics) when a cell-tile plan exists (fdts_set_cell_tiles; the Python layer only builds one on request: the kernel is slower than the per-cell one today)
  int64_t opt_dmma = 2;          // BBM P2/P3 Jacobian on the fp64 tensor cores: 2 = reference-tensor contraction (default), 1 = quadrature formulation, 0 = generic kernel
  double *bbm_T = nullptr;       // reference tensor of the tensor-contraction BBM Jacobian (built on first use)
  int64_t opt_overlap = 0;       // plan 2: warp-specialised kernel (element matrices of block b+1 overlap the gather of block b; needs two staging buffers in smem; slower than the two-phase kernel on every block shape measured so far)
  int64_t opt_cache_metric = 1;  // plan 2: phase A reads precomputed cell metrics (one TMA copy per block) instead of forming the geometry from the vertex table (30 % of its fp64 work, 13 conflict-prone loads per cell); costs (NP+1)*8 bytes per block cell of device memory and HBM reads
  int64_t opt_pipelined_readback = 1;  // fdts_ijacobian_host on the gather path: assemble in 8 block ranges, read range i back while range i+1 is assembled (A/B on one box at C4: 202.2 against 207.1 ms)
  int64_t opt_threads = 512;     // plan 2: threads per persistent CTA (512: 128 registers per thread, or 1024)
  int64_t opt_debug = 0;         // development only (timing experiments with phases switched off)
  int variant = -1;              // which baked RefT table matches cfg.R (-1: none, closed-form path off)
  // scratch
  double *h2d_x = nullptr, *h2d_xdot = nullptr, *d_f = nullptr, *d_vals = nullptr;  // host-variant staging
  double *scr_x = nullptr, *scr_xdot = nullptr;  // coefficient-scaled states of the diffusion family
  int64_t launches = 0;
  int64_t gather_b0 = 0, gather_b1 = 0;  // transient: block range of the next plan-2 Jacobian launch (0, 0 = all)
  cudaStream_t copy_stream = nullptr;    // host-buffer Jacobian: read-back of chunk i overlaps the kernel of chunk i+1
  cudaEvent_t chunk_event[16] = {};
  std::vector<int64_t> chunk_slot;       // first CSR slot of every chunk (+ nnz), for plan2.nblocks blocks
  int64_t chunk_for_blocks = -1;
  int64_t h2d_bytes = 0, d2h_bytes = 0;  // moved by the host-buffer entry points so far (bench.py's e2e record)
  cudaStream_t own_stream = nullptr;
};

void fdts_set_error(const std::string &msg);
#define FDTS_CHECK(call)                                                                              \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      fdts_set_error(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " at " + __FILE__ + \
                     ":" + std::to_string(__LINE__));                                                 \
      return 2;                                                                                       \
    }                                                                                                 \
  } while (0)

// dispatch entry points implemented per family (kernels_*.cu)
int fdts_launch_residual(fdts_engine *e, const double *x, const double *xdot, double *f, int64_t c0, int64_t c1,
                         cudaStream_t s);
int fdts_launch_jacobian(fdts_engine *e, const double *x, const double *xdot, double shift, double *vals, int64_t c0,
                         int64_t c1, cudaStream_t s);
int fdts_launch_heat(fdts_engine *e, int which, const double *x, const double *xdot, double shift, double *out,
                     int64_t c0, int64_t c1, int flags, cudaStream_t s);
int fdts_build_pattern(fdts_engine *e, cudaStream_t s);
struct KArgs;
int fdts_bbm_jacobian_dmma(fdts_engine *e, const KArgs &a, cudaStream_t s);
int fdts_heat_match_tables(int dim, int deg, const double *R_host);
int fdts_build_plan(fdts_engine *e, const int64_t *starts_host, int64_t nblocks);
int fdts_build_plan2(fdts_engine *e, const int64_t *starts_host, int64_t nblocks);
void fdts_free_plan(fdts_engine *e);
void fdts_free_plan2(fdts_engine *e);
int fdts_build_rplan(fdts_engine *e, const int64_t *tile_starts_host, int64_t ntiles);
void fdts_free_rplan(fdts_engine *e);
void fdts_free_halo(fdts_engine *e);

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ void red_add_f64(double *addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

template <int DIM>
struct Geom {
  double Jinv[DIM][DIM];  // Jinv[a][k] = d xi_a / d x_k
  double adet;
};

template <int DIM>
__device__ __forceinline__ void geometry_from_vertices(const double (&X)[DIM + 1][DIM], Geom<DIM> &g) {
  double J[DIM][DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k)
#pragma unroll
    for (int a = 0; a < DIM; ++a) J[k][a] = X[a + 1][k] - X[0][k];
  if constexpr (DIM == 1) {
    g.adet = fabs(J[0][0]);
    g.Jinv[0][0] = 1.0 / J[0][0];
  } else if constexpr (DIM == 2) {
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double r = 1.0 / det;
    g.Jinv[0][0] = J[1][1] * r;
    g.Jinv[0][1] = -J[0][1] * r;
    g.Jinv[1][0] = -J[1][0] * r;
    g.Jinv[1][1] = J[0][0] * r;
    g.adet = fabs(det);
  } else {
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    double r = 1.0 / det;
    g.Jinv[0][0] = c00 * r;
    g.Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
    g.Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    g.Jinv[1][0] = c01 * r;
    g.Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
    g.Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    g.Jinv[2][0] = c02 * r;
    g.Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    g.Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
    g.adet = fabs(det);
  }
}

template <int DIM>
__device__ __forceinline__ void compute_geometry(const double *__restrict__ coords,
                                                 const int32_t *__restrict__ cc, Geom<DIM> &g) {
  double X[DIM + 1][DIM];
#pragma unroll
  for (int v = 0; v <= DIM; ++v)
#pragma unroll
    for (int k = 0; k < DIM; ++k) X[v][k] = __ldg(coords + (int64_t)cc[v] * DIM + k);
  geometry_from_vertices<DIM>(X, g);
}

// ---- TMA 1-D bulk copy (global -> shared) completing on an mbarrier (SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// dof index helpers (layout: 0 interleaved, 1 blocked)
template <int NC, int LAYOUT>
__device__ __forceinline__ int64_t dof_index(int64_t node, int comp, int64_t n_nodes) {
  if constexpr (LAYOUT == 0) return node * NC + comp;
  return (int64_t)comp * n_nodes + node;
}

// dof-level CSR slot from the node-level pattern:
//   row (node rn, comp m), column (node position `pos` in node row, comp n)
template <int NC, int LAYOUT>
__device__ __forceinline__ int64_t dof_slot(int64_t nrp, int64_t len, int m, int pos, int n, int64_t node_nnz) {
  if constexpr (LAYOUT == 0) return nrp * (NC * NC) + (int64_t)m * len * NC + (int64_t)pos * NC + n;
  return (int64_t)m * NC * node_nnz + (int64_t)NC * nrp + (int64_t)n * len + pos;
}
