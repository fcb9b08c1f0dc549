/*
 * fdts.h -- C ABI of libfdts_b200.so, the B200-native assembly engine behind
 * firedrake-ts's IFunction / IJacobian callbacks.
 *
 * Every entry point returns 0 on success and a non-zero code on failure;
 * fdts_last_error() then describes the failure (thread local).  No exception
 * crosses this boundary; there is no global engine state (two solvers may
 * coexist, reference tests/test_two_daesolver.py:6-40): everything hangs off
 * the handle.  Plain pointers and sizes only.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   fdts_create / fdts_destroy
 *       _TSContext.__init__ + inherited _SNESContext.__init__, which build the
 *       assemblers, the residual Cofunction and (first touch of ctx._jac in
 *       set_ijacobian) the sparsity + Mat preallocation:
 *       firedrake_ts/solving_utils.py:77-111, :115-122.
 *   fdts_ifunction
 *       the body of _TSContext.form_function, i.e.
 *       ctx._assemble_residual(tensor=ctx._F) with zero_bc_nodes=True:
 *       firedrake_ts/solving_utils.py:236-267 (assembly at :258).
 *   fdts_ijacobian
 *       the body of _TSContext.form_jacobian, i.e. ctx._shift.assign(shift);
 *       ctx._assemble_jac(ctx._jac): firedrake_ts/solving_utils.py:270-317
 *       (assembly at :304-305).  J = shift*dF/dudot + dF/du,
 *       firedrake_ts/ts_solver.py:108.
 *   fdts_ifunction_host / fdts_ijacobian_host
 *       the same two callbacks when the caller's Vec/Mat live in host memory
 *       (the reference's default VECSEQ/MATSEQAIJ, firedrake_ts/ts_solver.py:239,
 *       solving_utils.py:122); copies are part of the call.
 *   fdts_halo_pack / fdts_halo_unpack
 *       PyOP2's global_to_local halo exchange of the READ Dats that the
 *       reference triggers by leaving dat.vec_wo (solving_utils.py:249-252);
 *       the wire transfer itself (NCCL send/recv) is driven by the host layer.
 *   fdts_get_csr / fdts_sizes
 *       ctx._jac.petscmat.getRowIJ() of the preallocated AIJ Mat.
 */
#ifndef FDTS_H
#define FDTS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdts_engine *fdts_handle;

/* FDTS_DIFFUSION is the two-sided form of the IMEX branch (reference examples/heat-explicit.py:13-14, assembled by
 * _TSContext.form_rhs_function / form_rhs_jacobian, firedrake_ts/solving_utils.py:320-384):
 *   params[1] (u_t, v) + params[2] (grad u, grad v) - params[0] (1, v)
 * F = (u_t, v) is params (0, 1, 0); G = -((grad u, grad v) - (1, v)) is params (-1, 0, -1). */
enum { FDTS_HEAT = 0, FDTS_BURGERS = 1, FDTS_BBM = 2, FDTS_CAHN_HILLIARD = 3, FDTS_DIFFUSION = 4 };
enum { FDTS_INTERLEAVED = 0, FDTS_BLOCKED = 1 };
enum { FDTS_MEM_HOST = 0, FDTS_MEM_DEVICE = 1 };
/* FDTS_QUADRILATERAL: Q1 cells with the isoparametric bilinear map (reference examples/cahn-hilliard.py:14-16,
 * UnitSquareMesh(96, 96, quadrilateral=True)); local node / vertex l = 2*ix + iy (tensor-product order, second
 * factor fastest); cell_coord has 4 columns and B is the Q1 tabulation at the tensor Gauss points. */
enum { FDTS_SIMPLEX = 0, FDTS_QUADRILATERAL = 1 };

typedef struct {
  int32_t dim;       /* 1, 2, 3 */
  int32_t degree;    /* 1, 2, 3 */
  int32_t nd;        /* nodes per cell of the scalar element */
  int32_t ncomp;     /* components (vector space) or fields (mixed space V*V) */
  int32_t layout;    /* FDTS_INTERLEAVED / FDTS_BLOCKED */
  int32_t family;    /* FDTS_HEAT ... */
  int32_t nq;        /* quadrature points */
  int32_t memspace;  /* where the arrays below live: FDTS_MEM_HOST / FDTS_MEM_DEVICE */
  int32_t device;    /* CUDA device ordinal */
  int32_t cell_type; /* FDTS_SIMPLEX / FDTS_QUADRILATERAL (dim 2, degree 1) */
  int64_t num_cells;          /* local cells, interior cells first */
  int64_t num_interior_cells; /* cells touching no ghost node */
  int64_t num_nodes;          /* local nodes: owned first, then ghosts */
  int64_t num_owned_nodes;
  int64_t num_coord_nodes;
  const double *coords;       /* num_coord_nodes x dim */
  const int32_t *cell_coord;  /* num_cells x (dim+1); quadrilaterals: num_cells x 4 */
  const int32_t *cell_node;   /* num_cells x nd */
  const double *B;            /* nq x (dim+1) x nd   basis values / reference derivatives */
  const double *w;            /* nq */
  const double *R;            /* (dim+1) x (dim+1) x nd x nd  reference tensors */
  const double *M0;           /* nd  integrals of the basis functions */
  double params[4];           /* family constants (source | nu | - | lambda) */
  int64_t num_bc_nodes;
  const int32_t *bc_nodes;    /* Dirichlet nodes (all components) */
} fdts_config;

typedef struct {
  int64_t num_rows;        /* owned dofs */
  int64_t num_cols;        /* local dofs (owned + ghost) */
  int64_t nnz;             /* dof-level non-zeros */
  int64_t node_nnz;        /* node-level non-zeros */
  int32_t max_row_nodes;   /* longest node row */
  int32_t reserved;
} fdts_sizes_t;

const char *fdts_last_error(void);
const char *fdts_version(void);

int fdts_create(const fdts_config *cfg, fdts_handle *out);
int fdts_destroy(fdts_handle h);
int fdts_sizes(fdts_handle h, fdts_sizes_t *out);

/* copy the dof-level CSR pattern into caller buffers (device pointers):
 * rowptr: (num_rows+1) int64, colidx: nnz int32; either may be NULL */
int fdts_get_csr(fdts_handle h, int64_t *rowptr_dev, int32_t *colidx_dev, void *stream);
/* node-level pattern (owned node rows x local node columns) */
int fdts_get_node_csr(fdts_handle h, int64_t *rowptr_dev, int32_t *colidx_dev, void *stream);

int fdts_set_param(fdts_handle h, int32_t index, double value);
/* Row blocks of the owner-computes gather path: starts_host[0..nblocks] are owned-node row
 * offsets of contiguous blocks (one CTA each); builds the gather plan on the device.
 * Replaces nothing in the reference: it is the engine's own scatter schedule (K6). */
int fdts_set_row_blocks(fdts_handle h, const int64_t *starts_host, int64_t nblocks);
/* Cell tiles of the residual's tile kernel: starts_host[0..ntiles] are cell offsets of contiguous tiles
 * (cells x nodes per cell <= 4096 per tile).  Builds the tile plan on the device: per tile the unique
 * nodes, vertex table and per-node gather lists, so that the element vectors are pre-reduced in shared
 * memory and one atomic per (tile, node) is issued.  The engine's own schedule (K1), replaces nothing. */
int fdts_set_cell_tiles(fdts_handle h, const int64_t *starts_host, int64_t ntiles);
/* options: "scatter" 0 = generic quadrature kernels + atomic scatter (REDG.ADD.F64),
 *                    1 = closed-form heat kernels + atomic scatter,
 *                    2 = closed-form heat kernels + owner-computes gather (deterministic, write-once);
 * schedule options (set before fdts_set_row_blocks; they never change a value beyond rounding):
 *   "plan_version" 2 (default: sector-grouped SELL + pattern dictionary) | 1 (windowed SELL, residual lists)
 *   "dedup" 1|0 share identical block tables; "optimize" 1|0 bank-conflict-aware ordering of the tables (at most 256
 *   unique patterns: on the host, store groups + lane orders; more, or no dictionary: lane orders on the GPU, one
 *   warp per slice; read back as "plan_optimized" 0|1|2)
 *   "sector" 4|2|1 CSR slots per store group (not set: 4, or 2 when fewer than 3/4 of the blocks repeat a pattern
 *   and "optimize" is off; read back as "plan_sector"); "threads" 512|1024 per persistent CTA
 *   "cache_metric" 1|0 gather Jacobian reads per-block cell metrics precomputed on first use (read back: "plan_metric_cached");
 *   "overlap" 0|1 warp-specialised Jacobian kernel; "residual_tiles" 1|0 use the tile plan when one exists
 *   "gather_residual" 0|1 (plan 1); "persist" 1|0 (plan 1)
 * read-only: "plan_cells_max", "plan_contributions", "plan_blocks", "plan_version", "plan_patterns",
 *   "plan_shared_contributions", "plan_vertices_max", "residual_tiles", "residual_patterns" */
int fdts_set_option(fdts_handle h, const char *key, int64_t value);
int fdts_get_option(fdts_handle h, const char *key, int64_t *value);

/* x, xdot: local vectors (owned + ghost dofs, space layout), device memory.
 * f: owned dofs, device memory.  vals: nnz values of the dof-level CSR, device memory. */
int fdts_ifunction(fdts_handle h, double t, const double *x, const double *xdot, double *f, void *stream);
int fdts_ijacobian(fdts_handle h, double t, const double *x, const double *xdot, double shift, double *vals,
                   void *stream);
/* cell-range variants used to overlap the halo exchange with interior cells:
 * flags bit0 = zero the output first, bit1 = apply the boundary-condition fix-up afterwards */
int fdts_ifunction_range(fdts_handle h, double t, const double *x, const double *xdot, double *f,
                         int64_t cell_begin, int64_t cell_end, int32_t flags, void *stream);
int fdts_ijacobian_range(fdts_handle h, double t, const double *x, const double *xdot, double shift, double *vals,
                         int64_t cell_begin, int64_t cell_end, int32_t flags, void *stream);

/* host-buffer variants: pinned or pageable host pointers; H2D / D2H copies are part of the call (the Jacobian of the
 * linear families does not read the state, so fdts_ijacobian_host uploads nothing for them; on the gather path it
 * assembles the row blocks in 8 ranges and reads each range back while the next is assembled, option
 * "pipelined_readback" 1|0); the bytes moved so far
 * are read back with fdts_get_option "h2d_bytes" / "d2h_bytes" */
int fdts_ifunction_host(fdts_handle h, double t, const double *x_host, const double *xdot_host, double *f_host);
int fdts_ijacobian_host(fdts_handle h, double t, const double *x_host, const double *xdot_host, double shift,
                        double *vals_host);

/* halo: gather `count` nodes (all components) of local vector x into buf / scatter buf into x.
 * nodes: device int32 node ids; buf layout [component][node]. */
int fdts_halo_pack(fdts_handle h, const double *x, const int32_t *nodes, int64_t count, double *buf, void *stream);
int fdts_halo_unpack(fdts_handle h, double *x, int64_t first_node, int64_t count, const double *buf, void *stream);

/* Ghost exchange inside the library (K5 fused around ncclSend/ncclRecv).
 * Replaces PyOP2's global_to_local_begin/_end halves that the reference triggers by leaving dat.vec_wo in
 * _TSContext.form_function / form_jacobian (firedrake_ts/solving_utils.py:249-252, 294-297) on the communicator
 * set up at firedrake_ts/ts_solver.py:233.  One rank per GPU; the library opens its own NCCL communicator
 * (ncclCommInitRank is collective: every rank calls fdts_halo_setup with the same 128-byte id, which rank 0
 * obtains from fdts_nccl_unique_id and the host layer broadcasts).  nccl_library: path of the libnccl.so.2
 * the process already uses (NULL: the loader's default).  A rank may name itself as a neighbour. */
typedef struct {
  int32_t rank, nranks;
  int32_t num_neighbours;
  int32_t reserved;
  const int32_t *neighbour_rank; /* [num_neighbours] */
  const int64_t *send_offset;    /* [num_neighbours + 1] offsets into send_nodes */
  const int32_t *send_nodes;     /* host: owned node ids to pack, per neighbour in the receiver's ghost order */
  const int64_t *recv_first;     /* [num_neighbours] first local (ghost) node of the neighbour's slice */
  const int64_t *recv_count;     /* [num_neighbours] */
  const char *nccl_library;
  const void *nccl_unique_id;    /* 128 bytes */
  /* optional (NULL: the contiguous slices above): explicit ghost node lists, for numberings whose ghosts are not
   * grouped by owner (PyOP2's): recv_nodes[recv_offset[r] .. recv_offset[r+1]) in the sender's packing order */
  const int64_t *recv_offset;    /* [num_neighbours + 1] */
  const int32_t *recv_nodes;     /* host */
} fdts_halo_config;
int fdts_nccl_unique_id(const char *nccl_library, void *id128);
int fdts_halo_setup(fdts_handle h, const fdts_halo_config *cfg);
/* refresh the ghost entries of the local vectors x and xdot (one message per neighbour carries both);
 * stream-ordered: returns after enqueueing, the ghosts are valid for later work on `stream` */
int fdts_halo_exchange(fdts_handle h, double *x, double *xdot, void *stream);
/* IFunction with the exchange hidden behind the interior cells: pack -> grouped send/recv on the library's
 * communication stream -> interior cells -> wait -> unpack -> remaining cells, all enqueued by one call */
int fdts_ifunction_exchange(fdts_handle h, double t, double *x, double *xdot, double *f, void *stream);

/* global (owned dofs, what PETSc's TS passes to form_function / form_jacobian as X, Xdot) -> head of the local
 * (owned + ghost) vector the kernels read; replaces the `X.copy(v)` into ctx._x.dat.vec_wo of the reference
 * (firedrake_ts/solving_utils.py:249-252, 294-297); follow with fdts_halo_exchange on more than one rank */
int fdts_owned_to_local(fdts_handle h, const double *owned, double *local, void *stream);

/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t fdts_launch_count(fdts_handle h);

/* CSR mat-vec / diagonal on device arrays (standalone Newton-Krylov harness; stands in for
 * PETSc's MatMult / MatGetDiagonal on the MATAIJCUSPARSE Mat the reference's KSP would hold) */
int fdts_csr_spmv(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *x,
                  double *y, void *stream);
/* the same product with one warp per row (kept as the comparison kernel for fdts_csr_spmv's CSR-stream scheme) */
int fdts_csr_spmv_warp(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals,
                       const double *x, double *y, void *stream);
int fdts_csr_diagonal(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals,
                      int64_t nown_nodes, int64_t nnodes, int32_t ncomp, int32_t layout, double *d, void *stream);

/* micro-benchmarks used for the roofline denominators: return GB/s resp. GFLOP/s */
int fdts_measure_fp64_peak(int32_t device, int32_t use_dmma, double *gflops);

#ifdef __cplusplus
}
#endif
#endif
