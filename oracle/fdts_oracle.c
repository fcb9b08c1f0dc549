/*
 * fdts_oracle.c -- CPU restatement of the firedrake-ts IFunction / IJacobian
 * assembly path.  TEST INFRASTRUCTURE ONLY: it may be used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
 * never by the product path (firedrake_ts_b200/ never imports it).
 *
 * PARITY UNPINNED: the reference's own tests pin no numbers for this path
 * (tests/test_examples_run.py:19-20 checks exit codes only,
 * tests/test_two_daesolver.py:37-40 only "does not raise"), and the arithmetic
 * lives in un-vendored, un-pinned Firedrake/PyOP2/TSFC/PETSc which are absent
 * here.  The oracle is therefore pinned against exact
 * known answers instead (tests/golden/, generated with sympy by
 * tests/golden/make_golden.py) and against the examples' soft oracles.
 *
 * What it follows:
 *   - callback semantics of firedrake_ts/solving_utils.py:236-267 (form_function)
 *     and :270-317 (form_jacobian): state in, time/shift scalars, assemble;
 *   - J = shift*dF/dudot + dF/du, firedrake_ts/ts_solver.py:108;
 *   - the weak forms as written in examples/heat.py:11, examples/burgers.py:21-25,
 *     examples/bbm.py:35-40, examples/cahn-hilliard.py:26-38;
 *   - the shape of the PyOP2 generated wrapper: per cell gather
 *     through the cell->node map, quadrature element kernel with physical
 *     gradients, `+=` scatter for the residual, per-row search-and-add into CSR
 *     (the MatSetValues cost model) for the Jacobian; Dirichlet rows of the
 *     residual are zeroed (zero_bc_nodes=True, solving_utils.py:105-107),
 *     Dirichlet rows+columns of the Jacobian are dropped and the diagonal set
 *     to one; ghost rows are not assembled (exec-halo redundant compute).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 3
#define MAXND 20
#define MAXC 3

enum { FAM_HEAT = 0, FAM_BURGERS = 1, FAM_BBM = 2, FAM_CH = 3, FAM_DIFFUSION = 4 };
enum { LAYOUT_INTERLEAVED = 0, LAYOUT_BLOCKED = 1 };

typedef struct {
  int32_t dim, nd, ncomp, layout, family, nq;
  int64_t num_cells, num_nodes, num_owned_nodes, num_coord_nodes;
  const double *coords;        /* num_coord_nodes x dim */
  const int32_t *cell_coord;   /* num_cells x (dim+1), quadrilaterals: x 4 */
  const int32_t *cell_node;    /* num_cells x nd */
  const double *B;             /* nq x (dim+1) x nd */
  const double *w;             /* nq */
  double params[4];
  int64_t num_bc_nodes;
  const int32_t *bc_nodes;     /* node ids with Dirichlet conditions (all components) */
  int32_t cell_type;           /* 0 simplex (affine), 1 quadrilateral with the isoparametric Q1 map
                                  (examples/cahn-hilliard.py:14: UnitSquareMesh(..., quadrilateral=True));
                                  cell_coord then has 4 columns and B's derivative rows are the map's */
} oracle_problem;

static inline int64_t dof_local(const oracle_problem *p, int64_t node, int comp) {
  return p->layout == LAYOUT_INTERLEAVED ? node * p->ncomp + comp : (int64_t)comp * p->num_nodes + node;
}
static inline int64_t dof_owned(const oracle_problem *p, int64_t node, int comp) {
  return p->layout == LAYOUT_INTERLEAVED ? node * p->ncomp + comp : (int64_t)comp * p->num_owned_nodes + node;
}

/* geometry at quadrature point q: Jinv[a][k] = d xi_a / d x_k, returns |det J|.
 * Simplices: affine, J[k][a] = x_a[k] - x_0[k] (the same at every point).  Quadrilaterals: the coordinate
 * field is Q1 on the cell's 4 vertices, J[k][a] = sum_v x_v[k] d phi_v / d xi_a at the point. */
static double geometry(const oracle_problem *p, int64_t c, int q, double Jinv[MAXD][MAXD]) {
  const int d = p->dim;
  double X[4][MAXD];
  double J[MAXD][MAXD];
  if (p->cell_type == 1) {
    const double *Bq = p->B + (int64_t)q * (d + 1) * p->nd;
    for (int v = 0; v < 4; ++v)
      for (int k = 0; k < d; ++k) X[v][k] = p->coords[(int64_t)p->cell_coord[c * 4 + v] * d + k];
    for (int k = 0; k < d; ++k)
      for (int a = 0; a < d; ++a) {
        double s = 0.0;
        for (int v = 0; v < 4; ++v) s += X[v][k] * Bq[(1 + a) * p->nd + v];
        J[k][a] = s;
      }
  } else {
    for (int v = 0; v <= d; ++v)
      for (int k = 0; k < d; ++k) X[v][k] = p->coords[(int64_t)p->cell_coord[c * (d + 1) + v] * d + k];
    for (int k = 0; k < d; ++k)
      for (int a = 0; a < d; ++a) J[k][a] = X[a + 1][k] - X[0][k];
  }
  double det;
  if (d == 1) {
    det = J[0][0];
    Jinv[0][0] = 1.0 / det;
  } else if (d == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    Jinv[0][0] = J[1][1] / det;  Jinv[0][1] = -J[0][1] / det;
    Jinv[1][0] = -J[1][0] / det; Jinv[1][1] = J[0][0] / det;
  } else {
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    Jinv[0][0] = c00 / det;
    Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    Jinv[1][0] = c01 / det;
    Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    Jinv[2][0] = c02 / det;
    Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  }
  return fabs(det);
}

/* values of all fields and their physical gradients at one quadrature point */
typedef struct {
  double phi[MAXND];
  double G[MAXND][MAXD]; /* physical gradient of phi_i */
  double u[MAXC], ud[MAXC];
  double gu[MAXC][MAXD], gud[MAXC][MAXD];
  double wq;
} qpoint;

static void eval_qpoint(const oracle_problem *p, int q, double adet, double Jinv[MAXD][MAXD],
                        const double *ul, const double *udl, qpoint *s) {
  const int d = p->dim, nd = p->nd, nc = p->ncomp;
  const double *Bq = p->B + (int64_t)q * (d + 1) * nd;
  for (int i = 0; i < nd; ++i) {
    s->phi[i] = Bq[i];
    for (int k = 0; k < d; ++k) {
      double g = 0.0;
      for (int a = 0; a < d; ++a) g += Jinv[a][k] * Bq[(1 + a) * nd + i];
      s->G[i][k] = g;
    }
  }
  for (int m = 0; m < nc; ++m) {
    s->u[m] = s->ud[m] = 0.0;
    for (int k = 0; k < d; ++k) s->gu[m][k] = s->gud[m][k] = 0.0;
    for (int i = 0; i < nd; ++i) {
      double a = ul[i * nc + m], b = udl[i * nc + m];
      s->u[m] += s->phi[i] * a;
      s->ud[m] += s->phi[i] * b;
      for (int k = 0; k < d; ++k) {
        s->gu[m][k] += s->G[i][k] * a;
        s->gud[m][k] += s->G[i][k] * b;
      }
    }
  }
  s->wq = p->w[q] * adet;
}

static inline double dotd(const double *a, const double *b, int d) {
  double s = 0.0;
  for (int k = 0; k < d; ++k) s += a[k] * b[k];
  return s;
}

/* element residual: Fe[i*nc + m] */
static void element_residual(const oracle_problem *p, int64_t c, const double *ul, const double *udl, double *Fe) {
  const int d = p->dim, nd = p->nd, nc = p->ncomp;
  double Jinv[MAXD][MAXD];
  for (int i = 0; i < nd * nc; ++i) Fe[i] = 0.0;
  qpoint s;
  double adet = 0.0;
  for (int q = 0; q < p->nq; ++q) {
    if (q == 0 || p->cell_type == 1) adet = geometry(p, c, q, Jinv); /* affine cells: once per cell */
    eval_qpoint(p, q, adet, Jinv, ul, udl, &s);
    for (int i = 0; i < nd; ++i) {
      const double ph = s.phi[i];
      const double *Gi = s.G[i];
      switch (p->family) {
        case FAM_HEAT: /* examples/heat.py:11 */
          Fe[i] += s.wq * (s.ud[0] * ph + dotd(s.gu[0], Gi, d) - p->params[0] * ph);
          break;
        case FAM_DIFFUSION: /* examples/heat-explicit.py:13-14: a (u_t, v) + k (grad u, grad v) - s (1, v) */
          Fe[i] += s.wq * (p->params[1] * s.ud[0] * ph + p->params[2] * dotd(s.gu[0], Gi, d) - p->params[0] * ph);
          break;
        case FAM_BBM: /* examples/bbm.py:35-40 */
          Fe[i] += s.wq * (s.ud[0] * ph + s.gu[0][0] * ph + s.u[0] * s.gu[0][0] * ph + s.gud[0][0] * Gi[0]);
          break;
        case FAM_BURGERS: /* examples/burgers.py:21-25 */
          for (int m = 0; m < nc; ++m) {
            double adv = 0.0;
            for (int k = 0; k < d; ++k) adv += s.u[k] * s.gu[m][k];
            Fe[i * nc + m] += s.wq * (s.ud[m] * ph + adv * ph + p->params[0] * dotd(s.gu[m], Gi, d));
          }
          break;
        case FAM_CH: { /* examples/cahn-hilliard.py:26-38 */
          const double cc = s.u[0], mu = s.u[1], lam = p->params[0];
          const double dfdc = 200.0 * cc * (1.0 - cc) * (1.0 - 2.0 * cc);
          Fe[i * nc + 0] += s.wq * (s.ud[0] * ph + dotd(s.gu[1], Gi, d));
          Fe[i * nc + 1] += s.wq * (mu * ph - dfdc * ph - lam * dotd(s.gu[0], Gi, d));
        } break;
      }
    }
  }
}

/* element Jacobian: Ae[(i*nc+m) * (nd*nc) + (j*nc+n)] */
static void element_jacobian(const oracle_problem *p, int64_t c, const double *ul, const double *udl, double shift,
                             double *Ae) {
  const int d = p->dim, nd = p->nd, nc = p->ncomp, ne = nd * nc;
  double Jinv[MAXD][MAXD];
  for (int i = 0; i < ne * ne; ++i) Ae[i] = 0.0;
  qpoint s;
  double adet = 0.0;
  for (int q = 0; q < p->nq; ++q) {
    if (q == 0 || p->cell_type == 1) adet = geometry(p, c, q, Jinv); /* affine cells: once per cell */
    eval_qpoint(p, q, adet, Jinv, ul, udl, &s);
    for (int i = 0; i < nd; ++i)
      for (int j = 0; j < nd; ++j) {
        const double pi = s.phi[i], pj = s.phi[j];
        const double *Gi = s.G[i], *Gj = s.G[j];
        const double gg = dotd(Gi, Gj, d);
        switch (p->family) {
          case FAM_HEAT:
            Ae[i * ne + j] += s.wq * (shift * pj * pi + gg);
            break;
          case FAM_DIFFUSION:
            Ae[i * ne + j] += s.wq * (shift * p->params[1] * pj * pi + p->params[2] * gg);
            break;
          case FAM_BBM:
            Ae[i * ne + j] += s.wq * (shift * (pj * pi + Gj[0] * Gi[0]) + Gj[0] * pi +
                                      (s.gu[0][0] * pj + s.u[0] * Gj[0]) * pi);
            break;
          case FAM_BURGERS: {
            double ugj = 0.0;
            for (int k = 0; k < d; ++k) ugj += s.u[k] * Gj[k];
            for (int m = 0; m < nc; ++m)
              for (int n = 0; n < nc; ++n) {
                double v = pj * s.gu[m][n] * pi; /* (du . grad) u */
                if (m == n) v += shift * pj * pi + ugj * pi + p->params[0] * gg;
                Ae[(i * nc + m) * ne + (j * nc + n)] += s.wq * v;
              }
          } break;
          case FAM_CH: {
            const double cc = s.u[0], lam = p->params[0];
            const double d2f = 200.0 * (1.0 - 6.0 * cc + 6.0 * cc * cc);
            Ae[(i * nc + 0) * ne + (j * nc + 0)] += s.wq * (shift * pj * pi);
            Ae[(i * nc + 0) * ne + (j * nc + 1)] += s.wq * gg;
            Ae[(i * nc + 1) * ne + (j * nc + 0)] += s.wq * (-d2f * pj * pi - lam * gg);
            Ae[(i * nc + 1) * ne + (j * nc + 1)] += s.wq * (pj * pi);
          } break;
        }
      }
  }
}

static void gather(const oracle_problem *p, int64_t c, const double *x, double *xl) {
  const int nd = p->nd, nc = p->ncomp;
  for (int i = 0; i < nd; ++i) {
    int64_t node = p->cell_node[c * nd + i];
    for (int m = 0; m < nc; ++m) xl[i * nc + m] = x[dof_local(p, node, m)];
  }
}

/* ---- IFunction: F (owned layout, num_owned_nodes*ncomp) from local u, udot ---- */
int oracle_ifunction(const oracle_problem *p, double t, const double *u, const double *udot, double *F) {
  (void)t;
  const int nd = p->nd, nc = p->ncomp;
  const int64_t nown = p->num_owned_nodes;
  memset(F, 0, sizeof(double) * (size_t)(nown * nc));
  double ul[MAXND * MAXC], udl[MAXND * MAXC], Fe[MAXND * MAXC];
  for (int64_t c = 0; c < p->num_cells; ++c) {
    gather(p, c, u, ul);
    gather(p, c, udot, udl);
    element_residual(p, c, ul, udl, Fe);
    for (int i = 0; i < nd; ++i) {
      int64_t node = p->cell_node[c * nd + i];
      if (node >= nown) continue; /* ghost rows are not assembled */
      for (int m = 0; m < nc; ++m) F[dof_owned(p, node, m)] += Fe[i * nc + m];
    }
  }
  for (int64_t b = 0; b < p->num_bc_nodes; ++b) {
    int64_t node = p->bc_nodes[b];
    if (node >= nown) continue;
    for (int m = 0; m < nc; ++m) F[dof_owned(p, node, m)] = 0.0;
  }
  return 0;
}

/* ---- sparsity (node level): rows = owned nodes, columns = local nodes, sorted ---- */
static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

/* pass 1 (colidx == NULL): fills rowptr[0..nown]; pass 2: fills colidx */
int oracle_node_pattern(const oracle_problem *p, int64_t *rowptr, int32_t *colidx) {
  const int nd = p->nd;
  const int64_t nown = p->num_owned_nodes, nc = p->num_cells;
  int64_t *cnt = calloc((size_t)nown + 1, sizeof(int64_t));
  for (int64_t c = 0; c < nc; ++c)
    for (int i = 0; i < nd; ++i) {
      int64_t n = p->cell_node[c * nd + i];
      if (n < nown) cnt[n + 1]++;
    }
  for (int64_t n = 0; n < nown; ++n) cnt[n + 1] += cnt[n];
  int64_t *adj = malloc(sizeof(int64_t) * (size_t)(cnt[nown] > 0 ? cnt[nown] : 1));
  int64_t *fill = malloc(sizeof(int64_t) * (size_t)(nown + 1));
  memcpy(fill, cnt, sizeof(int64_t) * (size_t)(nown + 1));
  for (int64_t c = 0; c < nc; ++c)
    for (int i = 0; i < nd; ++i) {
      int64_t n = p->cell_node[c * nd + i];
      if (n < nown) adj[fill[n]++] = c;
    }
  int32_t *tmp = malloc(sizeof(int32_t) * 4096);
  size_t cap = 4096;
  int64_t nnz = 0;
  rowptr[0] = 0;
  for (int64_t n = 0; n < nown; ++n) {
    size_t need = (size_t)(cnt[n + 1] - cnt[n]) * nd;
    if (need > cap) { cap = need * 2; tmp = realloc(tmp, sizeof(int32_t) * cap); }
    size_t k = 0;
    for (int64_t a = cnt[n]; a < cnt[n + 1]; ++a)
      for (int j = 0; j < nd; ++j) tmp[k++] = p->cell_node[adj[a] * nd + j];
    qsort(tmp, k, sizeof(int32_t), cmp_i32);
    size_t u = 0;
    for (size_t a = 0; a < k; ++a)
      if (a == 0 || tmp[a] != tmp[a - 1]) {
        if (colidx) colidx[nnz + u] = tmp[a];
        u++;
      }
    nnz += (int64_t)u;
    rowptr[n + 1] = nnz;
  }
  free(tmp); free(fill); free(adj); free(cnt);
  return 0;
}

/* dof-level CSR (what PETSc AIJ stores) from the node-level pattern */
int oracle_dof_pattern(const oracle_problem *p, const int64_t *nrowptr, const int32_t *ncolidx, int64_t *rowptr,
                       int32_t *colidx) {
  const int nc = p->ncomp;
  const int64_t nown = p->num_owned_nodes;
  const int64_t nnzn = nrowptr[nown];
  rowptr[0] = 0;
  if (p->layout == LAYOUT_INTERLEAVED) {
    for (int64_t n = 0; n < nown; ++n) {
      const int64_t len = nrowptr[n + 1] - nrowptr[n];
      for (int m = 0; m < nc; ++m) {
        int64_t row = n * nc + m;
        int64_t base = nrowptr[n] * nc * nc + (int64_t)m * len * nc;
        rowptr[row + 1] = base + len * nc;
        if (colidx)
          for (int64_t k = 0; k < len; ++k)
            for (int g = 0; g < nc; ++g) colidx[base + k * nc + g] = ncolidx[nrowptr[n] + k] * nc + g;
      }
    }
  } else {
    for (int m = 0; m < nc; ++m)
      for (int64_t n = 0; n < nown; ++n) {
        const int64_t len = nrowptr[n + 1] - nrowptr[n];
        int64_t row = (int64_t)m * nown + n;
        int64_t base = (int64_t)m * nc * nnzn + (int64_t)nc * nrowptr[n];
        rowptr[row + 1] = base + len * nc;
        if (colidx)
          for (int g = 0; g < nc; ++g)
            for (int64_t k = 0; k < len; ++k)
              colidx[base + g * len + k] = (int32_t)((int64_t)g * p->num_nodes + ncolidx[nrowptr[n] + k]);
      }
  }
  return 0;
}

/* MatSetValues-like insertion: binary search of the column in the (sorted) row */
static inline int64_t find_slot(const int64_t *rowptr, const int32_t *colidx, int64_t row, int32_t col) {
  int64_t lo = rowptr[row], hi = rowptr[row + 1] - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t v = colidx[mid];
    if (v == col) return mid;
    if (v < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* ---- IJacobian into dof-level CSR values (rows in owned layout, columns in local layout) ---- */
int oracle_ijacobian(const oracle_problem *p, double t, const double *u, const double *udot, double shift,
                     const int64_t *rowptr, const int32_t *colidx, double *vals) {
  (void)t;
  const int nd = p->nd, nc = p->ncomp, ne = nd * nc;
  const int64_t nown = p->num_owned_nodes;
  const int64_t nrows = nown * nc;
  memset(vals, 0, sizeof(double) * (size_t)rowptr[nrows]);
  char *isbc = calloc((size_t)p->num_nodes + 1, 1);
  for (int64_t b = 0; b < p->num_bc_nodes; ++b) isbc[p->bc_nodes[b]] = 1;
  double ul[MAXND * MAXC], udl[MAXND * MAXC];
  double *Ae = malloc(sizeof(double) * (size_t)ne * ne);
  int rc = 0;
  for (int64_t c = 0; c < p->num_cells; ++c) {
    gather(p, c, u, ul);
    gather(p, c, udot, udl);
    element_jacobian(p, c, ul, udl, shift, Ae);
    for (int i = 0; i < nd; ++i) {
      int64_t rn = p->cell_node[c * nd + i];
      if (rn >= nown || isbc[rn]) continue; /* ghost rows / Dirichlet rows dropped */
      for (int m = 0; m < nc; ++m) {
        int64_t row = dof_owned(p, rn, m);
        for (int j = 0; j < nd; ++j) {
          int64_t cn = p->cell_node[c * nd + j];
          if (isbc[cn]) continue; /* Dirichlet columns dropped */
          for (int n = 0; n < nc; ++n) {
            int64_t s = find_slot(rowptr, colidx, row, (int32_t)dof_local(p, cn, n));
            if (s < 0) { rc = 1; continue; }
            vals[s] += Ae[(i * nc + m) * ne + (j * nc + n)];
          }
        }
      }
    }
  }
  for (int64_t b = 0; b < p->num_bc_nodes; ++b) { /* unit diagonal on Dirichlet rows */
    int64_t node = p->bc_nodes[b];
    if (node >= nown) continue;
    for (int m = 0; m < nc; ++m) {
      int64_t s = find_slot(rowptr, colidx, dof_owned(p, node, m), (int32_t)dof_local(p, node, m));
      if (s < 0) rc = 1; else vals[s] = 1.0;
    }
  }
  free(Ae); free(isbc);
  return rc;
}

/* single element matrices/vectors, exposed for the known-answer tests */
int oracle_element(const oracle_problem *p, int64_t c, const double *u, const double *udot, double shift,
                   double *Fe, double *Ae) {
  double ul[MAXND * MAXC], udl[MAXND * MAXC];
  gather(p, c, u, ul);
  gather(p, c, udot, udl);
  if (Fe) element_residual(p, c, ul, udl, Fe);
  if (Ae) element_jacobian(p, c, ul, udl, shift, Ae);
  return 0;
}
