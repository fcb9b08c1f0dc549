// kernels_generic.cuh -- K1/K2 generic quadrature kernels (one thread per cell).
//
// These are the general-purpose residual / Jacobian kernels for every form
// family x dimension x degree of the catalogue:
// gather through the cell->node map (SoA, coalesced across the warp), affine
// geometry, quadrature with the basis tables staged in shared memory, and a
// scatter-add with `red.global.add.f64` into the device-resident Vec / CSR
// values.  Dirichlet rows (and columns, for the Jacobian) and ghost rows are
// skipped at scatter time, which reproduces PyOP2's negative-lgmap "drop"
// semantics without a separate pass.
//
// The headline configuration (P2 heat on tetrahedra) does not use these
// kernels; see kernels_heat.cu for the closed-form, owner-computes path.
#pragma once
#include "fdts_common.cuh"

struct KArgs {
  const double *coords;
  const int32_t *cell_coord_t;  // (DIM+1) x ncells_total  (SoA)
  const int32_t *cell_node_t;   // ND x ncells_total
  const double *B;              // nq x (DIM+1) x ND
  const double *w;              // nq
  const double *x, *xdot;
  double *out;
  const uint8_t *isbc;
  const uint32_t *rowmask;      // per cell: owned non-Dirichlet rows (bit per local node)
  const int64_t *nrowptr;
  const uint8_t *pos8_t;        // (ND*ND) x ncells_total
  int64_t ncells_total, c0, c1, nown, nnodes, node_nnz;
  int32_t nq;
  double shift;
  double p0, p1, p2;
#ifdef FDTS_DEV_SWITCHES  // timing experiments only (results wrong); never compiled into the shipped library
  int32_t dbg;
#endif
};

template <int DIM, int NC>
struct Point {
  double u[NC], ud[NC];
  double gu[NC][DIM], gud[NC][DIM];  // physical gradients
};

// pointwise residual fluxes: F_{i,m} += wq * (f0[m] phi_i + sum_k f1[m][k] d_k phi_i)
template <int FAM, int DIM, int NC>
__device__ __forceinline__ void fluxes(const Point<DIM, NC> &s, double p0, double p1, double p2, double (&f0)[NC],
                                       double (&f1)[NC][DIM]) {
  if constexpr (FAM == FDTS_HEAT) {  // examples/heat.py:11
    f0[0] = s.ud[0] - p0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) f1[0][k] = s.gu[0][k];
  } else if constexpr (FAM == FDTS_DIFFUSION) {  // examples/heat-explicit.py:13-14 (F and G are both of this shape)
    f0[0] = p1 * s.ud[0] - p0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) f1[0][k] = p2 * s.gu[0][k];
  } else if constexpr (FAM == FDTS_BBM) {  // examples/bbm.py:35-40
    f0[0] = s.ud[0] + s.gu[0][0] + s.u[0] * s.gu[0][0];
#pragma unroll
    for (int k = 0; k < DIM; ++k) f1[0][k] = 0.0;
    f1[0][0] = s.gud[0][0];
  } else if constexpr (FAM == FDTS_BURGERS) {  // examples/burgers.py:21-25
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      double adv = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) adv += s.u[k] * s.gu[m][k];
      f0[m] = s.ud[m] + adv;
#pragma unroll
      for (int k = 0; k < DIM; ++k) f1[m][k] = p0 * s.gu[m][k];
    }
  } else {  // Cahn-Hilliard, examples/cahn-hilliard.py:26-38
    const double c = s.u[0];
    f0[0] = s.ud[0];
    f0[1] = s.u[1] - 200.0 * c * (1.0 - c) * (1.0 - 2.0 * c);
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      f1[0][k] = s.gu[1][k];
      f1[1][k] = -p0 * s.gu[0][k];
    }
  }
}

// pointwise Jacobian entry for test (i, m), trial (j, n):  shift*dF/dudot + dF/du
template <int FAM, int DIM, int NC>
__device__ __forceinline__ double jac_entry(int m, int n, double pj, const double *Gj, double pi, const double *Gi,
                                            const Point<DIM, NC> &s, double shift, double p0, double p1,
                                            double p2) {
  double gg = 0.0;
#pragma unroll
  for (int k = 0; k < DIM; ++k) gg += Gj[k] * Gi[k];
  if constexpr (FAM == FDTS_HEAT) {
    return shift * pj * pi + gg;
  } else if constexpr (FAM == FDTS_DIFFUSION) {
    return shift * p1 * pj * pi + p2 * gg;
  } else if constexpr (FAM == FDTS_BBM) {
    return shift * (pj * pi + Gj[0] * Gi[0]) + Gj[0] * pi + (s.gu[0][0] * pj + s.u[0] * Gj[0]) * pi;
  } else if constexpr (FAM == FDTS_BURGERS) {
    double v = pj * s.gu[m][n] * pi;
    if (m == n) {
      double ugj = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) ugj += s.u[k] * Gj[k];
      v += shift * pj * pi + ugj * pi + p0 * gg;
    }
    return v;
  } else {
    const double c = s.u[0];
    if (m == 0 && n == 0) return shift * pj * pi;
    if (m == 0 && n == 1) return gg;
    if (m == 1 && n == 0) return -200.0 * (1.0 - 6.0 * c + 6.0 * c * c) * pj * pi - p0 * gg;
    return pj * pi;
  }
}

template <int DIM, int ND, int NC>
__device__ __forceinline__ void eval_point(const double *__restrict__ Bq, const Geom<DIM> &g, const double *ul,
                                           const double *udl, Point<DIM, NC> &s) {
  double gr[NC][DIM], gdr[NC][DIM];
#pragma unroll
  for (int m = 0; m < NC; ++m) {
    s.u[m] = s.ud[m] = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) gr[m][a] = gdr[m][a] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const double b0 = Bq[i];
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      s.u[m] += b0 * ul[i * NC + m];
      s.ud[m] += b0 * udl[i * NC + m];
    }
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      const double ba = Bq[(1 + a) * ND + i];
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        gr[m][a] += ba * ul[i * NC + m];
        gdr[m][a] += ba * udl[i * NC + m];
      }
    }
  }
#pragma unroll
  for (int m = 0; m < NC; ++m)
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; ++a) {
        a0 += g.Jinv[a][k] * gr[m][a];
        a1 += g.Jinv[a][k] * gdr[m][a];
      }
      s.gu[m][k] = a0;
      s.gud[m][k] = a1;
    }
}

// Geometry of cell type CELL: simplices are affine (computed once per cell, before the quadrature loop);
// quadrilaterals carry the isoparametric Q1 map, J[k][a] = sum_v x_v[k] d phi_v / d xi_a at every point
// (reference examples/cahn-hilliard.py:14-16).  NV = coordinate nodes per cell.
template <int DIM, int CELL>
struct CellGeom {
  static constexpr int NV = CELL == FDTS_QUADRILATERAL ? 4 : DIM + 1;
  double X[CELL == FDTS_QUADRILATERAL ? 4 : 1][DIM];  // vertex coordinates (quadrilaterals only)
  Geom<DIM> g;
  __device__ __forceinline__ void load(const double *__restrict__ coords, const int32_t *__restrict__ cell_coord_t,
                                       int64_t ncells_total, int64_t c) {
    int32_t cc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) cc[v] = cell_coord_t[v * ncells_total + c];
    if constexpr (CELL == FDTS_QUADRILATERAL) {
#pragma unroll
      for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int k = 0; k < DIM; ++k) X[v][k] = __ldg(coords + (int64_t)cc[v] * DIM + k);
    } else {
      compute_geometry<DIM>(coords, cc, g);
    }
  }
  // Bq: tabulation at the point, Bq[(1 + a) * ND + v] = d phi_v / d xi_a
  template <int ND>
  __device__ __forceinline__ void at_point(const double *__restrict__ Bq) {
    if constexpr (CELL == FDTS_QUADRILATERAL) {
      static_assert(DIM == 2 && ND == 4, "quadrilaterals: Q1 in two dimensions");
      double J[2][2];
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          double t = 0.0;
#pragma unroll
          for (int v = 0; v < 4; ++v) t += X[v][k] * Bq[(1 + a) * ND + v];
          J[k][a] = t;
        }
      const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      const double r = 1.0 / det;
      g.Jinv[0][0] = J[1][1] * r;
      g.Jinv[0][1] = -J[0][1] * r;
      g.Jinv[1][0] = -J[1][0] * r;
      g.Jinv[1][1] = J[0][0] * r;
      g.adet = fabs(det);
    }
  }
};

template <int DIM, int ND>
__device__ __forceinline__ void load_tables(double *sB, double *sw, const KArgs &a) {
  const int nB = a.nq * (DIM + 1) * ND;
  for (int i = threadIdx.x; i < nB; i += blockDim.x) sB[i] = a.B[i];
  for (int i = threadIdx.x; i < a.nq; i += blockDim.x) sw[i] = a.w[i];
  __syncthreads();
}

// ---------------------------------------------------------------- K1 residual
template <int FAM, int DIM, int ND, int NC, int LAYOUT, int CELL = FDTS_SIMPLEX>
__global__ void __launch_bounds__(128) residual_kernel(const KArgs a) {
  extern __shared__ double smem[];
  double *sB = smem;
  double *sw = smem + a.nq * (DIM + 1) * ND;
  load_tables<DIM, ND>(sB, sw, a);
  const int64_t c = a.c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.c1) return;
  int32_t nodes[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) nodes[i] = a.cell_node_t[i * a.ncells_total + c];
  CellGeom<DIM, CELL> cg;
  cg.load(a.coords, a.cell_coord_t, a.ncells_total, c);
  Geom<DIM> &g = cg.g;
  double ul[ND * NC], udl[ND * NC], Fe[ND * NC];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const int64_t d = dof_index<NC, LAYOUT>(nodes[i], m, a.nnodes);
      ul[i * NC + m] = __ldg(a.x + d);
      udl[i * NC + m] = __ldg(a.xdot + d);
      Fe[i * NC + m] = 0.0;
    }
  for (int q = 0; q < a.nq; ++q) {
    const double *Bq = sB + q * (DIM + 1) * ND;
    cg.template at_point<ND>(Bq);
    Point<DIM, NC> s;
    eval_point<DIM, ND, NC>(Bq, g, ul, udl, s);
    double f0[NC], f1[NC][DIM];
    fluxes<FAM, DIM, NC>(s, a.p0, a.p1, a.p2, f0, f1);
    const double wq = sw[q] * g.adet;
    double r[NC][DIM];  // fluxes pulled back to reference derivatives
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      f0[m] *= wq;
#pragma unroll
      for (int b = 0; b < DIM; ++b) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) t += g.Jinv[b][k] * f1[m][k];
        r[m][b] = t * wq;
      }
    }
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        double t = f0[m] * Bq[i];
#pragma unroll
        for (int b = 0; b < DIM; ++b) t += r[m][b] * Bq[(1 + b) * ND + i];
        Fe[i * NC + m] += t;
      }
  }
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const int64_t n = nodes[i];
    if (n >= a.nown || a.isbc[n]) continue;
#pragma unroll
    for (int m = 0; m < NC; ++m) red_add_f64(a.out + dof_index<NC, LAYOUT>(n, m, a.nown), Fe[i * NC + m]);
  }
}

// ---------------------------------------------------------------- K2 Jacobian
template <int FAM, int DIM, int ND, int NC, int LAYOUT, int CELL = FDTS_SIMPLEX>
__global__ void __launch_bounds__(128) jacobian_kernel(const KArgs a) {
  constexpr int NE = ND * NC;
  constexpr int RB = (48 / NE) < 1 ? 1 : ((48 / NE) > ND ? ND : (48 / NE));  // row nodes per pass
  extern __shared__ double smem[];
  double *sB = smem;
  double *sw = smem + a.nq * (DIM + 1) * ND;
  load_tables<DIM, ND>(sB, sw, a);
  const int64_t c = a.c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.c1) return;
  int32_t nodes[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) nodes[i] = a.cell_node_t[i * a.ncells_total + c];
  CellGeom<DIM, CELL> cg;
  cg.load(a.coords, a.cell_coord_t, a.ncells_total, c);
  Geom<DIM> &g = cg.g;
  double ul[ND * NC], udl[ND * NC];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const int64_t d = dof_index<NC, LAYOUT>(nodes[i], m, a.nnodes);
      ul[i * NC + m] = __ldg(a.x + d);
      udl[i * NC + m] = __ldg(a.xdot + d);
    }
  uint32_t colbc = 0;
#pragma unroll
  for (int j = 0; j < ND; ++j) colbc |= (a.isbc[nodes[j]] ? 1u : 0u) << j;

  for (int i0 = 0; i0 < ND; i0 += RB) {
    double acc[RB][NC][NE];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int m = 0; m < NC; ++m)
#pragma unroll
        for (int e = 0; e < NE; ++e) acc[r][m][e] = 0.0;
    for (int q = 0; q < a.nq; ++q) {
      const double *Bq = sB + q * (DIM + 1) * ND;
      cg.template at_point<ND>(Bq);
      Point<DIM, NC> s;
      eval_point<DIM, ND, NC>(Bq, g, ul, udl, s);
      const double wq = sw[q] * g.adet;
      double G[ND][DIM];
#pragma unroll
      for (int j = 0; j < ND; ++j)
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
          double t = 0.0;
#pragma unroll
          for (int b = 0; b < DIM; ++b) t += g.Jinv[b][k] * Bq[(1 + b) * ND + j];
          G[j][k] = t;
        }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int i = i0 + r;
        if (i < ND) {
          const double pi = Bq[i] * wq;
          double Gi[DIM];
#pragma unroll
          for (int k = 0; k < DIM; ++k) {  // dynamic row index: recompute instead of indexing G
            double t = 0.0;
#pragma unroll
            for (int b = 0; b < DIM; ++b) t += g.Jinv[b][k] * Bq[(1 + b) * ND + i];
            Gi[k] = t * wq;
          }
#pragma unroll
          for (int m = 0; m < NC; ++m)
#pragma unroll
            for (int j = 0; j < ND; ++j)
#pragma unroll
              for (int n = 0; n < NC; ++n)
                acc[r][m][j * NC + n] += jac_entry<FAM, DIM, NC>(m, n, Bq[j], G[j], pi, Gi, s, a.shift, a.p0, a.p1,
                                                              a.p2);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int i = i0 + r;
      if (i >= ND) continue;
      const int64_t rn = a.cell_node_t[i * a.ncells_total + c];
      if (rn >= a.nown || a.isbc[rn]) continue;
      const int64_t nrp = a.nrowptr[rn];
      const int64_t len = a.nrowptr[rn + 1] - nrp;
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        if ((colbc >> j) & 1u) continue;
        const int pos = a.pos8_t[(int64_t)(i * ND + j) * a.ncells_total + c];
#pragma unroll
        for (int m = 0; m < NC; ++m)
#pragma unroll
          for (int n = 0; n < NC; ++n)
            red_add_f64(a.out + dof_slot<NC, LAYOUT>(nrp, len, m, pos, n, a.node_nnz), acc[r][m][j * NC + n]);
      }
    }
  }
}

template <int FAM, int DIM, int ND, int NC, int LAYOUT, int CELL = FDTS_SIMPLEX>
int launch_generic(bool jac, const KArgs &a, cudaStream_t s) {
  const int64_t n = a.c1 - a.c0;
  if (n <= 0) return 0;
  const int threads = 128;
  const int64_t blocks = (n + threads - 1) / threads;
  const size_t smem = sizeof(double) * (size_t)a.nq * ((DIM + 1) * ND + 1);
  if (jac) {
    auto k = jacobian_kernel<FAM, DIM, ND, NC, LAYOUT, CELL>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)blocks, threads, smem, s>>>(a);
  } else {
    auto k = residual_kernel<FAM, DIM, ND, NC, LAYOUT, CELL>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)blocks, threads, smem, s>>>(a);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// scalar families: dims 1-3, degrees 1-3
template <int FAM>
int dispatch_scalar(bool jac, int dim, int degree, const KArgs &a, cudaStream_t s, int cell_type = FDTS_SIMPLEX) {
  if (cell_type == FDTS_QUADRILATERAL) {
    if (dim == 2 && degree == 1) return launch_generic<FAM, 2, 4, 1, 0, FDTS_QUADRILATERAL>(jac, a, s);
    return 3;
  }
#define CASE(D, K, N) \
  if (dim == D && degree == K) return launch_generic<FAM, D, N, 1, 0>(jac, a, s);
  CASE(1, 1, 2) CASE(1, 2, 3) CASE(1, 3, 4)
  CASE(2, 1, 3) CASE(2, 2, 6) CASE(2, 3, 10)
  CASE(3, 1, 4) CASE(3, 2, 10) CASE(3, 3, 20)
#undef CASE
  return 3;
}
