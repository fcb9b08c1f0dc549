// kernels_dmma.cu -- fp64 tensor-core (DMMA, mma.sync.m8n8k4.f64) element-matrix contraction for the
// BBM family (examples/bbm.py:35-40) on high-order simplices: configuration C5 of SURVEY.md
// section 8(d) (P3 tetrahedra), "K2 ... P3: DMMA contraction variant".
//
//   J_ij = sum_q W_q [ (shift + u_x) phi_i phi_j + (1 + u) phi_i psi_j + shift psi_i psi_j ],
//   psi = d/dx_0 = sum_a Jinv[a][0] d_a phi,  W_q = w_q |detJ|
//
// i.e. three dense products over the quadrature index with diagonal weights that depend on the state:
//   A = Phi^T D1 Phi + (Phi^T D2 + Psi^T D3) Psi .
// One warp owns one cell: the basis tables live in shared memory (zero padded to 8 x 4 tiles, row stride
// 28 doubles => conflict-free fragment loads), the weights D1..D3 of the cell are computed by the warp
// (u, u_x at the quadrature points), then 2 x NT8^2 DMMAs per 4 quadrature points accumulate the
// NT8 x NT8 tiles of the element matrix in registers; finally one REDG per entry through a cell-major
// copy of the row-relative slot map.  The generic kernel (thread per cell, 254 registers) recomputes the
// physical gradients for every row pass: ncu shows 73 % fp64-pipe utilisation for 10x the algorithmic
// flops; this kernel does the algorithmic 2 nq nd^2 x 3 flops on the tensor pipe.
#include "kargs.cuh"

namespace {

__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

constexpr int DMMA_STRIDE = 28;  // doubles per quadrature row of a table (>= 24, = 12 mod 16: see header)
constexpr int DMMA_WARPS = 16;

__global__ void pos8_cell_major(const uint8_t *__restrict__ pos8_t, int64_t nc, int nd2, uint8_t *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nd2) return;
  const int64_t c = t / nd2;
  const int e = (int)(t - c * nd2);
  out[t] = pos8_t[(int64_t)e * nc + c];
}

template <int DIM, int ND>
__global__ void __launch_bounds__(DMMA_WARPS * 32, 1) bbm_jacobian_dmma(const KArgs a, const uint8_t *__restrict__ pos8_cm) {
  constexpr int NT8 = (ND + 7) / 8, S = DMMA_STRIDE;
  static_assert(NT8 * 8 <= 24, "element too large for the table stride");
  extern __shared__ double sm[];
  const int nqp = (a.nq + 7) & ~7;
  double *sT = sm;                              // [(DIM+1)][nqp][S]  table alpha: 0 values, 1.. reference derivatives
  double *sw = sT + (size_t)(DIM + 1) * nqp * S; // [nqp]
  double *sdall = sw + nqp;                     // per warp [3][nqp]
  for (int i = threadIdx.x; i < (DIM + 1) * nqp * S; i += blockDim.x) {
    const int al = i / (nqp * S), r = i - al * nqp * S, q = r / S, j = r - q * S;
    sT[i] = (q < a.nq && j < ND) ? a.B[((size_t)q * (DIM + 1) + al) * ND + j] : 0.0;
  }
  for (int i = threadIdx.x; i < nqp; i += blockDim.x) sw[i] = i < a.nq ? a.w[i] : 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *sd = sdall + (size_t)warp * 3 * nqp;
  const int g4 = lane >> 2, t4 = lane & 3;
  for (int64_t c = a.c0 + (int64_t)blockIdx.x * DMMA_WARPS + warp; c < a.c1; c += (int64_t)gridDim.x * DMMA_WARPS) {
    // ---- cell data: nodes, state, geometry (all lanes redundantly: broadcast loads)
    int32_t node = 0;
    double uj = 0.0;
    bool bc = true;
    if (lane < ND) {
      node = a.cell_node_t[(int64_t)lane * a.ncells_total + c];
      uj = __ldg(a.x + node);
      bc = a.isbc[node] != 0;
    }
    const uint32_t colbc = __ballot_sync(0xffffffffu, bc);
    int32_t cc[DIM + 1];
#pragma unroll
    for (int v = 0; v <= DIM; ++v) cc[v] = a.cell_coord_t[(int64_t)v * a.ncells_total + c];
    Geom<DIM> g;
    compute_geometry<DIM>(a.coords, cc, g);
    double ca[DIM];
#pragma unroll
    for (int b = 0; b < DIM; ++b) ca[b] = g.Jinv[b][0];
    // ---- u and u_x at the quadrature points: lane = (q mod 8, quarter of the basis functions)
    {
      constexpr int JP = (ND + 3) / 4;  // basis functions per quarter
      double ul[JP];
#pragma unroll
      for (int jj = 0; jj < JP; ++jj) {
        const int j = t4 * JP + jj;
        ul[jj] = __shfl_sync(0xffffffffu, uj, j < ND ? j : 0);
        if (j >= ND) ul[jj] = 0.0;
      }
      for (int q0 = 0; q0 < nqp; q0 += 8) {
        const int q = q0 + g4;
        double s0 = 0.0, sx = 0.0;
#pragma unroll
        for (int jj = 0; jj < JP; ++jj) {
          const int j = t4 * JP + jj;
          if (j < ND) {
            const double *tp = sT + (size_t)q * S + j;
            s0 = fma(ul[jj], tp[0], s0);
            double ps = 0.0;
#pragma unroll
            for (int b = 0; b < DIM; ++b) ps = fma(ca[b], tp[(size_t)(1 + b) * nqp * S], ps);
            sx = fma(ul[jj], ps, sx);
          }
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        sx += __shfl_xor_sync(0xffffffffu, sx, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        sx += __shfl_xor_sync(0xffffffffu, sx, 2);
        if (t4 == 0) {
          const double W = sw[q] * g.adet;
          sd[q] = W * (a.shift + sx);
          sd[nqp + q] = W * (1.0 + s0);
          sd[2 * nqp + q] = W * a.shift;
        }
      }
    }
    __syncwarp();
    // ---- contraction on the tensor pipe
    double C[NT8][NT8][2];
#pragma unroll
    for (int mt = 0; mt < NT8; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) C[mt][nt][0] = C[mt][nt][1] = 0.0;
    for (int q0 = 0; q0 < nqp; q0 += 4) {
      const int q = q0 + t4;
      const double d1 = sd[q], d2 = sd[nqp + q], d3 = sd[2 * nqp + q];
      double phi[NT8], psi[NT8];
#pragma unroll
      for (int t = 0; t < NT8; ++t) {
        const double *tp = sT + (size_t)q * S + 8 * t + g4;
        phi[t] = tp[0];
        double ps = 0.0;
#pragma unroll
        for (int b = 0; b < DIM; ++b) ps = fma(ca[b], tp[(size_t)(1 + b) * nqp * S], ps);
        psi[t] = ps;
      }
#pragma unroll
      for (int mt = 0; mt < NT8; ++mt) {
        // A operands (row i = 8 mt + g4, column q): phi_i D1  and  phi_i D2 + psi_i D3
        const double a1 = phi[mt] * d1, a23 = fma(phi[mt], d2, psi[mt] * d3);
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
          dmma_m8n8k4(C[mt][nt], a1, phi[nt]);   // B operand (row q, column j = 8 nt + g4)
          dmma_m8n8k4(C[mt][nt], a23, psi[nt]);
        }
      }
    }
    // ---- scatter: C[mt][nt][h] is entry (i, j) = (8 mt + g4, 8 nt + 2 t4 + h)
#pragma unroll
    for (int mt = 0; mt < NT8; ++mt) {
      const int i = 8 * mt + g4;
      const int32_t rn = __shfl_sync(0xffffffffu, node, i < ND ? i : 0);
      const bool rowok = i < ND && rn < a.nown && !((colbc >> i) & 1u);
      const int64_t nrp = rowok ? a.nrowptr[rn] : 0;
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = 8 * nt + 2 * t4 + h;
          if (rowok && j < ND && !((colbc >> j) & 1u)) {
            const int pos = pos8_cm[(size_t)c * (ND * ND) + i * ND + j];
            red_add_f64(a.out + nrp + pos, C[mt][nt][h]);
          }
        }
    }
    __syncwarp();
  }
}


// ------------------------------------------------------------------ tensor-contraction variant (default)
// On an affine cell the BBM Jacobian is LINEAR in a short coefficient vector of the cell:
//   A_ij = sum_kappa c_kappa T[kappa][i][j],
//   kappa = (k, a):  c = |detJ| Jinv[a][0] u_k        T = int (d_a phi_k phi_j + phi_k d_a phi_j) phi_i      (u_x phi_j + u psi_j) phi_i
//   then             c = |detJ| shift                  T = int phi_j phi_i                                   shift * mass
//                    c = |detJ| Jinv[a][0]             T = int d_a phi_j phi_i                               psi_j phi_i
//                    c = |detJ| shift Jinv[a][0] Jinv[b][0] (a <= b)   T = int d_a phi_j d_b phi_i (+ a<->b) shift psi_j psi_i
// with psi = d/dx_0.  The reference tensor T (K x nd^2, K = nd*dim + 1 + dim + dim(dim+1)/2: 70 x 400 for P3
// tetrahedra) is built once from the engine's own B / w tables and lives in shared memory (row stride = 4 mod 8
// doubles: conflict-free fragment loads); the element matrices of 16 cells per warp are then ONE small GEMM
//   A[entry][cell] = T^T[entry][kappa] c[kappa][cell]
// on the fp64 tensor pipe (mma.sync.m8n8k4: M = 8 entries, N = 8 cells, K = 4 kappas), i.e. 2 K nd^2 flops per
// cell (57.6 kflop for P3 tetrahedra) instead of the 2 nq nd^2 x 3 of the quadrature formulation (168 kflop with
// the 70-point rule, 300 kflop with the 125-point rule of round 1), and no per-cell tabulation work at all.
// The C fragment of a lane holds one entry for two cells; it goes out with one REDG per value through the
// cell-major slot map.  Exact for the same reason the quadrature kernels are: all integrands are polynomials.
template <int DIM, int ND>
struct BbmTensorShape {
  static constexpr int NL = 1 + DIM + DIM * (DIM + 1) / 2;   // state-independent rows
  static constexpr int K = ND * DIM + NL;
  static constexpr int KS = (K + 3) / 4;                      // k-steps
  static constexpr int M = ND * ND, MT = (M + 7) / 8, STRIDE = MT * 8 + 4;
};

template <int DIM, int ND>
__global__ void bbm_build_tensor(const double *__restrict__ B, const double *__restrict__ w, int nq, double *__restrict__ T) {
  using S = BbmTensorShape<DIM, ND>;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S::K * S::STRIDE) return;
  const int kap = t / S::STRIDE, m = t - kap * S::STRIDE;
  double v = 0.0;
  if (m < S::M) {
    const int i = m / ND, j = m - i * ND;
    for (int q = 0; q < nq; ++q) {
      const double *b = B + (size_t)q * (DIM + 1) * ND;  // b[al * ND + node]
      const double pi = b[i], pj = b[j];
      double f;
      if (kap < ND * DIM) {
        const int k = kap / DIM, a = kap - k * DIM;
        f = (b[(1 + a) * ND + k] * pj + b[k] * b[(1 + a) * ND + j]) * pi;
      } else {
        int l = kap - ND * DIM;
        if (l == 0) {
          f = pj * pi;
        } else if (l <= DIM) {
          f = b[l * ND + j] * pi;
        } else {
          l -= 1 + DIM;
          int a = 0;
          while (l >= DIM - a) { l -= DIM - a; ++a; }
          const int bb = a + l;
          f = b[(1 + a) * ND + j] * b[(1 + bb) * ND + i];
          if (bb != a) f += b[(1 + bb) * ND + j] * b[(1 + a) * ND + i];
        }
      }
      v = fma(w[q], f, v);
    }
  }
  T[t] = v;
}

constexpr int BT_WARPS = 16;

template <int DIM, int ND>
__global__ void __launch_bounds__(BT_WARPS * 32, 1) bbm_jacobian_tensor(const KArgs a, const double *__restrict__ T,
                                                                        const uint8_t *__restrict__ pos8_cm) {
  using S = BbmTensorShape<DIM, ND>;
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < S::K * S::STRIDE; i += blockDim.x) sm[i] = T[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g4 = lane >> 2, t4 = lane & 3;
  const int64_t ngroups = (a.c1 - a.c0 + 15) / 16;
  for (int64_t grp = (int64_t)blockIdx.x * BT_WARPS + warp; grp < ngroups; grp += (int64_t)gridDim.x * BT_WARPS) {
    const int64_t cbase = a.c0 + grp * 16;
    // ---- coefficient fragments: lane (g4, t4) holds c[kappa = 4 s + t4] of cells cbase + g4 and cbase + 8 + g4
    double bf[2][S::KS];
    uint32_t colbc[2], rowmask[2];
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
      const int64_t c = cbase + 8 * tl + g4;
      const bool live = c < a.c1;
      double adet = 0.0, ca[DIM];
#pragma unroll
      for (int b = 0; b < DIM; ++b) ca[b] = 0.0;
      uint32_t bcm = 0;
      if (live) {
        int32_t cc[DIM + 1];
#pragma unroll
        for (int v = 0; v <= DIM; ++v) cc[v] = a.cell_coord_t[(int64_t)v * a.ncells_total + c];
        Geom<DIM> g;
        compute_geometry<DIM>(a.coords, cc, g);
        adet = g.adet;
#pragma unroll
        for (int b = 0; b < DIM; ++b) ca[b] = g.Jinv[b][0];
        // Dirichlet columns of the cell: the 4 lanes of a cell split its nodes
        for (int j = t4; j < ND; j += 4)
          if (a.isbc[a.cell_node_t[(int64_t)j * a.ncells_total + c]]) bcm |= 1u << j;
      }
      bcm |= __shfl_xor_sync(0xffffffffu, bcm, 1);
      bcm |= __shfl_xor_sync(0xffffffffu, bcm, 2);
      colbc[tl] = bcm;
      rowmask[tl] = live ? a.rowmask[c] : 0u;
#pragma unroll
      for (int s = 0; s < S::KS; ++s) {
        const int kap = 4 * s + t4;
        double v = 0.0;
        if (live && kap < S::K) {
          if (kap < ND * DIM) {
            const int k = kap / DIM, b = kap - k * DIM;
            v = adet * ca[b] * __ldg(a.x + a.cell_node_t[(int64_t)k * a.ncells_total + c]);
          } else {
            int l = kap - ND * DIM;
            if (l == 0) {
              v = adet * a.shift;
            } else if (l <= DIM) {
              v = adet * ca[l - 1];
            } else {
              l -= 1 + DIM;
              int p = 0;
              while (l >= DIM - p) { l -= DIM - p; ++p; }
              v = adet * a.shift * ca[p] * ca[p + l];
            }
          }
        }
        bf[tl][s] = v;
      }
    }
    // ---- scatter bookkeeping: this lane stores entries of cells cbase + 8 tl + 2 t4 + h
    int64_t cell_s[2][2];
    uint32_t cbc_s[2][2], rm_s[2][2];
#pragma unroll
    for (int tl = 0; tl < 2; ++tl)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n = 2 * t4 + h;  // cell inside the tile; its masks sit in the lanes with g4 == n
        cell_s[tl][h] = cbase + 8 * tl + n;
        cbc_s[tl][h] = __shfl_sync(0xffffffffu, colbc[tl], 4 * n);
        rm_s[tl][h] = __shfl_sync(0xffffffffu, rowmask[tl], 4 * n);
      }
    int cur_i = -1;
    int64_t rp[2][2];
    for (int mt = 0; mt < S::MT; ++mt) {
      double C[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
      const double *tp = sm + 8 * mt + g4;
#pragma unroll
      for (int s = 0; s < S::KS; ++s) {
        const int kap = 4 * s + t4;
        const double av = kap < S::K ? tp[(size_t)kap * S::STRIDE] : 0.0;
        dmma_m8n8k4(C[0], av, bf[0][s]);
        dmma_m8n8k4(C[1], av, bf[1][s]);
      }
      const int m = 8 * mt + g4;
      if (m < S::M) {
        const int i = m / ND, j = m - i * ND;
        if (i != cur_i) {  // row pointers of local row i for the lane's four cells (changes every ND entries)
          cur_i = i;
#pragma unroll
          for (int tl = 0; tl < 2; ++tl)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const bool ok = (rm_s[tl][h] >> i) & 1u;
              rp[tl][h] = ok ? a.nrowptr[a.cell_node_t[(int64_t)i * a.ncells_total + cell_s[tl][h]]] : -1;
            }
        }
#pragma unroll
        for (int tl = 0; tl < 2; ++tl)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (rp[tl][h] >= 0 && !((cbc_s[tl][h] >> j) & 1u)) {
              const int pos = pos8_cm[(size_t)cell_s[tl][h] * (ND * ND) + m];
              red_add_f64(a.out + rp[tl][h] + pos, C[tl][h]);
            }
      }
    }
  }
}

template <int DIM, int ND>
int launch_bbm_tensor(fdts_engine *e, const KArgs &a, cudaStream_t s) {
  using S = BbmTensorShape<DIM, ND>;
  const size_t tbytes = sizeof(double) * S::K * S::STRIDE;
  if (tbytes > 227 * 1024) return 3;
  if (!e->bbm_T) {
    if (cudaMalloc(&e->bbm_T, tbytes) != cudaSuccess) return 2;
    bbm_build_tensor<DIM, ND><<<(S::K * S::STRIDE + 255) / 256, 256, 0, s>>>(e->B, e->w, e->cfg.nq, e->bbm_T);
  }
  auto kern = bbm_jacobian_tensor<DIM, ND>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tbytes) != cudaSuccess) return 2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = ((a.c1 - a.c0 + 15) / 16 + BT_WARPS - 1) / BT_WARPS;
  kern<<<(unsigned)(want < sms ? want : sms), BT_WARPS * 32, tbytes, s>>>(a, e->bbm_T, e->pos8_cm);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

}  // namespace

// returns 0 when the DMMA kernel handled the call, 3 when it does not apply (caller falls back)
int fdts_bbm_jacobian_dmma(fdts_engine *e, const KArgs &a, cudaStream_t s) {
  const int dim = e->cfg.dim, nd = e->cfg.nd;
  if (e->cfg.family != FDTS_BBM || e->cfg.ncomp != 1 || !e->opt_dmma) return 3;
  if (!((dim == 3 && nd == 20) || (dim == 2 && nd == 10) || (dim == 3 && nd == 10))) return 3;
  const int64_t nc = e->cfg.num_cells;
  if (a.c1 <= a.c0) return 0;
  if (!e->pos8_cm) {
    if (cudaMalloc(&e->pos8_cm, (size_t)nc * nd * nd) != cudaSuccess) {
      fdts_set_error("DMMA Jacobian: out of device memory for the cell-major slot map");
      return 2;
    }
    pos8_cell_major<<<(unsigned)((nc * nd * nd + 255) / 256), 256, 0, s>>>(e->pos8, nc, nd * nd, e->pos8_cm);
  }
  if (e->opt_dmma >= 2) {  // tensor-contraction formulation (default)
    int rc = 3;
    if (dim == 3 && nd == 20) rc = launch_bbm_tensor<3, 20>(e, a, s);
    if (dim == 3 && nd == 10) rc = launch_bbm_tensor<3, 10>(e, a, s);
    if (dim == 2 && nd == 10) rc = launch_bbm_tensor<2, 10>(e, a, s);
    if (rc == 2) fdts_set_error(std::string("tensor-contraction Jacobian launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    if (rc != 3) return rc;
  }
  const int nqp = (a.nq + 7) & ~7;
  const size_t smem = sizeof(double) * ((size_t)(dim + 1) * nqp * DMMA_STRIDE + nqp + (size_t)DMMA_WARPS * 3 * nqp);
  if (smem > 227 * 1024) return 3;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (a.c1 - a.c0 + DMMA_WARPS - 1) / DMMA_WARPS;
  const unsigned grid = (unsigned)(want < sms ? want : sms);
  auto launch = [&](auto kern) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    kern<<<grid, DMMA_WARPS * 32, smem, s>>>(a, e->pos8_cm);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
  };
  int rc = 3;
  if (dim == 3 && nd == 20) rc = launch(bbm_jacobian_dmma<3, 20>);
  if (dim == 3 && nd == 10) rc = launch(bbm_jacobian_dmma<3, 10>);
  if (dim == 2 && nd == 10) rc = launch(bbm_jacobian_dmma<2, 10>);
  if (rc == 2) fdts_set_error(std::string("DMMA Jacobian launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  return rc;
}
