// kernels_heat.cu -- closed-form kernels for the constant-coefficient heat family
// (examples/heat.py:11; headline configuration C4 of SURVEY.md section 8(d)).
//
// On affine simplices the heat element matrix is a linear combination of
// reference tensors,
//     A = shift*|detJ| * R00 + sum_{a<=b} (|detJ| G_ab) * S_ab ,   G = Jinv Jinv^T,
// so no quadrature loop is needed at run time: the tensors are compile-time
// constants (ref_tables.cuh), fully unrolled into DFMA immediates with their
// zeros dropped.  Two scatter strategies share this element kernel:
//   path 1  one thread per cell, `red.global.add.f64` into Vec / CSR (atomic)
//   path 2  owner-computes gather (deterministic, write-once): plan 1 and the two-phase plan-2 kernel here,
//           the warp-specialised variant and the residual tile kernels in kernels_heat2.cu
#include <vector>

#include "heat_common.cuh"

namespace {

// ------------------------------------------------------------------ path 1: atomic
template <int DIM, int DEG, int VAR>
__global__ void __launch_bounds__(128) heat_jacobian_atomic(const KArgs a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP;
  const int64_t c = a.c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.c1) return;
  int32_t cc[DIM + 1];
#pragma unroll
  for (int v = 0; v <= DIM; ++v) cc[v] = a.cell_coord_t[v * a.ncells_total + c];
  Geom<DIM> g;
  compute_geometry<DIM>(a.coords, cc, g);
  double gm[NP];
  metric<DIM>(g, gm);
  const double m = a.shift * g.adet;
  int64_t rowbase[ND];
  uint32_t colbc = 0, rowok = 0;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const int32_t n = a.cell_node_t[i * a.ncells_total + c];
    const bool bc = a.isbc[n];
    colbc |= (bc ? 1u : 0u) << i;
    const bool ok = n < a.nown && !bc;
    rowok |= (ok ? 1u : 0u) << i;
    rowbase[i] = ok ? a.nrowptr[n] : 0;
  }
  static_for<ND>([&](auto I) {
    constexpr int i = decltype(I)::value;
    static_for<ND - i>([&](auto JJ) {
      constexpr int j = i + decltype(JJ)::value;
      const double v = heat_entry<T, i, j, NP>(m, gm);
      if (((rowok >> i) & 1u) && !((colbc >> j) & 1u))
        red_add_f64(a.out + rowbase[i] + a.pos8_t[(int64_t)(i * ND + j) * a.ncells_total + c], v);
      if (j != i && ((rowok >> j) & 1u) && !((colbc >> i) & 1u))
        red_add_f64(a.out + rowbase[j] + a.pos8_t[(int64_t)(j * ND + i) * a.ncells_total + c], v);
    });
  });
}

#ifndef RES_NT
#define RES_NT 128
#endif
#ifdef RES_MINB  // experiment builds only (firedrake_ts_b200/csrc/build.py --dev with FDTS_DEV_DEFS)
#define RES_BOUNDS __launch_bounds__(RES_NT, RES_MINB)
#else
#define RES_BOUNDS __launch_bounds__(RES_NT)
#endif
template <int DIM, int DEG, int VAR>
__global__ void RES_BOUNDS heat_residual_atomic(const KArgs a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP;
  const int64_t c = a.c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.c1) return;
  int32_t cc[DIM + 1], nodes[ND];
#pragma unroll
  for (int v = 0; v <= DIM; ++v) cc[v] = a.cell_coord_t[v * a.ncells_total + c];
#pragma unroll
  for (int i = 0; i < ND; ++i) nodes[i] = a.cell_node_t[i * a.ncells_total + c];
  const uint32_t rowok = a.rowmask[c];
  Geom<DIM> g;
#ifdef FDTS_DEV_SWITCHES
  if (FDTS_DEV(a, 64)) {  // experiment: no coordinate gathers
    double X[DIM + 1][DIM];
#pragma unroll
    for (int v = 0; v <= DIM; ++v)
#pragma unroll
      for (int k = 0; k < DIM; ++k) X[v][k] = 1e-3 * (double)(cc[v] % (97 + 13 * k)) + (v == k + 1 ? 1.0 : 0.0);
    geometry_from_vertices<DIM>(X, g);
  } else
#endif
  compute_geometry<DIM>(a.coords, cc, g);
  double gm[NP];
  metric<DIM>(g, gm);
  double u[ND], ud[ND];
#ifdef FDTS_DEV_SWITCHES
  if (FDTS_DEV(a, 32)) {  // experiment: no state gathers
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      u[i] = 1e-9 * (double)nodes[i];
      ud[i] = 2e-9 * (double)nodes[i];
    }
  } else
#endif
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    u[i] = __ldg(a.x + nodes[i]);
    ud[i] = __ldg(a.xdot + nodes[i]);
  }
#ifdef FDTS_DEV_SWITCHES
  if (FDTS_DEV(a, 128)) {  // experiment: sum-factorised element vector
    double f[ND];
    heat_element_residual<DIM, DEG, VAR, ND, NP>(gm, g.adet, a.p0, u, ud, f);
#pragma unroll
    for (int i = 0; i < ND; ++i)
      if (((rowok >> i) & 1u) && (!FDTS_DEV(a, 16) || f[i] == 1.2345e300)) red_add_f64(a.out + nodes[i], f[i]);
    return;
  }
#endif
  static_for<ND>([&](auto I) {
    constexpr int i = decltype(I)::value;
    double mass = -a.p0 * T::M0[i];
    static_for<ND>([&](auto J) {
      constexpr int j = decltype(J)::value;
      constexpr double r = T::R00[i][j];
      if constexpr (r != 0.0) mass = fma(r, ud[j], mass);
    });
    double f = mass * g.adet;
    static_for<NP>([&](auto P) {
      constexpr int p = decltype(P)::value;
      double s = 0.0;
      static_for<ND>([&](auto J) {
        constexpr int j = decltype(J)::value;
        constexpr double sv = T::S[p][i][j];
        if constexpr (sv != 0.0) s = fma(sv, u[j], s);
      });
      f = fma(gm[p], s, f);
    });
#ifdef FDTS_DEV_SWITCHES
    if (FDTS_DEV(a, 16) && f != 1.2345e300) return;  // experiment: no atomics (the value stays live)
#endif
    if ((rowok >> i) & 1u) red_add_f64(a.out + nodes[i], f);
  });
}

template <int DIM, int DEG, int VAR>
int launch_atomic(int which, const KArgs &a, cudaStream_t s) {
  const int64_t n = a.c1 - a.c0;
  if (n <= 0) return 0;
  if (which == 0) {
    const int threads = RES_NT;
    heat_residual_atomic<DIM, DEG, VAR><<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(a);
  } else {
    const int threads = 128;
    heat_jacobian_atomic<DIM, DEG, VAR><<<(unsigned)((n + threads - 1) / threads), threads, 0, s>>>(a);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// ------------------------------------------------------------------ path 2: owner-computes gather
struct GArgs {
  const double *coords;
  const int32_t *cell_coord_t, *cell_node_t;
  const double *x, *xdot;
  double *out;
  const uint8_t *isbc;
  const int64_t *nrowptr, *adj_ptr;
  const int64_t *bstart, *bbase;
  const int32_t *bcells, *bcount;
  const uint16_t *sdst, *cidx, *ridx;
  const uint32_t *soff;
  int64_t ncells_total;
  int32_t cmax;
  int32_t off_idx, off_end, off_bar, off_inA, inA_bytes, off_dst, off_soff, off_obuf, zaddr;  // smem layout (bytes)
  int32_t vmax, persist;
  int64_t nblocks;
  const double *bvcoords;
  const uint32_t *bcc;
  const BlockDesc *bdesc;
  double shift, p0;
};

constexpr int split_row(int nd) {
  int half = nd * (nd + 1) / 4, acc = 0, i = 0;
  while (i < nd && acc + (nd - i) <= half) {
    acc += nd - i;
    ++i;
  }
  return i < 1 ? 1 : i;
}

constexpr int GATHER_THREADS = 512;    // residual gather (one CTA per row block)
constexpr int JGATHER_THREADS = 1024;  // persistent Jacobian gather
constexpr int SELL_WINDOW = 2048;      // must match gather_plan.cu

// Persistent, software-pipelined Jacobian gather: one CTA per SM loops over row blocks.
//   phase A  the block's cells' element matrices -> shared memory (symmetric packing, odd stride);
//   phase B  SELL-32-sigma sweep: a warp owns a slice of 32 count-sorted slots and sums their staged
//            contributions column by column (coalesced 16-bit address loads, no divergence), results
//            go through a 2048-slot shared buffer and leave as coalesced 8-byte stores: every CSR
//            value is written exactly once, no atomics, no zero-fill, bitwise reproducible.
// Every input of a block arrives through TMA bulk copies launched one step ahead:
//   bars[1..2]  block-local vertex coordinates + per-cell local vertex ids (phase-A inputs, double
//               buffered, fetched during the previous block's work);
//   bars[0]     contribution addresses, slice offsets, destination offsets (phase-B inputs, fetched
//               while phase A computes);
// so neither phase waits on a dependent global load.
template <int DIM, int DEG, int VAR>
__global__ void __launch_bounds__(JGATHER_THREADS, 1) heat_jacobian_gather_persist(const GArgs a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP, NEP = (ND * (ND + 1) / 2) | 1;
  constexpr int NWARPS = JGATHER_THREADS / 32;
  extern __shared__ __align__(16) unsigned char smraw[];
  double *stage = reinterpret_cast<double *>(smraw);
  uint16_t *sidx = reinterpret_cast<uint16_t *>(smraw + a.off_idx);
  uint16_t *sdst = reinterpret_cast<uint16_t *>(smraw + a.off_dst);
  uint32_t *soff = reinterpret_cast<uint32_t *>(smraw + a.off_soff);
  double *obuf = reinterpret_cast<double *>(smraw + a.off_obuf);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smraw + a.off_bar);
  unsigned char *inA0 = smraw + a.off_inA;
  const uint32_t vbytes = (uint32_t)a.vmax * DIM * 8u;  // multiple of 16
  const uint32_t cbytes = (uint32_t)((a.cmax * 4 + 15) & ~15);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  auto issue_inA = [&](int64_t b, int slot) {
    unsigned char *dst = inA0 + slot * a.inA_bytes;
    mbar_expect_tx(&bars[1 + slot], vbytes + cbytes);
    bulk_g2s(dst, a.bvcoords + b * a.vmax * DIM, vbytes, &bars[1 + slot]);
    bulk_g2s(dst + vbytes, a.bcc + b * (int64_t)((a.cmax + 3) & ~3), cbytes, &bars[1 + slot]);
  };

  const int64_t stride = gridDim.x;
  int64_t b = blockIdx.x;
  if (b >= a.nblocks) return;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    issue_inA(b, 0);
    stage[a.zaddr] = 0.0;
  }
  BlockDesc d = a.bdesc[b];
  BlockDesc dn = a.bdesc[b + stride < a.nblocks ? b + stride : b];
  __syncthreads();
  for (int it = 0; b < a.nblocks; ++it, b += stride) {
    const int slot = it & 1;
    const int64_t bn = b + stride;
    const int64_t s0a = d.s0 & ~7ll;
    if (threadIdx.x == 0) {
      const uint32_t nb_idx = (uint32_t)d.ncontrib * 2u;
      const uint32_t nb_dst = (uint32_t)(((d.s0 + d.nslots + 7) & ~7ll) - s0a) * 2u;
      const uint32_t nb_off = (uint32_t)d.nslpad * 4u;
      mbar_expect_tx(&bars[0], nb_idx + nb_dst + nb_off);
      if (nb_idx) bulk_g2s(sidx, a.cidx + d.cbase, nb_idx, &bars[0]);
      if (nb_dst) bulk_g2s(sdst, a.sdst + s0a, nb_dst, &bars[0]);
      if (nb_off) bulk_g2s(soff, a.soff + d.slbase, nb_off, &bars[0]);
      if (bn < a.nblocks) issue_inA(bn, slot ^ 1);
    }
    const int64_t bnn = bn + stride;
    const BlockDesc dnn = a.bdesc[bnn < a.nblocks ? bnn : b];  // consumed two iterations from now
    // ---- phase A
    mbar_wait(&bars[1 + slot], (uint32_t)((it >> 1) & 1));
    const double *vtab = reinterpret_cast<const double *>(inA0 + slot * a.inA_bytes);
    const uint32_t *lcc = reinterpret_cast<const uint32_t *>(inA0 + slot * a.inA_bytes + vbytes);
    for (int k = threadIdx.x; k < d.ncell; k += JGATHER_THREADS) {
      const uint32_t packed = lcc[k];
      double X[DIM + 1][DIM];
#pragma unroll
      for (int v = 0; v <= DIM; ++v) {
        const int lv = (packed >> (8 * v)) & 255;
#pragma unroll
        for (int q = 0; q < DIM; ++q) X[v][q] = vtab[lv * DIM + q];
      }
      Geom<DIM> g;
      geometry_from_vertices<DIM>(X, g);
      double gm[NP];
      metric<DIM>(g, gm);
      const double m = a.shift * g.adet;
      double *st = stage + k * NEP;
      static_for<ND>([&](auto I) {
        constexpr int i = decltype(I)::value;
        static_for<ND - i>([&](auto JJ) {
          constexpr int j = i + decltype(JJ)::value;
          constexpr int e = i * ND - (i * (i - 1)) / 2 + (j - i);
          st[e] = heat_entry<T, i, j, NP>(m, gm);
        });
      });
    }
    __syncthreads();
    // ---- phase B
    mbar_wait(&bars[0], (uint32_t)(it & 1));
    const uint16_t *dstp = sdst + (d.s0 - s0a);
    int slice0 = 0;
    for (int wb = 0; wb < d.nslots; wb += SELL_WINDOW) {
      const int nw = d.nslots - wb < SELL_WINDOW ? d.nslots - wb : SELL_WINDOW;
      const int nsl = (nw + 31) >> 5;
      // slices are sorted by width: deal them to the warps in snake order to balance the load
      for (int r0 = 0, rnd = 0; r0 < nsl; r0 += NWARPS, ++rnd) {
        const int sl = r0 + ((rnd & 1) ? NWARPS - 1 - warp : warp);
        if (sl >= nsl) continue;
        const uint32_t off = soff[slice0 + sl];
        const int wdt = (int)((soff[slice0 + sl + 1] - off) >> 5);
        const uint16_t *ip = sidx + off + lane;
        double acc = 0.0;
#pragma unroll 4
        for (int k = 0; k < wdt; ++k) acc += stage[ip[k * 32]];
        const int q = sl * 32 + lane;
        if (q < nw) obuf[dstp[wb + q]] = acc;
      }
      __syncthreads();
      double *o = a.out + d.s0 + wb;
      for (int t = threadIdx.x; t < nw; t += JGATHER_THREADS) o[t] = obuf[t];
      __syncthreads();
      slice0 += nsl;
    }
    d = dn;
    dn = dnn;
  }
}

template <int DIM, int DEG, int VAR>
__global__ void __launch_bounds__(GATHER_THREADS) heat_residual_gather(const GArgs a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP, NDP = ND | 1;
  extern __shared__ __align__(16) unsigned char smraw[];
  double *stage = reinterpret_cast<double *>(smraw);
  uint16_t *sidx = reinterpret_cast<uint16_t *>(smraw + a.off_idx);
  int64_t *sptr = reinterpret_cast<int64_t *>(smraw + a.off_end);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smraw + a.off_bar);
  const int64_t b = blockIdx.x;
  const int64_t r0 = a.bstart[b], r1 = a.bstart[b + 1];
  const int64_t r0a = r0 & ~1ll;
  const int64_t a0 = a.adj_ptr[r0], a1 = a.adj_ptr[r1];
  const int64_t a0a = a0 & ~7ll;
  if (threadIdx.x == 0) {
    const uint32_t nb_idx = (uint32_t)(((a1 + 7) & ~7ll) - a0a) * 2u;
    const uint32_t nb_ptr = (uint32_t)(((r1 + 2) & ~1ll) - r0a) * 8u;
    mbar_init(bar, 1);
    mbar_expect_tx(bar, nb_idx + nb_ptr);
    if (nb_idx) bulk_g2s(sidx, a.ridx + a0a, nb_idx, bar);
    bulk_g2s(sptr, a.adj_ptr + r0a, nb_ptr, bar);
  }
  const int ncell = a.bcount[b];
  const int32_t *cells = a.bcells + b * a.cmax;
  for (int k = threadIdx.x; k < ncell; k += GATHER_THREADS) {
    const int64_t c = cells[k];
    int32_t cc[DIM + 1], nodes[ND];
#pragma unroll
    for (int v = 0; v <= DIM; ++v) cc[v] = a.cell_coord_t[v * a.ncells_total + c];
#pragma unroll
    for (int i = 0; i < ND; ++i) nodes[i] = a.cell_node_t[i * a.ncells_total + c];
    Geom<DIM> g;
    compute_geometry<DIM>(a.coords, cc, g);
    double gm[NP];
    metric<DIM>(g, gm);
    double u[ND], ud[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      u[i] = __ldg(a.x + nodes[i]);
      ud[i] = __ldg(a.xdot + nodes[i]);
    }
    double *st = stage + k * NDP;
    static_for<ND>([&](auto I) {
      constexpr int i = decltype(I)::value;
      double mass = -a.p0 * T::M0[i];
      static_for<ND>([&](auto J) {
        constexpr int j = decltype(J)::value;
        constexpr double r = T::R00[i][j];
        if constexpr (r != 0.0) mass = fma(r, ud[j], mass);
      });
      double f = mass * g.adet;
      static_for<NP>([&](auto P) {
        constexpr int p = decltype(P)::value;
        double sacc = 0.0;
        static_for<ND>([&](auto J) {
          constexpr int j = decltype(J)::value;
          constexpr double sv = T::S[p][i][j];
          if constexpr (sv != 0.0) sacc = fma(sv, u[j], sacc);
        });
        f = fma(gm[p], sacc, f);
      });
      st[i] = f;
    });
  }
  __syncthreads();
  mbar_wait(bar, 0);
  for (int64_t r = r0 + threadIdx.x; r < r1; r += GATHER_THREADS) {
    double acc = 0.0;
    if (!a.isbc[r]) {
      const int64_t kb = sptr[r - r0a] - a0a, ke = sptr[r + 1 - r0a] - a0a;
      for (int64_t k = kb; k < ke; ++k) acc += stage[sidx[k]];
    }
    a.out[r] = acc;
  }
}

template <int DIM, int DEG, int VAR>
int launch_gather(int which, GArgs a, const GatherPlan &p, cudaStream_t s) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND;
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  if (which == 0) {
    size_t off = up16(sizeof(double) * (size_t)p.cmax * (ND | 1));
    a.off_idx = (int)off;
    off += up16(2 * (size_t)(p.max_block_adj + 16));
    a.off_end = (int)off;
    off += up16(8 * (size_t)(p.max_block_rows + 4));
    a.off_bar = (int)off;
    const size_t smem = off + 48;
    if (smem > 227 * 1024) return 4;
    auto k = heat_residual_gather<DIM, DEG, VAR>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)p.nblocks, GATHER_THREADS, smem, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
  }
  size_t off = up16(sizeof(double) * ((size_t)p.cmax * ((ND * (ND + 1) / 2) | 1) + 1));
  a.off_idx = (int)off;
  off += up16(2 * (size_t)(p.max_block_contrib + 32));
  a.off_dst = (int)off;
  off += up16(2 * (size_t)(p.max_block_slots + 16));
  a.off_soff = (int)off;
  off += up16(4 * (size_t)(p.max_block_slices + 4));
  a.off_obuf = (int)off;
  off += sizeof(double) * SELL_WINDOW;
  a.off_bar = (int)off;
  off += 32;
  a.inA_bytes = (int)(up16((size_t)p.vmax * DIM * 8) + up16((size_t)p.cmax * 4));
  a.off_inA = (int)off;
  off += 2 * (size_t)a.inA_bytes;
  a.zaddr = p.zaddr;
  const size_t smem = off + 16;
  if (smem > 227 * 1024) return 4;
  auto k = heat_jacobian_gather_persist<DIM, DEG, VAR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t grid = p.nblocks < sms ? p.nblocks : sms;
  k<<<(unsigned)grid, JGATHER_THREADS, smem, s>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// ------------------------------------------------------------------ path 2, plan version 2
// Persistent owner-computes Jacobian gather over the sector-grouped SELL plan (gather_plan.cu, v2).
//   phase A  element matrices of the block's cells -> shared memory (symmetric packing, odd stride);
//   phase B  one flat, barrier-free sweep over the block's slices: lane l of a slice sums the staged
//            contributions of one CSR slot (16-bit addresses, column-major => conflict-free index
//            loads) and stores it itself; the 4 lanes of a lane group own one 32-byte sector, so
//            every store instruction writes full sectors and every CSR value is written exactly once
//            (no atomics, no zero-fill, no transposition buffer, bitwise reproducible).  Slots with
//            more than 8 contributions (vertex diagonals) are summed by 4 lanes + 2 shuffles.
// All inputs of a block arrive through TMA bulk copies issued one phase ahead: phase-B tables
// (addresses, slice table, sector ids) while phase A computes, phase-A tables (vertex coordinates,
// per-cell local vertex ids) of the next block while phase B sweeps.  With the pattern dictionary the
// index tables of a structured mesh are a few hundred KB in total and stay L2 resident.
// CM (option cache_metric): phase A reads the cell metrics |detJ| G_ab, |detJ| of the block, precomputed once per
// plan (block_metrics) and copied in by ONE bulk copy per block into a single buffer (issued right after the
// phase-A barrier of the previous block, so it lands behind phase B), instead of forming the geometry from
// the vertex table: 30 % less fp64 work in phase A and 7 conflict-free loads instead of 13 scattered ones.
template <int DIM>
__global__ void block_metrics(const BlockDesc2 *__restrict__ bdesc, const uint32_t *__restrict__ bcc,
                              const double *__restrict__ bvcoords, int cmaxp, double *__restrict__ out) {
  constexpr int NP = DIM * (DIM + 1) / 2;
  const int64_t b = blockIdx.x;
  const BlockDesc2 d = bdesc[b];
  double *o = out + b * (int64_t)(NP + 1) * cmaxp;
  for (int k = threadIdx.x; k < cmaxp; k += blockDim.x) {
    double gm[NP], adet = 0.0;
#pragma unroll
    for (int p = 0; p < NP; ++p) gm[p] = 0.0;
    if (k < d.ncell) {
      const uint32_t packed = bcc[d.cc_off + k];
      double X[DIM + 1][DIM];
#pragma unroll
      for (int v = 0; v <= DIM; ++v) {
        const int lv = (packed >> (8 * v)) & 255;
#pragma unroll
        for (int q = 0; q < DIM; ++q) X[v][q] = bvcoords[d.vc_off + lv * DIM + q];
      }
      Geom<DIM> g;
      geometry_from_vertices<DIM>(X, g);
      metric<DIM>(g, gm);
      adet = g.adet;
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) o[p * cmaxp + k] = gm[p];
    o[NP * cmaxp + k] = adet;
  }
}

template <int DIM, int DEG, int VAR, int NT, int G, bool CM = false>
__global__ void __launch_bounds__(NT, NT <= 256 ? 2 : 1) heat_jacobian_gather2(const G2Args a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP, NEP = (ND * (ND + 1) / 2) | 1;
  constexpr int NWARPS = NT / 32;
  constexpr int PER_NORM = 32 / G;
  extern __shared__ __align__(16) unsigned char smraw[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smraw + a.off_bar);
  BlockDesc2 *sdesc = reinterpret_cast<BlockDesc2 *>(smraw + a.off_desc);  // ring of 4 blocks
  const uint32_t sbase = smem_u32(smraw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // bulk copies and descriptor fetches are issued by lane 0 of the LAST warp: phase A fills the warps from
  // the front (one cell per thread), so this warp is normally idle there and the issue latency of the
  // copies (~1000 cycles measured) stays off the critical path
  const bool producer = threadIdx.x == NT - 32;
  const uint32_t idx0 = sbase + a.off_idx + lane * 4;
  const uint32_t gd0 = sbase + a.off_gdst;
  const uint32_t cls0 = sbase + a.off_meta;

  // phase-A inputs (vertex table + local vertex ids), double buffered: issued one block ahead
  auto issue_M = [&](int64_t blk) {  // CM: all (NP + 1) metric rows of block blk, single buffer
    const uint32_t mb = (uint32_t)((NP + 1) * a.cmaxp * 8);
    mbar_expect_tx(&bars[2], mb);
    bulk_g2s(smraw + a.off_met, a.bmetric + blk * (int64_t)(NP + 1) * a.cmaxp, mb, &bars[2]);
  };
  auto issue_A = [&](const BlockDesc2 &d, int buf) {
    const uint32_t vb = (uint32_t)(((d.nvert * DIM + 1) & ~1) * 8);
    const uint32_t cb = (uint32_t)(((d.ncell + 3) & ~3) * 4);
    mbar_expect_tx(&bars[2 + buf], vb + cb);
    if (vb) bulk_g2s(smraw + a.off_vtab + buf * a.inA_bytes, a.bvcoords + d.vc_off, vb, &bars[2 + buf]);
    if (cb) bulk_g2s(smraw + a.off_lcc + buf * a.inA_bytes, a.bcc + d.cc_off, cb, &bars[2 + buf]);
  };
  auto issue_B = [&](const BlockDesc2 &d) {  // SELL addresses, class table, sector ids
    const uint32_t nslp = (uint32_t)((d.nsl + 3) & ~3);
    const uint32_t ib = (uint32_t)d.nidx * 2u, mb = (uint32_t)((d.ncls + 1) & ~1) * 8u, gb = nslp * 2u * (uint32_t)PER_NORM;
    mbar_expect_tx(&bars[1], ib + mb + gb);
    if (ib) bulk_g2s(smraw + a.off_idx, a.cidx + d.idx_off, ib, &bars[1]);
    if (mb) bulk_g2s(smraw + a.off_meta, a.cmeta + d.cls_off, mb, &bars[1]);
    if (gb) bulk_g2s(smraw + a.off_gdst, a.gdst + d.meta_off * PER_NORM, gb, &bars[1]);
  };

  const int64_t stride = gridDim.x;
  int64_t b = a.block_first + blockIdx.x;
  if (b >= a.nblocks) return;
  // descriptors are fetched two blocks ahead (ring of four in shared memory), phase-A inputs one block
  // ahead, so no dependent global access sits on a block's critical path
  if (producer) {
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    sdesc[0] = a.bdesc[b];
    if (b + stride < a.nblocks) sdesc[1] = a.bdesc[b + stride];
    if constexpr (CM) issue_M(b); else issue_A(sdesc[0], 0);
  }
  if (threadIdx.x < 16) sts_f64(sbase + (a.zaddr + threadIdx.x) * 8, 0.0);  // one staged zero per 8-byte bank
  __syncthreads();
  // Pattern residency: consecutive blocks of a structured mesh mostly share one pattern, so the index
  // tables already in shared memory are reused and only reloaded when the pattern id changes.
  int resident = -1;
  uint32_t nloadB = 0;
  for (int it = 0; b < a.nblocks; ++it, b += stride) {
    const int64_t bn = b + stride, bnn = bn + stride;
    const BlockDesc2 &d = sdesc[it & 3];
    const int pattern = d.pattern;
    const bool loadB = pattern != resident;
    BlockDesc2 dnn;
    if (producer) {
      if (loadB) issue_B(d);
      if constexpr (!CM)
        if (bn < a.nblocks) issue_A(sdesc[(it + 1) & 3], (it + 1) & 1);
      if (bnn < a.nblocks) dnn = a.bdesc[bnn];
    }
    if (loadB) {
      resident = pattern;
      ++nloadB;
    }
    // ---- phase A
    if constexpr (CM) mbar_wait(&bars[2], (uint32_t)(it & 1));
    else mbar_wait(&bars[2 + (it & 1)], (uint32_t)((it >> 1) & 1));
    const int ncell = FDTS_DEV(a, 1) ? 0 : d.ncell;
    const uint32_t lcc0 = sbase + a.off_lcc + (it & 1) * a.inA_bytes, vt0 = sbase + a.off_vtab + (it & 1) * a.inA_bytes;
    const uint32_t met0 = sbase + a.off_met;
    for (int k = threadIdx.x; k < ncell; k += NT) {
      double gm[NP], adet;
      if constexpr (CM) {
#pragma unroll
        for (int p = 0; p < NP; ++p) gm[p] = lds_f64(met0 + (uint32_t)(p * a.cmaxp + k) * 8u);
        adet = lds_f64(met0 + (uint32_t)(NP * a.cmaxp + k) * 8u);
      } else {
        const uint32_t packed = lds_u32(lcc0 + k * 4);
        double X[DIM + 1][DIM];
#pragma unroll
        for (int v = 0; v <= DIM; ++v) {
          const uint32_t va = vt0 + ((packed >> (8 * v)) & 255u) * (DIM * 8);
#pragma unroll
          for (int q = 0; q < DIM; ++q) X[v][q] = lds_f64(va + q * 8);
        }
        Geom<DIM> g;
        geometry_from_vertices<DIM>(X, g);
        metric<DIM>(g, gm);
        adet = g.adet;
      }
      const uint32_t st = sbase + k * (NEP * 8);
      if constexpr (T::INT) {
        const double ms = a.shift * adet * (1.0 / T::RDEN);
#pragma unroll
        for (int p = 0; p < NP; ++p) gm[p] *= (1.0 / T::SDEN);
        static_for<ND>([&](auto I) {
          constexpr int i = decltype(I)::value;
          static_for<ND - i>([&](auto JJ) {
            constexpr int j = i + decltype(JJ)::value;
            constexpr int e = i * ND - (i * (i - 1)) / 2 + (j - i);
            sts_f64(st + e * 8, heat_entry_i<T, i, j, NP>(ms, gm));
          });
        });
      } else {
        const double m = a.shift * adet;
        static_for<ND>([&](auto I) {
          constexpr int i = decltype(I)::value;
          static_for<ND - i>([&](auto JJ) {
            constexpr int j = i + decltype(JJ)::value;
            constexpr int e = i * ND - (i * (i - 1)) / 2 + (j - i);
            sts_f64(st + e * 8, heat_entry<T, i, j, NP>(m, gm));
          });
        });
      }
    }
    if (producer && bnn < a.nblocks) sdesc[(it + 2) & 3] = dnn;
    __syncthreads();
    if constexpr (CM)  // every thread is done with the metric buffer: fetch the next block's behind phase B
      if (producer && bn < a.nblocks) issue_M(bn);
    // ---- phase B: class by class, no barrier in between
    if (loadB) mbar_wait(&bars[1], (nloadB - 1u) & 1u);
    const uint32_t lo = (uint32_t)(d.s0 & (G - 1)), span = FDTS_DEV(a, 8) ? 0u : (uint32_t)d.nslots;
    double *__restrict__ out = a.out + (d.s0 - lo);
    const int ncls = FDTS_DEV(a, 2) ? 0 : d.ncls;
    for (int c = 0; c < ncls; ++c) {
      uint32_t c0, c1;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(c0), "=r"(c1) : "r"(cls0 + c * 8));
      const uint32_t first = c0 & 0xffffu, count = c0 >> 16, off64 = c1 & 0xffffu, w = (c1 >> 16) & 0xffu;
      const bool split = c1 >= 0x1000000u;
#define SWEEP(W) sweep_class<W, G, NWARPS>(first, count, off64, split, warp, lane, idx0, gd0, sbase, out, lo, span)
      switch (w) {
        case 0: {  // slots without contributions (Dirichlet rows/columns): write zeros
          constexpr int LG = G == 4 ? 2 : (G == 2 ? 1 : 0);
          for (uint32_t i = first_slice_of<NWARPS>(warp, first); i < count; i += NWARPS) {
            const uint32_t gs = lds_u16(gd0 + ((first + i) * PER_NORM + (split ? (lane >> (LG + 2)) : (lane >> LG))) * 2u);
            const uint32_t slot = gs * G + (split ? ((lane >> 2) & (G - 1)) : (lane & (G - 1)));
            if ((!split || (lane & 3) == 0) && slot - lo < span) out[slot] = 0.0;
          }
          break;
        }
        case 1: SWEEP(1); break;
        case 2: SWEEP(2); break;
        case 3: SWEEP(3); break;
        case 4: SWEEP(4); break;
        case 5: SWEEP(5); break;
        case 6: SWEEP(6); break;
        case 7: SWEEP(7); break;
        case 8: SWEEP(8); break;
        default: {  // wider classes (high-valence vertices of unstructured meshes): generic loop
          constexpr int LG = G == 4 ? 2 : (G == 2 ? 1 : 0);
          const uint32_t pairs = (w + 1) >> 1;
          for (uint32_t i = first_slice_of<NWARPS>(warp, first); i < count; i += NWARPS) {
            const uint32_t ip = idx0 + (off64 + i * pairs) * 128u;
            const uint32_t gs = lds_u16(gd0 + ((first + i) * PER_NORM + (split ? (lane >> (LG + 2)) : (lane >> LG))) * 2u);
            double a0 = 0.0, a1 = 0.0;
            for (uint32_t q = 0; q < pairs; ++q) {
              const uint32_t pr = lds_u32(ip + q * 128);
              a0 += lds_f64(sbase + (pr & 0xffffu) * 8u);
              if (2 * q + 1 < w) a1 += lds_f64(sbase + (pr >> 16) * 8u);
            }
            double acc = a0 + a1;
            if (split) {
              acc += __shfl_xor_sync(0xffffffffu, acc, 1);
              acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            }
            const uint32_t slot = gs * G + (split ? ((lane >> 2) & (G - 1)) : (lane & (G - 1)));
            if ((!split || (lane & 3) == 0) && slot - lo < span) out[slot] = acc;
          }
        }
      }
#undef SWEEP
    }
    __syncthreads();
  }
}

template <int DIM, int DEG, int VAR, int NT, bool CM = false>
int launch_gather2_nt(G2Args a, const GatherPlan2 &p, cudaStream_t s) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NEP = (ND * (ND + 1) / 2) | 1;
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  if (p.nep != NEP) return 3;
  size_t off = up16(sizeof(double) * ((size_t)p.zaddr + 16));
  a.off_idx = (int)off;
  off += up16(2 * (size_t)p.max_nidx + 64);
  a.off_meta = (int)off;
  off += up16(8 * (size_t)PLAN2_MAX_CLASSES + 16);
  a.off_gdst = (int)off;
  off += up16(2 * (size_t)(32 / p.sector) * (size_t)p.max_nsl + 16);
  a.sector = p.sector;
  // phase-A inputs: two buffers of (vertex table, local vertex ids)
  a.inA_bytes = (int)(up16(8 * (size_t)p.vmax * DIM + 16) + up16(4 * (size_t)p.cmax + 16));
  a.off_vtab = (int)off;
  a.off_lcc = (int)(off + up16(8 * (size_t)p.vmax * DIM + 16));
  a.off_met = (int)off;
  a.cmaxp = p.cmaxp;
  if (CM)
    off += up16(8 * (size_t)(T::NP + 1) * p.cmaxp);  // one buffer of cached metrics replaces the two vertex-table buffers
  else
    off += 2 * (size_t)a.inA_bytes;
  a.off_bar = (int)off;
  off += 32;
  a.off_desc = (int)off;
  off += 4 * sizeof(BlockDesc2);
  a.zaddr = p.zaddr;
  const size_t smem = off;
  if (smem > 227 * 1024) return 4;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  auto launch = [&](auto kern) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem) != cudaSuccess || occ < 1) occ = 1;
    const int64_t slots = (int64_t)sms * occ, nlaunch = a.nblocks - a.block_first;
    const int64_t grid = nlaunch < slots ? nlaunch : slots;
    if (grid <= 0) return 0;
    kern<<<(unsigned)grid, NT, smem, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
  };
  if (p.sector == 4) return launch(heat_jacobian_gather2<DIM, DEG, VAR, NT, 4, CM>);
  if (p.sector == 2) return launch(heat_jacobian_gather2<DIM, DEG, VAR, NT, 2, CM>);
  return launch(heat_jacobian_gather2<DIM, DEG, VAR, NT, 1, CM>);
}

template <int DIM, int DEG, int VAR>
int launch_gather2(G2Args a, const GatherPlan2 &p, int threads, cudaStream_t s) {
  if (threads == 512) {
    if (a.bmetric) {  // cached metrics: falls back to the vertex-table kernel when the metric buffer does not fit
      const int rc = launch_gather2_nt<DIM, DEG, VAR, 512, true>(a, p, s);
      if (rc != 4) return rc;
    }
    return launch_gather2_nt<DIM, DEG, VAR, 512>(a, p, s);
  }
  return launch_gather2_nt<DIM, DEG, VAR, 1024>(a, p, s);  // (1024 threads with cached metrics: 6.84 ms at C4, not kept)
}

// builds the plan's metric cache on first use (option cache_metric); a failed allocation just leaves it off
template <int DIM>
void ensure_metrics(fdts_engine *e, cudaStream_t s) {
  GatherPlan2 &p = e->plan2;
  if (p.bmetric || !e->opt_cache_metric || p.cmaxp <= 0) return;
  constexpr int NP = DIM * (DIM + 1) / 2;
  const size_t bytes = sizeof(double) * (size_t)p.nblocks * (NP + 1) * p.cmaxp + 64;
  if (cudaMalloc(&p.bmetric, bytes) != cudaSuccess) {
    cudaGetLastError();
    p.bmetric = nullptr;
    e->opt_cache_metric = 0;
    return;
  }
  block_metrics<DIM><<<(unsigned)p.nblocks, 128, 0, s>>>(p.bdesc, p.bcc, p.bvcoords, p.cmaxp, p.bmetric);
  e->launches++;
}

__global__ void set_values_k(double *__restrict__ v, const int64_t *__restrict__ slots, int64_t n, double val) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && slots[t] >= 0) v[slots[t]] = val;
}

}  // namespace

template <int DIM, int DEG, int VAR>
static bool tables_match(const double *R) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND;
  const int d1 = DIM + 1;
  double err = 0.0;
  for (int i = 0; i < ND; ++i)
    for (int j = 0; j < ND; ++j) {
      err = fmax(err, fabs(R[i * ND + j] - T::R00[i][j]));
      int p = 0;
      for (int a = 1; a <= DIM; ++a)
        for (int b = a; b <= DIM; ++b, ++p) {
          const double rab = R[((a * d1 + b) * ND + i) * ND + j], rba = R[((b * d1 + a) * ND + i) * ND + j];
          const double s = a == b ? rab : rab + rba;
          const double st = a == b ? R[((a * d1 + b) * ND + j) * ND + i] : R[((a * d1 + b) * ND + j) * ND + i] + R[((b * d1 + a) * ND + j) * ND + i];
          err = fmax(err, fabs(0.5 * (s + st) - T::S[p][i][j]));
        }
    }
  return err < 1e-13;
}

// returns the RefT variant whose compile-time tensors equal the run-time tensors, or -1
int fdts_heat_match_tables(int dim, int deg, const double *R) {
#define CASE(D, K, V) \
  if (dim == D && deg == K && tables_match<D, K, V>(R)) return V;
  CASE(1, 1, 1) CASE(1, 2, 1) CASE(1, 3, 0) CASE(1, 3, 1)
  CASE(2, 1, 1) CASE(2, 2, 1) CASE(2, 3, 0) CASE(2, 3, 1)
  CASE(3, 1, 1) CASE(3, 2, 1) CASE(3, 3, 0) CASE(3, 3, 1)
#undef CASE
  return -1;
}

// which: 0 residual, 1 Jacobian.  Full call: zero, assemble, boundary fix-up.
// flags: bit0 zero the output first (atomic paths), bit1 apply the Dirichlet diagonal afterwards (Jacobian)
static int launch_heat_impl(fdts_engine *e, int which, const double *x, const double *xdot, double shift, double *out,
                            int64_t c0, int64_t c1, int flags, cudaStream_t s) {
  KArgs a = fdts_make_kargs(e, x, xdot, shift, out, c0, c1);
  const int dim = e->cfg.dim, deg = e->cfg.degree, var = e->variant;
  if (e->opt_scatter == 2 && var < 0) {
    fdts_set_error("closed-form heat path: run-time reference tensors do not match the compiled tables");
    return 3;
  }
  if (e->opt_scatter == 2 && which == 1 && !e->plan2.ready && !e->plan.ready) {
    fdts_set_error("gather path requested but no gather plan is built (fdts_set_row_blocks)");
    return 3;
  }
  if (e->opt_scatter == 2 && which == 1 && e->plan2.ready) {
    if (dim == 1) ensure_metrics<1>(e, s);
    if (dim == 2) ensure_metrics<2>(e, s);
    if (dim == 3) ensure_metrics<3>(e, s);
    const GatherPlan2 &p = e->plan2;
    G2Args g;
    g.bdesc = p.bdesc; g.cidx = p.cidx; g.gdst = p.gdst; g.cmeta = p.cmeta; g.bcc = p.bcc; g.bvcoords = p.bvcoords;
    g.bmetric = e->opt_cache_metric && !e->opt_overlap ? p.bmetric : nullptr;
    g.out = out; g.nblocks = p.nblocks; g.shift = shift;
    g.block_first = 0;
    if (e->gather_b1 > e->gather_b0 && !e->opt_overlap) {  // block range requested by fdts_ijacobian_host (pipelined read-back)
      g.block_first = e->gather_b0;
      g.nblocks = e->gather_b1 < p.nblocks ? e->gather_b1 : p.nblocks;
    }
#ifdef FDTS_DEV_SWITCHES
    g.dbg = (int32_t)e->opt_debug;
#endif
    int rc = 3;
#define G2CASE(D, K, V)                                                      \
  if (dim == D && deg == K && (K < 3 || var == V)) {                         \
    rc = e->opt_overlap ? fdts_heat_gather3(D, K, V, g, p, s) : 4;           \
    if (rc == 4) rc = launch_gather2<D, K, V>(g, p, (int)e->opt_threads, s); \
  }
    G2CASE(1, 1, 1) G2CASE(1, 2, 1) G2CASE(1, 3, 0) G2CASE(1, 3, 1)
    G2CASE(2, 1, 1) G2CASE(2, 2, 1) G2CASE(2, 3, 0) G2CASE(2, 3, 1)
    G2CASE(3, 1, 1) G2CASE(3, 2, 1) G2CASE(3, 3, 0) G2CASE(3, 3, 1)
#undef G2CASE
    if (rc == 4) fdts_set_error("gather path: a row block needs more shared memory than one SM has (use smaller row blocks)");
    if (rc == 3) fdts_set_error("gather path: unsupported dimension/degree");
    if (rc == 2) fdts_set_error(std::string("gather kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    if (rc) return rc;
    e->launches++;
    if (e->cfg.num_bc_nodes > 0) {
      const int64_t n = e->cfg.num_bc_nodes;
      set_values_k<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, e->bc_diag_slot, n, 1.0);
      e->launches++;
    }
    FDTS_CHECK(cudaGetLastError());
    return 0;
  }
  if (e->opt_scatter == 2 && (which == 1 || e->opt_gather_residual) && e->plan.ready) {
    if (!e->plan.ready) {
      fdts_set_error("gather path requested but no gather plan is built (fdts_set_row_blocks)");
      return 3;
    }
    const GatherPlan &p = e->plan;
    GArgs g;
    g.coords = e->coords; g.cell_coord_t = e->cell_coord; g.cell_node_t = e->cell_node;
    g.x = x; g.xdot = xdot; g.out = out; g.isbc = e->isbc;
    g.nrowptr = e->nrowptr; g.adj_ptr = e->adj_ptr; g.bstart = p.bstart; g.bbase = p.bbase;
    g.bcells = p.bcells; g.bcount = p.bcount; g.sdst = p.sdst; g.soff = p.soff; g.cidx = p.cidx; g.ridx = p.ridx;
    g.ncells_total = e->cfg.num_cells; g.cmax = p.cmax; g.shift = shift; g.p0 = e->cfg.params[0];
    g.vmax = p.vmax; g.nblocks = p.nblocks; g.bvcoords = p.bvcoords; g.bcc = p.bcc; g.bdesc = p.bdesc;
    g.persist = 1; g.off_inA = 0; g.inA_bytes = 0; g.off_dst = g.off_soff = g.off_obuf = 0; g.zaddr = p.zaddr;
    int rc = 3;
#define GCASE(D, K, V) \
  if (dim == D && deg == K && (K < 3 || var == V)) rc = launch_gather<D, K, V>(which, g, p, s);
    GCASE(1, 1, 1) GCASE(1, 2, 1) GCASE(1, 3, 0) GCASE(1, 3, 1)
    GCASE(2, 1, 1) GCASE(2, 2, 1) GCASE(2, 3, 0) GCASE(2, 3, 1)
    GCASE(3, 1, 1) GCASE(3, 2, 1) GCASE(3, 3, 0) GCASE(3, 3, 1)
#undef GCASE
    if (rc == 4) fdts_set_error("gather path: a row block needs more shared memory than one SM has");
    if (rc == 3) fdts_set_error("gather path: unsupported dimension/degree");
    if (rc == 2) fdts_set_error(std::string("gather kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
    if (rc) return rc;
    e->launches++;
    if (which == 1 && e->cfg.num_bc_nodes > 0) {
      const int64_t n = e->cfg.num_bc_nodes;
      set_values_k<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, e->bc_diag_slot, n, 1.0);
      e->launches++;
    }
    FDTS_CHECK(cudaGetLastError());
    return 0;
  }
  if (var < 0) {
    fdts_set_error("closed-form heat path: run-time reference tensors do not match the compiled tables");
    return 3;
  }
  if (which == 0 && e->rplan.ready && e->opt_residual_tiles) {
    // cell range -> tile range (only when both ends are tile boundaries; otherwise the per-cell kernel)
    const ResidualPlan &p = e->rplan;
    const int64_t *ts = p.tstart_host;
    const int64_t *lo = std::lower_bound(ts, ts + p.ntiles + 1, c0), *hi = std::lower_bound(ts, ts + p.ntiles + 1, c1);
    if (lo != ts + p.ntiles + 1 && hi != ts + p.ntiles + 1 && *lo == c0 && *hi == c1) {
      if (flags & 1) FDTS_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * e->cfg.num_owned_nodes, s));
      RTArgs r;
      r.tdesc = p.tdesc; r.cidx = p.cidx; r.gdst = p.gdst; r.tlid = p.tlid; r.cmeta = p.cmeta; r.bcc = p.bcc;
      r.tnode = p.tnode; r.bvcoords = p.bvcoords; r.x = x; r.xdot = xdot; r.out = out; r.p0 = e->cfg.params[0];
      r.cmaxp = p.cmaxp; r.umaxp = p.umaxp; r.ndp = p.ndp; r.zaddr = p.zaddr;
      r.tstart = p.tstart_dev; r.cell_node_t = e->cell_node; r.cell_coord_t = e->cell_coord; r.coords = e->coords;
      r.nc = e->cfg.num_cells; r.direct = e->opt_residual_tiles >= 2 ? (int)e->opt_residual_tiles - 1 : 0;
      int rc = 3;
      rc = fdts_heat_residual_tiles(dim, deg, var, r, p, lo - ts, hi - ts, s);
      if (rc == 2) fdts_set_error(std::string("residual tile kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
      if (rc == 3 || rc == 4) fdts_set_error("residual tile kernel: unsupported element or tile too large for shared memory");
      if (rc) return rc;
      e->launches++;
      return 0;
    }
  }
  if (flags & 1) {
    if (which == 0)
      FDTS_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * e->cfg.num_owned_nodes, s));
    else
      FDTS_CHECK(cudaMemsetAsync(out, 0, sizeof(double) * e->node_nnz, s));
  }
  int rc = 3;
#define CASE(D, K, V) \
  if (dim == D && deg == K && (K < 3 || var == V)) rc = launch_atomic<D, K, V>(which, a, s);
  CASE(1, 1, 1) CASE(1, 2, 1) CASE(1, 3, 0) CASE(1, 3, 1)
  CASE(2, 1, 1) CASE(2, 2, 1) CASE(2, 3, 0) CASE(2, 3, 1)
  CASE(3, 1, 1) CASE(3, 2, 1) CASE(3, 3, 0) CASE(3, 3, 1)
#undef CASE
  if (rc == 3) fdts_set_error("closed-form heat kernel: unsupported dimension/degree");
  if (rc == 2) fdts_set_error(std::string("heat kernel launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  if (rc) return rc;
  e->launches++;
  if (which == 1 && (flags & 2) && e->cfg.num_bc_nodes > 0) {
    const int64_t n = e->cfg.num_bc_nodes;
    set_values_k<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, e->bc_diag_slot, n, 1.0);
    e->launches++;
  }
  FDTS_CHECK(cudaGetLastError());
  return 0;
}

namespace {
__global__ void scale_copy_k(const double *__restrict__ in, double *__restrict__ out, int64_t n, double c) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = c * in[i];
}
__global__ void scale_inplace_k(double *__restrict__ v, int64_t n, double c) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= c;
}
}  // namespace

// Heat family: straight through.  Diffusion family  a (u_t, v) + k (grad u, grad v) - s (1, v)
// (params = (s, a, k); the two sides of the reference's IMEX split, examples/heat-explicit.py:13-14) on the
// same closed-form kernels by linearity:
//   residual:  F(u, udot) = F_heat(k u, a udot; source s)        (two scaled copies of the state)
//   Jacobian:  shift a M + k K = k (shift a / k  M + K)          (k = 1: nothing extra; else one scaling pass
//                                                                 over the values, Dirichlet diagonal re-set)
// Both are exact for the coefficients 0 and +-1 of the split.
int fdts_launch_heat(fdts_engine *e, int which, const double *x, const double *xdot, double shift, double *out,
                     int64_t c0, int64_t c1, int flags, cudaStream_t s) {
  if (e->cfg.ncomp != 1 || (e->cfg.family != FDTS_HEAT && e->cfg.family != FDTS_DIFFUSION)) {
    fdts_set_error("closed-form path: heat / diffusion families only");
    return 3;
  }
  if (e->cfg.family == FDTS_HEAT) return launch_heat_impl(e, which, x, xdot, shift, out, c0, c1, flags, s);
  const double a = e->cfg.params[1], k = e->cfg.params[2];
  const int64_t nloc = e->cfg.num_nodes;
  if (which == 0) {
    if (a != 1.0 || k != 1.0) {
      if (!e->scr_x) {
        FDTS_CHECK(cudaMalloc(&e->scr_x, sizeof(double) * (nloc + 1)));
        FDTS_CHECK(cudaMalloc(&e->scr_xdot, sizeof(double) * (nloc + 1)));
      }
      const unsigned grid = (unsigned)((nloc + 255) / 256);
      if (k != 1.0) {
        scale_copy_k<<<grid, 256, 0, s>>>(x, e->scr_x, nloc, k);
        x = e->scr_x;
        e->launches++;
      }
      if (a != 1.0) {
        scale_copy_k<<<grid, 256, 0, s>>>(xdot, e->scr_xdot, nloc, a);
        xdot = e->scr_xdot;
        e->launches++;
      }
    }
    return launch_heat_impl(e, 0, x, xdot, 0.0, out, c0, c1, flags, s);
  }
  if (k == 0.0 || (k != 1.0 && !(c0 == 0 && c1 == e->cfg.num_cells && (flags & 1)))) {
    fdts_set_error("closed-form diffusion Jacobian: k = 0 or a partial range with k != 1 belongs to the quadrature kernels");
    return 3;
  }
  int rc = launch_heat_impl(e, 1, x, xdot, shift * a / k, out, c0, c1, flags, s);
  if (rc || k == 1.0) return rc;
  scale_inplace_k<<<(unsigned)((e->node_nnz + 255) / 256), 256, 0, s>>>(out, e->node_nnz, k);
  e->launches++;
  if (e->cfg.num_bc_nodes > 0 && ((flags & 2) || e->opt_scatter == 2)) {
    const int64_t n = e->cfg.num_bc_nodes;
    set_values_k<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, e->bc_diag_slot, n, 1.0);
    e->launches++;
  }
  FDTS_CHECK(cudaGetLastError());
  return 0;
}
