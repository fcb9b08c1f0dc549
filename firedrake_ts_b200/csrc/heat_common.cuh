// heat_common.cuh -- device helpers shared by the closed-form heat kernels (kernels_heat.cu: atomic paths,
// plan-1 gather, two-phase plan-2 gather; kernels_heat2.cu: warp-specialised gather, residual tile kernels).
#pragma once
#include <algorithm>
#include <type_traits>
#include <utility>

#include "kargs.cuh"
#include "ref_tables.cuh"

struct G2Args {
  const BlockDesc2 *bdesc;
  const uint16_t *cidx, *gdst;
  const uint2 *cmeta;
  const uint32_t *bcc;
  const double *bvcoords;
  const double *bmetric;  // cached cell metrics (cache_metric): block b at b * (NP + 1) * cmaxp
  double *out;
  int64_t nblocks;       // end of the block range of this launch (exclusive)
  int64_t block_first;   // first block of this launch (0 unless the caller pipelines the output in chunks)
  double shift;
  int32_t cmaxp, off_met;
  int32_t off_idx, off_meta, off_gdst, off_vtab, off_lcc, off_bar, off_desc, zaddr, sector, inA_bytes, stage_bytes;
#ifdef FDTS_DEV_SWITCHES  // timing experiments only (phases switched off, results wrong); never compiled into the shipped library
  int32_t dbg;
#endif
};


// FDTS_DEV(a, bit): development switch `bit` of the launch arguments; constant false in the shipped build
#ifdef FDTS_DEV_SWITCHES
#define FDTS_DEV(a, bit) (((a).dbg & (bit)) != 0)
#else
#define FDTS_DEV(a, bit) false
#endif

#ifndef RT_MINB
#define RT_MINB 4
#endif
struct RTArgs {
  const BlockDesc2 *tdesc;
  const uint16_t *cidx, *gdst, *tlid;
  const uint2 *cmeta;
  const uint32_t *bcc;
  const int32_t *tnode;
  const double *bvcoords, *x, *xdot;
  double *out;
  const int64_t *tstart;        // direct variant: first cell of every tile, maps, coordinates
  const int32_t *cell_node_t, *cell_coord_t;
  const double *coords;
  int64_t nc;
  int64_t t0;  // first tile of this launch
  int32_t direct;
  double p0;
  int32_t cmaxp, umaxp, ndp, zaddr, off_su, off_sud, off_vtab;
  int32_t off_sidx, off_sgd, off_scls, off_stn;  // variant 3: the tile's reduction tables in shared memory
};


namespace {

template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
  static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// g[p] = |detJ| * (Jinv Jinv^T)_{ab} for the pairs a<=b in the order of RefT::S
template <int DIM>
__device__ __forceinline__ void metric(const Geom<DIM> &g, double (&gm)[DIM * (DIM + 1) / 2]) {
  int p = 0;
#pragma unroll
  for (int a = 0; a < DIM; ++a)
#pragma unroll
    for (int b = a; b < DIM; ++b) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) t += g.Jinv[a][k] * g.Jinv[b][k];
      gm[p++] = t * g.adet;
    }
}

template <class T, int I, int J, int NP>
__device__ __forceinline__ double heat_entry(double m, const double (&gm)[NP]) {
  double v = 0.0;
  constexpr double r = T::R00[I][J];
  if constexpr (r != 0.0) v = m * r;
  static_for<NP>([&](auto P) {
    constexpr int p = decltype(P)::value;
    constexpr double s = T::S[p][I][J];
    if constexpr (s != 0.0) v = fma(gm[p], s, v);
  });
  return v;
}

// Sum-factorised element residual of the heat form on a P2 tetrahedron (local node order of fem.py:
// vertices 0-3, then edges (2,3) (1,3) (1,2) (0,3) (0,2) (0,1)).  Same result as the table form
//   f_i = adet * sum_j R00_ij ud_j + sum_p gm_p sum_j S_pij u_j - p0 M0_i adet
// in ~230 instead of ~460 fp64 operations and fewer live registers:
//  * grad u is linear on the cell, so it is represented by its values at the 4 vertices
//    (d_a^v = 3 u_a if a == v else 4 u_(a,v) - u_a in barycentric components), mapped through the metric
//    (q = G grad u, 36 FMA) and tested with the (equally linear) basis gradients via the P1 mass matrix
//    int lambda_v lambda_w = (1 + delta_vw) / 120;
//  * the P2 mass matrix 1/2520 [6,1 | -4,-6 | 32,16,8] is applied through vertex / edge sums.
__device__ __forceinline__ void p2tet_heat_residual(const double (&gm)[6], double adet, double p0, const double (&u)[10],
                                                    const double (&ud)[10], double (&f)[10]) {
  constexpr int E[4][4] = {{-1, 9, 8, 7}, {9, -1, 6, 5}, {8, 6, -1, 4}, {7, 5, 4, -1}};  // edge node of (a, b)
  double d[4][4];  // d[a][v]: barycentric derivative a of u at vertex v
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int v = 0; v < 4; ++v) d[a][v] = a == v ? 3.0 * u[a] : fma(4.0, u[E[a][v]], -u[a]);
  const double s120 = 1.0 / 120.0;
  const double G[3][3] = {{gm[0] * s120, gm[1] * s120, gm[2] * s120},
                          {gm[1] * s120, gm[3] * s120, gm[4] * s120},
                          {gm[2] * s120, gm[4] * s120, gm[5] * s120}};
  double Q[3][4], SQ[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double q[4], sq = 0.0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      double t = 0.0;
#pragma unroll
      for (int e = 0; e < 3; ++e) t = fma(G[c][e], d[e + 1][v] - d[0][v], t);
      q[v] = t;
      sq += t;
    }
    SQ[c] = 0.0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      Q[c][v] = q[v] + sq;
      SQ[c] += Q[c][v];
    }
  }
  f[0] = -(fma(4.0, Q[0][0], -SQ[0]) + fma(4.0, Q[1][0], -SQ[1]) + fma(4.0, Q[2][0], -SQ[2]));
#pragma unroll
  for (int a = 1; a < 4; ++a) f[a] = fma(4.0, Q[a - 1][a], -SQ[a - 1]);
#pragma unroll
  for (int a = 1; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) f[E[a][b]] = 4.0 * (Q[a - 1][b] + Q[b - 1][a]);
#pragma unroll
  for (int a = 1; a < 4; ++a) f[E[0][a]] = 4.0 * (Q[a - 1][0] - (Q[0][a] + Q[1][a] + Q[2][a]));
  const double SV = (ud[0] + ud[1]) + (ud[2] + ud[3]);
  const double SE = ((ud[4] + ud[5]) + (ud[6] + ud[7])) + (ud[8] + ud[9]);
  const double ms = adet * (1.0 / 2520.0);
  const double cv = fma(-6.0, SE, SV), ce = fma(16.0, SE, -6.0 * SV);
  const double src_v = p0 * adet * (1.0 / 120.0), src_e = -p0 * adet * (1.0 / 30.0);  // -p0 * M0_i * adet
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double adj = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (b != a) adj += ud[E[a][b]];
    f[a] += fma(fma(5.0, ud[a], fma(2.0, adj, cv)), ms, src_v);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      int c = -1, dd = -1;  // opposite edge = the other two vertices
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (v != a && v != b) {
          if (c < 0) c = v; else dd = v;
        }
      const int e = E[a][b];
      f[e] += fma(fma(16.0, ud[e], fma(-8.0, ud[E[c][dd]], fma(2.0, ud[a] + ud[b], ce))), ms, src_e);
    }
}

// element residual of one cell, table form or (P2 tetrahedra) sum-factorised
template <int DIM, int DEG, int VAR, int ND, int NP>
__device__ __forceinline__ void heat_element_residual(const double (&gm)[NP], double adet, double p0, const double (&u)[ND],
                                                      const double (&ud)[ND], double (&f)[ND]) {
  using T = RefT<DIM, DEG, VAR>;
  if constexpr (DIM == 3 && DEG == 2) {
    p2tet_heat_residual(gm, adet, p0, u, ud, f);
  } else {
    static_for<ND>([&](auto I) {
      constexpr int i = decltype(I)::value;
      double mass = -p0 * T::M0[i];
      static_for<ND>([&](auto J) {
        constexpr int j = decltype(J)::value;
        constexpr double r = T::R00[i][j];
        if constexpr (r != 0.0) mass = fma(r, ud[j], mass);
      });
      double fi = mass * adet;
      static_for<NP>([&](auto P) {
        constexpr int p = decltype(P)::value;
        double sacc = 0.0;
        static_for<ND>([&](auto J) {
          constexpr int j = decltype(J)::value;
          constexpr double sv = T::S[p][i][j];
          if constexpr (sv != 0.0) sacc = fma(sv, u[j], sacc);
        });
        fi = fma(gm[p], sacc, fi);
      });
      f[i] = fi;
    });
  }
}


__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// loads the W 16-bit staging addresses of one slot (two per 32-bit word); ip = shared address of the
// lane's first word
template <int W>
__device__ __forceinline__ void load_pairs(uint32_t ip, uint32_t (&pr)[(W + 1) / 2]) {
#pragma unroll
  for (int q = 0; q < (W + 1) / 2; ++q) pr[q] = lds_u32(ip + q * 128);
}

// sums the W staged contributions addressed by pr (units of 8 bytes from sbase)
template <int W>
__device__ __forceinline__ double sum_pairs(const uint32_t (&pr)[(W + 1) / 2], uint32_t sbase) {
  double v[W];
#pragma unroll
  for (int k = 0; k < W; ++k)
    v[k] = lds_f64(__byte_perm(pr[k >> 1], 0u, (k & 1) ? 0x4432u : 0x4410u) * 8u + sbase);
  if constexpr (W == 1) return v[0];
  double a0 = v[0], a1 = v[1];
#pragma unroll
  for (int k = 2; k < W; ++k) {
    if (k & 1) a1 += v[k]; else a0 += v[k];
  }
  return a0 + a1;
}

// element-matrix entry from the integer form of the reference tensors (ref_tables.cuh):
//   A_ij = ms * IR_ij + sum_p gs_p * IS_pij,   ms = shift |detJ| / RDEN,  gs_p = |detJ| G_p / SDEN.
// The handful of distinct integer multipliers stay in uniform registers; the plain tables would cost one
// 64-bit immediate (two UMOVs) per use.
template <class T, int I, int J, int NP>
__device__ __forceinline__ double heat_entry_i(double ms, const double (&gs)[NP]) {
  double v = 0.0;
  constexpr int r = T::IR[I][J];
  if constexpr (r != 0) v = ms * (double)r;
  static_for<NP>([&](auto P) {
    constexpr int p = decltype(P)::value;
    constexpr int c = T::IS[p][I][J];
    if constexpr (c != 0) v = fma(gs[p], (double)c, v);
  });
  return v;
}

// slices are dealt round-robin over the whole block (slice s -> warp s mod NWARPS): index inside the
// class of the first slice that belongs to `warp`
template <int NWARPS>
__device__ __forceinline__ uint32_t first_slice_of(int warp, uint32_t first) {
  if constexpr ((NWARPS & (NWARPS - 1)) == 0) return (uint32_t)(warp - (int)first) & (NWARPS - 1);
  const int r = (warp - (int)(first % NWARPS)) % NWARPS;
  return (uint32_t)(r < 0 ? r + NWARPS : r);
}

// one class of slices (equal width W, equal split flag): warp `warp` takes the slices congruent to it
template <int W, int G, int NWARPS, bool PIPE = true>
__device__ __forceinline__ void sweep_class(uint32_t first, uint32_t count, uint32_t off64, bool split, int warp, int lane,
                                            uint32_t idx0, uint32_t gd0, uint32_t sbase, double *__restrict__ out,
                                            uint32_t lo, uint32_t span) {
  constexpr int LG = G == 4 ? 2 : (G == 2 ? 1 : 0), PER_NORM = 32 / G, PAIRS = (W + 1) / 2;
  uint32_t i = first_slice_of<NWARPS>(warp, first);
  const uint32_t gsel = split ? (uint32_t)(lane >> (LG + 2)) : (uint32_t)(lane >> LG);
  const uint32_t sub = split ? (uint32_t)((lane >> 2) & (G - 1)) : (uint32_t)(lane & (G - 1));
  const bool wr = split ? (lane & 3) == 0 : true;
  uint32_t ip = idx0 + (off64 + i * PAIRS) * 128u;
  uint32_t gd = gd0 + ((first + i) * PER_NORM + gsel) * 2u;
  if (i >= count) return;
  if constexpr (!PIPE) {  // plain loop (fits 64 registers: used when the CTA runs 32 warps)
    for (; i < count; i += NWARPS, ip += NWARPS * PAIRS * 128u, gd += NWARPS * PER_NORM * 2u) {
      uint32_t pr[PAIRS];
      load_pairs<W>(ip, pr);
      const uint32_t gs = lds_u16(gd);
      double acc = sum_pairs<W>(pr, sbase);
      if (split) {
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      }
      const uint32_t slot = gs * G + sub;
      if (wr && slot - lo < span) out[slot] = acc;
    }
    return;
  }
  // software pipeline: the addresses and the sector id of the warp's next slice are fetched while the
  // current slice is summed
  uint32_t pr[PAIRS];
  load_pairs<W>(ip, pr);
  uint32_t gs = lds_u16(gd);
  for (;;) {
    const uint32_t in = i + NWARPS;
    const bool has_next = in < count;
    uint32_t prn[PAIRS], gsn = 0;
    ip += NWARPS * PAIRS * 128u;
    gd += NWARPS * PER_NORM * 2u;
    if (has_next) {
      load_pairs<W>(ip, prn);
      gsn = lds_u16(gd);
    }
    double acc = sum_pairs<W>(pr, sbase);
    if (split) {
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    }
    const uint32_t slot = gs * G + sub;  // sector id 0xffff (none) lands far outside [lo, lo + span)
    if (wr && slot - lo < span) out[slot] = acc;
    if (!has_next) break;
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) pr[q] = prn[q];
    gs = gsn;
    i = in;
  }
}


}  // namespace

// implemented in kernels_heat2.cu (return codes: 0 ok, 2 launch failure, 3 unsupported element, 4 does not fit)
int fdts_heat_gather3(int dim, int deg, int var, const G2Args &g, const GatherPlan2 &p, cudaStream_t s);
int fdts_heat_residual_tiles(int dim, int deg, int var, const RTArgs &r, const ResidualPlan &p, int64_t t0, int64_t t1,
                             cudaStream_t s);
