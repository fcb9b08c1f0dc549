// kernels_heat2.cu -- the optional schedules of the closed-form heat path: warp-specialised Jacobian gather
// (producer / consumer / copy warps) and the residual tile kernels (per-tile pre-reduction of the atomics).
#include "heat_common.cuh"

namespace {

// ------------------------------------------------------------------ path 2, plan 2, warp-specialised
// Same plan, same arithmetic, same summation order as heat_jacobian_gather2, but the two phases run
// concurrently on different warps of one persistent CTA with two staging buffers:
//   NPW producer warps   element matrices of block b+1 -> stage[(b+1)&1]        (fp64 pipe)
//   NCW consumer warps   gather/sum/store of block b from stage[b&1]            (shared-memory pipe)
//   1 copy warp          lane 0 issues every TMA bulk copy and fetches the block descriptors
// Hand-over through mbarriers: full[s] (producers -> consumers), empty[s] (consumers -> producers),
// inA[s] (copy warp -> producers); no CTA-wide barrier after the prologue.  The index tables of the
// resident pattern are only reloaded (by the consumers themselves) when the pattern changes.
template <int DIM, int DEG, int VAR, int NPW, int NCW, int G>
__global__ void __launch_bounds__((NPW + NCW + 1) * 32, 1) heat_jacobian_gather3(const G2Args a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP, NEP = (ND * (ND + 1) / 2) | 1;
  constexpr int PER_NORM = 32 / G;
  constexpr int NDESC = 8;
  extern __shared__ __align__(16) unsigned char smraw[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smraw + a.off_bar);
  constexpr int NA = 4;  // ring of phase-A input buffers: copies run three blocks ahead of the producers
  uint64_t *inA = bars, *freeA = bars + NA, *full = bars + 2 * NA, *empty = bars + 2 * NA + 2, *barB = bars + 2 * NA + 4;
  BlockDesc2 *sdesc = reinterpret_cast<BlockDesc2 *>(smraw + a.off_desc);  // ring of NDESC blocks
  const uint32_t sbase = smem_u32(smraw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stride = gridDim.x;
  const int64_t b0 = blockIdx.x;
  if (b0 >= a.nblocks) return;
  const int64_t nmine = (a.nblocks - b0 + stride - 1) / stride;  // blocks of this CTA

  auto issue_A = [&](const BlockDesc2 &d, int buf) {
    const uint32_t vb = (uint32_t)(((d.nvert * DIM + 1) & ~1) * 8);
    const uint32_t cb = (uint32_t)(((d.ncell + 3) & ~3) * 4);
    mbar_expect_tx(&inA[buf], vb + cb);
    if (vb) bulk_g2s(smraw + a.off_vtab + buf * a.inA_bytes, a.bvcoords + d.vc_off, vb, &inA[buf]);
    if (cb) bulk_g2s(smraw + a.off_lcc + buf * a.inA_bytes, a.bcc + d.cc_off, cb, &inA[buf]);
  };

  // ---- prologue (the only CTA-wide barrier)
  if (threadIdx.x == 0) {
    for (int q = 0; q < NA; ++q) {
      mbar_init(&inA[q], 1);
      mbar_init(&freeA[q], NPW);
    }
    mbar_init(&full[0], NPW);
    mbar_init(&full[1], NPW);
    mbar_init(&empty[0], NCW);
    mbar_init(&empty[1], NCW);
    mbar_init(barB, 1);
    for (int q = 0; q < NA && q < nmine; ++q) {
      sdesc[q] = a.bdesc[b0 + q * stride];
      if (q < NA - 1) issue_A(sdesc[q], q);
    }
  }
  if (threadIdx.x < 32) {  // one staged zero per 8-byte bank, in both staging buffers
    const int q = threadIdx.x & 15, buf = threadIdx.x >> 4;
    sts_f64(sbase + buf * a.stage_bytes + (a.zaddr + q) * 8, 0.0);
  }
  __syncthreads();

  if (warp == NPW + NCW) {
    // ================= copy warp
    if (lane == 0) {
      for (int64_t it = 0; it < nmine; ++it) {
        // the phase-A inputs of block it+NA-1 reuse the buffer of block it-1: wait until the producers have
        // released it (freeA cannot run ahead of this wait: its next phase needs the refill issued here);
        // the copy is issued first, the (slow, dependent) descriptor fetch for block it+NA afterwards
        if (it >= 1) mbar_wait(&freeA[(it - 1) % NA], (uint32_t)(((it - 1) / NA) & 1));
        if (it + NA - 1 < nmine) issue_A(sdesc[(it + NA - 1) % NDESC], (int)((it + NA - 1) % NA));
        if (it + NA < nmine) sdesc[(it + NA) % NDESC] = a.bdesc[b0 + (it + NA) * stride];
      }
    }
  } else if (warp < NPW) {
    // ================= producers: element matrices
    const int tid = threadIdx.x;  // 0 .. NPW*32-1
    for (int64_t it = 0; it < nmine; ++it) {
      const int s = (int)(it & 1);
      if (it >= 2) mbar_wait(&empty[s], (uint32_t)((((it >> 1) - 1)) & 1));
      const int qa = (int)(it % NA);
      mbar_wait(&inA[qa], (uint32_t)((it / NA) & 1));
      const BlockDesc2 &d = sdesc[it % NDESC];
      const int ncell = FDTS_DEV(a, 1) ? 0 : d.ncell;
      const uint32_t lcc0 = sbase + a.off_lcc + qa * a.inA_bytes, vt0 = sbase + a.off_vtab + qa * a.inA_bytes;
      const uint32_t stb = sbase + s * a.stage_bytes;
      for (int k = tid; k < ncell; k += NPW * 32) {
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
        double gm[NP];
        metric<DIM>(g, gm);
        const uint32_t st = stb + k * (NEP * 8);
        if constexpr (T::INT) {
          const double ms = a.shift * g.adet * (1.0 / T::RDEN);
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
          const double m = a.shift * g.adet;
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
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&freeA[qa]);
        mbar_arrive(&full[s]);
      }
    }
  } else {
    // ================= consumers: gather, sum, store
    const int cw = warp - NPW;  // consumer warp id
    const uint32_t idx0 = sbase + a.off_idx + lane * 4;
    const uint32_t gd0 = sbase + a.off_gdst;
    const uint32_t cls0 = sbase + a.off_meta;
    int resident = -1;
    uint32_t nloadB = 0;
    for (int64_t it = 0; it < nmine; ++it) {
      const int s = (int)(it & 1);
      mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));  // also orders the descriptor written by the copy warp
      const BlockDesc2 &d = sdesc[it % NDESC];
      const int pattern = d.pattern;
      if (pattern != resident) {
        // all consumers are done with the old tables (named barrier), then one of them reloads
        asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
        if (cw == 0 && lane == 0) {
          const uint32_t nslp = (uint32_t)((d.nsl + 3) & ~3);
          const uint32_t ib = (uint32_t)d.nidx * 2u, mb = (uint32_t)((d.ncls + 1) & ~1) * 8u, gb = nslp * 2u * (uint32_t)PER_NORM;
          mbar_expect_tx(barB, ib + mb + gb);
          if (ib) bulk_g2s(smraw + a.off_idx, a.cidx + d.idx_off, ib, barB);
          if (mb) bulk_g2s(smraw + a.off_meta, a.cmeta + d.cls_off, mb, barB);
          if (gb) bulk_g2s(smraw + a.off_gdst, a.gdst + d.meta_off * PER_NORM, gb, barB);
        }
        mbar_wait(barB, nloadB & 1u);
        ++nloadB;
        resident = pattern;
      }
      const uint32_t lo = (uint32_t)(d.s0 & (G - 1)), span = FDTS_DEV(a, 8) ? 0u : (uint32_t)d.nslots;
      double *__restrict__ out = a.out + (d.s0 - lo);
      const uint32_t stb = sbase + s * a.stage_bytes;
      const int ncls = FDTS_DEV(a, 2) ? 0 : d.ncls;
      for (int c = 0; c < ncls; ++c) {
        uint32_t c0, c1;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(c0), "=r"(c1) : "r"(cls0 + c * 8));
        const uint32_t first = c0 & 0xffffu, count = c0 >> 16, off64 = c1 & 0xffffu, w = (c1 >> 16) & 0xffu;
        const bool split = c1 >= 0x1000000u;
#define SWEEP(W) sweep_class<W, G, NCW, (NCW <= 16)>(first, count, off64, split, cw, lane, idx0, gd0, stb, out, lo, span)
        switch (w) {
          case 0: {  // slots without contributions (Dirichlet rows/columns): write zeros
            constexpr int LG = G == 4 ? 2 : (G == 2 ? 1 : 0);
            for (uint32_t i = first_slice_of<NCW>(cw, first); i < count; i += NCW) {
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
            for (uint32_t i = first_slice_of<NCW>(cw, first); i < count; i += NCW) {
              const uint32_t ip = idx0 + (off64 + i * pairs) * 128u;
              const uint32_t gs = lds_u16(gd0 + ((first + i) * PER_NORM + (split ? (lane >> (LG + 2)) : (lane >> LG))) * 2u);
              double a0 = 0.0, a1 = 0.0;
              for (uint32_t q = 0; q < pairs; ++q) {
                const uint32_t pr = lds_u32(ip + q * 128);
                a0 += lds_f64(stb + (pr & 0xffffu) * 8u);
                if (2 * q + 1 < w) a1 += lds_f64(stb + (pr >> 16) * 8u);
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

// shared-memory layout of the warp-specialised kernel; returns the total size (0: does not fit)
template <int DIM>
size_t layout_gather3(G2Args &a, const GatherPlan2 &p) {
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  a.stage_bytes = (int)up16(sizeof(double) * ((size_t)p.zaddr + 16));
  size_t off = 2 * (size_t)a.stage_bytes;
  a.off_idx = (int)off;
  off += up16(2 * (size_t)p.max_nidx + 64);
  a.off_meta = (int)off;
  off += up16(8 * (size_t)PLAN2_MAX_CLASSES + 16);
  a.off_gdst = (int)off;
  off += up16(2 * (size_t)(32 / p.sector) * (size_t)p.max_nsl + 16);
  a.sector = p.sector;
  a.inA_bytes = (int)(up16(8 * (size_t)p.vmax * DIM + 16) + up16(4 * (size_t)p.cmax + 16));
  a.off_vtab = (int)off;
  a.off_lcc = (int)(off + up16(8 * (size_t)p.vmax * DIM + 16));
  off += 4 * (size_t)a.inA_bytes;
  a.off_bar = (int)off;
  off += 128;
  a.off_desc = (int)off;
  off += 8 * sizeof(BlockDesc2);
  a.zaddr = p.zaddr;
  return off <= 227 * 1024 ? off : 0;
}

template <int DIM, int DEG, int VAR>
int launch_gather3(G2Args a, const GatherPlan2 &p, cudaStream_t s) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NEP = (ND * (ND + 1) / 2) | 1;
  constexpr int NPW = 6, NCW = 25;  // 192 cells per pass of the producers; 32 warps, 64 registers each
  if (p.nep != NEP) return 3;
  const size_t smem = layout_gather3<DIM>(a, p);
  if (smem == 0) return 4;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t grid = p.nblocks < sms ? p.nblocks : sms;
  auto launch = [&](auto kern) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    kern<<<(unsigned)grid, (NPW + NCW + 1) * 32, smem, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
  };
  if (p.sector == 4) return launch(heat_jacobian_gather3<DIM, DEG, VAR, NPW, NCW, 4>);
  if (p.sector == 2) return launch(heat_jacobian_gather3<DIM, DEG, VAR, NPW, NCW, 2>);
  return launch(heat_jacobian_gather3<DIM, DEG, VAR, NPW, NCW, 1>);
}

// ------------------------------------------------------------------ residual, tile plan
// One CTA per cell tile (gather_plan.cu: residual tile plan).  The tile's node values and vertex
// coordinates come in with coalesced loads and live in shared memory, every thread computes the element
// vector of one cell (closed form), the vectors are pre-reduced per unique node of the tile through
// SELL lists (lane = node) and ONE `red.global.add.f64` per (tile, node) goes out -- 5x fewer atomics and
// no scattered 8-byte gathers of u, udot and the coordinates (the per-cell atomic kernel is bound by
// exactly those: 776 M L1 sector requests per call, ncu l1tex 69 % of peak).
template <int DIM, int DEG, int VAR, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) heat_residual_tiles(const RTArgs a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP, NDP = ND | 1, NWARPS = NT / 32;
  extern __shared__ __align__(16) unsigned char smraw[];
  double *stage = reinterpret_cast<double *>(smraw);
  double *su = reinterpret_cast<double *>(smraw + a.off_su);
  double *sud = reinterpret_cast<double *>(smraw + a.off_sud);
  double *vtab = reinterpret_cast<double *>(smraw + a.off_vtab);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t = a.t0 + blockIdx.x;
  const BlockDesc2 d = a.tdesc[t];
  const int64_t pat = d.pattern;
  const int32_t *__restrict__ tn = a.tnode + pat * a.umaxp;
  // ---- stage node values and vertex coordinates
  for (int q = threadIdx.x; q < d.nslots; q += NT) {
    const int64_t node = d.s0 + (tn[q] & 0x7fffffff);
    su[q] = __ldg(a.x + node);
    sud[q] = __ldg(a.xdot + node);
  }
  for (int q = threadIdx.x; q < d.nvert * DIM; q += NT) vtab[q] = __ldg(a.bvcoords + d.vc_off + q);
  if (threadIdx.x < 16) stage[a.zaddr + threadIdx.x] = 0.0;
  __syncthreads();
  // ---- element vectors
  for (int k = threadIdx.x; k < d.ncell; k += NT) {
    const uint32_t packed = a.bcc[d.cc_off + k];
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
    // two passes to keep the register footprint low (4 CTAs per SM): mass with udot, then stiffness with u
    uint16_t lid[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) lid[i] = a.tlid[(pat * ND + i) * a.cmaxp + k];
    double *st = stage + k * NDP;
    {
      double ud[ND];
#pragma unroll
      for (int i = 0; i < ND; ++i) ud[i] = sud[lid[i]];
      static_for<ND>([&](auto I) {
        constexpr int i = decltype(I)::value;
        double mass = -a.p0 * T::M0[i];
        static_for<ND>([&](auto J) {
          constexpr int j = decltype(J)::value;
          constexpr double r = T::R00[i][j];
          if constexpr (r != 0.0) mass = fma(r, ud[j], mass);
        });
        st[i] = mass * g.adet;
      });
    }
    {
      double u[ND];
#pragma unroll
      for (int i = 0; i < ND; ++i) u[i] = su[lid[i]];
      static_for<ND>([&](auto I) {
        constexpr int i = decltype(I)::value;
        double f = st[i];
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
  }
  __syncthreads();
  // ---- per-node sums (SELL sweep, lane = unique node of the tile), one atomic per node
  const uint16_t *__restrict__ idx = a.cidx + d.idx_off;
  const uint16_t *__restrict__ gd = a.gdst + d.meta_off * 32;
  const uint2 *__restrict__ cls = a.cmeta + d.cls_off;
  for (int c = 0; c < d.ncls; ++c) {
    const uint2 ce = cls[c];
    const int first = (int)(ce.x & 0xffffu), count = (int)(ce.x >> 16), w = (int)((ce.y >> 16) & 0xffu);
    const bool split = (ce.y >> 24) != 0;
    const uint32_t pairs = (uint32_t)((w + 1) >> 1);
    for (int i = first_slice_of<NWARPS>(warp, (uint32_t)first); i < count; i += NWARPS) {
      const uint32_t *ip = reinterpret_cast<const uint32_t *>(idx) + ((size_t)(ce.y & 0xffffu) + (size_t)i * pairs) * 32 + lane;
      double a0 = 0.0, a1 = 0.0;
      for (uint32_t q = 0; q < pairs; ++q) {
        const uint32_t pr = __ldg(ip + q * 32);
        a0 += stage[pr & 0xffffu];
        if (2 * q + 1 < (uint32_t)w) a1 += stage[pr >> 16];
      }
      double acc = a0 + a1;
      int gsel = lane;
      bool wr = true;
      if (split) {
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        gsel = lane >> 2;
        wr = (lane & 3) == 0;
      }
      const uint32_t ln = gd[(size_t)(first + i) * 32 + gsel];
      if (wr && ln != 0xffffu) {
        const int32_t rel = tn[ln];
        if (rel >= 0) red_add_f64(a.out + d.s0 + rel, acc);
      }
    }
  }
}

// variant: the cells gather their inputs directly (like the per-cell kernel: high memory-level
// parallelism, no staging phase), only the element vectors are staged and pre-reduced per tile
template <int DIM, int DEG, int VAR, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) heat_residual_tiles_direct(const RTArgs a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP, NDP = ND | 1, NWARPS = NT / 32;
  extern __shared__ __align__(16) unsigned char smraw[];
  double *stage = reinterpret_cast<double *>(smraw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t = a.t0 + blockIdx.x;
  const BlockDesc2 d = a.tdesc[t];
  const int64_t c0 = a.tstart[t];
  if (threadIdx.x < 16) stage[a.zaddr + threadIdx.x] = 0.0;
  for (int k = threadIdx.x; k < d.ncell; k += NT) {
    const int64_t c = c0 + k;
    int32_t cc[DIM + 1], nodes[ND];
#pragma unroll
    for (int v = 0; v <= DIM; ++v) cc[v] = a.cell_coord_t[v * a.nc + c];
#pragma unroll
    for (int i = 0; i < ND; ++i) nodes[i] = a.cell_node_t[i * a.nc + c];
    Geom<DIM> g;
    compute_geometry<DIM>(a.coords, cc, g);
    double gm[NP];
    metric<DIM>(g, gm);
    double u[ND], ud[ND], f[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      u[i] = __ldg(a.x + nodes[i]);
      ud[i] = __ldg(a.xdot + nodes[i]);
    }
    heat_element_residual<DIM, DEG, VAR, ND, NP>(gm, g.adet, a.p0, u, ud, f);
    double *st = stage + k * NDP;
#pragma unroll
    for (int i = 0; i < ND; ++i) st[i] = f[i];
  }
  __syncthreads();
  const int64_t pat = d.pattern;
  const int32_t *__restrict__ tn = a.tnode + pat * a.umaxp;
  const uint16_t *__restrict__ idx = a.cidx + d.idx_off;
  const uint16_t *__restrict__ gd = a.gdst + d.meta_off * 32;
  const uint2 *__restrict__ cls = a.cmeta + d.cls_off;
  for (int c = 0; c < d.ncls; ++c) {
    const uint2 ce = cls[c];
    const int first = (int)(ce.x & 0xffffu), count = (int)(ce.x >> 16), w = (int)((ce.y >> 16) & 0xffu);
    const bool split = (ce.y >> 24) != 0;
    const uint32_t pairs = (uint32_t)((w + 1) >> 1);
    for (int i = first_slice_of<NWARPS>(warp, (uint32_t)first); i < count; i += NWARPS) {
      const uint32_t *ip = reinterpret_cast<const uint32_t *>(idx) + ((size_t)(ce.y & 0xffffu) + (size_t)i * pairs) * 32 + lane;
      double a0 = 0.0, a1 = 0.0;
      for (uint32_t q = 0; q < pairs; ++q) {
        const uint32_t pr = __ldg(ip + q * 32);
        a0 += stage[pr & 0xffffu];
        if (2 * q + 1 < (uint32_t)w) a1 += stage[pr >> 16];
      }
      double acc = a0 + a1;
      int gsel = lane;
      bool wr = true;
      if (split) {
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        gsel = lane >> 2;
        wr = (lane & 3) == 0;
      }
      const uint32_t ln = gd[(size_t)(first + i) * 32 + gsel];
      if (wr && ln != 0xffffu) {
        const int32_t rel = tn[ln];
        if (rel >= 0) red_add_f64(a.out + d.s0 + rel, acc);
      }
    }
  }
}

// variant 3: like the direct variant, but the tile's reduction tables (SELL addresses, slice classes, node
// ids) are fetched into shared memory with cp.async while the element vectors are computed, so the
// pre-reduction after the barrier reads shared memory only (in the direct variant every step of the sweep
// waits for a dependent L2 access: ~10 k cycles per tile, which made it slower than the per-cell kernel)
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

template <int DIM, int DEG, int VAR, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) heat_residual_tiles_direct3(const RTArgs a) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int ND = T::ND, NP = T::NP, NDP = ND | 1, NWARPS = NT / 32;
  extern __shared__ __align__(16) unsigned char smraw[];
  double *stage = reinterpret_cast<double *>(smraw);
  uint32_t *sidx = reinterpret_cast<uint32_t *>(smraw + a.off_sidx);
  uint16_t *sgd = reinterpret_cast<uint16_t *>(smraw + a.off_sgd);
  uint2 *scls = reinterpret_cast<uint2 *>(smraw + a.off_scls);
  int32_t *stn = reinterpret_cast<int32_t *>(smraw + a.off_stn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t = a.t0 + blockIdx.x;
  const BlockDesc2 d = a.tdesc[t];
  const int64_t c0 = a.tstart[t];
  const int64_t pat = d.pattern;
  {  // tables: issued now, complete behind the element-vector phase
    const unsigned char *g = reinterpret_cast<const unsigned char *>(a.cidx + d.idx_off);
    for (int q = threadIdx.x; q < d.nidx / 8; q += NT) cp_async16(reinterpret_cast<unsigned char *>(sidx) + q * 16, g + q * 16);
    g = reinterpret_cast<const unsigned char *>(a.gdst + d.meta_off * 32);
    for (int q = threadIdx.x; q < d.nsl * 4; q += NT) cp_async16(reinterpret_cast<unsigned char *>(sgd) + q * 16, g + q * 16);
    const uint32_t *gc = reinterpret_cast<const uint32_t *>(a.cmeta + d.cls_off);
    for (int q = threadIdx.x; q < d.ncls * 2; q += NT) cp_async4(reinterpret_cast<uint32_t *>(scls) + q, gc + q);
    const int32_t *gt = a.tnode + pat * a.umaxp;
    for (int q = threadIdx.x; q < d.nslots; q += NT) cp_async4(stn + q, gt + q);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (threadIdx.x < 16) stage[a.zaddr + threadIdx.x] = 0.0;
  for (int k = threadIdx.x; k < d.ncell; k += NT) {
    const int64_t c = c0 + k;
    int32_t cc[DIM + 1], nodes[ND];
#pragma unroll
    for (int v = 0; v <= DIM; ++v) cc[v] = a.cell_coord_t[v * a.nc + c];
#pragma unroll
    for (int i = 0; i < ND; ++i) nodes[i] = a.cell_node_t[i * a.nc + c];
    Geom<DIM> g;
    compute_geometry<DIM>(a.coords, cc, g);
    double gm[NP];
    metric<DIM>(g, gm);
    double u[ND], ud[ND], f[ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      u[i] = __ldg(a.x + nodes[i]);
      ud[i] = __ldg(a.xdot + nodes[i]);
    }
    heat_element_residual<DIM, DEG, VAR, ND, NP>(gm, g.adet, a.p0, u, ud, f);
    double *st = stage + k * NDP;
#pragma unroll
    for (int i = 0; i < ND; ++i) st[i] = f[i];
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int c = 0; c < d.ncls; ++c) {
    const uint2 ce = scls[c];
    const int first = (int)(ce.x & 0xffffu), count = (int)(ce.x >> 16), w = (int)((ce.y >> 16) & 0xffu);
    const bool split = (ce.y >> 24) != 0;
    const uint32_t pairs = (uint32_t)((w + 1) >> 1);
    for (int i = first_slice_of<NWARPS>(warp, (uint32_t)first); i < count; i += NWARPS) {
      const uint32_t *ip = sidx + ((size_t)(ce.y & 0xffffu) + (size_t)i * pairs) * 32 + lane;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
      for (uint32_t q = 0; q < pairs; ++q) {
        const uint32_t pr = ip[q * 32];
        a0 += stage[pr & 0xffffu];
        if (2 * q + 1 < (uint32_t)w) a1 += stage[pr >> 16];
      }
      double acc = a0 + a1;
      int gsel = lane;
      bool wr = true;
      if (split) {
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        gsel = lane >> 2;
        wr = (lane & 3) == 0;
      }
      const uint32_t ln = sgd[(size_t)(first + i) * 32 + gsel];
      if (wr && ln != 0xffffu) {
        const int32_t rel = stn[ln];
        if (rel >= 0) red_add_f64(a.out + d.s0 + rel, acc);
      }
    }
  }
}

template <int DIM, int DEG, int VAR>
int launch_residual_tiles(RTArgs a, const ResidualPlan &p, int64_t t0, int64_t t1, cudaStream_t s) {
  using T = RefT<DIM, DEG, VAR>;
  constexpr int NT = 192;
  if (p.ndp != (T::ND | 1)) return 3;
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t off = up16(sizeof(double) * ((size_t)p.zaddr + 16));
  a.off_su = (int)off;
  off += up16(8 * (size_t)p.umaxp);
  a.off_sud = (int)off;
  off += up16(8 * (size_t)p.umaxp);
  a.off_vtab = (int)off;
  off += up16(8 * (size_t)p.vmax * DIM + 16);
  const size_t smem = off;
  if (smem > 227 * 1024) return 4;
  a.t0 = t0;
  if (a.direct == 2) {
    size_t o = up16(sizeof(double) * ((size_t)p.zaddr + 16));
    a.off_sidx = (int)o;
    o += up16(2 * (size_t)p.max_nidx + 64);
    a.off_sgd = (int)o;
    o += up16(64 * (size_t)p.max_nsl + 64);
    a.off_scls = (int)o;
    o += up16(8 * (size_t)PLAN2_MAX_CLASSES);
    a.off_stn = (int)o;
    o += up16(4 * (size_t)p.umaxp + 16);
    if (o > 227 * 1024) return 4;
    auto k3 = heat_residual_tiles_direct3<DIM, DEG, VAR, NT, RT_MINB>;
    if (o > 48 * 1024 && cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)o) != cudaSuccess) return 2;
    if (t1 > t0) k3<<<(unsigned)(t1 - t0), NT, o, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
  }
  if (a.direct) {
    const size_t sm2 = up16(sizeof(double) * ((size_t)p.zaddr + 16));
    auto k2 = heat_residual_tiles_direct<DIM, DEG, VAR, NT, RT_MINB>;
    if (sm2 > 48 * 1024 && cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2) != cudaSuccess) return 2;
    if (t1 > t0) k2<<<(unsigned)(t1 - t0), NT, sm2, s>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
  }
  auto k = heat_residual_tiles<DIM, DEG, VAR, NT, RT_MINB>;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
  if (t1 > t0) k<<<(unsigned)(t1 - t0), NT, smem, s>>>(a);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

}  // namespace

#define HEAT_CASES(M)                                 \
  M(1, 1, 1) M(1, 2, 1) M(1, 3, 0) M(1, 3, 1)         \
  M(2, 1, 1) M(2, 2, 1) M(2, 3, 0) M(2, 3, 1)         \
  M(3, 1, 1) M(3, 2, 1) M(3, 3, 0) M(3, 3, 1)

int fdts_heat_gather3(int dim, int deg, int var, const G2Args &g, const GatherPlan2 &p, cudaStream_t s) {
  int rc = 3;
#define G3CASE(D, K, V) \
  if (dim == D && deg == K && (K < 3 || var == V)) rc = launch_gather3<D, K, V>(g, p, s);
  HEAT_CASES(G3CASE)
#undef G3CASE
  return rc;
}

int fdts_heat_residual_tiles(int dim, int deg, int var, const RTArgs &r, const ResidualPlan &p, int64_t t0, int64_t t1,
                             cudaStream_t s) {
  int rc = 3;
#define RTCASE(D, K, V) \
  if (dim == D && deg == K && (K < 3 || var == V)) rc = launch_residual_tiles<D, K, V>(r, p, t0, t1, s);
  HEAT_CASES(RTCASE)
#undef RTCASE
  return rc;
}
