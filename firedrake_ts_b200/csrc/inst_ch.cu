#include "kargs.cuh"
// mixed space V*V, field-blocked dofs (examples/cahn-hilliard.py:15-16)
int fdts_generic_ch(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s) {
  const int dim = e->cfg.dim, degree = e->cfg.degree;
  if (e->cfg.ncomp != 2 || e->cfg.layout != FDTS_BLOCKED) return 3;
  if (e->cfg.cell_type == FDTS_QUADRILATERAL) {  // the literal example: Q1 x Q1 on quadrilaterals (cahn-hilliard.py:14-16)
    if (dim == 2 && degree == 1) return launch_generic<FDTS_CAHN_HILLIARD, 2, 4, 2, 1, FDTS_QUADRILATERAL>(jac, a, s);
    return 3;
  }
#define CASE(D, K, N) \
  if (dim == D && degree == K) return launch_generic<FDTS_CAHN_HILLIARD, D, N, 2, 1>(jac, a, s);
  CASE(1, 1, 2) CASE(1, 2, 3)
  CASE(2, 1, 3) CASE(2, 2, 6) CASE(2, 3, 10)
  CASE(3, 1, 4) CASE(3, 2, 10)
#undef CASE
  return 3;
}
