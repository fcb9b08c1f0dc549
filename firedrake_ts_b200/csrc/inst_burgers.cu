#include "kargs.cuh"
// vector space, interleaved components, NC = DIM (examples/burgers.py:9)
int fdts_generic_burgers(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s) {
  const int dim = e->cfg.dim, degree = e->cfg.degree;
  if (e->cfg.ncomp != dim || e->cfg.layout != FDTS_INTERLEAVED) return 3;
#define CASE(D, K, N) \
  if (dim == D && degree == K) return launch_generic<FDTS_BURGERS, D, N, D, 0>(jac, a, s);
  CASE(1, 1, 2) CASE(1, 2, 3) CASE(1, 3, 4)
  CASE(2, 1, 3) CASE(2, 2, 6) CASE(2, 3, 10)
  CASE(3, 1, 4) CASE(3, 2, 10)
#undef CASE
  return 3;
}
