#include "kargs.cuh"
int fdts_generic_heat(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s) {
  return dispatch_scalar<FDTS_HEAT>(jac, e->cfg.dim, e->cfg.degree, a, s, e->cfg.cell_type);
}
