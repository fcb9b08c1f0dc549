#include "kargs.cuh"
int fdts_generic_diffusion(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s) {
  return dispatch_scalar<FDTS_DIFFUSION>(jac, e->cfg.dim, e->cfg.degree, a, s, e->cfg.cell_type);
}
