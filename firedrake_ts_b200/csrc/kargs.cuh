// kargs.cuh -- fills the generic kernel argument block from the engine state.
#pragma once
#include "kernels_generic.cuh"

struct fdts_engine_ext;  // fwd
KArgs fdts_make_kargs(fdts_engine *e, const double *x, const double *xdot, double shift, double *out, int64_t c0,
                      int64_t c1);
int fdts_generic_heat(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s);
int fdts_generic_diffusion(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s);
int fdts_generic_bbm(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s);
int fdts_generic_burgers(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s);
int fdts_generic_ch(bool jac, fdts_engine *e, const KArgs &a, cudaStream_t s);
