// microbench.cu -- fp64 pipe micro-benchmarks used as roofline denominators for
// the compute-bound P3 configuration ("builder must measure
// a DFMA and a DMMA micro-benchmark on the box").
#include "fdts_common.cuh"

namespace {
__global__ void dfma_chain(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__global__ void dmma_chain(double *out, int iters) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
  for (int i = 0; i < iters; ++i) {
    dmma(c0, a, b);
    dmma(c1, a, b);
    dmma(c2, a, b);
    dmma(c3, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}
}  // namespace

extern "C" int fdts_measure_fp64_peak(int32_t device, int32_t use_dmma, double *gflops) {
  FDTS_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  FDTS_CHECK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 4096;
  double *out = nullptr;
  FDTS_CHECK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    if (use_dmma)
      dmma_chain<<<blocks, threads>>>(out, iters);
    else
      dfma_chain<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    FDTS_CHECK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  // DFMA: 8 fma per thread-iteration (2 flop each); DMMA m8n8k4: 2*8*8*4 = 512 flop per warp instruction, 4 per iteration
  const double flop = use_dmma ? (double)blocks * (threads / 32) * iters * 4.0 * 512.0
                               : (double)blocks * threads * iters * 8.0 * 2.0;
  *gflops = flop / (best * 1e-3) / 1e9;
  cudaFree(out);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}
