// misc.cu — the FP64 roofline denominator.  MEASURED_PEAKS.json holds HBM and bf16 peaks only; the
// k-means / KNN kernels are bound by the FP64 FMA pipe, so its throughput is measured here with a
// register-resident DFMA loop (8 independent chains per thread, every SM full) — SURVEY.md §2b note.
#include "kernels.cuh"

namespace flgp {

namespace {

__global__ void __launch_bounds__(256) dfma_peak_kernel(int iters, double seed, double* out) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c);
    a1 = fma(a1, m, c);
    a2 = fma(a2, m, c);
    a3 = fma(a3, m, c);
    a4 = fma(a4, m, c);
    a5 = fma(a5, m, c);
    a6 = fma(a6, m, c);
    a7 = fma(a7, m, c);
  }
  double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 12345.678) out[0] = r;  // never true; keeps the chains alive
}

}  // namespace

double dfma_peak_run(Ctx* c, int iters) {
  if (iters < 1) iters = 4096;
  DevBuf<double> out(1);
  const int grid = c->sm_count * 8;
  cudaEvent_t e0, e1;
  FLGP_CUDA(cudaEventCreate(&e0));
  FLGP_CUDA(cudaEventCreate(&e1));
  FLGP_LAUNCH(c, dfma_peak_kernel, grid, 256, 0, iters, 1.0, out.p);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    FLGP_CUDA(cudaEventRecord(e0, c->stream));
    FLGP_LAUNCH(c, dfma_peak_kernel, grid, 256, 0, iters, 1.0 + rep, out.p);
    FLGP_CUDA(cudaEventRecord(e1, c->stream));
    FLGP_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FLGP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = 2.0 * 8.0 * (double)iters * 256.0 * grid;
  return flops / (best * 1e-3) / 1e12;
}

}  // namespace flgp
