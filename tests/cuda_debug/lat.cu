// lat.cu — dependent-chain latencies of the fp64 operations the eigensolver and LAE kernels sit on (B200 microbenchmark).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, double a, double b, int n, long long* cyc) {
  double x = a + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = fma(x, b, a);
    if (OP == 1) x = x + b;
    if (OP == 2) x = a / x + b;
    if (OP == 3) x = sqrt(x) + b;
    if (OP == 4) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + b; }
    if (OP == 5) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); r = fma(fma(-x, r, 1.0), r, r); r = fma(fma(-x, r, 1.0), r, r); x = fma(-a, r, b); }
    if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1) + b;
    if (OP == 7) { x = (x < b) ? a : x; x = x + b; }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const char* names[] = {"dfma", "dadd", "ieee div + dadd", "sqrt + dadd", "rcp.approx + dadd", "fast_rcp(2 newton) + dfma", "shfl64 + dadd", "dsetp/sel + dadd"};
  const int n = 4096;
  for (int threads = 32; threads <= 256; threads *= 8)
    for (int op = 0; op < 8; ++op) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: chain<0><<<1, threads>>>(out, 1.0000001, 0.9999999, n, cyc); break;
          case 1: chain<1><<<1, threads>>>(out, 1.0, 1e-3, n, cyc); break;
          case 2: chain<2><<<1, threads>>>(out, 1.7, 0.3, n, cyc); break;
          case 3: chain<3><<<1, threads>>>(out, 1.7, 0.3, n, cyc); break;
          case 4: chain<4><<<1, threads>>>(out, 1.7, 0.3, n, cyc); break;
          case 5: chain<5><<<1, threads>>>(out, 1.7, 0.3, n, cyc); break;
          case 6: chain<6><<<1, threads>>>(out, 1.7, 0.3, n, cyc); break;
          case 7: chain<7><<<1, threads>>>(out, 1.7, 0.3, n, cyc); break;
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("threads %4d  %-28s %.1f cycles/iter\n", threads, names[op], (double)h / n);
    }
  return 0;
}
