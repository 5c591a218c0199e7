// Dependent-issue latency of DFMA / DMUL / DADD / DMMA.8x8x4 on sm_100a (one warp per SM, one chain): cycles per op.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH> __global__ void k_fma(double* out, long long* cyc, int iters, double x) {
  double a[CH];
  for (int i = 0; i < CH; i++) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = fma(a[i], x, 1e-9);
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < CH; i++) s += a[i];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH> __global__ void k_mma(double* out, long long* cyc, int iters, double x) {
  double c0[CH], c1[CH];
  for (int i = 0; i < CH; i++) { c0[i] = i; c1[i] = -i; }
  double a = x + threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int i = 0; i < CH; i++) dmma884(c0[i], c1[i], a, a);
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < CH; i++) s += c0[i] + c1[i];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sincos(double* out, long long* cyc, int iters, double x) {
  double a = x * (threadIdx.x + 1) * 1e-3, acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) { double s, c; sincos(a, &s, &c); a = s * 0.5 + c * 0.25; acc += a; }
  long long t1 = clock64();
  out[threadIdx.x] = acc; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const int iters = 4096; long long h;
#define RUN(name, k, ops) k<<<1, 32>>>(out, cyc, iters, 0.999); cudaDeviceSynchronize(); k<<<1, 32>>>(out, cyc, iters, 0.999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s %7.2f cycles per op (1 warp)\n", name, double(h) / iters / (ops));
  RUN("DFMA dependent chain", k_fma<1>, 1)
  RUN("DFMA 2 chains", k_fma<2>, 2)
  RUN("DFMA 4 chains", k_fma<4>, 4)
  RUN("DFMA 8 chains", k_fma<8>, 8)
  RUN("DMMA dependent chain", k_mma<1>, 1)
  RUN("DMMA 2 chains", k_mma<2>, 2)
  RUN("DMMA 4 chains", k_mma<4>, 4)
  RUN("DMMA 6 chains", k_mma<6>, 6)
  RUN("sincos dependent chain", k_sincos, 1)
  return 0;
}
