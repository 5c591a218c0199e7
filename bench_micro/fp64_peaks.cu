// Micro-benchmark: FP64 FMA pipe, FP64 tensor (DMMA) pipe, their overlap, and sincos
// throughput on sm_100a. Output feeds DESIGN.md's FP64 roofline denominators
// (MEASURED_PEAKS.json only holds HBM + bf16).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

template <int NCH>
__global__ void k_dfma(double* out, int iters, double x) {
  double acc[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) acc[i] = threadIdx.x * 1e-3 + i;
  double m = x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) acc[i] = fma(acc[i], m, 1e-9);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_dmma884(double* out, int iters, double x) {
  double c0[NT], c1[NT];
#pragma unroll
  for (int i = 0; i < NT; i++) { c0[i] = i; c1[i] = -i; }
  double a = x + threadIdx.x * 1e-6, b = x - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_dmma16816(double* out, int iters, double x) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  double a[8], b[4];
  for (int j = 0; j < 8; j++) a[j] = x + threadIdx.x * 1e-6 * j;
  for (int j = 0; j < 4; j++) b[j] = x - threadIdx.x * 1e-6 * j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void k_dmma1684(double* out, int iters, double x) {
  double c[NT][4];
#pragma unroll
  for (int i = 0; i < NT; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  double a[2] = {x + threadIdx.x * 1e-6, x}; double b = x - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++) dmma1684(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// interleave: per iteration NT dmma884 + NF dfma (independent chains) -> do the pipes overlap?
template <int NT, int NF>
__global__ void k_mix(double* out, int iters, double x) {
  double c0[NT > 0 ? NT : 1], c1[NT > 0 ? NT : 1], acc[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < NT; i++) { c0[i] = i; c1[i] = -i; }
#pragma unroll
  for (int i = 0; i < NF; i++) acc[i] = threadIdx.x * 1e-3 + i;
  double a = x + threadIdx.x * 1e-6, b = x - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < (NT > NF ? NT : NF); i++) {
      if (i < NT) dmma884(c0[i], c1[i], a, b);
      if (i < NF) acc[i] = fma(acc[i], x, 1e-9);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_sincos(double* out, int iters, double x) {
  double acc = 0, arg = x * (threadIdx.x + 1) * 1e-3;
  for (int it = 0; it < iters; it++) {
    double s, c; sincos(arg, &s, &c); acc += s * c; arg += 1e-4;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_sqrtdiv(double* out, int iters, double x) {
  double acc = 0, arg = x * (threadIdx.x + 1) * 1e-3 + 1.0;
  for (int it = 0; it < iters; it++) { acc += sqrt(arg) / (arg + 2.0); arg += 1e-4; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_atan(double* out, int iters, double x) {
  double acc = 0, arg = x * (threadIdx.x + 1) * 1e-3;
  for (int it = 0; it < iters; it++) { acc += atan(arg); arg += 1e-4; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F> float timeit(F f, int rep = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < rep; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 4096;
  double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  double nthreads = double(blocks) * threads, nwarps = nthreads / 32;
  float ms;
  ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 0.999); });
  printf("DFMA8           : %8.3f ms  %8.2f TFLOP/s\n", ms, 2.0 * nthreads * iters * 8 / ms * 1e-9);
  ms = timeit([&] { k_dfma<16><<<blocks, threads>>>(out, iters, 0.999); });
  printf("DFMA16          : %8.3f ms  %8.2f TFLOP/s\n", ms, 2.0 * nthreads * iters * 16 / ms * 1e-9);
  ms = timeit([&] { k_dmma884<4><<<blocks, threads>>>(out, iters, 0.999); });
  printf("DMMA m8n8k4 x4  : %8.3f ms  %8.2f TFLOP/s\n", ms, 2.0 * 256 * nwarps * iters * 4 / ms * 1e-9);
  ms = timeit([&] { k_dmma884<10><<<blocks, threads>>>(out, iters, 0.999); });
  printf("DMMA m8n8k4 x10 : %8.3f ms  %8.2f TFLOP/s\n", ms, 2.0 * 256 * nwarps * iters * 10 / ms * 1e-9);
  ms = timeit([&] { k_dmma1684<6><<<blocks, threads>>>(out, iters, 0.999); });
  printf("DMMA m16n8k4 x6 : %8.3f ms  %8.2f TFLOP/s\n", ms, 2.0 * 512 * nwarps * iters * 6 / ms * 1e-9);
  ms = timeit([&] { k_dmma16816<4><<<blocks, threads>>>(out, iters / 4, 0.999); });
  printf("DMMA m16n8k16 x4: %8.3f ms  %8.2f TFLOP/s\n", ms, 2.0 * 2048 * nwarps * (iters / 4) * 4 / ms * 1e-9);
  ms = timeit([&] { k_mix<8, 0><<<blocks, threads>>>(out, iters, 0.999); });
  float ms_t = ms;
  printf("MIX dmma8 only  : %8.3f ms\n", ms);
  ms = timeit([&] { k_mix<0, 8><<<blocks, threads>>>(out, iters, 0.999); });
  printf("MIX dfma8 only  : %8.3f ms\n", ms);
  float ms_f = ms;
  ms = timeit([&] { k_mix<8, 8><<<blocks, threads>>>(out, iters, 0.999); });
  printf("MIX dmma8+dfma8 : %8.3f ms (sum-of-parts %.3f, max-of-parts %.3f)\n", ms, ms_t + ms_f, ms_t > ms_f ? ms_t : ms_f);
  ms = timeit([&] { k_mix<8, 64><<<blocks, threads>>>(out, iters / 4, 0.999); });
  printf("MIX dmma8+dfma64 (iters/4): %8.3f ms\n", ms);
  ms = timeit([&] { k_mix<0, 64><<<blocks, threads>>>(out, iters / 4, 0.999); });
  printf("MIX dfma64 only (iters/4) : %8.3f ms\n", ms);
  ms = timeit([&] { k_sincos<<<blocks, threads>>>(out, iters, 0.7); });
  printf("sincos          : %8.3f ms  %8.2f G sincos/s  (~%.1f dfma-slots each)\n", ms, nthreads * iters / ms * 1e-6,
         (ms / (nthreads * iters)) / (ms_f / (nthreads * iters * 8.0)));
  ms = timeit([&] { k_sqrtdiv<<<blocks, threads>>>(out, iters, 0.7); });
  printf("sqrt+div        : %8.3f ms  %8.2f G/s\n", ms, nthreads * iters / ms * 1e-6);
  ms = timeit([&] { k_atan<<<blocks, threads>>>(out, iters, 0.7); });
  printf("atan            : %8.3f ms  %8.2f G/s\n", ms, nthreads * iters / ms * 1e-6);
  return 0;
}
