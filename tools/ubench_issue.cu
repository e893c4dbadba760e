// ubench_issue.cu -- does a non-FP64 instruction issue "for free" between two DFMAs on B200, or does every FP64
// warp-instruction hold the scheduler's dispatch slot for both of its cycles?
//   per iteration and thread: 64 independent DFMA (8 chains x 8) and M integer LOP3/IADD on 8 other chains
//   model A (free):      cycles/iter/warp = max(2*64, 64 + M)        model B (not free): 2*64 + M
#include <cstdio>
#include <cuda_runtime.h>
template <int M>
__global__ void __launch_bounds__(256) k(double *out, int iters, double s) {
  double a[8];
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = s + i; x[i] = threadIdx.x * 2654435761u + i; }
  const double m = 1.0000001, b = 1e-9;
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, b);
#pragma unroll
      for (int j = 0; j < M / 8; ++j) {
#pragma unroll
        for (int i = 0; i < 8 / 8 + 0; ++i) {}
        x[(u + j) & 7] = (x[(u + j) & 7] ^ (x[(u + j + 1) & 7] >> 3)) + 0x9e3779b9u;   // ~2 integer instructions
      }
    }
  }
  double r = 0; unsigned y = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { r += a[i]; y ^= x[i]; }
  if (r == 123.456 || y == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r + y;
}
template <int M> void run() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * 8, tpb = 256, iters = 2048;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * tpb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<M><<<blocks, tpb>>>(out, 64, 1.0);
  cudaEventRecord(e0); k<M><<<blocks, tpb>>>(out, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double warps_per_smsp = 8.0 * 8 / 4;  // 8 blocks x 8 warps per SM over 4 schedulers
  double cyc_per_iter_per_smsp = ms * 1e-3 * clk * 1e3 / iters / warps_per_smsp;
  printf("M=%3d integer statements per 64 DFMA: %.1f cycles per warp-iteration per scheduler (2*64 = 128; DFMA rate %.2f T/s)\n", M,
         cyc_per_iter_per_smsp, 64.0 * iters * blocks * tpb / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main() { run<0>(); run<16>(); run<32>(); run<64>(); run<128>(); return 0; }
