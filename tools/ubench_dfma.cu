// ubench_dfma.cu -- how fast does the B200 FP64 pipe go for different operand mixes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_dfma tools/ubench_dfma.cu && tools/ubench_dfma
// variant 0: a = fma(a, m, b)      m, b shared by all chains (operand reuse friendly)  [the roofline denominator]
// variant 1: a_i = fma(a_i, b_i, c_i)  three distinct 64-bit register operands per DFMA
// variant 2: a_i = fma(a_i, b_i, K)    two register operands + one constant-bank operand
// variant 3: mix 50% DFMA(3 reg) / 25% DMUL / 25% DADD
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void __launch_bounds__(256) k(double *out, int iters, double s, const double kc) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = s + i; b[i] = 1.0 + 1e-9 * (i + threadIdx.x); c[i] = 1e-9 * (i + 1); }
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (V == 0) a[i] = fma(a[i], b[0], c[0]);
        if (V == 1) a[i] = fma(a[i], b[i], c[i]);
        if (V == 2) a[i] = fma(a[i], b[i], kc);
        if (V == 3) { if ((u & 3) < 2) a[i] = fma(a[i], b[i], c[i]); else if ((u & 3) == 2) a[i] = a[i] * b[i]; else a[i] = a[i] + c[i]; }
      }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i];
  if (r == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int V> void run(const char *name, int wpb) {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * (wpb <= 2 ? 4 : 8), tpb = wpb * 32, iters = 4096;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * tpb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<blocks, tpb>>>(out, 64, 1.0, 1e-9);
  cudaEventRecord(e0); k<V><<<blocks, tpb>>>(out, iters, 1.0, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double instr = 64.0 * iters * blocks * tpb;
  printf("%-34s warps/block %d blocks %d: %.2f T FP64-instr/s  (x2 = %.2f TFLOP/s if all FMA)\n", name, wpb, blocks, instr / (ms * 1e-3) / 1e12, 2 * instr / (ms * 1e-3) / 1e12);
  cudaFree(out);
}
int main() {
  for (int wpb : {2, 8}) {
    run<0>("DFMA shared m,b (reuse)", wpb);
    run<1>("DFMA 3 distinct reg operands", wpb);
    run<2>("DFMA 2 reg + constant", wpb);
    run<3>("50% DFMA3 / 25% DMUL / 25% DADD", wpb);
  }
  return 0;
}
