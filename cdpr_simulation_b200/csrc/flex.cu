#include "launch.h"
#include "step_flex.cuh"
namespace cdpr {
void flex_launch(int nc, unsigned grid, int tpb, size_t smem, const StepArgs &A, cudaStream_t st) {
  if (nc == 4) k_step_flex<4><<<grid, tpb, smem, st>>>(A);
  else k_step_flex<8><<<grid, tpb, smem, st>>>(A);
}
void flex_prepare(int nc, size_t smem) {
  const void *f = (nc == 4) ? (const void *)k_step_flex<4> : (const void *)k_step_flex<8>;
  cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
size_t flex_smem_bytes(int nc, int ps, int ds, int tpb) { return sizeof(double) * flex_smem_doubles(nc, ps, ds, tpb); }
}  // namespace cdpr
