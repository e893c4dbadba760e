#include "launch.h"
#include "step_general.cuh"
namespace cdpr {
void general_launch(int dmax, unsigned grid, const StepArgs &A, cudaStream_t st) {
  if (dmax <= 2) k_step_general<2><<<grid, kTpb, 0, st>>>(A);
  else k_step_general<4><<<grid, kTpb, 0, st>>>(A);
}
}  // namespace cdpr
