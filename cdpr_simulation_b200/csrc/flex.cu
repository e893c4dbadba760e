#include "launch.h"
#include "step_flex.cuh"
namespace cdpr {
void flex_prepare_u2(int nc, int nf);
void flex_prepare_u4(int nc, int nf);
void flex_launch_u2(int nc, int nf, unsigned grid, const StepArgs &A, cudaStream_t st);
void flex_launch_u4(int nc, int nf, unsigned grid, const StepArgs &A, cudaStream_t st);
int flex_tpb() { return 32; }
int flex_stage_slots(int ps, int ds) { const int m = ps > ds ? ps : ds; return m == 0 ? 0 : (m == 1 ? 1 : 4); }
size_t flex_smem_bytes(int nc, int nf) {
  if (nc == 4) return nf == 0 ? FlexSmem<4, 32, 0>::bytes : nf == 1 ? FlexSmem<4, 32, 1>::bytes : FlexSmem<4, 32, 4>::bytes;
  return nf == 0 ? FlexSmem<8, 32, 0>::bytes : nf == 1 ? FlexSmem<8, 32, 1>::bytes : FlexSmem<8, 32, 4>::bytes;
}
// unroll = 4 when a cable can never change Pid inside a run (no hold), 2 otherwise (see step_flex.cuh)
void flex_prepare(int nc, int nf, int unroll) { if (unroll >= 4) flex_prepare_u4(nc, nf); else flex_prepare_u2(nc, nf); }
void flex_launch(int nc, int nf, int unroll, unsigned grid, const StepArgs &A, cudaStream_t st) {
  if (unroll >= 4) flex_launch_u4(nc, nf, grid, A, st); else flex_launch_u2(nc, nf, grid, A, st);
}
}  // namespace cdpr
