#include "launch.h"
#include "step_flex.cuh"
namespace cdpr {
void flex_prepare_u2(int nc, int nf, int lanes);
void flex_prepare_u4(int nc, int nf, int lanes);
void flex_launch_u2(int nc, int nf, int lanes, unsigned grid, const StepArgs &A, cudaStream_t st);
void flex_launch_u4(int nc, int nf, int lanes, unsigned grid, const StepArgs &A, cudaStream_t st);
int flex_tpb() { return 32; }
int flex_stage_slots(int ps, int ds) { const int m = ps > ds ? ps : ds; return m == 0 ? 0 : (m == 1 ? 1 : 4); }
// lanes that share one robot: 1, 2 (or 4 at 8 cables); the shared memory of a block is that of NC / lanes cables per thread
int flex_lanes_supported(int nc, int lanes) { return lanes >= 4 && nc == 8 ? 4 : (lanes >= 2 ? 2 : 1); }
size_t flex_smem_bytes(int nc, int nf, int lanes) {
  const int cpl = nc / flex_lanes_supported(nc, lanes);
#define CDPR_B(CPL_) (nf == 0 ? FlexSmem<CPL_, 32, 0>::bytes : nf == 1 ? FlexSmem<CPL_, 32, 1>::bytes : FlexSmem<CPL_, 32, 4>::bytes)
  return cpl == 8 ? CDPR_B(8) : cpl == 4 ? CDPR_B(4) : cpl == 2 ? CDPR_B(2) : CDPR_B(1);
#undef CDPR_B
}
// unroll = 4 when a cable can never change Pid inside a run (no hold), 2 otherwise (see step_flex.cuh)
void flex_prepare(int nc, int nf, int unroll, int lanes) { if (unroll >= 4) flex_prepare_u4(nc, nf, lanes); else flex_prepare_u2(nc, nf, lanes); }
void flex_launch(int nc, int nf, int unroll, int lanes, unsigned grid, const StepArgs &A, cudaStream_t st) {
  if (unroll >= 4) flex_launch_u4(nc, nf, lanes, grid, A, st); else flex_launch_u2(nc, nf, lanes, grid, A, st);
}
}  // namespace cdpr
