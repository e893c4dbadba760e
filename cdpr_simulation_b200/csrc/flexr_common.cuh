// flexr_common.cuh -- launch / attribute helpers shared by the translation units that hold k_step_flexr instances
#pragma once
#include "launch.h"
#include "step_flexr.cuh"
namespace cdpr {
constexpr int kFlexrTpb = 32;  // one warp per block, as in k_step_flex
// the isotropic-inertia instance when the robot allows it (rc.spec, detected at cdpr_create), else the general one
template <int NC, int NF, bool HOLD, int LANES> static void flexr_go(unsigned grid, const StepArgs &A, cudaStream_t st) {
  if (A.rc.spec & SPEC_ISO) k_step_flexr<NC, kFlexrTpb, NF, HOLD, LANES, true><<<grid, kFlexrTpb, FlexRSmem<NC / LANES, kFlexrTpb, NF, LANES>::bytes, st>>>(A);
  else k_step_flexr<NC, kFlexrTpb, NF, HOLD, LANES, false><<<grid, kFlexrTpb, FlexRSmem<NC / LANES, kFlexrTpb, NF, LANES>::bytes, st>>>(A);
}
template <int NC, int NF, bool HOLD, int LANES> static void flexr_prep() {
  for (const void *f : {(const void *)k_step_flexr<NC, kFlexrTpb, NF, HOLD, LANES, true>, (const void *)k_step_flexr<NC, kFlexrTpb, NF, HOLD, LANES, false>}) {
    cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FlexRSmem<NC / LANES, kFlexrTpb, NF, LANES>::bytes);
    cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
}
// the four (NF, HOLD) instances of one (NC, LANES) shape
#define CDPR_FLEXR_UNIT(NAME, NC_, L_)                                                                           \
  void flexr_prepare_##NAME(int nf, bool hold) {                                                                 \
    if (nf == 0) { if (hold) flexr_prep<NC_, 0, true, L_>(); else flexr_prep<NC_, 0, false, L_>(); }             \
    else { if (hold) flexr_prep<NC_, 1, true, L_>(); else flexr_prep<NC_, 1, false, L_>(); }                     \
  }                                                                                                              \
  void flexr_launch_##NAME(int nf, bool hold, unsigned grid, const StepArgs &A, cudaStream_t st) {               \
    if (nf == 0) { if (hold) flexr_go<NC_, 0, true, L_>(grid, A, st); else flexr_go<NC_, 0, false, L_>(grid, A, st); } \
    else { if (hold) flexr_go<NC_, 1, true, L_>(grid, A, st); else flexr_go<NC_, 1, false, L_>(grid, A, st); }   \
  }                                                                                                              \
  size_t flexr_smem_##NAME(int nf) { return nf == 0 ? FlexRSmem<NC_ / L_, kFlexrTpb, 0, L_>::bytes : FlexRSmem<NC_ / L_, kFlexrTpb, 1, L_>::bytes; }
}  // namespace cdpr
