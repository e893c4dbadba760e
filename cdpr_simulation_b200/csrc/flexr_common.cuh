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
// the six (NF, HOLD) instances of one (NC, LANES) shape
#define CDPR_FLEXR_BY_NF(WHAT, NC_, L_, ...)                                                                     \
  do {                                                                                                           \
    if (nf == 0) { if (hold) WHAT<NC_, 0, true, L_>(__VA_ARGS__); else WHAT<NC_, 0, false, L_>(__VA_ARGS__); }   \
    else if (nf == 1) { if (hold) WHAT<NC_, 1, true, L_>(__VA_ARGS__); else WHAT<NC_, 1, false, L_>(__VA_ARGS__); } \
    else { if (hold) WHAT<NC_, 2, true, L_>(__VA_ARGS__); else WHAT<NC_, 2, false, L_>(__VA_ARGS__); }           \
  } while (0)
#define CDPR_FLEXR_UNIT(NAME, NC_, L_)                                                                           \
  void flexr_prepare_##NAME(int nf, bool hold) { CDPR_FLEXR_BY_NF(flexr_prep, NC_, L_); }                        \
  void flexr_launch_##NAME(int nf, bool hold, unsigned grid, const StepArgs &A, cudaStream_t st) { CDPR_FLEXR_BY_NF(flexr_go, NC_, L_, grid, A, st); } \
  size_t flexr_smem_##NAME(int nf) {                                                                             \
    return nf == 0 ? FlexRSmem<NC_ / L_, kFlexrTpb, 0, L_>::bytes : nf == 1 ? FlexRSmem<NC_ / L_, kFlexrTpb, 1, L_>::bytes : FlexRSmem<NC_ / L_, kFlexrTpb, 2, L_>::bytes; \
  }
}  // namespace cdpr
