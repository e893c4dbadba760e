#include "launch.h"
#include "step_flex.cuh"
namespace cdpr {
// one warp per block: shared memory is the resource that bounds residency, and 32-thread blocks pack it best
constexpr int kFlexTpb = 32;
template <int NC, int NF, int UNR, int LANES> static void flex_go(unsigned grid, const StepArgs &A, cudaStream_t st) {
  k_step_flex<NC, kFlexTpb, NF, UNR, LANES><<<grid, kFlexTpb, FlexSmem<NC / LANES, kFlexTpb, NF>::bytes, st>>>(A);
}
template <int NC, int NF, int UNR, int LANES> static void flex_prep() {
  const void *f = (const void *)k_step_flex<NC, kFlexTpb, NF, UNR, LANES>;
  cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FlexSmem<NC / LANES, kFlexTpb, NF>::bytes);
  cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
#define CDPR_FLEX_NF(WHAT, NC_, L_)                                                                                  \
  do { if (nf == 0) WHAT<NC_, 0, UNR_, L_>(ARGS); else if (nf == 1) WHAT<NC_, 1, UNR_, L_>(ARGS); else WHAT<NC_, 4, UNR_, L_>(ARGS); } while (0)
#define CDPR_FLEX_DISPATCH(WHAT)                                                                                     \
  do {                                                                                                               \
    if (nc == 4) { if (lanes >= 2) CDPR_FLEX_NF(WHAT, 4, 2); else CDPR_FLEX_NF(WHAT, 4, 1); }                        \
    else { if (lanes >= 4) CDPR_FLEX_NF(WHAT, 8, 4); else if (lanes >= 2) CDPR_FLEX_NF(WHAT, 8, 2); else CDPR_FLEX_NF(WHAT, 8, 1); } \
  } while (0)
#define UNR_ 2
void flex_prepare_u2(int nc, int nf, int lanes) {
#define ARGS
  CDPR_FLEX_DISPATCH(flex_prep);
#undef ARGS
}
void flex_launch_u2(int nc, int nf, int lanes, unsigned grid, const StepArgs &A, cudaStream_t st) {
#define ARGS grid, A, st
  CDPR_FLEX_DISPATCH(flex_go);
#undef ARGS
}
}  // namespace cdpr
