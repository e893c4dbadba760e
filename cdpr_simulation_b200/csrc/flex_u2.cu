#include "launch.h"
#include "step_flex.cuh"
namespace cdpr {
// one warp per block: shared memory is the resource that bounds residency, and 32-thread blocks pack it best
constexpr int kFlexTpb = 32;
template <int NC, int NF, int UNR> static void flex_go(unsigned grid, const StepArgs &A, cudaStream_t st) {
  k_step_flex<NC, kFlexTpb, NF, UNR><<<grid, kFlexTpb, FlexSmem<NC, kFlexTpb, NF>::bytes, st>>>(A);
}
template <int NC, int NF, int UNR> static void flex_prep() {
  const void *f = (const void *)k_step_flex<NC, kFlexTpb, NF, UNR>;
  cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FlexSmem<NC, kFlexTpb, NF>::bytes);
  cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
#define CDPR_FLEX_DISPATCH(WHAT)                                                                             \
  do {                                                                                                       \
    if (nc == 4) { if (nf == 0) WHAT<4, 0, UNR_>(ARGS); else if (nf == 1) WHAT<4, 1, UNR_>(ARGS); else WHAT<4, 4, UNR_>(ARGS); } \
    else { if (nf == 0) WHAT<8, 0, UNR_>(ARGS); else if (nf == 1) WHAT<8, 1, UNR_>(ARGS); else WHAT<8, 4, UNR_>(ARGS); }         \
  } while (0)
#define UNR_ 2
void flex_prepare_u2(int nc, int nf) {
#define ARGS
  CDPR_FLEX_DISPATCH(flex_prep);
#undef ARGS
}
void flex_launch_u2(int nc, int nf, unsigned grid, const StepArgs &A, cudaStream_t st) {
#define ARGS grid, A, st
  CDPR_FLEX_DISPATCH(flex_go);
#undef ARGS
}
}  // namespace cdpr
