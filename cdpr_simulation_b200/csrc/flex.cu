#include "launch.h"
#include "step_flex.cuh"
namespace cdpr {
// one warp per block: shared memory is the resource that bounds residency, and 32-thread blocks pack it best
constexpr int kFlexTpb = 32;
template <int NC, int NF> static const void *flex_func() { return (const void *)k_step_flex<NC, kFlexTpb, NF>; }
template <int NC, int NF> static void flex_go(unsigned grid, const StepArgs &A, cudaStream_t st) {
  k_step_flex<NC, kFlexTpb, NF><<<grid, kFlexTpb, FlexSmem<NC, kFlexTpb, NF>::bytes, st>>>(A);
}
#define CDPR_FLEX_DISPATCH(WHAT)                                                     \
  do {                                                                               \
    if (nc == 4) { if (nf == 0) WHAT(4, 0); else if (nf == 1) WHAT(4, 1); else WHAT(4, 4); } \
    else { if (nf == 0) WHAT(8, 0); else if (nf == 1) WHAT(8, 1); else WHAT(8, 4); }         \
  } while (0)
int flex_tpb() { return kFlexTpb; }
int flex_stage_slots(int ps, int ds) { const int m = ps > ds ? ps : ds; return m == 0 ? 0 : (m == 1 ? 1 : 4); }
size_t flex_smem_bytes(int nc, int nf) {
  size_t b = 0;
#define CDPR_FLEX_BYTES(NC_, NF_) b = FlexSmem<NC_, kFlexTpb, NF_>::bytes
  CDPR_FLEX_DISPATCH(CDPR_FLEX_BYTES);
  return b;
}
void flex_prepare(int nc, int nf) {
  const void *f = nullptr;
#define CDPR_FLEX_FUNC(NC_, NF_) f = flex_func<NC_, NF_>()
  CDPR_FLEX_DISPATCH(CDPR_FLEX_FUNC);
  cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flex_smem_bytes(nc, nf));
  cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
void flex_launch(int nc, int nf, unsigned grid, const StepArgs &A, cudaStream_t st) {
#define CDPR_FLEX_GO(NC_, NF_) flex_go<NC_, NF_>(grid, A, st)
  CDPR_FLEX_DISPATCH(CDPR_FLEX_GO);
}
}  // namespace cdpr
