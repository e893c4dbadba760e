#include "flexr_common.cuh"
namespace cdpr {
// eight cables in one thread: the biquad state of eight cables does not fit in registers, so no filter slots here
void flexr_prepare_nc8l1(bool hold) { if (hold) flexr_prep<8, 0, true, 1>(); else flexr_prep<8, 0, false, 1>(); }
void flexr_launch_nc8l1(bool hold, unsigned grid, const StepArgs &A, cudaStream_t st) {
  if (hold) flexr_go<8, 0, true, 1>(grid, A, st); else flexr_go<8, 0, false, 1>(grid, A, st);
}
size_t flexr_smem_nc8l1() { return FlexRSmem<8, kFlexrTpb, 0, 1>::bytes; }
}
