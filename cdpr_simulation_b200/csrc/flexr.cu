#include "launch.h"
namespace cdpr {
void flexr_prepare_nc8l2(int nf, bool hold);
void flexr_launch_nc8l2(int nf, bool hold, unsigned grid, const StepArgs &A, cudaStream_t st);
size_t flexr_smem_nc8l2(int nf);
void flexr_prepare_nc4l1(int nf, bool hold);
void flexr_launch_nc4l1(int nf, bool hold, unsigned grid, const StepArgs &A, cudaStream_t st);
size_t flexr_smem_nc4l1(int nf);
void flexr_prepare_nc8l1(bool hold);
void flexr_launch_nc8l1(bool hold, unsigned grid, const StepArgs &A, cudaStream_t st);
size_t flexr_smem_nc8l1();

// shapes compiled: 4 cables in one thread; 8 cables in two lanes of 4; 8 cables in one thread without filter slots (tuning
// runs).  Measured at NC=8, 2^20 x 1000 steps, ms for steady / hold transitions / hold + 1 P + 1 D stage: two lanes 80 / 99 / 134,
// one lane 82 / 122 / -, four lanes of 2 cables (16 resident warps at 128 registers, platform update four times) 89 / 128 / 174.
int flexr_lanes(int nc, int nf, int lanes_wanted) {
  if (nf > 2 || (nc != 4 && nc != 8)) return 0;
  if (nc == 4) return 1;
  return (lanes_wanted == 1 && nf == 0) ? 1 : 2;
}
size_t flexr_smem_bytes(int nc, int nf, int lanes) { return nc == 4 ? flexr_smem_nc4l1(nf) : (lanes == 1 ? flexr_smem_nc8l1() : flexr_smem_nc8l2(nf)); }
void flexr_prepare(int nc, int nf, bool hold, int lanes) {
  if (nc == 4) flexr_prepare_nc4l1(nf, hold); else if (lanes == 1) flexr_prepare_nc8l1(hold); else flexr_prepare_nc8l2(nf, hold);
}
void flexr_launch(int nc, int nf, bool hold, int lanes, unsigned grid, const StepArgs &A, cudaStream_t st) {
  if (nc == 4) flexr_launch_nc4l1(nf, hold, grid, A, st); else if (lanes == 1) flexr_launch_nc8l1(hold, grid, A, st); else flexr_launch_nc8l2(nf, hold, grid, A, st);
}
}  // namespace cdpr
