#include "launch.h"
namespace cdpr {
void flexr_prepare_nc8l2(int nf, bool hold);
void flexr_launch_nc8l2(int nf, bool hold, unsigned grid, const StepArgs &A, cudaStream_t st);
size_t flexr_smem_nc8l2(int nf);
void flexr_prepare_nc4l1(int nf, bool hold);
void flexr_launch_nc4l1(int nf, bool hold, unsigned grid, const StepArgs &A, cudaStream_t st);
size_t flexr_smem_nc4l1(int nf);
void flexr_prepare_nc8l4(int nf, bool hold);
void flexr_launch_nc8l4(int nf, bool hold, unsigned grid, const StepArgs &A, cudaStream_t st);
size_t flexr_smem_nc8l4(int nf);
void flexr_prepare_nc8l1(bool hold);
void flexr_launch_nc8l1(bool hold, unsigned grid, const StepArgs &A, cudaStream_t st);
size_t flexr_smem_nc8l1();

// shapes compiled: 4 cables in one thread; 8 cables in two lanes of 4; 8 cables in one thread without filter slots
int flexr_lanes(int nc, int nf, int lanes_wanted) {
  if (nf > 1 || (nc != 4 && nc != 8)) return 0;
  if (nc == 4) return 1;
  if (lanes_wanted >= 4) return 4;
  return (lanes_wanted == 1 && nf == 0) ? 1 : 2;
}
size_t flexr_smem_bytes(int nc, int nf, int lanes) { return nc == 4 ? flexr_smem_nc4l1(nf) : (lanes == 1 ? flexr_smem_nc8l1() : lanes == 4 ? flexr_smem_nc8l4(nf) : flexr_smem_nc8l2(nf)); }
void flexr_prepare(int nc, int nf, bool hold, int lanes) {
  if (nc == 4) flexr_prepare_nc4l1(nf, hold); else if (lanes == 1) flexr_prepare_nc8l1(hold); else if (lanes == 4) flexr_prepare_nc8l4(nf, hold); else flexr_prepare_nc8l2(nf, hold);
}
void flexr_launch(int nc, int nf, bool hold, int lanes, unsigned grid, const StepArgs &A, cudaStream_t st) {
  if (nc == 4) flexr_launch_nc4l1(nf, hold, grid, A, st); else if (lanes == 1) flexr_launch_nc8l1(hold, grid, A, st); else if (lanes == 4) flexr_launch_nc8l4(nf, hold, grid, A, st); else flexr_launch_nc8l2(nf, hold, grid, A, st);
}
}  // namespace cdpr
