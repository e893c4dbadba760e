// fast_inst.cuh -- included by the fast_nc*_*.cu units: turns a list of template arguments into FastEntry rows.
#pragma once
#include "launch.h"
#include "step_fast.cuh"

namespace cdpr {

template <int NC, int MODE, bool DM, int SP>
static void fast_launch_one(unsigned grid, const StepArgs &A, cudaStream_t st) {
  k_step_fast<NC, 11, MODE, DM, SP><<<grid, FastCfg<NC, SP>::tpb, fast_smem_bytes<NC, 11, SP>(), st>>>(A);
}
template <int NC, int MODE, bool DM, int SP>
static FastEntry fast_entry() {
  return {NC, MODE, DM, SP, FastCfg<NC, SP>::tpb, fast_smem_bytes<NC, 11, SP>(), &fast_launch_one<NC, MODE, DM, SP>,
          (const void *)k_step_fast<NC, 11, MODE, DM, SP>};
}
// the five (mode, D-term form) combinations of one SPEC
template <int NC, int SP>
static void fast_entries_all_modes(std::vector<FastEntry> &out) {
  out.push_back(fast_entry<NC, MODE_FORCE, false, SP>());
  out.push_back(fast_entry<NC, MODE_POSITION, false, SP>());
  out.push_back(fast_entry<NC, MODE_POSITION, true, SP>());
  out.push_back(fast_entry<NC, MODE_VELOCITY, false, SP>());
  out.push_back(fast_entry<NC, MODE_VELOCITY, true, SP>());
}
// the robot-constant specialisations of the velocity / position kernels with the moment D-term
template <int NC>
static void fast_entries_spec(std::vector<FastEntry> &out) {
  out.push_back(fast_entry<NC, MODE_VELOCITY, true, SPEC_DIAG | SPEC_BZ0>());
  out.push_back(fast_entry<NC, MODE_VELOCITY, true, SPEC_DIAG | SPEC_ISO>());
  out.push_back(fast_entry<NC, MODE_VELOCITY, true, SPEC_DIAG | SPEC_ISO | SPEC_BZ0>());
  out.push_back(fast_entry<NC, MODE_POSITION, true, SPEC_DIAG | SPEC_ISO | SPEC_BZ0>());
  out.push_back(fast_entry<NC, MODE_VELOCITY, true, SPEC_DIAG | SPEC_ISO | SPEC_BZ0 | SPEC_NOFF | SPEC_UTGT>());
}
// ... and with the cables in pairs that share a platform anchor (the 8-cable cube)
template <int NC>
static void fast_entries_pair(std::vector<FastEntry> &out) {
  out.push_back(fast_entry<NC, MODE_VELOCITY, true, SPEC_DIAG | SPEC_ISO | SPEC_BZ0 | SPEC_PAIR>());
  out.push_back(fast_entry<NC, MODE_POSITION, true, SPEC_DIAG | SPEC_ISO | SPEC_BZ0 | SPEC_PAIR>());
  out.push_back(fast_entry<NC, MODE_VELOCITY, true, SPEC_DIAG | SPEC_ISO | SPEC_BZ0 | SPEC_PAIR | SPEC_NOFF | SPEC_UTGT>());
}

}  // namespace cdpr
