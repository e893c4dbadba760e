// launch.h -- the step kernels are compiled in several translation units (one per group of template instances, so the
// library builds in parallel); each unit lists its instances in a table that api.cu searches by key.
#pragma once
#include <cstddef>
#include <vector>

#include "common.cuh"

namespace cdpr {

// one k_step_fast<NC, 11, MODE, DMOM, SPEC> instance
struct FastEntry {
  int nc, mode;
  bool dmom;
  int spec;
  int tpb;
  size_t smem;
  void (*launch)(unsigned grid, const StepArgs &A, cudaStream_t st);
  const void *func;
};

void fast_entries_nc4_base(std::vector<FastEntry> &out);
void fast_entries_nc4_diag(std::vector<FastEntry> &out);
void fast_entries_nc4_spec(std::vector<FastEntry> &out);
void fast_entries_nc8_base(std::vector<FastEntry> &out);
void fast_entries_nc8_diag(std::vector<FastEntry> &out);
void fast_entries_nc8_spec(std::vector<FastEntry> &out);
void fast_entries_nc8_pair(std::vector<FastEntry> &out);

// k_step_general<DMAX> (step_general.cuh), DMAX in {2, 4}
void general_launch(int dmax, unsigned grid, const StepArgs &A, cudaStream_t st);

// k_step_flex<NC, 32, NF, UNR, LANES> (step_flex.cuh), NC in {4, 8}; nf = biquad stages held on chip per filter (0, 1 or 4);
// unroll = cable-loop unroll factor of the hot body (2 or 4); lanes = threads that share one robot (1, 2, or 4 at 8 cables)
int flex_tpb();
int flex_stage_slots(int p_stages, int d_stages);
int flex_lanes_supported(int nc, int lanes);
size_t flex_smem_bytes(int nc, int nf, int lanes);
void flex_prepare(int nc, int nf, int unroll, int lanes);
void flex_launch(int nc, int nf, int unroll, int lanes, unsigned grid, const StepArgs &A, cudaStream_t st);

// k_step_flexr<NC, 32, NF, HOLD, LANES, ISO> (step_flexr.cuh): the rebuilt form of the flex kernel, NF <= 2, ISO = isotropic body inertia.
// flexr_lanes: lanes per robot of the instance that would run (0 = shape not compiled)
int flexr_lanes(int nc, int nf, int lanes_wanted);
size_t flexr_smem_bytes(int nc, int nf, int lanes);
void flexr_prepare(int nc, int nf, bool hold, int lanes);
void flexr_launch(int nc, int nf, bool hold, int lanes, unsigned grid, const StepArgs &A, cudaStream_t st);

}  // namespace cdpr
