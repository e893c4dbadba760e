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

// k_step_general<DMAX> (step_general.cuh), DMAX in {2, 4}
void general_launch(int dmax, unsigned grid, const StepArgs &A, cudaStream_t st);

// k_step_flex<NC> (step_flex.cuh), NC in {4, 8}; tpb threads per block, smem bytes of dynamic shared memory
void flex_launch(int nc, unsigned grid, int tpb, size_t smem, const StepArgs &A, cudaStream_t st);
void flex_prepare(int nc, size_t smem);
size_t flex_smem_bytes(int nc, int ps, int ds, int tpb);

}  // namespace cdpr
