#include "fast_inst.cuh"
namespace cdpr { void fast_entries_nc4_diag(std::vector<FastEntry> &out) { fast_entries_all_modes<4, SPEC_DIAG>(out); } }
