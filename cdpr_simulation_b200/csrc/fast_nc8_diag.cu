#include "fast_inst.cuh"
namespace cdpr { void fast_entries_nc8_diag(std::vector<FastEntry> &out) { fast_entries_all_modes<8, SPEC_DIAG>(out); } }
