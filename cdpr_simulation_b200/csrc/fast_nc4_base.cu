#include "fast_inst.cuh"
namespace cdpr { void fast_entries_nc4_base(std::vector<FastEntry> &out) { fast_entries_all_modes<4, 0>(out); } }
