#include "fast_inst.cuh"
namespace cdpr { void fast_entries_nc4_spec(std::vector<FastEntry> &out) { fast_entries_spec<4>(out); } }
