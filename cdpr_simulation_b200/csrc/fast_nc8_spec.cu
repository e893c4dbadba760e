#include "fast_inst.cuh"
namespace cdpr { void fast_entries_nc8_spec(std::vector<FastEntry> &out) { fast_entries_spec<8>(out); } }
