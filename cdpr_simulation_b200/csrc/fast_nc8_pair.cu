#include "fast_inst.cuh"
namespace cdpr { void fast_entries_nc8_pair(std::vector<FastEntry> &out) { fast_entries_pair<8>(out); } }
