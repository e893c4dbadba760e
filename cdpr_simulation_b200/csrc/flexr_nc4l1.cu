#include "flexr_common.cuh"
namespace cdpr {
CDPR_FLEXR_UNIT(nc4l1, 4, 1)
}
