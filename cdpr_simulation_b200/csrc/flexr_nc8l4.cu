#include "flexr_common.cuh"
namespace cdpr {
CDPR_FLEXR_UNIT(nc8l4, 8, 4)
}
