// bvh_lbvh.cu — device LBVH builder (placeholder until the Karras build lands).
#include "rt_internal.h"

namespace rtb {
int build_bvh_lbvh_device(const std::vector<Aabb> &bounds, HostBvh &out, float *ms_device) {
    (void) bounds; (void) out; (void) ms_device;
    return -1;
}
}  // namespace rtb
