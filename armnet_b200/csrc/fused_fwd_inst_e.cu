// Fused-forward kernel instances, group B: generic field buckets (runtime F <= FP), nemb <= 12,
// forward and backward.
#include "fused_fwd.cuh"
namespace armnet {
#define ARMNET_F_BUCKETS(EC, ES)                                                                              \
    ARMNET_FWD_BWD_INSTANCE(4, 0, EC, ES), ARMNET_FWD_BWD_INSTANCE(8, 0, EC, ES),         \
        ARMNET_FWD_BWD_INSTANCE(12, 0, EC, ES), ARMNET_FWD_BWD_INSTANCE(16, 0, EC, ES),   \
        ARMNET_FWD_BWD_INSTANCE(24, 0, EC, ES), ARMNET_FWD_BWD_INSTANCE(32, 0, EC, ES),   \
        ARMNET_FWD_BWD_INSTANCE(40, 0, EC, ES), ARMNET_FWD_BWD_INSTANCE(48, 0, EC, ES),   \
        ARMNET_FWD_BWD_INSTANCE(64, 0, EC, ES)
extern const FwdInstance kFwdInstancesE[] = {ARMNET_F_BUCKETS(12, 1)};
extern const int kNumFwdInstancesE = sizeof(kFwdInstancesE) / sizeof(kFwdInstancesE[0]);
}  // namespace armnet
