// Fused-forward kernel instances, group B: generic field buckets (runtime F <= FP), nemb <= 10, <= 12 and <= 16.
#include "fused_fwd.cuh"
namespace armnet {
#define ARMNET_F_BUCKETS(EC, ES)                                                                              \
    ARMNET_FWD_INSTANCE(4, 0, EC, ES), ARMNET_FWD_INSTANCE(8, 0, EC, ES), ARMNET_FWD_INSTANCE(12, 0, EC, ES), \
        ARMNET_FWD_INSTANCE(16, 0, EC, ES), ARMNET_FWD_INSTANCE(24, 0, EC, ES),                               \
        ARMNET_FWD_INSTANCE(32, 0, EC, ES), ARMNET_FWD_INSTANCE(40, 0, EC, ES),                               \
        ARMNET_FWD_INSTANCE(48, 0, EC, ES), ARMNET_FWD_INSTANCE(64, 0, EC, ES)
extern const FwdInstance kFwdInstancesB[] = {ARMNET_F_BUCKETS(10, 1), ARMNET_F_BUCKETS(12, 1), ARMNET_F_BUCKETS(16, 1)};
extern const int kNumFwdInstancesB = sizeof(kFwdInstancesB) / sizeof(kFwdInstancesB[0]);
}  // namespace armnet
