// Fused-forward kernel instances, group C: generic field buckets, small embeddings (nemb <= 2, <= 4, <= 6, <= 8).
#include "fused_fwd.cuh"
namespace armnet {
#define ARMNET_F_BUCKETS(EC, ES)                                                                              \
    ARMNET_FWD_INSTANCE(4, 0, EC, ES), ARMNET_FWD_INSTANCE(8, 0, EC, ES), ARMNET_FWD_INSTANCE(12, 0, EC, ES), \
        ARMNET_FWD_INSTANCE(16, 0, EC, ES), ARMNET_FWD_INSTANCE(24, 0, EC, ES),                               \
        ARMNET_FWD_INSTANCE(32, 0, EC, ES), ARMNET_FWD_INSTANCE(40, 0, EC, ES),                               \
        ARMNET_FWD_INSTANCE(48, 0, EC, ES), ARMNET_FWD_INSTANCE(64, 0, EC, ES)
extern const FwdInstance kFwdInstancesC[] = {ARMNET_F_BUCKETS(2, 1), ARMNET_F_BUCKETS(4, 1), ARMNET_F_BUCKETS(6, 1), ARMNET_F_BUCKETS(8, 1)};
extern const int kNumFwdInstancesC = sizeof(kFwdInstancesC) / sizeof(kFwdInstancesC[0]);
}  // namespace armnet
