// Fused-forward kernel instances, group D: wide embeddings, a row split over ES lanes (nemb <= 32, 64, 128).
#include "fused_fwd.cuh"
namespace armnet {
#define ARMNET_F_BUCKETS_WIDE(EC, ES)                                                                          \
    ARMNET_FWD_INSTANCE(8, 0, EC, ES), ARMNET_FWD_INSTANCE(16, 0, EC, ES), ARMNET_FWD_INSTANCE(24, 0, EC, ES), \
        ARMNET_FWD_INSTANCE(40, 0, EC, ES), ARMNET_FWD_INSTANCE(64, 0, EC, ES)
extern const FwdInstance kFwdInstancesD[] = {ARMNET_F_BUCKETS_WIDE(16, 2), ARMNET_F_BUCKETS_WIDE(16, 4),
                                             ARMNET_F_BUCKETS_WIDE(16, 8)};
extern const int kNumFwdInstancesD = sizeof(kFwdInstancesD) / sizeof(kFwdInstancesD[0]);
}  // namespace armnet
