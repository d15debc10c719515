// Fused-forward kernel instances, group A: the BASELINE.json shapes, compiled with exact field counts.
#include "fused_fwd.cuh"
namespace armnet {
extern const FwdInstance kFwdInstancesA[] = {
    ARMNET_FWD_BWD_INSTANCE(39, 1, 10, 1),  // C2a/C2b: Criteo shape, nemb 10
    ARMNET_FWD_BWD_INSTANCE(39, 1, 16, 1),  // C4: Criteo shape, nemb 16
    ARMNET_FWD_BWD_INSTANCE(10, 1, 10, 1),  // C1: Frappe, nemb 10
    ARMNET_FWD_BWD_INSTANCE(22, 1, 28, 4),  // C3: Avazu shape, nemb 100 (4 lanes x 28)
};
extern const int kNumFwdInstancesA = sizeof(kFwdInstancesA) / sizeof(kFwdInstancesA[0]);
}  // namespace armnet
