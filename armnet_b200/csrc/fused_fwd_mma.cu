// Tensor-core fused-forward kernel instances (fused_fwd_mma.cuh): NT 8-field blocks, EK MMA steps of 8 embedding
// lanes + ER leftover lanes, rows E_STRIDE floats apart in shared memory.
#include "fused_fwd_mma.cuh"
namespace armnet {
extern const MmaInstance kMmaInstances[] = {
    ARMNET_MMA_INSTANCE(5, 1, 2, 12, 0),  // C2a / C2b: 33..40 fields, nemb 10
    ARMNET_MMA_INSTANCE(5, 2, 0, 16, 1),  // C4: 33..40 fields, nemb 16 -- default: 2.0x armnet_fwd_kernel at this shape
};
extern const int kNumMmaInstances = sizeof(kMmaInstances) / sizeof(kMmaInstances[0]);
}  // namespace armnet
