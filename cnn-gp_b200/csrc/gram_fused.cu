// gram_fused.cu -- register-resident fused Gram kernel (placeholder until the kernel lands).
#include "plan.h"

namespace cnngp {

struct FusedPlan { int unused; };

FusedPlan *fused_plan_create(const Plan *) { return nullptr; }
void fused_plan_destroy(FusedPlan *fp) { delete fp; }

int launch_fused_gram(const Plan *, const void *, int64_t, const void *, int64_t, int32_t, const void *,
                      const void *, int32_t, int32_t, int32_t, void *, int64_t, void *) {
    set_error("fused kernel not built");
    return 4;
}

}  // namespace cnngp
