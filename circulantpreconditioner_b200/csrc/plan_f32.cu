// plan_f32.cu -- complex64 instantiation (the fp32 option of BASELINE.json's north_star).
#include "plan_impl.cuh"

namespace cpc {
PlanBase *make_plan_f32() { return new PlanT<float>(); }
}  // namespace cpc
