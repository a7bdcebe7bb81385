// plan_f32.cu -- complex64 instantiation (the fp32 option of BASELINE.json's north_star).
#include "pencil_impl.cuh"

namespace cpc {
PlanBase *make_plan_f32() { return new PlanT<float>(); }
PlanBase *make_pencil_plan_f32(int p_rows, int p_cols) { return new PencilPlanT<float>(p_rows, p_cols); }
}  // namespace cpc
