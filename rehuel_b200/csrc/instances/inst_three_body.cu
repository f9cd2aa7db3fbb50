// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_three_body() { return &dispatch_method_cta<ThreeBodyF>; }
EvalFn evaluator_three_body() { return &eval_functor<ThreeBodyF, true>; }
} // namespace rehuel
