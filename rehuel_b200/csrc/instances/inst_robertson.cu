// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_robertson() { return &dispatch_method<RobertsonF>; }
EvalFn evaluator_robertson() { return &eval_functor<RobertsonF, false>; }
} // namespace rehuel
