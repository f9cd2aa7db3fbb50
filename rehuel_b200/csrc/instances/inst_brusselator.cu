// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_brusselator() { return &dispatch_method<BrusselatorF>; }
EvalFn evaluator_brusselator() { return &eval_functor<BrusselatorF, false>; }
} // namespace rehuel
