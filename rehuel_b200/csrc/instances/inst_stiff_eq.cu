// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_stiff_eq() { return &dispatch_method<StiffEqF>; }
EvalFn evaluator_stiff_eq() { return &eval_functor<StiffEqF, false>; }
} // namespace rehuel
