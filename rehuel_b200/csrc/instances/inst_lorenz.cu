// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_lorenz() { return &dispatch_method<LorenzF>; }
EvalFn evaluator_lorenz() { return &eval_functor<LorenzF, false>; }
} // namespace rehuel
