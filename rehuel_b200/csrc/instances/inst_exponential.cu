// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_exponential() { return &dispatch_method<ExponentialF>; }
EvalFn evaluator_exponential() { return &eval_functor<ExponentialF, false>; }
} // namespace rehuel
