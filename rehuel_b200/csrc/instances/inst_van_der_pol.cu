// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_van_der_pol() { return &dispatch_method<VanDerPolF>; }
EvalFn evaluator_van_der_pol() { return &eval_functor<VanDerPolF, false>; }
} // namespace rehuel
