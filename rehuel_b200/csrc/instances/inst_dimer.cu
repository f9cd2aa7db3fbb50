// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_dimer() { return &dispatch_method<DimerF>; }
EvalFn evaluator_dimer() { return &eval_functor<DimerF, false>; }
} // namespace rehuel
