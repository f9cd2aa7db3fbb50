// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_bruss1d_12() { return &dispatch_method_cta<Brusselator1dF<6>>; }
EvalFn evaluator_bruss1d_12() { return &eval_functor<Brusselator1dF<6>, true>; }
} // namespace rehuel
