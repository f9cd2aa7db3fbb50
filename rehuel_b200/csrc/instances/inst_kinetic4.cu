// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_kinetic4() { return &dispatch_method<Kinetic4F>; }
EvalFn evaluator_kinetic4() { return &eval_functor<Kinetic4F, false>; }
} // namespace rehuel
