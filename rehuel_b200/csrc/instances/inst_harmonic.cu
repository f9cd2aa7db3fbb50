// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_harmonic() { return &dispatch_method<HarmonicF>; }
EvalFn evaluator_harmonic() { return &eval_functor<HarmonicF, false>; }
} // namespace rehuel
