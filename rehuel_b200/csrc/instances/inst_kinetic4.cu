// Kernel instantiations for one built-in device functor (own translation unit: the builds run in parallel).
#include "../functors.cuh"
#include "../launch.cuh"

namespace rehuel {
LaunchFn launcher_kinetic4() { return &dispatch_method<Kinetic4F>; }
} // namespace rehuel
