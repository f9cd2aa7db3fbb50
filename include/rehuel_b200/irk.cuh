// User-functor overload of the batched irk::odeint.  Include from a .cu file compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a  ... -lrehuel_b200
//
// The reference's functor concept (functor.hpp:35-73: fun(t,y), jac(t,y) with J(i,j) = df_i/dy_j)
// becomes a struct of static __device__ members; per-member parameters replace data members:
//
//   struct my_ode {
//       static constexpr int N = 2;   // equations (1..4 for the thread-per-system kernels)
//       static constexpr int P = 1;   // parameters per member (0 allowed; then p has one unused slot)
//       __device__ static void fun(double t, const double (&y)[N], const double (&p)[P], double (&f)[N]);
//       __device__ static void jac(double t, const double (&y)[N], const double (&p)[P], double (&J)[N][N]);
//   };
//   auto sol = irk::odeint(my_ode{}, t0, t1, y0, /*broadcast*/true, params, M, s_opts, out_opts, irk::RADAU_IIA_53);
//
// Larger systems (4 < N <= 64) are worked on by a whole warp or thread block, so their functor is
// COMPONENTWISE (irk_warp.cuh: banded; irk_wdense.cuh: dense, N <= 16; irk_cta.cuh: dense, N > 16):
//
//   struct my_pde {
//       static constexpr int N = 64, P = 3;
//       static constexpr int kBand = 2;   // optional: J(i,j) == 0 for |i-j| > kBand  ->  warp-per-system band kernels
//       __device__ static double fun(int i, double t, const double *y, const double *p);                 // f_i(t, y)
//       __device__ static void jac_row(int i, double t, const double *y, const double *p, double *row);  // row i of J, pre-zeroed
//       // optional with kBand, faster: band[kBand + o] = J(i, i + o), o = -kBand .. kBand, statically indexed
//       __device__ static void jac_band(int i, double t, const double *y, const double *p, double (&band)[2 * kBand + 1]);
//   };
//
// The kernel template is instantiated HERE, in the user's translation unit; the host pipeline
// (uploads, chunked launches, result download, statistics) is rehuel::ensemble_host in the library.
#ifndef REHUEL_B200_IRK_CUH
#define REHUEL_B200_IRK_CUH

#include "../../rehuel_b200/csrc/launch.cuh"
#include "irk.hpp"

namespace irk {

template <class F, typename = decltype(F::N)>
ensemble_output odeint(F, double t0, double t1, const double *y0, bool y0_broadcast, const double *params,
                       std::size_t M, solver_options solver_opts, const output_options &output_opts,
                       int method = RADAU_IIA_53, double dt = 1e-6, int n_gpus = 1)
{
	rehuel_options o = to_c_options(solver_opts, output_opts);
	ensemble_output out;
	rehuel::LaunchFn launch;
	if constexpr (F::N <= 4) launch = &rehuel::dispatch_method<F>; // thread-per-system
	else launch = &rehuel::dispatch_method_cta<F>;                 // warp- or CTA-per-system
	check(rehuel::ensemble_host(launch, F::N, F::P, method, t0, t1, dt, M, y0, y0_broadcast ? 1 : 0,
	                            params, &o, output_opts.store_in_vectors() ? output_opts.max_samples : 0, n_gpus,
	                            &out.raw()));
	return out;
}

} // namespace irk

#endif // REHUEL_B200_IRK_CUH
