// A USER-defined ODE through the batched irk::odeint: the reference's `dimer` test equation
// (test_equations.hpp:180-203), which is not among the library's built-in functors, with its
// rate as a per-member parameter; compared against the closed-form equilibrium.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -Iinclude examples/user_functor.cu -Lrehuel_b200 -lrehuel_b200
#include <cmath>
#include <cstdio>
#include <vector>

#include "rehuel_b200/irk.cuh"

struct dimer {
	static constexpr int N = 2;
	static constexpr int P = 1;
	__device__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		const double rate = p[0], irate = 1.0 / p[0];
		f[0] = -2 * rate * y[0] * y[0] + 2 * irate * y[1];
		f[1] = rate * y[0] * y[0] - irate * y[1];
	}
	__device__ static void jac(double, const double (&y)[N], const double (&p)[P], double (&J)[N][N])
	{
		const double rate = p[0], irate = 1.0 / p[0];
		J[0][0] = -4 * rate * y[0];
		J[0][1] = 2 * irate;
		J[1][0] = 2 * rate * y[0];
		J[1][1] = -irate;
	}
};

int main()
{
	const std::size_t M = 1000;
	std::vector<double> rate(M);
	for (std::size_t i = 0; i < M; ++i) rate[i] = std::pow(10.0, -1.0 + 3.0 * i / (M - 1.0)); // 0.1 .. 100
	const double y0[2] = {1.0, 0.0};
	irk::solver_options so = irk::default_solver_options();
	newton::options no;
	so.rel_tol = so.abs_tol = 1e-8;
	no.tol = 0.1 * 1e-8;
	so.newton_opts = &no;
	output_options oo;
	irk::ensemble_output sol = irk::odeint(dimer{}, 0.0, 1e4, y0, true, rate.data(), M, so, oo, irk::LOBATTO_IIIC_43);
	double worst = 0.0;
	for (std::size_t i = 0; i < M; ++i) {
		// equilibrium: rate*m^2 = d/rate with m + 2d = 1  =>  2 rate^2 m^2 + m - 1 = 0
		const double r2 = rate[i] * rate[i];
		const double m = (-1.0 + std::sqrt(1.0 + 8.0 * r2)) / (4.0 * r2);
		worst = std::fmax(worst, std::fabs(sol.y(i, 0) - m));
		worst = std::fmax(worst, std::fabs(sol.y(i, 0) + 2.0 * sol.y(i, 1) - 1.0)); // monomer count is conserved
	}
	std::printf("user functor `dimer`: %zu members, status %d, %.1f%% of %llu attempts accepted, max error vs equilibrium %.2e\n",
	            M, sol.status(), 100.0 * sol.accept_frac(), sol.count().attempt, worst);
	return (sol.status() == 0 && worst < 1e-6) ? 0 : 1;
}
