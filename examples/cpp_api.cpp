// The reference's canonical driver shape (examples/example_equations.cpp:267-312 int_equation,
// 316-338 solve_robertson) against the BATCHED irk::odeint: fill solver_options / newton::options /
// output_options, call odeint, check status, print elapsed time, accept fraction and counters.
//   c++ -std=c++11 -Iinclude examples/cpp_api.cpp -Lrehuel_b200 -lrehuel_b200
#include <cstdio>
#include <vector>

#include "rehuel_b200/irk.hpp"

int main()
{
	const std::size_t M = 2000;
	std::vector<double> params(3 * M);
	for (std::size_t i = 0; i < M; ++i) {
		params[0 * M + i] = 0.04;
		params[1 * M + i] = 1e4 * (0.5 + (double)i / M);
		params[2 * M + i] = 3e7;
	}
	const double y0[3] = {1.0, 0.0, 0.0};
	irk::solver_options s_opts = irk::default_solver_options();
	newton::options n_opts;
	s_opts.newton_opts = &n_opts;           // the user must set this pointer, as in the reference
	s_opts.rel_tol = 1e-5;                  // driver defaults, examples/example_equations.cpp:41-43
	s_opts.abs_tol = 1e-4;
	n_opts.tol = 0.1 * std::min(s_opts.rel_tol, s_opts.abs_tol);
	output_options output_opts;
	output_opts.max_samples = 256;          // keep (up to) 256 samples per member, like store_in_vectors
	int method = irk::name_to_method("RADAU_IIA_53");
	irk::ensemble_output sol = irk::odeint(REHUEL_PROBLEM_ROBERTSON, 0.0, 1e3, y0, true, params.data(), M, s_opts,
	                                       output_opts, method, 1e-2);
	if (sol.status()) {
		std::fprintf(stderr, "Got an error integrating ODE. :/\n");
		return 2;
	}
	std::printf("Solved %zu equations in %g ms.\nAccepted %g%% of the steps\n%llu function evaluations.\n"
	            "%llu Jacobi matrix evaluations.\nmember 0: %zu samples, t_end = %g, y = %g %g %g\n",
	            M, sol.elapsed_time(), 100 * sol.accept_frac(), sol.count().fun_evals, sol.count().jac_evals,
	            sol.n_samples(0), sol.t_final(0), sol.y(0, 0), sol.y(0, 1), sol.y(0, 2));
	// The "sane default" form, irk.hpp:915-926: nothing but the problem and the span (tests/golden/odeint_defaults.json
	// holds what the reference's own 4-argument odeint returns for this system).
	const double nominal[3] = {0.04, 1e4, 3e7};
	irk::ensemble_output dflt = irk::odeint(REHUEL_PROBLEM_ROBERTSON, 0.0, 40.0, y0, true, nominal, 1);
	if (dflt.status()) return 3;
	std::printf("defaults: y = %a %a %a attempts %llu fun_evals %llu\n", dflt.y(0, 0), dflt.y(0, 1), dflt.y(0, 2),
	            dflt.count().attempt, dflt.count().fun_evals);
	return 0;
}
