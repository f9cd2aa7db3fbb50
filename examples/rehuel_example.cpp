// rehuel_example -- the reference's example driver (examples/example_equations.cpp) on the GPU library,
// with ensembles.  Same options, same output conventions:
//   * the trajectory goes to stdout (or --output-file) as lines "t y0 y1 ...", irk.hpp:837-843 /
//     example_equations.cpp:195-207;
//   * the performance summary goes to stderr in the reference's words (example_equations.cpp:244-263).
// New here: --ensemble M solves M members at once (parameters swept log-uniformly over --sweep
// decades around the nominal values, the same hash as the benchmark), --member picks the member
// whose trajectory is printed, --gpus N shards the ensemble, --dense K dt prints dense output on a
// uniform grid instead of the accepted steps, --param overrides the nominal parameters (the
// reference asks for them on stdin).
//
// Build: make examples   (c++ -std=c++11 -Iinclude examples/rehuel_example.cpp -Lrehuel_b200 -lrehuel_b200)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "rehuel_b200/irk.hpp"

namespace {

struct equation {
	const char *name;      // --equation value, example_equations.cpp:475-494
	std::vector<double> p; // nominal parameters
	std::vector<double> y0;
	int n_hint;
};

// nominal parameters / initial values of the reference driver (example_equations.cpp:316-450)
const equation kEquations[] = {
    {"exponential", {-0.4}, {1.0}, 0},
    {"stiff-equation", {}, {1.0, 0.0}, 0},
    {"robertson", {0.04, 1e4, 3e7}, {1.0, 0.0, 0.0}, 0},
    {"van-der-pol", {1e-3}, {2.0, 0.0}, 0},
    {"brusselator", {1.0, 2.5}, {1.0, 1.0}, 0},
    {"lorenz", {10.0, 28.0, 8.0 / 3.0}, {1.0, 1.0, 1.0}, 0},
    {"harmonic", {1.0}, {0.0, 1.0}, 0},
    {"dimer", {1.0}, {1.0, 0.0}, 0},
    {"kinetic-4", {0.5, 0.3, 0.2}, {1.0, 0.0, 0.0, 0.0}, 0},
    {"three-body", {1.0, 1.0, 1.0}, {1.0, 0.0, 3.0, 2.0, 0.0, 3.0, 0.0, 2.0, -1.0, 0.0, -1.0, 0.0}, 0}, // example_equations.cpp:341-343
    {"brusselator-1d", {0.02, 1.0, 3.0}, {}, 64},
};

struct user_options {
	std::string eq = "exponential", method = "LOBATTO_IIIC_85", output_fname, stream_fname;
	double t0 = 0.0, t1 = 10.0, dt = 1e-2, rel_tol = 1e-5, abs_tol = 1e-4; // example_equations.cpp:40-43
	int out_interval = 0, time_internals = 0, gpus = 1;
	std::size_t ensemble = 1, member = 0, max_samples = 0, dense = 0; // max_samples 0: chosen from the ensemble size
	double sweep = 0.0, dense_dt = 0.0;
	std::vector<double> params;
};

void print_usage()
{
	std::cerr << "Usage:\n"
	             "./rehuel_example --equation <equation> --method <method>\n"
	             "                 --time-span <t0> <t1> --dt <dt>\n"
	             "                 --rel-tol <rtol> --abs-tol <atol>\n"
	             "                 --out-interval <interval> --time-internals <0/1>\n"
	             "                 --output-file <fname> [--write-to-file <fname>]\n"
	             "                 [--ensemble <M>] [--sweep <decades>] [--member <i>] [--gpus <N>]\n"
	             "                 [--param <value>]... [--dense <K> <dt>] [--max-samples <rows>]\n"
	             "with\n"
	             "\t<equation>: exponential, stiff-equation, robertson, van-der-pol, brusselator, lorenz,\n"
	             "\t            harmonic, dimer, kinetic-4, three-body, brusselator-1d\n"
	             "\t<method>:   IMPLICIT_EULER, RADAU_IIA_32/53/95/137, LOBATTO_IIIC_43/64/85 (default LOBATTO_IIIC_85)\n"
	             "\t<M>:        number of ensemble members; member i gets parameter_j * 10^(sweep*(u_ij - 1/2))\n";
}

// one table instead of an if-chain: option -> number of values it takes
int parse_opts(int argc, char **argv, user_options &u)
{
	const std::map<std::string, int> arity = {
	    {"--equation", 1}, {"--method", 1}, {"--time-span", 2}, {"--dt", 1}, {"--rel-tol", 1}, {"--abs-tol", 1},
	    {"--out-interval", 1}, {"--time-internals", 1}, {"--output-file", 1}, {"--write-to-file", 1}, {"--ensemble", 1}, {"--sweep", 1},
	    {"--member", 1}, {"--gpus", 1}, {"--param", 1}, {"--dense", 2}, {"--max-samples", 1}};
	for (int i = 1; i < argc;) {
		const std::string arg = argv[i];
		if (arg == "-h" || arg == "--help") {
			print_usage();
			return 1;
		}
		const auto it = arity.find(arg);
		if (it == arity.end()) {
			std::cerr << "Unrecognized arg \"" << arg << "\"!\n";
			return -1;
		}
		if (i + it->second >= argc) {
			std::cerr << "Option \"" << arg << "\" needs " << (it->second == 1 ? "a value" : "two values") << "!\n";
			return -1;
		}
		const char *v = argv[i + 1];
		if (arg == "--equation") u.eq = v;
		else if (arg == "--method") u.method = v;
		else if (arg == "--time-span") { u.t0 = std::atof(v); u.t1 = std::atof(argv[i + 2]); }
		else if (arg == "--dt") u.dt = std::atof(v);
		else if (arg == "--rel-tol") u.rel_tol = std::atof(v);
		else if (arg == "--abs-tol") u.abs_tol = std::atof(v);
		else if (arg == "--out-interval") u.out_interval = std::atoi(v);
		else if (arg == "--time-internals") {
			u.time_internals = std::atoi(v);
			if (u.time_internals != 0 && u.time_internals != 1) {
				std::cerr << "Got an ambiguous value for a bool! Please use 0 or 1 only!\n";
				return -2;
			}
		} else if (arg == "--output-file") u.output_fname = v;
		else if (arg == "--write-to-file") u.stream_fname = v; // output_options::WRITE_TO_FILE inside the integrator (irk.hpp:837-843)
		else if (arg == "--ensemble") u.ensemble = (std::size_t)std::atoll(v);
		else if (arg == "--sweep") u.sweep = std::atof(v);
		else if (arg == "--member") u.member = (std::size_t)std::atoll(v);
		else if (arg == "--gpus") u.gpus = std::atoi(v);
		else if (arg == "--param") u.params.push_back(std::atof(v));
		else if (arg == "--dense") { u.dense = (std::size_t)std::atoll(v); u.dense_dt = std::atof(argv[i + 2]); }
		else if (arg == "--max-samples") u.max_samples = (std::size_t)std::atoll(v);
		i += 1 + it->second;
	}
	return 0;
}

// draw k of the splitmix64 stream seeded 0x5EED, in [0,1)  (rehuel_b200/ensembles.py)
double sweep_uniform(std::uint64_t k)
{
	std::uint64_t z = 0x5EEDull + (k + 1) * 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	return (double)(z >> 11) * std::ldexp(1.0, -53);
}

} // namespace

int main(int argc, char **argv)
{
	std::cerr << "Welcome to the Rehuel example program (B200 ensemble edition).\n";
	user_options u;
	if (argc <= 1) {
		std::cerr << "\n";
		print_usage();
		return 3;
	}
	const int parse_status = parse_opts(argc, argv, u);
	if (parse_status < 0) {
		std::cerr << "Error parsing user options!\n";
		return -1;
	} else if (parse_status > 0) {
		return 0;
	}
	const equation *eq = nullptr;
	for (const equation &e : kEquations)
		if (u.eq == e.name) eq = &e;
	const int problem = rehuel_name_to_problem(u.eq.c_str());
	if (!eq || !problem) {
		std::cerr << "Equation \"" << u.eq << "\" not recognized!\n";
		return -1;
	}
	const int method = irk::name_to_method(u.method);
	int n = 0, p = 0;
	if (!method || rehuel_problem_dims(problem, eq->n_hint, &n, &p) != REHUEL_OK || rehuel_get_coefficients(method, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr) == 0) {
		std::cerr << "Method \"" << u.method << "\" not recognized!\n"; // example_equations.cpp:310
		return -1;
	}
	const std::size_t M = u.ensemble ? u.ensemble : 1;
	if (u.member >= M) {
		std::cerr << "--member must be smaller than --ensemble\n";
		return -1;
	}
	std::vector<double> nominal = eq->p;
	for (std::size_t j = 0; j < u.params.size() && j < nominal.size(); ++j) nominal[j] = u.params[j];
	std::vector<double> params((std::size_t)(p ? p : 1) * M, 0.0);
	for (std::size_t i = 0; i < M; ++i)
		for (int j = 0; j < p; ++j)
			params[(std::size_t)j * M + i] = nominal[j] * std::pow(10.0, u.sweep * (sweep_uniform(i * p + j) - 0.5));
	std::vector<double> y0 = eq->y0;
	if (y0.empty()) { // brusselator-1d: u = 1 + sin(2 pi x), v = 3
		const int ng = n / 2;
		y0.resize(n);
		for (int g = 0; g < ng; ++g) {
			y0[2 * g] = 1.0 + std::sin(2.0 * M_PI * (g + 1) / (ng + 1.0));
			y0[2 * g + 1] = 3.0;
		}
	}

	// example_equations.cpp:211-222
	irk::solver_options s_opts = irk::default_solver_options();
	newton::options n_opts;
	s_opts.newton_opts = &n_opts;
	s_opts.rel_tol = u.rel_tol;
	s_opts.abs_tol = u.abs_tol;
	s_opts.out_interval = u.out_interval;
	s_opts.time_internals = u.time_internals != 0;
	s_opts.trace_member = u.member;
	n_opts.tol = 0.1 * std::min(s_opts.rel_tol, s_opts.abs_tol);
	output_options output_opts;
	// rows of sample storage per member: the whole trajectory of a single system; for an ensemble as many rows as 2^24
	// stored states allow (the storage is page-locked host memory: 65 536 rows x 5000 members would be 13 GB)
	std::size_t rows = u.max_samples;
	if (rows == 0) rows = std::min<std::size_t>(std::size_t(1) << 16, std::max<std::size_t>(64, (std::size_t(1) << 24) / M));
	output_opts.max_samples = u.dense ? 0 : rows;
	if (u.dense) {
		output_opts.dense_samples = u.dense;
		output_opts.dense_t0 = u.t0;
		output_opts.dense_dt = u.dense_dt;
	}
	if (!u.stream_fname.empty()) output_opts.set_output_file(u.stream_fname);
	std::ofstream file;
	if (!u.output_fname.empty()) {
		std::cerr << "Writing to file\n";
		file.open(u.output_fname);
	}
	std::ostream &os = u.output_fname.empty() ? std::cout : file;

	try {
		irk::ensemble_output sol = irk::odeint(problem, u.t0, u.t1, y0.data(), true, params.data(), M, s_opts, output_opts,
		                                       method, u.dt, u.gpus, eq->n_hint);
		if (sol.status()) {
			std::cerr << "Got an error integrating ODE. :/\n";
			return 2;
		}
		const std::size_t i = u.member;
		if (u.out_interval > 0) std::cout << sol.log(); // output_opts.log_out is std::cout in the reference (options.hpp:106)
		if (u.time_internals) std::cerr << sol.timing_breakdown(irk::method_to_name(method)); // irk.hpp:866
		if (u.dense) {
			for (std::size_t g = 0; g < sol.n_dense(); ++g) {
				os << u.t0 + (double)g * u.dense_dt;
				for (int k = 0; k < n; ++k) os << " " << sol.y_dense(i, g, k);
				os << "\n";
			}
		} else {
			for (std::size_t s = 0; s < sol.n_samples(i); ++s) { // example_equations.cpp:195-207
				os << sol.t_vals(i, s);
				for (int k = 0; k < n; ++k) os << " " << sol.y_vals(i, s, k);
				os << "\n";
			}
		}
		// example_equations.cpp:244-263
		double elapsed = sol.elapsed_time();
		std::string units = "ms";
		if (elapsed > 1000) {
			elapsed /= 1000;
			units = "s";
		} else if (elapsed < 1e-3) {
			elapsed *= 1000;
			units = "us";
		}
		std::cerr << "Solved " << (M > 1 ? std::to_string(M) + " equations" : std::string("equation")) << " in " << elapsed << " "
		          << units << ".\n";
		std::cerr << "Accepted " << 100 * sol.accept_frac() << "% of the steps\n";
		std::cerr << sol.count().fun_evals << " function evaluations.\n"
		          << sol.count().jac_evals << " Jacobi matrix evaluations.\n";
		if (M > 1)
			std::cerr << "GPU kernel time " << sol.kernel_time() << " ms: " << (double)M / (sol.kernel_time() * 1e-3)
			          << " systems/s.\n";
	} catch (const std::exception &e) {
		std::cerr << e.what() << "\n";
		return 2;
	}
	return 0;
}
