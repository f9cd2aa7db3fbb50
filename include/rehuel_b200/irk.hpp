// C++ face of rehuel_b200: the reference's option / output structs and the BATCHED overload of
// irk::odeint, layered on the C ABI (include/rehuel_b200.h).  Plain C++11, no CUDA needed to
// include it; link with -lrehuel_b200.
//
// Mirrors (reference file:line):
//   common_solver_options / irk::solver_options   options.hpp:40-85, irk.hpp:95-124
//   newton::options, newton ret codes             newton.hpp:45-78
//   output_options                                options.hpp:93-133
//   irk::rk_methods, odeint_status_codes          enums.hpp:38-58, 103-127
//   irk::rk_output (+ counters)                   irk.hpp:140-163   -> irk::ensemble_output
//   irk::odeint (8- and 4-argument forms)         irk.hpp:884-926   -> irk::odeint(problem, ...)
//   irk::merge_rk_output                          irk.cpp:750-783   -> irk::merge_rk_output
//   irk::get_coefficients / name maps             irk.cpp:215-728
// The user-functor overload (functor.hpp:35-73 as a __device__ callable) is in irk.cuh.
#ifndef REHUEL_B200_IRK_HPP
#define REHUEL_B200_IRK_HPP

#include <algorithm>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../rehuel_b200.h"

namespace newton {
enum newton_solve_ret_codes { SUCCESS = 0, INCREMENT_DIVERGE, ITERATION_ERROR_TOO_LARGE, MAXIT_EXCEEDED, GENERIC_ERROR };
// newton.hpp:57-78; only tol, dx_delta, maxit and refresh_jac are read by the irk path.
struct options {
	options() : tol(1e-4), dx_delta(1e-4), maxit(5000), time_internals(false), max_step(-1), refresh_jac(10),
	            precondition(false), limit_step(false) {}
	double tol, dx_delta;
	int maxit;
	bool time_internals;
	double max_step;
	int refresh_jac;
	bool precondition, limit_step;
};
} // namespace newton

enum odeint_status_codes {
	SUCCESS = 0, DT_TOO_SMALL = 1, INTERNAL_SOLVE_FEW_ITERS = 2, GENERAL_ERROR = 4, DT_TOO_LARGE = 8,
	INTERNAL_SOLVE_FAILURE = 16, ERROR_LARGER_THAN_RELTOL = 32, ERROR_LARGER_THAN_ABSTOL = 64,
	ERROR_MAX_STEPS_EXCEEDED = 128
};

struct common_solver_options {
	common_solver_options() : internal_solver(1), rel_tol(1e-4), abs_tol(10 * rel_tol), max_dt(0.0), max_steps(-1),
	                          newton_opts(nullptr), out_interval(0), time_internals(false) {}
	int internal_solver;
	double rel_tol, abs_tol, max_dt;
	long long int max_steps;
	const newton::options *newton_opts; // must be set, like in the reference (asserted at irk.hpp:548)
	int out_interval;
	bool time_internals;
};

struct output_options {
	enum output_bits { NO_OUTPUT = 0, STORE_IN_VECTORS = 1, WRITE_TO_FILE = 2 };
	std::size_t output_mode = STORE_IN_VECTORS;
	std::size_t output_interval = 1;
	std::size_t max_samples = 0; // NEW: rows of sample storage per member (0 = final state only)
	// options.hpp:107,119-123: the reference holds an ostream*; across the C ABI the stream is a path
	std::string output_file;
	bool write_to_file() const { return output_mode & 2; }
	void set_output_file(const std::string &path)
	{
		output_file = path;
		output_mode |= 2;
	}
	// NEW: dense output on the uniform grid dense_t0 + k*dense_dt, k < dense_samples (project_b, irk.cpp:731-747)
	std::size_t dense_samples = 0;
	double dense_t0 = 0.0, dense_dt = 0.0;
	bool store_in_vectors() const { return output_mode & 1; }
	void enable_store_in_vectors() { output_mode |= 1; }
	void disable_store_in_vectors() { output_mode &= ~std::size_t(1); }
};

namespace irk {

enum rk_methods {
	IMPLICIT_EULER = 200, RADAU_IIA_32 = 210, RADAU_IIA_53 = 211, RADAU_IIA_95 = 212, RADAU_IIA_137 = 213,
	LOBATTO_IIIA_43 = 220, LOBATTO_IIIA_85 = 221, LOBATTO_IIIA_127 = 222, LOBATTO_IIIC_43 = 230, LOBATTO_IIIC_64 = 231,
	LOBATTO_IIIC_85 = 232, LOBATTO_IIIC_127 = 233, GAUSS_LEGENDRE_42 = 240, GAUSS_LEGENDRE_63 = 241,
	GAUSS_LEGENDRE_105 = 242, GAUSS_LEGENDRE_147 = 243
};

struct solver_options : common_solver_options {
	solver_options() : adaptive_step_size(true), use_newton_iters_adaptive_step(true), verbose_newton(false),
	                   extrapolate_stage(false), keep_state(false), handout_order(REHUEL_ORDER_LOCALITY),
	                   shard_mode(REHUEL_SHARD_DYNAMIC), trace_member(0), result_mask(REHUEL_OUT_ALL) {}
	bool adaptive_step_size, use_newton_iters_adaptive_step, verbose_newton, extrapolate_stage;
	bool keep_state; // NEW: keep the controller state so that irk::odeint_continue can carry on (restore_state, irk.cpp:786)
	int handout_order; // NEW: REHUEL_ORDER_* (scheduling of the batched overload only; results do not depend on it)
	int shard_mode;    // NEW: REHUEL_SHARD_* (how n_gpus > 1 deals members to the devices)
	unsigned long long trace_member; // NEW: the member out_interval / WRITE_TO_FILE report on
	unsigned result_mask;            // NEW: REHUEL_OUT_* -- which per-member arrays come back (default: all of rk_output)
};
inline solver_options default_solver_options() { return solver_options(); }

inline int name_to_method(const std::string &name) { return rehuel_name_to_method(name.c_str()); }
inline const char *method_to_name(int method) { return rehuel_method_to_name(method); }

// Flatten the reference's three option structs into the C ABI's POD.
inline rehuel_options to_c_options(const solver_options &so, const output_options &oo)
{
	if (!so.newton_opts) throw std::invalid_argument("Newton solver options not set!"); // irk.hpp:548
	rehuel_options o;
	rehuel_default_options(&o);
	o.rel_tol = so.rel_tol;
	o.abs_tol = so.abs_tol;
	o.max_dt = so.max_dt;
	o.max_steps = so.max_steps;
	o.adaptive_step_size = so.adaptive_step_size ? 1 : 0;
	o.out_interval = so.out_interval;
	o.newton_tol = so.newton_opts->tol;
	o.newton_dx_delta = so.newton_opts->dx_delta;
	o.newton_maxit = so.newton_opts->maxit;
	o.newton_refresh_jac = so.newton_opts->refresh_jac;
	o.output_mode = (int)oo.output_mode;
	o.output_interval = oo.output_interval;
	o.keep_state = so.keep_state ? 1 : 0;
	o.dense_samples = oo.dense_samples;
	o.dense_t0 = oo.dense_t0;
	o.dense_dt = oo.dense_dt;
	o.handout_order = so.handout_order;
	o.shard_mode = so.shard_mode;
	o.trace_member = so.trace_member;
	o.output_file = oo.write_to_file() ? oo.output_file.c_str() : nullptr; // borrowed: `oo` outlives the call
	o.result_mask = so.result_mask;
	o.time_internals = so.time_internals ? 1 : 0;
	return o;
}

// rk_output for an ensemble: owns a rehuel_result (SoA over members).
class ensemble_output {
public:
	ensemble_output() : r_() {}
	ensemble_output(const ensemble_output &) = delete;
	ensemble_output &operator=(const ensemble_output &) = delete;
	ensemble_output(ensemble_output &&o) noexcept : r_(o.r_) { o.r_ = rehuel_result(); }
	ensemble_output &operator=(ensemble_output &&o) noexcept
	{
		if (this != &o) {
			rehuel_result_free(&r_);
			r_ = o.r_;
			o.r_ = rehuel_result();
		}
		return *this;
	}
	~ensemble_output() { rehuel_result_free(&r_); }

	std::size_t members() const { return r_.M; }
	int neq() const { return r_.n; }
	int status(std::size_t i) const { return r_.status[i]; }
	int status() const // OR over members, like merge_rk_output ORs legs
	{
		int s = 0;
		for (std::size_t i = 0; i < r_.M; ++i) s |= r_.status[i];
		return s;
	}
	double t_final(std::size_t i) const { return r_.t_final[i]; } // needs REHUEL_OUT_T_FINAL (the default)
	// the reference's per-attempt log (irk.hpp:575-577, 805-810) of solver_opts.trace_member, as text
	std::string log(int precision = 6) const
	{
		std::string text(rehuel_result_format_log(&r_, precision, nullptr, 0) + 1, '\0');
		rehuel_result_format_log(&r_, precision, &text[0], text.size());
		text.pop_back();
		return text;
	}
	// print_timing_breakdown (irk.hpp:328-360) for solver_opts.time_internals
	std::string timing_breakdown(const char *method_name) const
	{
		std::string text(rehuel_result_format_timings(&r_, method_name, nullptr, 0) + 1, '\0');
		rehuel_result_format_timings(&r_, method_name, &text[0], text.size());
		text.pop_back();
		return text;
	}
	double y(std::size_t i, int k) const { return r_.y[(std::size_t)k * r_.M + i]; }
	std::vector<double> y(std::size_t i) const
	{
		std::vector<double> v(r_.n);
		for (int k = 0; k < r_.n; ++k) v[k] = y(i, k);
		return v;
	}
	// rk_output::t_vals / y_vals / err of member i (first entry is (t0, y0): irk.hpp:581-585)
	std::size_t n_samples(std::size_t i) const { return r_.n_samples ? std::min<std::size_t>(r_.n_samples[i], r_.n_samples_cap) : 0; }
	double t_vals(std::size_t i, std::size_t s) const { return r_.t_samples[s * r_.M + i]; }
	double y_vals(std::size_t i, std::size_t s, int k) const { return r_.y_samples[(s * r_.n + k) * r_.M + i]; }
	double err(std::size_t i, std::size_t s) const { return r_.err_samples[s * r_.M + i]; }
	// dense output of member i at grid time dense_t0 + g*dense_dt (NaN where the member never got)
	std::size_t n_dense() const { return r_.n_dense; }
	double y_dense(std::size_t i, std::size_t g, int k) const { return r_.y_dense[(g * r_.n + k) * r_.M + i]; }
	double elapsed_time() const { return r_.elapsed_ms; } // ms, my_timer.hpp:142-149
	double kernel_time() const { return r_.kernel_ms; }
	double accept_frac() const { return r_.accept_frac; }
	const rehuel_stats &count() const { return r_.sum; } // ensemble sums of rk_output::counters
	rehuel_result &raw() { return r_; }
	const rehuel_result &raw() const { return r_; }

private:
	rehuel_result r_;
};

inline void check(int rc)
{
	if (rc != REHUEL_OK) throw std::runtime_error(std::string("rehuel_b200: ") + rehuel_last_error());
}

// Batched irk::odeint for a built-in device functor (see REHUEL_PROBLEM_*):
//   y0      [n][M] SoA, or [n] when y0_broadcast
//   params  [p][M] SoA
inline ensemble_output odeint(int problem, double t0, double t1, const double *y0, bool y0_broadcast,
                       const double *params, std::size_t M, solver_options solver_opts,
                       const output_options &output_opts, int method = RADAU_IIA_53, double dt = 1e-6, int n_gpus = 1,
                       int n_hint = 0)
{
	rehuel_options o = to_c_options(solver_opts, output_opts);
	ensemble_output out;
	check(rehuel_irk_ensemble(problem, method, t0, t1, dt, M, n_hint, y0, y0_broadcast ? 1 : 0, params, &o,
	                          output_opts.store_in_vectors() ? output_opts.max_samples : 0, n_gpus, &out.raw()));
	return out;
}

// The "sane default" form (irk.hpp:915-926): RADAU_IIA_53, dt = 1e-6, refresh_jac = 25,
// newton tol = 0.1*min(abs_tol, rel_tol).
inline ensemble_output odeint(int problem, double t0, double t1, const double *y0, bool y0_broadcast,
                              const double *params, std::size_t M)
{
	solver_options s_opts = default_solver_options();
	newton::options n_opts;
	n_opts.refresh_jac = 25;
	n_opts.tol = 0.1 * std::min(s_opts.abs_tol, s_opts.rel_tol);
	s_opts.newton_opts = &n_opts;
	output_options output_opts;
	return odeint(problem, t0, t1, y0, y0_broadcast, params, M, s_opts, output_opts);
}

// What irk::restore_state (irk.cpp:786-803, an empty stub) was meant for: carry an ensemble that ran with
// solver_opts.keep_state on to a later end time, every member from its own (t, y, controller state).
inline ensemble_output odeint_continue(int problem, const ensemble_output &prev, double t1, const double *params,
                                       solver_options solver_opts, const output_options &output_opts,
                                       int method = RADAU_IIA_53, int n_gpus = 1, int n_hint = 0)
{
	rehuel_options o = to_c_options(solver_opts, output_opts);
	ensemble_output out;
	check(rehuel_irk_ensemble_continue(problem, method, t1, n_hint, params, &o,
	                                   output_opts.store_in_vectors() ? output_opts.max_samples : 0, n_gpus, &prev.raw(), &out.raw()));
	return out;
}

// irk::merge_rk_output (irk.cpp:750-783), batched: appends leg 2 to leg 1.
inline ensemble_output merge_rk_output(ensemble_output &&sol1, const ensemble_output &sol2)
{
	check(rehuel_result_merge(&sol1.raw(), &sol2.raw()));
	return std::move(sol1);
}

} // namespace irk

#endif // REHUEL_B200_IRK_HPP
