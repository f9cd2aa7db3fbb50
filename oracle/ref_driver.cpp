// TEST INFRASTRUCTURE ONLY (oracle).  Never linked into or called from the product library.
//
// C-callable driver around the UNMODIFIED reference implicit-RK solver.  The reference sources
// are compiled where they lie (/root/reference/{irk.cpp,newton.cpp} + the headers included
// below) against oracle/arma_shim/armadillo; see oracle/Makefile.  Output: oracle/_ref/
// librehuel_ref.so.  Built WITHOUT -DNDEBUG: the LU call sits inside an assert (irk.hpp:432).
//
// What is ours in this file: parameterised user functors written to the reference's functor
// concept (functor.hpp:35-73) -- the reference hard-codes the Robertson rate constants
// (test_equations.hpp:155-157) and has no N=64 Brusselator (test_equations.hpp:546 "TODO") --
// and the loops that call irk::odeint / irk::newton_solve_stages once per ensemble member.
#include <cstring>
#include <sstream>
#include <iostream>
#include <chrono>

#include "irk.hpp"
#include "functor.hpp"
#include "test_equations.hpp"

namespace {

// Robertson with the three rate constants as members.  Expression shapes follow
// test_equations.hpp:153-175 so that (0.04, 1e4, 3e7) reproduces test_equations::rober bit for bit.
struct rober_p : public functor {
	typedef mat_type jac_type;
	double k1, k2, k3;
	rober_p(double a, double b, double c) : k1(a), k2(b), k3(c) {}
	vec_type fun(double, const vec_type &y)
	{
		return { -k1*y[0] + k2 * y[1]*y[2],
		          k1*y[0] - k2 * y[1]*y[2] - k3*y[1]*y[1],
		          k3*y[1]*y[1] };
	}
	jac_type jac(double, const vec_type &y)
	{
		jac_type J(3,3);
		J(0,0) = -k1;
		J(0,1) = k2*y[2];
		J(0,2) = k2*y[1];
		J(1,0) = k1;
		J(1,1) = -k2*y[2] - (2.0*k3)*y[1];
		J(1,2) = -k2*y[1];
		J(2,0) = J(2,2) = 0.0;
		J(2,1) = (2.0*k3)*y[1];
		return J;
	}
};

// 1-D Brusselator, method of lines, Ng interior points, interleaved (u_i, v_i); SURVEY.md §8(d)
// config 4.  Boundary values u=1, v=3.  params = (alpha, A, B).
struct bruss1d : public functor {
	typedef mat_type jac_type;
	int Ng;
	double alpha, A, B;
	bruss1d(int ng, double al, double a, double b) : Ng(ng), alpha(al), A(a), B(b) {}
	vec_type fun(double, const vec_type &y)
	{
		vec_type f(2*Ng);
		const double d = alpha * (Ng + 1.0) * (Ng + 1.0);
		for (int i = 0; i < Ng; ++i) {
			double u = y[2*i], v = y[2*i+1];
			double ul = (i == 0) ? 1.0 : y[2*(i-1)];
			double ur = (i == Ng-1) ? 1.0 : y[2*(i+1)];
			double vl = (i == 0) ? 3.0 : y[2*(i-1)+1];
			double vr = (i == Ng-1) ? 3.0 : y[2*(i+1)+1];
			f[2*i]   = A + u*u*v - (B + 1.0)*u + d*(ul - 2.0*u + ur);
			f[2*i+1] = B*u - u*u*v + d*(vl - 2.0*v + vr);
		}
		return f;
	}
	jac_type jac(double, const vec_type &y)
	{
		jac_type J = arma::zeros(2*Ng, 2*Ng);
		const double d = alpha * (Ng + 1.0) * (Ng + 1.0);
		for (int i = 0; i < Ng; ++i) {
			double u = y[2*i], v = y[2*i+1];
			J(2*i, 2*i)     = 2.0*u*v - (B + 1.0) - 2.0*d;
			J(2*i, 2*i+1)   = u*u;
			J(2*i+1, 2*i)   = B - 2.0*u*v;
			J(2*i+1, 2*i+1) = -u*u - 2.0*d;
			if (i > 0)      { J(2*i, 2*(i-1)) = d; J(2*i+1, 2*(i-1)+1) = d; }
			if (i < Ng - 1) { J(2*i, 2*(i+1)) = d; J(2*i+1, 2*(i+1)+1) = d; }
		}
		return J;
	}
};

struct cout_silencer {
	std::ostringstream sink; // declared first: must be alive before cout is pointed at it
	std::streambuf *old;
	std::streamsize old_prec;
	cout_silencer() : sink(), old(std::cout.rdbuf(sink.rdbuf())), old_prec(std::cout.precision(17)) {}
	~cout_silencer() { std::cout.rdbuf(old); std::cout.precision(old_prec); }
};

} // namespace

extern "C" {

// Mirrors include/rehuel_b200.h: rehuel_options (kept as a private copy: the oracle must not
// depend on product headers).
struct ref_options {
	double rel_tol, abs_tol, max_dt;
	long long max_steps;
	int adaptive_step_size;
	int out_interval;
	double newton_tol, newton_dx_delta;
	int newton_maxit, newton_refresh_jac;
	int output_mode;
	unsigned long long output_interval;
};

enum { REF_ROBERTSON = 1, REF_VDPOL = 2, REF_BRUSS1D = 3, REF_ROBER_HARDCODED = 4,
       REF_EXPONENTIAL = 5, REF_STIFF_EQ = 6, REF_BRUSS2 = 7, REF_LORENZ = 8, REF_HARMONIC = 9, REF_DIMER = 10,
       REF_KINETIC4 = 11, REF_THREE_BODY = 12 };

int ref_problem_dims(int problem, int n_hint, int *n, int *p)
{
	switch (problem) {
	case REF_ROBERTSON: *n = 3; *p = 3; return 0;
	case REF_VDPOL: *n = 2; *p = 1; return 0;
	case REF_BRUSS1D: *n = n_hint; *p = 3; return 0;
	case REF_ROBER_HARDCODED: *n = 3; *p = 0; return 0;
	case REF_EXPONENTIAL: *n = 1; *p = 1; return 0;
	case REF_STIFF_EQ: *n = 2; *p = 0; return 0;
	case REF_BRUSS2: *n = 2; *p = 2; return 0;
	case REF_LORENZ: *n = 3; *p = 3; return 0;
	case REF_HARMONIC: *n = 2; *p = 1; return 0;
	case REF_DIMER: *n = 2; *p = 1; return 0;
	case REF_KINETIC4: *n = 4; *p = 3; return 0;
	case REF_THREE_BODY: *n = 12; *p = 3; return 0;
	}
	return -1;
}

// Tableau of the reference (irk.cpp:215-689) plus the derived weights of irk.hpp:599-601.
// A is returned row-major s*s.  Returns the number of stages, 0 if the method is unknown.
int ref_get_coefficients(int method, double *A, double *b, double *c, double *b2,
                         double *gamma, int *order, int *order2, double *d, double *d2)
{
	std::streambuf *olderr = std::cerr.rdbuf();
	std::ostringstream sink;
	std::cerr.rdbuf(sink.rdbuf());
	irk::solver_coeffs sc = irk::get_coefficients(method);
	std::cerr.rdbuf(olderr);
	if (!irk::verify_solver_coeffs(sc)) return 0;
	int s = (int)sc.b.size();
	for (int i = 0; i < s; ++i) {
		for (int j = 0; j < s; ++j) A[i*s + j] = sc.A(i, j);
		b[i] = sc.b(i);
		c[i] = sc.c(i);
		b2[i] = sc.b2.size() ? sc.b2(i) : 0.0;
	}
	*gamma = sc.gamma;
	*order = sc.order;
	*order2 = sc.order2;
	try {
		mat_type Ai = arma::inv(sc.A);
		vec_type dw = (Ai.t())*sc.b;
		for (int i = 0; i < s; ++i) d[i] = dw(i);
		if (sc.b2.size()) {
			vec_type dw2 = (Ai.t())*sc.b2;
			for (int i = 0; i < s; ++i) d2[i] = dw2(i);
		}
	} catch (std::exception &) {
		for (int i = 0; i < s; ++i) d[i] = d2[i] = NAN;
	}
	return s;
}

int ref_name_to_method(const char *name) { return irk::name_to_method(name); }
const char *ref_method_to_name(int m) { return irk::method_to_name(m); }

// irk::project_b (irk.cpp:731-747) for the dense-output tests.
int ref_project_b(int method, double theta, double *out)
{
	irk::solver_coeffs sc = irk::get_coefficients(method);
	if (sc.b_interp.size() == 0) return 0;
	vec_type bs = irk::project_b(theta, sc);
	for (std::size_t i = 0; i < bs.size(); ++i) out[i] = bs(i);
	return (int)bs.size();
}

} // extern "C" (templates cannot have C linkage)

template <typename F>
static void run_member(F &func, int method, double t0, double t1, double dt0, const vec_type &y0,
                       const ref_options *o, std::size_t i, std::size_t M, int n,
                       double *y, double *t_final, double *err_last, int *status,
                       unsigned long long *counters, double *accept_frac,
                       // optional single-trajectory capture:
                       std::size_t traj_cap, std::size_t *traj_len, double *traj_t, double *traj_y,
                       double *traj_err)
{
	irk::solver_options so = irk::default_solver_options();
	newton::options no;
	so.rel_tol = o->rel_tol;
	so.abs_tol = o->abs_tol;
	so.max_dt = o->max_dt;
	so.max_steps = o->max_steps;
	so.adaptive_step_size = o->adaptive_step_size != 0;
	so.out_interval = o->out_interval;
	no.tol = o->newton_tol;
	no.dx_delta = o->newton_dx_delta;
	no.maxit = o->newton_maxit;
	no.refresh_jac = o->newton_refresh_jac;
	so.newton_opts = &no;
	output_options oo;
	oo.output_mode = o->output_mode;
	oo.output_interval = o->output_interval;

	irk::rk_output sol = irk::odeint(func, t0, t1, y0, so, oo, method, dt0);

	std::size_t ns = sol.t_vals.size();
	status[i] = sol.status;
	if (ns) {
		t_final[i] = sol.t_vals[ns-1];
		for (int k = 0; k < n; ++k) y[(std::size_t)k*M + i] = sol.y_vals[ns-1](k);
		err_last[i] = sol.err[ns-1];
	}
	counters[0*M+i] = sol.count.attempt;
	counters[1*M+i] = sol.count.reject_newton;
	counters[2*M+i] = sol.count.reject_err;
	counters[3*M+i] = sol.count.newton_success;
	counters[4*M+i] = sol.count.newton_incr_diverge;
	counters[5*M+i] = sol.count.newton_iter_error_too_large;
	counters[6*M+i] = sol.count.newton_maxit_exceed;
	counters[7*M+i] = sol.count.fun_evals;
	counters[8*M+i] = sol.count.jac_evals;
	accept_frac[i] = sol.accept_frac;
	if (traj_cap && traj_len) {
		std::size_t L = ns < traj_cap ? ns : traj_cap;
		*traj_len = ns;
		for (std::size_t s = 0; s < L; ++s) {
			traj_t[s] = sol.t_vals[s];
			traj_err[s] = sol.err[s];
			for (int k = 0; k < n; ++k) traj_y[s*n + k] = sol.y_vals[s](k);
		}
	}
}

extern "C" {

// Ensemble = a loop over independent irk::odeint calls (irk.hpp:884-898), one per member.
//   y0      [n][M] (SoA) or [n] when y0_broadcast
//   params  [P][M] (SoA)
//   members i = start, start+stride, ... (< M) are solved; other entries are left untouched.
//   outputs y[n][M], t_final[M], err_last[M], status[M], counters[9][M], accept_frac[M].
// Output mode must have STORE_IN_VECTORS set (final state is read from sol.y_vals.back()); use a
// huge output_interval to keep only (t0,y0) plus nothing else -> then the final state is not
// available, so the driver forces output_interval=1 unless a trajectory is requested.
// Returns 0, or -1 on bad arguments, -2 if the reference threw (what() in ref_last_error()).
static thread_local std::string g_err;
const char *ref_last_error(void) { return g_err.c_str(); }

int ref_irk_ensemble(int problem, int method, double t0, double t1, double dt0,
                     std::size_t M, int n_hint, const double *y0, int y0_broadcast,
                     const double *params, const ref_options *opts,
                     std::size_t start, std::size_t stride,
                     double *y, double *t_final, double *err_last, int *status,
                     unsigned long long *counters, double *accept_frac, double *elapsed_ms,
                     std::size_t traj_cap, std::size_t *traj_len, double *traj_t, double *traj_y,
                     double *traj_err, char *log_buf, std::size_t log_cap)
{
	int n = 0, P = 0;
	if (ref_problem_dims(problem, n_hint, &n, &P) || !opts || stride == 0) return -1;
	cout_silencer quiet;
	auto tic = std::chrono::steady_clock::now();
	try {
		for (std::size_t i = start; i < M; i += stride) {
			vec_type Y0(n);
			for (int k = 0; k < n; ++k) Y0(k) = y0_broadcast ? y0[k] : y0[(std::size_t)k*M + i];
			auto par = [&](int j) { return params[(std::size_t)j*M + i]; };
#define RUN(F) run_member(F, method, t0, t1, dt0, Y0, opts, i, M, n, y, t_final, err_last, status, \
                          counters, accept_frac, traj_cap, traj_len, traj_t, traj_y, traj_err)
			switch (problem) {
			case REF_ROBERTSON: { rober_p f(par(0), par(1), par(2)); RUN(f); break; }
			case REF_VDPOL: { test_equations::vdpol f(par(0)); RUN(f); break; }
			case REF_BRUSS1D: { bruss1d f(n/2, par(0), par(1), par(2)); RUN(f); break; }
			case REF_ROBER_HARDCODED: { test_equations::rober f; RUN(f); break; }
			case REF_EXPONENTIAL: { test_equations::exponential f(par(0)); RUN(f); break; }
			case REF_STIFF_EQ: { test_equations::stiff_eq f; RUN(f); break; }
			case REF_BRUSS2: { test_equations::bruss f(par(0), par(1)); RUN(f); break; }
			case REF_LORENZ: { test_equations::lorenz f(par(0), par(1), par(2)); RUN(f); break; }
			case REF_HARMONIC: { test_equations::harmonic f(par(0)); RUN(f); break; }
			case REF_DIMER: { test_equations::dimer f(par(0)); RUN(f); break; }
			case REF_KINETIC4: { test_equations::kinetic_4 f(par(0), par(1), par(2)); RUN(f); break; }
			case REF_THREE_BODY: { test_equations::three_body f(par(0), par(1), par(2)); RUN(f); break; }
			}
#undef RUN
		}
	} catch (std::exception &e) {
		g_err = e.what();
		return -2;
	}
	auto toc = std::chrono::steady_clock::now();
	if (elapsed_ms) *elapsed_ms = std::chrono::duration<double, std::milli>(toc - tic).count();
	if (log_buf && log_cap) {
		std::string s = quiet.sink.str();
		std::size_t L = s.size() < log_cap - 1 ? s.size() : log_cap - 1;
		std::memcpy(log_buf, s.data(), L);
		log_buf[L] = 0;
	}
	return 0;
}

// The reference's 4-argument "sane default" irk::odeint(func, t0, t1, y0) (irk.hpp:915-926) for ONE system: whatever
// defaults it picks (method, first step, tolerances, Newton options) are the reference's, none are passed in.
int ref_odeint_defaults(int problem, int n_hint, double t0, double t1, const double *y0, const double *params,
                        double *y, double *t_final, double *err_last, int *status, unsigned long long *counters)
{
	int n = 0, P = 0;
	if (ref_problem_dims(problem, n_hint, &n, &P)) return -1;
	cout_silencer quiet;
	vec_type Y0(n);
	for (int k = 0; k < n; ++k) Y0(k) = y0[k];
	irk::rk_output sol;
	try {
		switch (problem) {
		case REF_ROBERTSON: { rober_p f(params[0], params[1], params[2]); sol = irk::odeint(f, t0, t1, Y0); break; }
		case REF_VDPOL: { test_equations::vdpol f(params[0]); sol = irk::odeint(f, t0, t1, Y0); break; }
		case REF_ROBER_HARDCODED: { test_equations::rober f; sol = irk::odeint(f, t0, t1, Y0); break; }
		case REF_EXPONENTIAL: { test_equations::exponential f(params[0]); sol = irk::odeint(f, t0, t1, Y0); break; }
		case REF_BRUSS2: { test_equations::bruss f(params[0], params[1]); sol = irk::odeint(f, t0, t1, Y0); break; }
		default: return -1;
		}
	} catch (std::exception &e) {
		g_err = e.what();
		return -2;
	}
	const std::size_t ns = sol.t_vals.size();
	if (!ns) return -3;
	*status = sol.status;
	*t_final = sol.t_vals[ns-1];
	*err_last = sol.err[ns-1];
	for (int k = 0; k < n; ++k) y[k] = sol.y_vals[ns-1](k);
	const unsigned long long c[9] = {sol.count.attempt, sol.count.reject_newton, sol.count.reject_err, sol.count.newton_success,
	                                 sol.count.newton_incr_diverge, sol.count.newton_iter_error_too_large,
	                                 sol.count.newton_maxit_exceed, sol.count.fun_evals, sol.count.jac_evals};
	for (int k = 0; k < 9; ++k) counters[k] = c[k];
	return 0;
}

// One call of irk::newton_solve_stages (irk.hpp:398-493) + the solution update of
// irk.hpp:695-705, i.e. exactly what the reference's own known-answer test does
// (test/irk.cpp:250-296).  variant 0: default template arguments <F,true,false> (damped,
// dense solve) as in the reference test; variant 1: <F,false,true> as called by irk_guts.
int ref_newton_solve_stages(int problem, int method, int variant, int n_hint, const double *params,
                            const double *y_in, double t, double dt, int maxit, int refresh_jac,
                            double xtol, double Rtol,
                            double *Y_out, double *y1_out, int *iters, double *res,
                            unsigned long long *fun_evals, unsigned long long *jac_evals)
{
	int n = 0, P = 0;
	if (ref_problem_dims(problem, n_hint, &n, &P)) return -1;
	irk::solver_coeffs sc = irk::get_coefficients(method);
	vec_type y(n);
	for (int k = 0; k < n; ++k) y(k) = y_in[k];
	std::size_t Ns = sc.b.size(), NN = Ns * n;
	arma::mat J(n, n);
	vec_type Y(NN);
	newton::status stats;
	std::size_t fc = 0, jc = 0;
	int status = -1;
	try {
#define CALL(F) status = variant == 0                                                              \
	? irk::newton_solve_stages(F, y, t, dt, sc, maxit, refresh_jac, xtol, Rtol, Y, J, stats, fc, jc)  \
	: irk::newton_solve_stages<decltype(F), false, true>(F, y, t, dt, sc, maxit, refresh_jac, xtol,  \
	                                                     Rtol, Y, J, stats, fc, jc)
		switch (problem) {
		case REF_ROBERTSON: { rober_p f(params[0], params[1], params[2]); CALL(f); break; }
		case REF_VDPOL: { test_equations::vdpol f(params[0]); CALL(f); break; }
		case REF_BRUSS1D: { bruss1d f(n/2, params[0], params[1], params[2]); CALL(f); break; }
		case REF_ROBER_HARDCODED: { test_equations::rober f; CALL(f); break; }
		case REF_EXPONENTIAL: { test_equations::exponential f(params[0]); CALL(f); break; }
		case REF_STIFF_EQ: { test_equations::stiff_eq f; CALL(f); break; }
		case REF_BRUSS2: { test_equations::bruss f(params[0], params[1]); CALL(f); break; }
		case REF_LORENZ: { test_equations::lorenz f(params[0], params[1], params[2]); CALL(f); break; }
		case REF_HARMONIC: { test_equations::harmonic f(params[0]); CALL(f); break; }
		case REF_DIMER: { test_equations::dimer f(params[0]); CALL(f); break; }
		case REF_KINETIC4: { test_equations::kinetic_4 f(params[0], params[1], params[2]); CALL(f); break; }
		case REF_THREE_BODY: { test_equations::three_body f(params[0], params[1], params[2]); CALL(f); break; }
		}
#undef CALL
	} catch (std::exception &e) {
		g_err = e.what();
		return -2;
	}
	arma::mat YYs = arma::reshape(Y, n, Ns);
	mat_type Ai = arma::inv(sc.A);
	vec_type d_weights = (Ai.t())*sc.b;
	arma::vec y1 = y + YYs*d_weights;
	for (std::size_t k = 0; k < NN; ++k) Y_out[k] = Y(k);
	for (int k = 0; k < n; ++k) y1_out[k] = y1(k);
	*iters = stats.iters;
	*res = stats.res;
	*fun_evals = fc;
	*jac_evals = jc;
	return status;
}

// Sanctioned fixed-step oracle: the reference's own fixed-step irk_guts path throws
// (delta_alt is never assigned, irk.hpp:698-702), so fixed-step parity is defined as a loop over
// irk::newton_solve_stages<F,false,true> + y += reshape(Y,n,s)*d_weights, exactly what
// test/irk.cpp:271-286 does for one step.  Returns the number of failed Newton solves.
int ref_fixed_step(int problem, int method, int n_hint, const double *params, const double *y_in,
                   double t0, double dt, long long nsteps, int maxit, int refresh_jac, double xtol,
                   double Rtol, double *y_out, unsigned long long *newton_iters)
{
	int n = 0, P = 0;
	if (ref_problem_dims(problem, n_hint, &n, &P)) return -1;
	irk::solver_coeffs sc = irk::get_coefficients(method);
	std::size_t Ns = sc.b.size(), NN = Ns * n;
	mat_type Ai = arma::inv(sc.A);
	vec_type d_weights = (Ai.t())*sc.b;
	vec_type y(n);
	for (int k = 0; k < n; ++k) y(k) = y_in[k];
	int failed = 0;
	*newton_iters = 0;
	double t = t0;
	try {
		for (long long st = 0; st < nsteps; ++st) {
			arma::mat J(n, n);
			vec_type Y(NN);
			newton::status stats;
			std::size_t fc = 0, jc = 0;
			int status = -1;
#define CALL(F) status = irk::newton_solve_stages<decltype(F), false, true>(F, y, t, dt, sc, maxit, \
	refresh_jac, xtol, Rtol, Y, J, stats, fc, jc)
			switch (problem) {
			case REF_ROBERTSON: { rober_p f(params[0], params[1], params[2]); CALL(f); break; }
			case REF_VDPOL: { test_equations::vdpol f(params[0]); CALL(f); break; }
			case REF_BRUSS1D: { bruss1d f(n/2, params[0], params[1], params[2]); CALL(f); break; }
			case REF_ROBER_HARDCODED: { test_equations::rober f; CALL(f); break; }
			case REF_EXPONENTIAL: { test_equations::exponential f(params[0]); CALL(f); break; }
			case REF_STIFF_EQ: { test_equations::stiff_eq f; CALL(f); break; }
			case REF_BRUSS2: { test_equations::bruss f(params[0], params[1]); CALL(f); break; }
			case REF_LORENZ: { test_equations::lorenz f(params[0], params[1], params[2]); CALL(f); break; }
			case REF_HARMONIC: { test_equations::harmonic f(params[0]); CALL(f); break; }
			case REF_DIMER: { test_equations::dimer f(params[0]); CALL(f); break; }
			case REF_KINETIC4: { test_equations::kinetic_4 f(params[0], params[1], params[2]); CALL(f); break; }
			case REF_THREE_BODY: { test_equations::three_body f(params[0], params[1], params[2]); CALL(f); break; }
			}
#undef CALL
			if (status != newton::SUCCESS) ++failed;
			*newton_iters += stats.iters;
			arma::mat YYs = arma::reshape(Y, n, Ns);
			y = y + YYs*d_weights;
			t += dt;
		}
	} catch (std::exception &e) {
		g_err = e.what();
		return -2;
	}
	for (int k = 0; k < n; ++k) y_out[k] = y(k);
	return failed;
}

// fun(t,y) and jac(t,y) of a reference test functor as they stand (functor.hpp:35-73); J is returned row-major,
// J[i*n + j] = jac(i,j).  Lets the device functors be checked against the reference's own expressions -- the only
// meaningful check for three_body, whose jac() is not the derivative of its fun().
int ref_eval_functor(int problem, int n_hint, const double *params, double t, const double *y_in, double *f_out,
                     double *J_out)
{
	int n = 0, P = 0;
	if (ref_problem_dims(problem, n_hint, &n, &P)) return -1;
	vec_type y(n);
	for (int k = 0; k < n; ++k) y(k) = y_in[k];
	vec_type f;
	arma::mat J;
	try {
#define CALL(F) do { f = F.fun(t, y); J = F.jac(t, y); } while (0)
		switch (problem) {
		case REF_ROBERTSON: { rober_p fn(params[0], params[1], params[2]); CALL(fn); break; }
		case REF_VDPOL: { test_equations::vdpol fn(params[0]); CALL(fn); break; }
		case REF_BRUSS1D: { bruss1d fn(n/2, params[0], params[1], params[2]); CALL(fn); break; }
		case REF_ROBER_HARDCODED: { test_equations::rober fn; CALL(fn); break; }
		case REF_EXPONENTIAL: { test_equations::exponential fn(params[0]); CALL(fn); break; }
		case REF_STIFF_EQ: { test_equations::stiff_eq fn; CALL(fn); break; }
		case REF_BRUSS2: { test_equations::bruss fn(params[0], params[1]); CALL(fn); break; }
		case REF_LORENZ: { test_equations::lorenz fn(params[0], params[1], params[2]); CALL(fn); break; }
		case REF_HARMONIC: { test_equations::harmonic fn(params[0]); CALL(fn); break; }
		case REF_DIMER: { test_equations::dimer fn(params[0]); CALL(fn); break; }
		case REF_KINETIC4: { test_equations::kinetic_4 fn(params[0], params[1], params[2]); CALL(fn); break; }
		case REF_THREE_BODY: { test_equations::three_body fn(params[0], params[1], params[2]); CALL(fn); break; }
		}
#undef CALL
	} catch (std::exception &e) {
		g_err = e.what();
		return -2;
	}
	for (int i = 0; i < n; ++i) {
		f_out[i] = f(i);
		for (int j = 0; j < n; ++j) J_out[i*n + j] = J(i, j);
	}
	return 0;
}

// irk::merge_rk_output (irk.cpp:750-783) on two synthetic legs: returns merged counters and
// accept_frac so that the product's batched merge can be compared (test/irk.cpp:185-216).
int ref_merge_counts(const double *t1v, std::size_t n1, double af1, const unsigned long long *c1,
                     const double *t2v, std::size_t n2, double af2, const unsigned long long *c2,
                     int st1, int st2, double *t_out, double *af_out, unsigned long long *c_out,
                     int *st_out)
{
	irk::rk_output a, b;
	a.status = st1; b.status = st2;
	a.t_vals.assign(t1v, t1v + n1);
	b.t_vals.assign(t2v, t2v + n2);
	a.accept_frac = af1; b.accept_frac = af2;
	a.elapsed_time = b.elapsed_time = 0.0;
	auto fill = [](irk::rk_output &s, const unsigned long long *c) {
		s.count.attempt = c[0]; s.count.reject_newton = c[1]; s.count.reject_err = c[2];
		s.count.newton_success = c[3]; s.count.newton_incr_diverge = c[4];
		s.count.newton_iter_error_too_large = c[5]; s.count.newton_maxit_exceed = c[6];
		s.count.fun_evals = c[7]; s.count.jac_evals = c[8];
	};
	fill(a, c1); fill(b, c2);
	irk::rk_output m = irk::merge_rk_output(a, b);
	for (std::size_t i = 0; i < m.t_vals.size(); ++i) t_out[i] = m.t_vals[i];
	*af_out = m.accept_frac;
	c_out[0] = m.count.attempt; c_out[1] = m.count.reject_newton; c_out[2] = m.count.reject_err;
	c_out[3] = m.count.newton_success; c_out[4] = m.count.newton_incr_diverge;
	c_out[5] = m.count.newton_iter_error_too_large; c_out[6] = m.count.newton_maxit_exceed;
	c_out[7] = m.count.fun_evals; c_out[8] = m.count.jac_evals;
	*st_out = m.status;
	return (int)m.t_vals.size();
}

} // extern "C"
