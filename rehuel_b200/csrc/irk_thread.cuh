// Thread-per-system ensemble integrator for small systems (n <= 4): one CUDA thread runs one
// member's whole adaptive integration, lanes refill from a global work queue when their member
// finishes (persistent warps, warp-aggregated atomics).
//
// One loop trip of a lane = one ATTEMPT of the reference time loop irk.hpp:604-860:
//   newton_solve_stages<F,false,true>  irk.hpp:398-493   -> newton_solve()
//   construct_R                         irk.hpp:366-386   -> residual()
//   solution update                     irk.hpp:691-707
//   embedded error estimate (8.19/8.20) irk.hpp:713-731
//   error norm                          irk.hpp:733-758
//   accept / reject                     irk.hpp:761-766, 813-846
//   step-size controller                irk.hpp:770-801, 850-855
//   Newton-failure policy               irk.hpp:634-665
//
// The one structural difference from the reference: the (s*n)x(s*n) system
//   (I - dt*kron(A,J)) dY = -R                                   (irk.hpp:426-432, 461-462)
// is not factorised densely.  With inv(A) = T Lambda T^-1 (tableau_eig.inc) it decouples into
//   (mu_r I - J) W_r = -mu_r Z_r,            mu_r = gamma_hat/dt           (one real  n x n LU)
//   (mu_c I - J) w_c = -mu_c z_c,            mu_c = (alpha+i*beta)/dt      (one complex n x n LU per pair)
// with Z = (T^-1 (x) I) R and dY = (T (x) I) W.  Same linear system, same Newton iterates up to
// rounding; 27 instead of 81 doubles of factor storage for n = s = 3, so everything lives in
// registers.  All LUs are row-partial-pivoted like LAPACK's (predicated swaps, static indexing).
#pragma once

#ifdef REHUEL_PHASE_CLOCKS
#include <cstdio>
#endif
#include <cuda_runtime.h>
#include <math.h>

#include <type_traits>

#include "../../include/rehuel_b200.h"
#include "tableau.hpp"

namespace rehuel {

struct KernelArgs {
	double t0, t1, dt0;
	double rtol, atol, max_dt;
	long long max_steps;
	int adaptive;
	double xtol2, Rtol2; // dx_delta^2, tol^2 (irk.hpp:441-442)
	int maxit0, refresh_jac;
	int store_samples;
	unsigned long long output_interval;
	unsigned long long M;
	rehuel_device_io io;
	// warp-per-system kernels only (filled in by launch_warp): global slabs for the pivoted band fallback, one per
	// warp of the grid; force_pivot != 0 sends every attempt through that fallback (rehuel_debug_force_pivot)
	double *pivot_scratch;
	int force_pivot;
	int batch_lanes; // thread-per-system kernels (filled in by launch_thread): waiting lanes that trigger a service stop, see kBatchLanes
	int trace_interval; // io.trace: log the attempts whose accepted-step count is a multiple of this (irk.hpp:805-806)
};

constexpr double kMachinePrecision = 1e-17; // enums.hpp:129
constexpr int kNewtonSuccess = 0, kNewtonMaxit = 3; // newton.hpp:45-51
#ifndef REHUEL_FAST_ITERS
#define REHUEL_FAST_ITERS 64
#endif
constexpr int kFastIters = REHUEL_FAST_ITERS; // Newton iterations a lane may take before it is parked for a slow trip
#ifndef REHUEL_PARK_LANES
#define REHUEL_PARK_LANES 16
#endif
constexpr int kParkLanes = REHUEL_PARK_LANES; // parked lanes per warp that trigger a slow trip
#ifndef REHUEL_BATCH_LANES
#define REHUEL_BATCH_LANES 32
#endif
#ifndef REHUEL_COHORT
#define REHUEL_COHORT 1
#endif
#ifndef REHUEL_IDLE_SHIFT
#define REHUEL_IDLE_SHIFT 3
#endif
#ifndef REHUEL_WAIT_SHIFT
#define REHUEL_WAIT_SHIFT 3
#endif
// GANG SCHEDULING.  Writing results and handing out new members is executed by the whole warp however few lanes need
// it, and lanes that hold members at the same stage of their integration need the same number of Newton iterations:
// so a warp takes its members 32 at a time.  A lane whose member is finished keeps the results in its registers and
// sits out until all lanes of the warp wait (kBatchLanes) -- or until it has waited an eighth (2^-kWaitShift) of its
// own member's attempts, whichever comes first, so that a warp whose members differ widely in length (an unsorted
// long-span ensemble) does not idle behind its slowest member; then every waiting lane is served at once.  Served at
// once, too, when nothing else runs.  Measured on the 2^20-member sweep: 1.71 -> 1.49 ms (Morton hand-out), 1.92 ->
// 1.64 ms (index order) against serving four lanes at a time.
// COHORTS (REHUEL_COHORT): a served lane is not given a new member while other lanes of the warp still integrate --
// the warp's 32 members start together and so stay at the same stage -- unless it has been idle for an eighth
// (2^-REHUEL_IDLE_SHIFT) of the attempts its last member took.  Long span (t1 = 1e11, 2^18 members, Morton hand-out)
// 21.3 -> 19.9 ms, short span unchanged; in index order (neighbours unrelated) 24.2 -> 25.0 ms, which is what the idle
// bound is for (without it 27.0 ms).  A gang costs a warp as many trips as its longest
// member, so a launch with only a few members per lane is quantised in whole gangs (1.7 members per lane take two
// gang times): launch_thread falls back to serving four lanes at a time (kSmallBatchLanes) below kGangMinPerLane
// members per lane, and the host layer cuts its chunks in multiples of the grid's lanes.
constexpr int kBatchLanes = REHUEL_BATCH_LANES;
constexpr int kSmallBatchLanes = 4;
constexpr int kGangMinPerLane = 4;
constexpr int kWaitShift = REHUEL_WAIT_SHIFT;

// Dev builds (-DREHUEL_PHASE_CLOCKS) of the CTA and warp kernels: one thread of block 0 adds up clock64() between marks
// and prints the shares at the end of the kernel (the meaning of the nine slots is the kernel's).
#ifdef REHUEL_PHASE_CLOCKS
struct PhaseClock {
	long long t, acc[9];
	bool on;
	__device__ void start() { if (on) t = clock64(); }
	__device__ void mark(int i) { if (on) { const long long now = clock64(); acc[i] += now - t; t = now; } }
};
#define REHUEL_PHASE_MARK(pc, i) (pc).mark(i)
#else
struct PhaseClock {};
#define REHUEL_PHASE_MARK(pc, i) ((void)0)
#endif

// Reciprocal: MUFU.RCP64H seed (rcp.approx.f64: the upper 20 mantissa bits) + ONE cubically convergent step
// r (1 + e + e^2), e = 1 - x r (2^-20 -> 2^-60, i.e. correctly rounded up to ~1 ulp): 4 instructions and no slow
// path instead of the ~14-instruction IEEE division (two Newton steps would be 5).  Used for the LU pivots and 1/dt (the rounding of the LINEAR SOLVES is not part of the
// parity contract: the decoupled solve already differs from the reference's dense LU in the last bits)
// and, for the same reason -- everything downstream of the solve carries that last-bit difference --,
// for the scaled error norm and the step-size proposal (propose_dt).  0, Inf and NaN arguments still
// produce Inf/NaN, so a singular pivot fails the Newton iteration as it does in the reference.
__device__ __forceinline__ double fast_rcp(double x)
{
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	const double e = fma(-x, r, 1.0); // |e| <= 2^-20: the seed carries the upper 20 mantissa bits
	const double t = fma(e, e, e);    // r (1 + e + e^2): the neglected e^3 is below 2^-60
	return fma(r, t, r);
}

// One row of the per-attempt log of member io.trace_member, irk.hpp:805-810 ("step t dt err iters"; the reference prints
// it for the attempts that reach the error estimate, accepted or not, when step % out_interval == 0).  Attempts whose
// Newton iteration failed are logged too, with err = -1 (the reference has that message commented out, irk.hpp:649-656).
// Called by the one thread that speaks for the member.
__device__ __forceinline__ void trace_row(const KernelArgs &a, unsigned long long mem, long long step, double t, double dt,
                                          double err, int iters)
{
	const rehuel_device_io &io = a.io;
	if (io.trace && mem == io.trace_member && (a.trace_interval <= 1 || step % a.trace_interval == 0)) {
		const unsigned r = atomicAdd(io.trace_len, 1u);
		if (r < io.trace_cap) {
			double *row = io.trace + 5 * (size_t)r;
			row[0] = (double)step;
			row[1] = t;
			row[2] = dt;
			row[3] = err;
			row[4] = (double)iters;
		}
	}
}

// ------------------------------------------------------------------ small dense LU, registers
// LAPACK-style row partial pivoting needs row interchanges; on statically indexed registers an
// interchange is a pair of predicated selects per entry, ~45 % of the instructions of a 3x3
// factorisation -- and for mu*I - J they are almost never needed (the shift mu = lambda/dt makes the
// diagonal heavy).  The factorisation is therefore SPECULATIVE (REHUEL_PIVOT_MODE 2):
//   * plain elimination first (PIVOT = false), which only CHECKS, with compares, whether partial
//     pivoting would have exchanged rows anywhere;
//   * if any lane of the warp says so, the warp redoes the factorisations with the pivoted routine
//     (PIVOT = true) and remembers to start with it next time (see the kernel).
// Where no interchange is needed the two routines perform the same operations, so the result is
// bit-identical to always pivoting (mode 1).  The interchanges of the pivoted path are recorded
// as 2-bit offsets p-k in one packed word and applied to right-hand sides by apply_row_swaps.
#ifndef REHUEL_PIVOT_MODE
#define REHUEL_PIVOT_MODE 2
#endif

template <int N>
struct LuReal {
	double a[N][N]; // unit-lower L below the diagonal, U above, RECIPROCAL pivots on the diagonal
};
template <int N>
struct LuCplx {
	double re[N][N], im[N][N];
};

// bit offset of the interchange of elimination step k of matrix `slot` (N <= 4: 3 steps x 2 bits)
__device__ __forceinline__ constexpr int piv_shift(int slot, int k) { return (slot * 3 + k) * 2; }

// PIVOT = false: returns true if partial pivoting would have exchanged rows (the factors are then
//                not the ones LAPACK would produce and the caller falls back);
// PIVOT = true:  records the interchanges in `pivs`, returns true if there was any.
// ZCOL0: bit i set = entry (i, 0), i > 0, is structurally zero (the functor's kJacZeroMask): its multiplier is zero and
// the row update of the first elimination step is skipped.  Only meaningful without interchanges (PIVOT = false).
template <bool PIVOT, unsigned ZCOL0 = 0u, int N>
__device__ __forceinline__ bool lu_factor(LuReal<N> &m, unsigned &pivs, int slot)
{
	static_assert(N <= 4, "interchanges are packed as 2-bit offsets");
	static_assert(!PIVOT || ZCOL0 == 0u, "structural zeros move when rows are interchanged");
	bool flag = false;
#pragma unroll
	for (int k = 0; k < N; ++k) {
		if (k + 1 < N) {
			int p = k;
			double best = fabs(m.a[k][k]);
#pragma unroll
			for (int i = k + 1; i < N; ++i) {
				if (k == 0 && (ZCOL0 >> i & 1u)) continue;
				const double v = fabs(m.a[i][k]);
				if (v > best) {
					best = v;
					p = i;
				}
			}
			flag |= (p != k);
			if (PIVOT) {
				pivs |= (unsigned)(p - k) << piv_shift(slot, k);
#pragma unroll
				for (int i = k + 1; i < N; ++i) {
					const bool sw = (p == i);
#pragma unroll
					for (int j = 0; j < N; ++j) {
						const double u = m.a[k][j], w = m.a[i][j];
						m.a[k][j] = sw ? w : u;
						m.a[i][j] = sw ? u : w;
					}
				}
			}
		}
		const double rp = fast_rcp(m.a[k][k]);
		m.a[k][k] = rp;
#pragma unroll
		for (int i = k + 1; i < N; ++i) {
			if (k == 0 && (ZCOL0 >> i & 1u)) {
				m.a[i][k] = 0.0;
				continue;
			}
			m.a[i][k] *= rp;
#pragma unroll
			for (int j = k + 1; j < N; ++j) m.a[i][j] = fma(-m.a[i][k], m.a[k][j], m.a[i][j]);
		}
	}
	return flag;
}

// x <- P x for the interchanges recorded for matrix `slot`
template <int N>
__device__ __forceinline__ void apply_row_swaps(unsigned pivs, int slot, double (&x)[N])
{
#pragma unroll
	for (int k = 0; k + 1 < N; ++k) {
		const int d = (int)((pivs >> piv_shift(slot, k)) & 3u);
#pragma unroll
		for (int i = k + 1; i < N; ++i) {
			const bool sw = (d == i - k);
			const double u = x[k], w = x[i];
			x[k] = sw ? w : u;
			x[i] = sw ? u : w;
		}
	}
}

// x <- U^-1 L^-1 x (the interchanges, if any, have been applied to x already)
template <unsigned ZCOL0 = 0u, int N>
__device__ __forceinline__ void lu_solve(const LuReal<N> &m, double (&x)[N])
{
#pragma unroll
	for (int i = 1; i < N; ++i)
#pragma unroll
		for (int k = 0; k < i; ++k) {
			if (k == 0 && (ZCOL0 >> i & 1u)) continue; // multiplier structurally zero
			x[i] = fma(-m.a[i][k], x[k], x[i]);
		}
#pragma unroll
	for (int i = N - 1; i >= 0; --i) {
#pragma unroll
		for (int k = i + 1; k < N; ++k) x[i] = fma(-m.a[i][k], x[k], x[i]);
		x[i] *= m.a[i][i];
	}
}

// The complex Newton matrices are (mr + i mi) I - J: when they arrive here their imaginary part is mi on the diagonal
// and ZERO elsewhere.  DIAGIM = true tells the first elimination step so (multipliers and updates of that step then take
// half the arithmetic: the compiler cannot drop a multiplication by a literal 0.0 itself, NaN/-0 semantics); only without
// interchanges, which would move the diagonal.  After that step row 0 of U still has a real off-diagonal part, which
// lu_solve<..., true> uses.
template <bool PIVOT, unsigned ZCOL0 = 0u, bool DIAGIM = false, int N>
__device__ __forceinline__ bool lu_factor(LuCplx<N> &m, unsigned &pivs, int slot)
{
	static_assert(N <= 4, "interchanges are packed as 2-bit offsets");
	static_assert(!PIVOT || (ZCOL0 == 0u && !DIAGIM), "structural zeros move when rows are interchanged");
	bool flag = false;
#pragma unroll
	for (int k = 0; k < N; ++k) {
		const bool first = DIAGIM && k == 0;
		if (k + 1 < N) {
			int p = k;
			double best = fabs(m.re[k][k]) + fabs(m.im[k][k]); // izamax's |re|+|im|
#pragma unroll
			for (int i = k + 1; i < N; ++i) {
				if (k == 0 && (ZCOL0 >> i & 1u)) continue;
				const double v = first ? fabs(m.re[i][k]) : fabs(m.re[i][k]) + fabs(m.im[i][k]);
				if (v > best) {
					best = v;
					p = i;
				}
			}
			flag |= (p != k);
			if (PIVOT) {
				pivs |= (unsigned)(p - k) << piv_shift(slot, k);
#pragma unroll
				for (int i = k + 1; i < N; ++i) {
					const bool sw = (p == i);
#pragma unroll
					for (int j = 0; j < N; ++j) {
						double u = m.re[k][j], w = m.re[i][j];
						m.re[k][j] = sw ? w : u;
						m.re[i][j] = sw ? u : w;
						u = m.im[k][j];
						w = m.im[i][j];
						m.im[k][j] = sw ? w : u;
						m.im[i][j] = sw ? u : w;
					}
				}
			}
		}
		// reciprocal pivot 1/(a+ib) = (a-ib)/(a^2+b^2)
		const double pr = m.re[k][k], pi = m.im[k][k];
		const double rden = fast_rcp(fma(pr, pr, pi * pi));
		const double rr = pr * rden, ri = -pi * rden;
		m.re[k][k] = rr;
		m.im[k][k] = ri;
#pragma unroll
		for (int i = k + 1; i < N; ++i) {
			if (k == 0 && (ZCOL0 >> i & 1u)) {
				m.re[i][k] = 0.0;
				m.im[i][k] = 0.0;
				continue;
			}
			double lr, li;
			if (first) { // im[i][0] == 0
				lr = m.re[i][k] * rr;
				li = m.re[i][k] * ri;
			} else {
				lr = fma(m.re[i][k], rr, -m.im[i][k] * ri);
				li = fma(m.re[i][k], ri, m.im[i][k] * rr);
			}
			m.re[i][k] = lr;
			m.im[i][k] = li;
#pragma unroll
			for (int j = k + 1; j < N; ++j) {
				if (first) { // im[0][j] == 0; im[i][j] is mi on the diagonal, 0 elsewhere
					m.re[i][j] = fma(-lr, m.re[k][j], m.re[i][j]);
					m.im[i][j] = (i == j) ? fma(-li, m.re[k][j], m.im[i][j]) : -li * m.re[k][j];
				} else {
					m.re[i][j] = fma(-lr, m.re[k][j], fma(li, m.im[k][j], m.re[i][j]));
					m.im[i][j] = fma(-lr, m.im[k][j], fma(-li, m.re[k][j], m.im[i][j]));
				}
			}
		}
	}
	return flag;
}

// ROW0RE: row 0 of U has no imaginary off-diagonal part (the matrix came through lu_factor<false, ..., true>)
template <unsigned ZCOL0 = 0u, bool ROW0RE = false, int N>
__device__ __forceinline__ void lu_solve(const LuCplx<N> &m, double (&xr)[N], double (&xi)[N])
{
#pragma unroll
	for (int i = 1; i < N; ++i)
#pragma unroll
		for (int k = 0; k < i; ++k) {
			if (k == 0 && (ZCOL0 >> i & 1u)) continue; // multiplier structurally zero
			xr[i] = fma(-m.re[i][k], xr[k], fma(m.im[i][k], xi[k], xr[i]));
			xi[i] = fma(-m.re[i][k], xi[k], fma(-m.im[i][k], xr[k], xi[i]));
		}
#pragma unroll
	for (int i = N - 1; i >= 0; --i) {
#pragma unroll
		for (int k = i + 1; k < N; ++k) {
			if (ROW0RE && i == 0) {
				xr[i] = fma(-m.re[i][k], xr[k], xr[i]);
				xi[i] = fma(-m.re[i][k], xi[k], xi[i]);
			} else {
				xr[i] = fma(-m.re[i][k], xr[k], fma(m.im[i][k], xi[k], xr[i]));
				xi[i] = fma(-m.re[i][k], xi[k], fma(-m.im[i][k], xr[k], xi[i]));
			}
		}
		const double tr = xr[i], ti = xi[i];
		xr[i] = fma(tr, m.re[i][i], -ti * m.im[i][i]);
		xi[i] = fma(tr, m.im[i][i], ti * m.re[i][i]);
	}
}

// Optional functor trait: `static constexpr unsigned kJacZeroMask`: bit i*N + j set promises that J(i, j) is identically
// zero.  The kernel uses the entries of column 0 (multipliers that are zero in the first elimination step of mu I - J).
template <class F, typename = void>
struct jac_zero_col0 { static constexpr unsigned value = 0u; };
template <class F>
struct jac_zero_col0<F, decltype((void)F::kJacZeroMask)> {
	static constexpr unsigned make()
	{
		unsigned m = 0u;
		for (int i = 1; i < F::N; ++i)
			if (F::kJacZeroMask >> (i * F::N) & 1u) m |= 1u << i;
		return m;
	}
	static constexpr unsigned value = make();
};

// Optional functor trait: `static constexpr bool kAutonomous = true` promises f(t, y) does not depend
// on t.  Every attempt starts its Newton iteration from Y = 0 (irk.hpp:415), where all S stage
// evaluations of the first residual are the same f(t + c_i dt, y) = f(y): the kernel then evaluates
// it once (the evaluation counters still book the S calls the reference makes) and reuses it for
// the gamma*dt*f(t, y) term of the error estimate (irk.hpp:702).
#ifndef REHUEL_STRUCT_ZEROS
#define REHUEL_STRUCT_ZEROS 1 // use the structural zeros of mu I - J in the unpivoted factorisations and solves
#endif
#ifndef REHUEL_TREE_NORMS
#define REHUEL_TREE_NORMS 1 // |R|^2 and |dY|^2 as S partial sums (dependency chains of n + S instead of n S FMAs)
#endif
#ifndef REHUEL_CANONICAL_TABLEAU
#define REHUEL_CANONICAL_TABLEAU 1 // use d = e_s and the unit last row of T (Tableau::canonical, tableau.hpp) instead of multiplying by 1 and 0
#endif
#ifndef REHUEL_PEEL_FIRST
#define REHUEL_PEEL_FIRST 1 // first Newton iteration as its own instance: Y = dY instead of Y = 0; Y += dY
#endif
#ifndef REHUEL_ROOT_CONTROLLER
#define REHUEL_ROOT_CONTROLLER 1 // step-size proposal from err^2 with one k-th root (propose_dt_sq) instead of three square roots
#endif
#ifndef REHUEL_SKIP_DEAD_RESIDUAL
#define REHUEL_SKIP_DEAD_RESIDUAL 1
#endif
#ifndef REHUEL_AUTONOMOUS_OPT
#define REHUEL_AUTONOMOUS_OPT 2 // 0: off, 1: first residual only, 2: + error estimate
#endif
template <class F, typename = void>
struct is_autonomous { static constexpr bool value = false; };
template <class F>
struct is_autonomous<F, decltype((void)F::kAutonomous)> { static constexpr bool value = F::kAutonomous; };

// max(|a|, |b|) (irk.hpp:741: std::max(std::abs(y), std::abs(y_n))) on the integer pipe: non-negative doubles order like
// their bit patterns.  (fmax is a DSETP + selects + moves; a NaN operand gives NaN here, the member is lost either way.)
__device__ __forceinline__ double max_abs(double a, double b)
{
	const long long ia = __double_as_longlong(a) & 0x7fffffffffffffffll, ib = __double_as_longlong(b) & 0x7fffffffffffffffll;
	return __longlong_as_double(ia > ib ? ia : ib);
}

// ------------------------------------------------------------------ controller power
// err^(1/(1+min_order)) (irk.hpp:774-782).  The exponent is a tableau constant; the common
// values get exact-root forms (each correctly rounded step, <= 1 ulp from pow()).
__device__ __forceinline__ double ctrl_pow(double x, int min_order, double expo)
{
	if (min_order == 3) return sqrt(sqrt(x));
	if (min_order == 2) return cbrt(x);
	if (min_order == 5) return sqrt(cbrt(x));
	return pow(x, expo);
}

// Step-size proposal, irk.hpp:770-801:
//   fac = 0.9 (maxit+1)/(maxit+iters);  s27 = (1/err)^e;  s28 = s27 (dts0/dts1) (errs1/errs0)^e;
//   new_dt = fac dt min(8, min(s27, s28));   e = 1/(1+min(order, order2)),  errs0 == err.
// Evaluated with ONE root instead of two powers: with q = err^e,  s27 = 1/q  and
// (errs1/errs0)^e = q_prev/q = q_prev * s27, so the kernels carry q_prev = errs1^e in place of errs1
// (errs[] entries are never 0: err is floored at 1e-17 and the history starts at 0.9, irk.hpp:573,752).
// Reciprocals go through fast_rcp (<= 1 ulp); the proposal differs from the reference expression
// by a few ulp, like everything downstream of the decoupled Newton solve.
__device__ __forceinline__ double propose_dt(double dt, double err, double &q_prev, double dts0, double dts1, int maxit,
                                             int iters, int min_order, double expo, double max_dt)
{
	const double q = ctrl_pow(err, min_order, expo);
	const double s27 = fast_rcp(q);
	const double s28 = s27 * (dts0 * fast_rcp(dts1)) * (q_prev * s27);
	q_prev = q; // irk.hpp:756-758: the history shifts on every attempt that reaches the estimator
	const double fac = (0.9 * (maxit + 1.0)) * fast_rcp((double)(maxit + iters));
	const double msc = (s28 < s27) ? s28 : s27;
	double new_dt = fac * dt * (msc < 8.0 ? msc : 8.0);
	if (max_dt > 0.0) new_dt = new_dt < max_dt ? new_dt : max_dt;
	return new_dt;
}

// The same proposal from the SQUARED error norm x = err^2 (the thread-per-system kernels): with k = 2 (1 + min_order),
//   s27 = (1/err)^e = x^(-1/k)    and    q = err^e = x^(1/k) = x * (x^(-1/k))^(k-1),
// so the three chained square roots of err = sqrt(x), q = sqrt(sqrt(err)) (and the reciprocal of q) become ONE k-th
// root: x is scaled by an exact power of two into [2^(-k/2), 2^(k/2)], a float-precision seed
// exp2(-log2(x)/k) (MUFU.LG2/EX2, ~2^-22) is refined by one cubically convergent step on r^-k = x, and the power
// of two is put back by exponent arithmetic.  About 12 FP64 instructions in two short dependency chains instead
// of ~40 in one long one; the result differs from the
// sqrt/reciprocal form by an ulp or two, like everything else in this function.  x must be positive and finite
// (the caller floors it at 1e-34 and sends non-finite values through propose_dt).
template <int K>
__device__ __forceinline__ double pow_int(double r)
{
	if (K == 1) return r;
	const double h = pow_int<(K > 1 ? K / 2 : 1)>(r);
	return (K & 1) ? h * h * r : h * h;
}

template <int K>
__device__ __forceinline__ void inv_root(double x, double &r_out, double &q_out)
{
	const int hi = __double2hiint(x);
	// j = round(E / K), E = floor(log2 x) = biased exponent - 1023: unsigned arithmetic (the +1024 K keeps it positive)
	const unsigned eb = ((unsigned)hi >> 20) & 0x7ffu;
	const int j = (int)((eb + (unsigned)(1024 * K - 1023 + K / 2)) / (unsigned)K) - 1024;
	const double xs = __hiloint2double(hi - K * j * 0x100000, __double2loint(x)); // x 2^(-K j), exact, in [2^(-K/2), 2^(K/2+1))
	float lg, sd;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float)xs)); // no denormals here: the guarded log2f/exp2f are not needed
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sd) : "f"(lg * (-1.0f / K)));
	double r = (double)sd;
	{ // ONE cubically convergent step:  r (1 - e)^(-1/K) = r (1 + e/K + (K+1) e^2 / (2 K^2) + O(e^3)),  e = 1 - x r^K.
	  // The seed is good to ~2^-22, so |e| <= K 2^-22 and the neglected term (K+1)(2K+1)/(6 K^3) e^3 is below 2^-58.
		const double e = fma(-xs, pow_int<K>(r), 1.0);
		const double t = fma(e, (K + 1.0) / (2.0 * K * K), 1.0 / K) * e;
		r = fma(r, t, r);
	}
	const double q = xs * pow_int<K - 1>(r);
	r_out = __hiloint2double(__double2hiint(r) - j * 0x100000, __double2loint(r)); // r 2^(-j)
	q_out = __hiloint2double(__double2hiint(q) + j * 0x100000, __double2loint(q)); // q 2^(+j)
}

// min(order, order2) of the tableau a kernel instantiation is built for (launch.cuh: one tableau per (S, NR, NC, EST));
// 0: not known at compile time (use the run-time value).  irk.cpp:215-689.
template <int S, int NR, int NC, int EST>
struct static_min_order {
	static constexpr int value = (S == 2 && NC == 1) ? 2                           // RADAU_IIA_32 (3, 2)
	                             : (S == 3 && EST == EST_REUSE) ? 3                // RADAU_IIA_53 (5, 3)
	                             : (S == 3 && EST == EST_IDENTITY) ? 2             // LOBATTO_IIIC_43 (4, 2)
	                             : (S == 4) ? 4                                    // LOBATTO_IIIC_64 (6, 4)
	                             : (S == 5) ? 5                                    // RADAU_IIA_95 (9, 5), LOBATTO_IIIC_85 (8, 5)
	                             : (S == 7) ? 7                                    // RADAU_IIA_137 (13, 7)
	                                        : 0;
};

template <int MO> // MO: compile-time min(order, order2), or 0: use `min_order`
__device__ __forceinline__ double propose_dt_sq(double dt, double errsq, double &q_prev, double dts0, double dts1, int maxit,
                                                int iters, int min_order, double expo, double max_dt)
{
	double s27, q;
	const int mo = MO ? MO : min_order;
	if (!(errsq <= 1e300)) { // Inf or NaN: keep the reference's arithmetic (the step is lost anyway)
		q = ctrl_pow(sqrt(errsq), mo, expo);
		s27 = fast_rcp(q);
	} else if (mo == 3) {
		inv_root<8>(errsq, s27, q);
	} else if (mo == 2) {
		inv_root<6>(errsq, s27, q);
	} else if (mo == 5) {
		inv_root<12>(errsq, s27, q);
	} else if (mo == 4) {
		inv_root<10>(errsq, s27, q);
	} else if (mo == 7) {
		inv_root<16>(errsq, s27, q);
	} else {
		q = ctrl_pow(sqrt(errsq), mo, expo);
		s27 = fast_rcp(q);
	}
	const double s28 = s27 * (dts0 * fast_rcp(dts1)) * (q_prev * s27);
	q_prev = q; // irk.hpp:756-758: the history shifts on every attempt that reaches the estimator
	const double fac = (0.9 * (maxit + 1.0)) * fast_rcp((double)(maxit + iters));
	const double msc = (s28 < s27) ? s28 : s27;
	double new_dt = fac * dt * (msc < 8.0 ? msc : 8.0);
	if (max_dt > 0.0) new_dt = new_dt < max_dt ? new_dt : max_dt;
	return new_dt;
}

// Dense-output weights of one grid time inside an accepted step:
// w_i(theta) = sum_k Wd[i][k] theta^(S-k) by Horner, so that y(t + theta dt) = y + sum_i w_i Y_i.
template <int S, int NC>
__device__ __forceinline__ void dense_weights(const DevTableau<S, NC> &tb, double theta, double (&w)[S])
{
#pragma unroll
	for (int i = 0; i < S; ++i) {
		double acc = tb.Wd[i][0];
#pragma unroll
		for (int k = 1; k < S; ++k) acc = fma(acc, theta, tb.Wd[i][k]);
		w[i] = acc * theta;
	}
}

// ------------------------------------------------------------------ the kernel
template <class F, int S, int NR, int NC, int EST, int BLOCK, int MIN_BLOCKS, bool TOLV = false, bool DENSE = false>
#ifdef REHUEL_MAXNREG
__global__ void __maxnreg__(REHUEL_MAXNREG)
#else
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
#endif
ensemble_irk_thread(const __grid_constant__ KernelArgs a, const __grid_constant__ DevTableau<S, NC> tb)
{
	static_assert(S == NR + 2 * NC, "stage count must match the eigen-structure");
	static_assert(NR <= 1, "at most one real eigenvalue");
	constexpr int N = F::N;
	constexpr int PA = F::P > 0 ? F::P : 1;
	constexpr int kAuto = is_autonomous<F>::value ? REHUEL_AUTONOMOUS_OPT : 0;
	constexpr unsigned ZC = REHUEL_STRUCT_ZEROS ? jac_zero_col0<F>::value : 0u; // rows i with J(i, 0) == 0
	constexpr bool DIAGIM = REHUEL_STRUCT_ZEROS != 0;
	const unsigned lane = threadIdx.x & 31u;
	const unsigned long long M = a.M;
	const rehuel_device_io &io = a.io;
	// ensemble sums (io.stats): a finishing member adds its counters to the CTA's shared-memory totals, the CTA adds
	// them to global memory once at the end -- no separate reduction pass over the counter arrays
	__shared__ unsigned long long s_stats[REHUEL_NSTATS];
	if (threadIdx.x < REHUEL_NSTATS) s_stats[threadIdx.x] = 0ull;
	__syncthreads();

	// ---- lane-persistent member state
	// lane state flags in one word (as separate bools they were spilled byte by byte)
	struct LaneFlags {
		unsigned have : 1;       // the lane holds a member
		unsigned done : 1;       // the member is finished, its results are still in this lane's registers (see kBatchLanes)
		unsigned exhausted : 1;  // the queue had nothing left for this lane
		unsigned parked : 1;     // this lane's attempt ran past kFastIters Newton iterations: redo it in a slow trip
		unsigned prefer_piv : 1; // the warp's last factorisations needed row interchanges: skip the speculative plain attempt
		unsigned alt : 1;        // alternative_error_formula, irk.hpp:587, 763, 845
		unsigned waited : 16;    // trips this lane has sat out with a finished member, or idle (see kBatchLanes)
	} fl{};
	fl.alt = 1;
	unsigned long long mem = 0;
	double y[N], p[PA];
	double t = 0.0, dt = 0.0;
	double dts0 = 0.0, dts1 = 0.0, q_prev = tb.q_init, err = 0.0; // q_prev = errs[1]^expo, see propose_dt
	int maxit = a.maxit0;
	// accepted steps of THIS call and since the Newton budget was last raised (irk.hpp:644-648); a continued member's
	// earlier legs stay in io.state_in and are only added where the total is needed (max_steps, log, samples, state_out)
	unsigned steps = 0, since_relax = 0;
	int status = REHUEL_SUCCESS; // kAborted marks an attempt that was counted but neither accepted nor rejected
	constexpr int kAborted = 1 << 30;
	unsigned c_attempt = 0, c_rej_newton = 0, c_rej_err = 0, c_fun = 0, c_jac = 0,
	         c_iters = 0, n_stored = 0;
	unsigned kd = 0; // dense output: index of the next grid time (DENSE kernels only)
#ifdef REHUEL_COUNT_SKIPS
	unsigned c_skips = 0;
#endif
	// per-member tolerances (TOLV kernels only; otherwise the uniform values of the options)
	double m_rtol = a.rtol, m_atol = a.atol, m_Rtol2 = a.Rtol2;

	auto total_step = [&]() -> long long { // irk_guts' `step`: accepted steps including earlier legs
		return (long long)steps + (io.state_in ? (long long)io.state_in[REHUEL_STATE_STEP * M + mem] : 0ll);
	};
	for (;;) {
		// ---------------- service stop: write the results of finished members, refill idle lanes from the global queue
		// (warp-aggregated).  Warp-uniform decision, see kBatchLanes.
		const unsigned wants = __ballot_sync(0xffffffffu, (fl.have && fl.done) || (!fl.have && !fl.exhausted));
		const unsigned busy = __ballot_sync(0xffffffffu, fl.have && !fl.done);
		bool impatient = false, idle_long = false;
		if (fl.have && fl.done) {
			const unsigned patience = (c_attempt >> kWaitShift) > 2u ? (c_attempt >> kWaitShift) : 2u;
			impatient = fl.waited >= patience;
			if (fl.waited < 0xffffu) fl.waited = fl.waited + 1u;
		} else if (REHUEL_COHORT && !fl.have && !fl.exhausted) { // c_attempt: the member this lane had before
			const unsigned patience = (c_attempt >> REHUEL_IDLE_SHIFT) > 2u ? (c_attempt >> REHUEL_IDLE_SHIFT) : 2u;
			idle_long = fl.waited >= patience;
			if (fl.waited < 0xffffu) fl.waited = fl.waited + 1u;
		}
		const bool refill = !REHUEL_COHORT || a.batch_lanes < 32 || busy == 0u || __any_sync(0xffffffffu, idle_long);
		if (wants != 0u && (__popc(wants) >= a.batch_lanes || busy == 0u || __any_sync(0xffffffffu, impatient) || (REHUEL_COHORT && refill))) {
			if (fl.have && fl.done) {
#pragma unroll
				for (int k = 0; k < N; ++k) io.y[k * M + mem] = y[k];
				io.t_final[mem] = t;
				if (io.err_last) io.err_last[mem] = REHUEL_ROOT_CONTROLLER ? sqrt(err) : err;
				io.status[mem] = status & ~kAborted;
				const unsigned steps_here = steps;
				// every attempt ends accepted, rejected or -- once, ending the member -- aborted
				const unsigned c_newton_ok = c_attempt - c_rej_newton - ((status & kAborted) ? 1u : 0u);
				if (io.counters) {
					io.counters[REHUEL_CNT_ATTEMPT * M + mem] = c_attempt;
					io.counters[REHUEL_CNT_REJECT_NEWTON * M + mem] = c_rej_newton;
					io.counters[REHUEL_CNT_REJECT_ERR * M + mem] = c_rej_err;
					io.counters[REHUEL_CNT_NEWTON_SUCCESS * M + mem] = c_newton_ok;
#ifdef REHUEL_COUNT_SKIPS
					io.counters[REHUEL_CNT_NEWTON_MAXIT_EXCEED * M + mem] = c_skips;
					c_skips = 0;
#else
					io.counters[REHUEL_CNT_NEWTON_MAXIT_EXCEED * M + mem] = c_rej_newton;
#endif
					io.counters[REHUEL_CNT_FUN_EVALS * M + mem] = c_fun;
					io.counters[REHUEL_CNT_JAC_EVALS * M + mem] = c_jac;
					io.counters[REHUEL_CNT_NEWTON_ITERS * M + mem] = c_iters;
					io.counters[REHUEL_CNT_STEPS * M + mem] = steps_here;
				}
				if (io.stats) {
					// the lanes that finish a member in this trip add up first (REDUX), their leader adds the sums to the
					// CTA's 64-bit shared-memory totals, which go to the global words at the end of the kernel
					const unsigned fm = __activemask();
					const unsigned v[REHUEL_NCOUNTERS] = {c_attempt, c_rej_newton, c_rej_err, c_newton_ok, c_rej_newton, c_fun, c_jac, c_iters, steps_here};
					unsigned long long w[REHUEL_NCOUNTERS];
#pragma unroll
					for (int c = 0; c < REHUEL_NCOUNTERS; ++c) // 16-bit halves: a sum over 32 lanes cannot wrap
						w[c] = (unsigned long long)__reduce_add_sync(fm, v[c] & 0xffffu) + ((unsigned long long)__reduce_add_sync(fm, v[c] >> 16) << 16);
					const unsigned nfail = __reduce_add_sync(fm, (status & ~kAborted) != 0 ? 1u : 0u);
					const unsigned amax = __reduce_max_sync(fm, c_attempt);
					if ((int)lane == __ffs(fm) - 1) {
#pragma unroll
						for (int c = 0; c < REHUEL_NCOUNTERS; ++c) atomicAdd(&s_stats[c], w[c]);
						atomicAdd(&s_stats[REHUEL_NCOUNTERS], (unsigned long long)nfail);
						atomicMax(&s_stats[REHUEL_NCOUNTERS + 1], (unsigned long long)amax);
					}
				}
				if (io.n_samples) io.n_samples[mem] = n_stored;
				if (io.state_out) {
					double *so = io.state_out + mem;
					so[REHUEL_STATE_DT * M] = dt;
					so[REHUEL_STATE_DTS0 * M] = dts0;
					so[REHUEL_STATE_DTS1 * M] = dts1;
					so[REHUEL_STATE_ERRQ * M] = q_prev;
					so[REHUEL_STATE_ALT * M] = fl.alt ? 1.0 : 0.0;
					so[REHUEL_STATE_MAXIT * M] = (double)maxit;
					so[REHUEL_STATE_STEP * M] = (double)total_step();
					so[REHUEL_STATE_LAST_RELAX * M] = (double)(total_step() - (long long)since_relax);
				}
				fl.have = false;
				fl.done = false;
				fl.waited = 0u;
			}
			const unsigned need = __ballot_sync(0xffffffffu, !fl.have && !fl.exhausted && refill);
			if (need) {
				const int leader = __ffs(need) - 1;
				unsigned long long base = 0;
				if ((int)lane == leader) base = atomicAdd(io.queue, (unsigned long long)__popc(need));
				base = __shfl_sync(0xffffffffu, base, leader);
				if (!fl.have && !fl.exhausted) {
					mem = base + __popc(need & ((1u << lane) - 1u));
					if (mem < M) {
						if (io.order) mem = io.order[mem]; // caller-chosen hand-out order (e.g. costly members first)
#pragma unroll
						for (int k = 0; k < N; ++k) y[k] = io.y0_broadcast ? io.y0[k] : io.y0[k * M + mem];
#pragma unroll
						for (int k = 0; k < F::P; ++k) p[k] = io.params[k * M + mem];
						if (F::P == 0) p[0] = 0.0;
						if (TOLV) {
							m_rtol = io.rtol[mem];
							m_atol = io.atol[mem];
							const double nt = io.newton_tol[mem];
							m_Rtol2 = nt * nt;
						}
						t = io.t_start ? io.t_start[mem] : a.t0;
						dt = a.dt0;
						if (t + dt > a.t1) dt = a.t1 - t; // irk.hpp:514-520
						dts0 = dts1 = dt;                 // irk.hpp:572
						q_prev = tb.q_init;               // irk.hpp:573: errs = 0.9
						err = 0.0;
						fl.alt = true;                       // irk.hpp:587
						maxit = a.maxit0;
						steps = 0;
						since_relax = 0;
						if (io.state_in) { // continue a stopped integration: the loop variables as they were at exit
							const double *si = io.state_in + mem;
							dt = si[REHUEL_STATE_DT * M];
							dts0 = si[REHUEL_STATE_DTS0 * M];
							dts1 = si[REHUEL_STATE_DTS1 * M];
							q_prev = si[REHUEL_STATE_ERRQ * M];
							fl.alt = si[REHUEL_STATE_ALT * M] != 0.0;
							maxit = (int)si[REHUEL_STATE_MAXIT * M];
							const double gap = si[REHUEL_STATE_STEP * M] - si[REHUEL_STATE_LAST_RELAX * M];
							since_relax = gap < 4e9 ? (unsigned)gap : 4000000000u;
						}
						if (DENSE) { // grid times before the start are not this call's; one AT the start is the initial state
							kd = 0;
							while (kd < io.n_dense && fma((double)kd, io.dense_dt, io.dense_t0) < t) ++kd;
							if (kd < io.n_dense && fma((double)kd, io.dense_dt, io.dense_t0) == t) {
#pragma unroll
								for (int k = 0; k < N; ++k) io.y_dense[((size_t)kd * N + k) * M + mem] = y[k];
								++kd;
							}
						}
						status = REHUEL_SUCCESS;
						c_attempt = c_rej_newton = c_rej_err = c_fun = c_jac = c_iters = 0;
						n_stored = 0;
						fl.parked = false;
						fl.have = true;
						if (REHUEL_COHORT) fl.waited = 0u;
						if (a.store_samples && io.n_samples_cap) { // sample 0 = (t0, y0), irk.hpp:581-585
							io.t_samples[mem] = t;
							if (io.err_samples) io.err_samples[mem] = 0.0;
#pragma unroll
							for (int k = 0; k < N; ++k) io.y_samples[k * M + mem] = y[k];
						}
						n_stored = a.store_samples ? 1u : 0u;
					} else {
						fl.exhausted = true;
					}
				}
			}
		}
		if (__all_sync(0xffffffffu, !fl.have)) break;

		// ---------------- slow-lane handling.  A failing simplified Newton iteration spins to maxit
		// (4999 iterations by default; the reference has no divergence test, irk.hpp:458-488) while
		// a normal one needs 1-3, so one such lane would idle the other 31 for thousands of
		// iterations.  A lane whose attempt is still unconverged after kFastIters iterations
		// abandons it and PARKS (keeps its member, sits out); once kParkLanes lanes of the warp are
		// parked -- or nothing else is left -- the warp runs a "slow trip" in which only the parked
		// lanes redo their attempt with the full iteration budget, side by side.  The redo starts
		// from Y = 0 with the same (t, y, dt) as the reference does, so results and counters are
		// unchanged; only the schedule differs.
		const unsigned pmask = __ballot_sync(0xffffffffu, fl.have && !fl.done && fl.parked);
		const unsigned cmask = __ballot_sync(0xffffffffu, fl.have && !fl.done && !fl.parked);
		const bool slow_trip = pmask != 0u && (__popc(pmask) >= kParkLanes || cmask == 0u);

		// ---------------- classify this trip: finish the member, or run one attempt
		bool finished = false, run = false;
		if (slow_trip) {
			run = fl.have && !fl.done && fl.parked; // attempt already counted and clamped when it was first tried
		} else if (fl.have && !fl.done && !fl.parked) {
			if (!(t < a.t1)) { // irk.hpp:604 (also covers t0 >= t1: zero attempts)
				finished = true;
			} else {
				if (t + dt > a.t1) dt = a.t1 - t; // irk.hpp:607-609
				++c_attempt;
				if (a.max_steps >= 0 && total_step() > a.max_steps) { // irk.hpp:612-616
					status = REHUEL_ERROR_MAX_STEPS_EXCEEDED | kAborted;
					finished = true;
				} else if (!(dt >= 2.2250738585072014e-308)) {
					// DELIBERATE DEVIATION: a subnormal dt (endless Newton back-offs: 0.7*denorm_min
					// rounds back to denorm_min, so dt never even reaches 0) or a NaN dt (a NaN error
					// estimate is "accepted" by irk.hpp:761) makes the reference loop forever; a GPU
					// lane must terminate.  Flag the member instead.
					status = REHUEL_GENERAL_ERROR | REHUEL_DT_TOO_SMALL | kAborted;
					finished = true;
				} else {
					run = true;
				}
			}
		}
		// Lanes that run an attempt this trip.  The mask is taken while all 32 lanes are converged
		// and is used to RE-converge them after the data-dependent Newton loop: without it the
		// compiler lets every group of lanes that leaves the loop at a different iteration run
		// the whole error-estimate/controller tail on its own (ncu: 2.4x executions at 13/32 lanes).
		const unsigned amask = __ballot_sync(0xffffffffu, run);
		if (run) {
			{
				// -------- simplified Newton on the stage increments, irk.hpp:398-493
				double Y[S][N], R[S][N];
				LuReal<N> lur;  // (mu_r I - J)        (dummy when NR == 0 and EST != EST_OWN_LU)
				LuCplx<N> luc[NC > 0 ? NC : 1];
				LuReal<N> lue;  // E/gam = (1/gam) I - J, only for EST_OWN_LU
				const double rdt = fast_rcp(dt);
				// EST_REUSE: mu_r = 1/(gamma*dt) instead of gamma_hat/dt (equal to 1e-15) so that
				// the same factors also serve E = I - gamma*dt*J of the error estimate.
				const double mur = (EST == EST_REUSE) ? tb.rgamma * rdt : tb.gamma_hat * rdt;
				// Jacobian (irk.hpp:425), the decoupled Newton matrices and their factors.  Slots of the
				// interchange word: 0 = real block, 1.. = complex blocks, 4 = the estimator's own matrix.
				unsigned pivs = 0u;
				auto factor_all = [&](auto pivot_tag) -> bool {
					constexpr bool PIV = decltype(pivot_tag)::value;
					double J[N][N];
					F::jac(t, y, p, J);
					bool flag = false;
					if (NR) {
#pragma unroll
						for (int i = 0; i < N; ++i)
#pragma unroll
							for (int j = 0; j < N; ++j) lur.a[i][j] = (i == j ? mur : 0.0) - J[i][j];
						flag |= lu_factor<PIV, PIV ? 0u : ZC>(lur, pivs, 0);
					}
#pragma unroll
					for (int c = 0; c < NC; ++c) {
						const double mr = tb.alpha[c] * rdt, mi = tb.beta[c] * rdt;
#pragma unroll
						for (int i = 0; i < N; ++i)
#pragma unroll
							for (int j = 0; j < N; ++j) {
								luc[c].re[i][j] = (i == j ? mr : 0.0) - J[i][j];
								luc[c].im[i][j] = (i == j ? mi : 0.0);
							}
						flag |= lu_factor<PIV, PIV ? 0u : ZC, !PIV && DIAGIM>(luc[c], pivs, 1 + c);
					}
					if (EST == EST_OWN_LU) {
						const double mue = tb.rgamma * rdt;
#pragma unroll
						for (int i = 0; i < N; ++i)
#pragma unroll
							for (int j = 0; j < N; ++j) lue.a[i][j] = (i == j ? mue : 0.0) - J[i][j];
						flag |= lu_factor<PIV, PIV ? 0u : ZC>(lue, pivs, 4);
					}
					return flag;
				};
				++c_jac;
				// warp-uniform decisions (every lane of amask is here): plain elimination unless the
				// warp's previous factorisations interchanged rows; pivoted redo if any lane needs it
				bool piv_any = (REHUEL_PIVOT_MODE == 1) || __any_sync(amask, fl.prefer_piv);
				if (!piv_any) piv_any = __any_sync(amask, factor_all(std::false_type()));
				if (piv_any) fl.prefer_piv = __any_sync(amask, factor_all(std::true_type()));

				// residual R = Y - dt*(A (x) I) F(y + Y), irk.hpp:366-386
				auto residual = [&]() -> double {
					double Fv[S][N];
#pragma unroll
					for (int i = 0; i < S; ++i) {
						double yi[N];
#pragma unroll
						for (int k = 0; k < N; ++k) yi[k] = y[k] + Y[i][k];
						F::fun(fma(tb.c[i], dt, t), yi, p, Fv[i]);
					}
					c_fun += S;
					double r2[S]; // one partial sum per stage: short dependency chains
#pragma unroll
					for (int i = 0; i < S; ++i) {
						r2[i] = 0.0;
#pragma unroll
						for (int k = 0; k < N; ++k) {
							double acc = 0.0;
#pragma unroll
							for (int j = 0; j < S; ++j) acc = fma(tb.A[i][j], Fv[j][k], acc);
							R[i][k] = fma(-dt, acc, Y[i][k]);
							r2[REHUEL_TREE_NORMS ? i : 0] = fma(R[i][k], R[i][k], r2[REHUEL_TREE_NORMS ? i : 0]);
						}
					}
					double tot = r2[0];
#pragma unroll
					for (int i = 1; i < S; ++i)
						if (REHUEL_TREE_NORMS) tot += r2[i];
					return tot;
				};

				if (!(REHUEL_PEEL_FIRST && kAuto)) { // irk.hpp:415 (otherwise the first iteration writes Y = dY)
#pragma unroll
					for (int i = 0; i < S; ++i)
#pragma unroll
						for (int k = 0; k < N; ++k) Y[i][k] = 0.0;
				}
				double f0[N]; // f(t, y) (autonomous functors)
				double Rnorm2;
				if (kAuto) { // Y = 0: R_i = -dt * sum_j A_ij f(y)
					F::fun(t, y, p, f0);
					c_fun += S;
					Rnorm2 = 0.0; // (the reference does not test the residual before the first update either, irk.hpp:443-458)
					if (!REHUEL_PEEL_FIRST) {
#pragma unroll
						for (int i = 0; i < S; ++i) {
							const double ai = -dt * tb.Asum[i];
#pragma unroll
							for (int k = 0; k < N; ++k) R[i][k] = ai * f0[k];
						}
					}
				} else {
					Rnorm2 = residual(); // irk.hpp:443-446
				}
				int nstat = kNewtonMaxit;
				int it = 1;
				int refresh_in = a.refresh_jac; // iterations until the next (no-op) Jacobian refresh, irk.hpp:485-487
				const int cap = (slow_trip || maxit <= kFastIters) ? maxit : kFastIters + 1;
				// One iteration; returns true when the loop is over (converged, or the iterate is beyond recovery).
				// FIRST: Y is still zero, so Y = dY (the zero fill and the additions are never executed).
				auto iteration = [&](auto swaps_tag, auto first_tag) -> bool {
					constexpr bool SWAPS = decltype(swaps_tag)::value;
					constexpr bool FIRST = decltype(first_tag)::value;
#if defined(REHUEL_COUNT_SKIPS) && REHUEL_COUNT_SKIPS == 2
					if (slow_trip) ++c_skips; // dev builds (mode 2): Newton iterations actually EXECUTED in slow trips
#endif
					// dY = -(I - dt A(x)J)^-1 R through the decoupled blocks
					double Z[S][N];
					if (FIRST) {
						// Y = 0 and f autonomous: R_i = -dt Asum_i f(y) has rank one, so does Z = (Ti (x) I) R
#pragma unroll
						for (int i = 0; i < S; ++i) {
							const double gi = -dt * tb.TiAsum[i];
#pragma unroll
							for (int k = 0; k < N; ++k) Z[i][k] = gi * f0[k];
						}
					} else {
#pragma unroll
						for (int i = 0; i < S; ++i)
#pragma unroll
							for (int k = 0; k < N; ++k) {
								double acc = 0.0;
#pragma unroll
								for (int j = 0; j < S; ++j) acc = fma(tb.Ti[i][j], R[j][k], acc);
								Z[i][k] = acc;
							}
					}
					if (SWAPS) { // the right-hand sides follow the recorded row interchanges
						if (NR) apply_row_swaps(pivs, 0, Z[0]);
#pragma unroll
						for (int c = 0; c < NC; ++c) {
							apply_row_swaps(pivs, 1 + c, Z[NR + 2 * c]);
							apply_row_swaps(pivs, 1 + c, Z[NR + 2 * c + 1]);
						}
					}
					if (NR) {
						lu_solve<SWAPS ? 0u : ZC>(lur, Z[0]);
#pragma unroll
						for (int k = 0; k < N; ++k) Z[0][k] *= -mur;
					}
#pragma unroll
					for (int c = 0; c < NC; ++c) {
						const double mr = tb.alpha[c] * rdt, mi = tb.beta[c] * rdt;
						lu_solve<SWAPS ? 0u : ZC, !SWAPS && DIAGIM>(luc[c], Z[NR + 2 * c], Z[NR + 2 * c + 1]);
#pragma unroll
						for (int k = 0; k < N; ++k) { // w = -(mr + i mi) * w
							const double wr = Z[NR + 2 * c][k], wi = Z[NR + 2 * c + 1][k];
							Z[NR + 2 * c][k] = fma(-mr, wr, mi * wi);
							Z[NR + 2 * c + 1][k] = fma(-mr, wi, -mi * wr);
						}
					}
					double x2[S]; // |dY|^2 as one partial sum per stage
#pragma unroll
					for (int i = 0; i < S; ++i) {
						x2[i] = 0.0;
#pragma unroll
						for (int k = 0; k < N; ++k) {
							double acc = 0.0;
							if (REHUEL_CANONICAL_TABLEAU && i == S - 1) {
								// last row of T: 1 for the real block, (1, 0) for every complex pair (Tableau::canonical)
								acc = Z[0][k];
#pragma unroll
								for (int j = NR ? 1 : 2; j < S; j += 2) acc += Z[j][k];
							} else {
#pragma unroll
								for (int j = 0; j < S; ++j) acc = fma(tb.T[i][j], Z[j][k], acc);
							}
							x2[REHUEL_TREE_NORMS ? i : 0] = fma(acc, acc, x2[REHUEL_TREE_NORMS ? i : 0]); // irk.hpp:466
							if (FIRST) Y[i][k] = acc;       // Y was 0 (irk.hpp:415)
							else Y[i][k] += acc;            // irk.hpp:468 (step == 1)
						}
					}
					double xnorm2 = x2[0];
#pragma unroll
					for (int i = 1; i < S; ++i)
						if (REHUEL_TREE_NORMS) xnorm2 += x2[i];
					// The reference evaluates R(Y) (irk.hpp:469-472) and then stops on |R|^2 < tol^2 (473-476) OR
					// |dY|^2 < dx_delta^2 (477-480), both with the same status and iteration count, and never
					// looks at that residual again.  When the step test already holds -- with the default
					// dx_delta = 1e-4 against tol ~ 1e-7 it almost always fires first -- the evaluation is
					// dead work: skip it, book its function calls.
					if (REHUEL_SKIP_DEAD_RESIDUAL && xnorm2 < a.xtol2) {
#if defined(REHUEL_COUNT_SKIPS) && REHUEL_COUNT_SKIPS == 1
						++c_skips; // dev builds (mode 1): how often the step test fires first (reported in the MAXIT_EXCEED row)
#endif
						c_fun += S;
						nstat = kNewtonSuccess;
						return true;
					}
					Rnorm2 = residual();                    // irk.hpp:469-472
					if (Rnorm2 < (TOLV ? m_Rtol2 : a.Rtol2)) { nstat = kNewtonSuccess; return true; } // irk.hpp:473-476
					if (xnorm2 < a.xtol2) { nstat = kNewtonSuccess; return true; } // irk.hpp:477-480
					if (--refresh_in == 0) { // it % refresh_jac == 0, irk.hpp:485-487 (a no-op refactorisation: only counted)
						++c_jac;
						refresh_in = a.refresh_jac;
					}
					if (!(Rnorm2 <= 1.79769313486231570e308) && !(xnorm2 <= 1.79769313486231570e308)) {
						// A non-finite update makes Y non-finite, and Y + dY can never become finite again: neither
						// test can pass any more, the reference spins to maxit.  Account the remaining evaluations
						// and stop.  (|R|^2 alone may overflow with R and Y finite -- then the iteration goes on.)
						const int rest = maxit - 1 - it;
						if (rest > 0) {
							c_fun += (unsigned)rest * S;
							if (a.refresh_jac > 0) c_jac += (unsigned)((maxit - 1) / a.refresh_jac - it / a.refresh_jac);
						}
						it = maxit;
						return true;
					}
					return false;
				};
				// The iteration loop exists twice: the common instance knows at compile time that no row was
				// interchanged; the rare one permutes the right-hand sides.  (One loop with a run-time test
				// costs ~25 register moves per iteration around the skipped block.)
				auto newton_loop = [&](auto swaps_tag) { // irk.hpp:458 (cap == maxit unless this is a fast trip)
					if (REHUEL_PEEL_FIRST && kAuto) {
						if (!(it < cap)) return;
						if (iteration(swaps_tag, std::true_type())) return;
						++it;
					}
					for (; it < cap; ++it)
						if (iteration(swaps_tag, std::false_type())) return;
				};
				if (piv_any) newton_loop(std::true_type());
				else newton_loop(std::false_type());

				__syncwarp(amask);
#if defined(REHUEL_COUNT_SKIPS) && REHUEL_COUNT_SKIPS == 3
				c_skips += __reduce_max_sync(amask, (unsigned)it); // dev builds (mode 3): iterations the WARP ran in this lane's attempts
#endif

				if (nstat != kNewtonSuccess && it < maxit) {
					// fast-trip budget exhausted: forget this try (counters back to the attempt's
					// start) and wait for a slow trip
					// (this try evaluated S residual stages before and in each of its cap-1 iterations,
					// one Jacobian, and one no-op refresh per refresh_jac iterations)
					c_fun -= (unsigned)(S * cap);
					c_jac -= 1u + (a.refresh_jac > 0 ? (unsigned)((cap - 1) / a.refresh_jac) : 0u);
					fl.parked = true;
				} else if (nstat != kNewtonSuccess) { // irk.hpp:634-665
					fl.parked = false;
					if (io.trace) trace_row(a, mem, total_step(), t, dt, -1.0, it);
					if (!a.adaptive) {
						status = REHUEL_GENERAL_ERROR | kAborted;
						finished = true;
					} else {
						dt *= 0.7;
						if (since_relax > 15u) { // step - last_maxit_relax_step > 15
							maxit += a.maxit0;
							since_relax = 0u;
						}
						++c_rej_newton;
					}
				} else {
					fl.parked = false;
					c_iters += (unsigned)it;
					// -------- new solution + embedded error, irk.hpp:691-731
					const double gam = tb.gamma * dt;
					double dy[N], dalt[N], fy[N], e[N], yn[N];
#pragma unroll
					for (int k = 0; k < N; ++k) {
						double s1 = 0.0, s2 = 0.0;
#pragma unroll
						for (int i = 0; i < S; ++i) {
							if (!REHUEL_CANONICAL_TABLEAU) s1 = fma(Y[i][k], tb.d[i], s1);
							s2 = fma(Y[i][k], tb.d2[i], s2);
						}
						dy[k] = REHUEL_CANONICAL_TABLEAU ? Y[S - 1][k] : s1; // d = A^-T b = e_s: stiffly accurate (Tableau::canonical)
						dalt[k] = s2;
					}
					if (EST != EST_IDENTITY) { // irk.hpp:702 (multiplied by gam == 0 when EST_IDENTITY)
						if (kAuto >= 2) {
#pragma unroll
							for (int k = 0; k < N; ++k) fy[k] = f0[k];
						} else {
							F::fun(t, y, p, fy);
						}
					}
					++c_fun;
#pragma unroll
					for (int k = 0; k < N; ++k) {
						yn[k] = y[k] + dy[k];
						const double dyalt = (EST != EST_IDENTITY) ? fma(gam, fy[k], dalt[k]) : dalt[k];
						e[k] = dyalt - dy[k]; // delta_delta, irk.hpp:707
					}
					// err_est = dt * E^-1 * (.), E = I - gam*J = gam*((1/gam) I - J), so
					// dt*E^-1 = (dt/gam) ((1/gam) I - J)^-1 = (1/gamma) (mu I - J)^-1
					auto apply_Einv_dt = [&](double (&v)[N]) {
						if (EST != EST_IDENTITY && piv_any) apply_row_swaps(pivs, EST == EST_REUSE ? 0 : 4, v);
						if (EST == EST_REUSE) {
							if (ZC != 0u && !piv_any) lu_solve<ZC>(lur, v);
							else lu_solve(lur, v);
						} else if (EST == EST_OWN_LU) {
							if (ZC != 0u && !piv_any) lu_solve<ZC>(lue, v);
							else lu_solve(lue, v);
						}
						const double sc = (EST == EST_IDENTITY) ? dt : tb.rgamma;
#pragma unroll
						for (int k = 0; k < N; ++k) v[k] *= sc;
					};
					apply_Einv_dt(e); // 8.19, irk.hpp:718
					if (fl.alt) {        // 8.20, irk.hpp:722-731
						double y2[N], f2[N];
#pragma unroll
						for (int k = 0; k < N; ++k) y2[k] = y[k] + e[k];
						if (EST != EST_IDENTITY) F::fun(t, y2, p, f2);
						++c_fun;
#pragma unroll
						for (int k = 0; k < N; ++k) {
							const double v = (EST != EST_IDENTITY) ? fma(gam, f2[k], dalt[k]) : dalt[k];
							e[k] = v - dy[k];
						}
						apply_Einv_dt(e);
					}
					// scaled RMS norm, irk.hpp:733-754
					double tot = 0.0;
#pragma unroll
					for (int k = 0; k < N; ++k) {
						const double sc = fma(TOLV ? m_rtol : a.rtol, max_abs(y[k], yn[k]), TOLV ? m_atol : a.atol);
						const double q = e[k] * fast_rcp(sc);
						tot = fma(q, q, tot);
					}
					// `err` holds the SQUARE of the error norm with REHUEL_ROOT_CONTROLLER (the root is only taken where the
					// norm itself goes out: trace, samples, err_last); err > 1 <=> err^2 > 1, the floor 1e-17 becomes 1e-34
					if (REHUEL_ROOT_CONTROLLER) {
						err = tot * (1.0 / (double)N);
						if (err < kMachinePrecision * kMachinePrecision) err = kMachinePrecision * kMachinePrecision;
					} else {
						err = sqrt(tot * (1.0 / (double)N));
						if (err < kMachinePrecision) err = kMachinePrecision;
					}
					const bool rejected = a.adaptive && (err > 1.0); // irk.hpp:761
					if (rejected) {
						fl.alt = true;
						++c_rej_err;
					}
					// -------- step-size proposal, irk.hpp:770-801 (shifts the error history, irk.hpp:756-758)
					const double new_dt = REHUEL_ROOT_CONTROLLER
					                          ? propose_dt_sq<static_min_order<S, NR, NC, EST>::value>(dt, err, q_prev, dts0, dts1, maxit, it, tb.min_order, tb.expo, a.max_dt)
					                          : propose_dt(dt, err, q_prev, dts0, dts1, maxit, it, tb.min_order, tb.expo, a.max_dt);

					if (io.trace) trace_row(a, mem, total_step(), t, dt, REHUEL_ROOT_CONTROLLER ? sqrt(err) : err, it); // irk.hpp:805-810
					if (!rejected) { // irk.hpp:813-846
						if (DENSE) { // grid times inside (t, t + dt] from this step's collocation polynomial
							const double t_new = t + dt;
							while (kd < io.n_dense) {
								const double tk = fma((double)kd, io.dense_dt, io.dense_t0);
								if (tk > t_new) break;
								double w[S];
								dense_weights<S, NC>(tb, (tk - t) * rdt, w);
#pragma unroll
								for (int k = 0; k < N; ++k) {
									double acc = y[k];
#pragma unroll
									for (int i = 0; i < S; ++i) acc = fma(w[i], Y[i][k], acc);
									io.y_dense[((size_t)kd * N + k) * M + mem] = (tk == t_new) ? yn[k] : acc;
								}
								++kd;
							}
						}
#pragma unroll
						for (int k = 0; k < N; ++k) y[k] = yn[k];
						const double t_old = t;
						t += dt;
						++steps;
						if (since_relax < 0xffffffffu) ++since_relax;
						if (t == t_old) { // dt below ulp(t): the reference would accept steps forever
							status = REHUEL_GENERAL_ERROR | REHUEL_DT_TOO_SMALL;
							finished = true;
						}
						if (a.store_samples && (a.output_interval == 1ull || (unsigned long long)total_step() % a.output_interval == 0ull)) {
							if (n_stored < io.n_samples_cap) {
								const size_t row = n_stored;
								io.t_samples[row * M + mem] = t;
								if (io.err_samples) io.err_samples[row * M + mem] = REHUEL_ROOT_CONTROLLER ? sqrt(err) : err;
#pragma unroll
								for (int k = 0; k < N; ++k) io.y_samples[(row * N + k) * M + mem] = y[k];
							}
							++n_stored;
						}
						fl.alt = false;
						if (!(t < a.t1)) finished = true;
					}
					if (a.adaptive) dt = new_dt; // irk.hpp:850-852
					dts1 = dts0;                 // irk.hpp:853-855 (dts[2] is never read)
					dts0 = dt;
				}
			}
		}

		if (finished) fl.done = true; // results are written when the warp next stops for service (top of the loop)
	}
	__syncthreads(); // every warp of the CTA has left the loop
	if (io.stats && threadIdx.x < REHUEL_NSTATS && s_stats[threadIdx.x]) {
		if (threadIdx.x == REHUEL_NCOUNTERS + 1) atomicMax(io.stats + threadIdx.x, s_stats[threadIdx.x]);
		else atomicAdd(io.stats + threadIdx.x, s_stats[threadIdx.x]);
	}
}

} // namespace rehuel
