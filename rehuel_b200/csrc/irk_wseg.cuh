// Segment-per-system ensemble integrator for the SMALLEST dense systems beyond the thread kernel (4 < n <= 8): FOUR
// members share a warp, each in its own segment of 8 lanes; persistent warps pull members from the global queue.
// (The segment width is a template parameter: two members per warp in 16-lane segments work as well, but at
// 8 < n <= 16 the matrices of one member already fill the warp in irk_wdense.cuh and nothing is gained, so
// launch.cuh only routes n <= 8 here.)
//
// Same algorithm, statement for statement, as the other kernels (irk_thread.cuh, irk_warp.cuh, irk_wdense.cuh,
// irk_cta.cuh) and hence as the reference loop irk.hpp:604-860 / newton_solve_stages irk.hpp:398-493; the linear
// algebra is irk_wdense.cuh's (a matrix lives in the registers of its segment, row i on lane i of the segment, LU
// and solves by segmented shuffles).  What changes is who sits next to whom: in irk_wdense.cuh the 2..4 matrices of
// ONE member sit side by side and the componentwise work (stage functions, transforms, norms) uses n of 32 lanes;
// here the same matrix of the four (two) members sits side by side and every phase uses the whole warp (real matrices go
// through a real-only routine, since a pass holds one kind of matrix).
//
// The members of a warp advance in LOCKSTEP, one attempt per trip of the outer loop, like the lanes of the thread
// kernel: every collective step (shuffles, __syncwarp) is executed by all lanes, a segment that has nothing to do in
// it (no member left, Newton iteration already converged, attempt failed) is switched off by a predicate, and a
// segment whose member is finished writes its results and takes the next member from the queue.  The Newton loop
// runs until the slowest of them has converged (or hit maxit).  Which lanes integrate a member, and next to whom,
// does not enter the arithmetic: a member's result is bit-identical wherever it runs in this kernel, and equal to
// irk_wdense.cuh's up to the rounding of the real-only LU.
//
// Not built here (launch.cuh routes those calls to irk_wdense.cuh): dense output, stored trajectory samples.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "../../include/rehuel_b200.h"
#include "irk_thread.cuh"
#include "irk_wdense.cuh"
#include "tableau.hpp"

namespace rehuel {

template <class F, int S>
struct WsSmem {
	static constexpr int N = F::N;
	static_assert(N > 4 && N <= 16, "one member per segment of 8 or 16 lanes");
	static constexpr int SEG = (N <= 8) ? 8 : 16, MPW = 32 / SEG;
	static constexpr int PA = ((F::P > 0 ? F::P : 1) + 1) & ~1;
	static constexpr int o_y = 0;
	static constexpr int o_p = o_y + N;
	static constexpr int o_w = o_p + PA;                          // work vectors [S][N]
	static constexpr int stride = o_w + (S > 1 ? S : 2) * N;      // doubles per member
	static constexpr size_t bytes = (size_t)stride * MPW * 8;
};

// ------------------------------------------------------------------ real-only twins of wd_factor / wd_solve
template <int N, int SEG>
__device__ __forceinline__ void wd_factor_real(double (&ar)[N], bool active, int &rsrc)
{
	const int wi = threadIdx.x % SEG;
	rsrc = wi;
#pragma unroll
	for (int k = 0; k < N; ++k) {
		double v = (active && wi >= k && wi < N) ? fabs(ar[k]) : -1.0;
		int bi = wi;
#pragma unroll
		for (int o = SEG / 2; o > 0; o >>= 1) {
			const double ov = __shfl_xor_sync(0xffffffffu, v, o, SEG);
			const int oi = __shfl_xor_sync(0xffffffffu, bi, o, SEG);
			if (ov > v || (ov == v && oi < bi)) {
				v = ov;
				bi = oi;
			}
		}
		const int p = active ? bi : k;
		if (__any_sync(0xffffffffu, p != k)) {
			const int src = (wi == k) ? p : (wi == p ? k : wi);
#pragma unroll
			for (int j = 0; j < N; ++j) ar[j] = __shfl_sync(0xffffffffu, ar[j], src, SEG);
			rsrc = __shfl_sync(0xffffffffu, rsrc, src, SEG);
		}
		const double rr = fast_rcp(__shfl_sync(0xffffffffu, ar[k], k, SEG));
		const bool below = wi > k && wi < N;
		const double l = below ? ar[k] * rr : 0.0;
		if (below) ar[k] = l;
		if (wi == k) ar[k] = rr;
#pragma unroll
		for (int j = k + 1; j < N; ++j) ar[j] = fma(-l, __shfl_sync(0xffffffffu, ar[j], k, SEG), ar[j]);
	}
}

template <int N, int SEG>
__device__ __forceinline__ void wd_solve_real(const double (&ar)[N], int rsrc, double *x, bool active)
{
	const int wi = threadIdx.x % SEG;
	const bool mine = active && wi < N;
	double v = mine ? x[rsrc] : 0.0; // P x
	__syncwarp();
#pragma unroll
	for (int k = 0; k + 1 < N; ++k) {
		const double xk = __shfl_sync(0xffffffffu, v, k, SEG);
		if (wi > k) v = fma(-ar[k], xk, v);
	}
#pragma unroll
	for (int k = N - 1; k >= 0; --k) {
		const double xk = __shfl_sync(0xffffffffu, v * ar[k], k, SEG);
		if (wi == k) v = xk;
		if (wi < k) v = fma(-ar[k], xk, v);
	}
	if (mine) x[wi] = v;
}

template <int SEG>
__device__ __forceinline__ double seg_sum(double v)
{
#pragma unroll
	for (int o = SEG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, SEG);
	return v;
}

// ------------------------------------------------------------------ the kernel: one warp per block, 32/SEG members per warp
template <class F, int S, int NR, int NC, int EST>
__global__ void __launch_bounds__(32, 10)
ensemble_irk_wseg(const __grid_constant__ KernelArgs a, const __grid_constant__ DevTableau<S, NC> tb)
{
	static_assert(S == NR + 2 * NC && NR <= 1, "eigen-structure");
	static_assert(EST != EST_REUSE || NR == 1, "E's factors are the real Newton block's");
	using L = WsSmem<F, S>;
	constexpr int N = F::N, SEG = L::SEG, NCA = NC > 0 ? NC : 1;
	constexpr unsigned FULLM = 0xffffffffu;
	extern __shared__ double sm[];
	const int lane = threadIdx.x, seg = lane / SEG, wi = lane % SEG;
	double *y = sm + seg * L::stride + L::o_y, *par = sm + seg * L::stride + L::o_p, *work = sm + seg * L::stride + L::o_w;
	const bool row = wi < N; // this lane owns component wi of its member and row wi of its matrices
	const unsigned long long M = a.M;
	const rehuel_device_io &io = a.io;

	// ---- per-member state, identical in the lanes of a segment
	bool has = false, exhausted = false;
	unsigned long long mem = 0;
	double yr = 0.0; // y, component wi; the shared copy is for the other lanes (f and J couple components)
	double t = 0.0, dt = 0.0, dts0 = 0.0, dts1 = 0.0, q_prev = tb.q_init, err = 0.0;
	bool alt = true;
	int maxit = a.maxit0;
	long long step = 0, last_relax = 0, step0 = 0;
	int status = REHUEL_SUCCESS;
	double m_rtol = a.rtol, m_atol = a.atol, m_Rtol2 = a.Rtol2; // per-member tolerances when io.rtol is given
	unsigned c_attempt = 0, c_rej_newton = 0, c_rej_err = 0, c_newton_ok = 0, c_fun = 0, c_jac = 0, c_iters = 0, n_stored = 0;
	// ---- this lane's rows of its member's factors
	double ar[N], cr[NCA][N], ci[NCA][N], er[N];
	int rs_r = 0, rs_c[NCA], rs_e = 0;

	for (;;) {
		// ---------------- segments without a member take the next one from the queue
		if (__any_sync(FULLM, !has && !exhausted)) {
			unsigned long long q = 0;
			if (!has && !exhausted && wi == 0) {
				q = atomicAdd(io.queue, 1ull);
				if (q < M && io.order) q = io.order[q];
			}
			q = __shfl_sync(FULLM, q, 0, SEG);
			if (!has && !exhausted) {
				if (q >= M) {
					exhausted = true;
				} else {
					has = true;
					mem = q;
					yr = row ? (io.y0_broadcast ? io.y0[wi] : io.y0[wi * M + mem]) : 0.0;
					if (row) y[wi] = yr;
					if (wi < F::P) par[wi] = io.params[wi * M + mem];
					static_assert(F::P <= SEG, "parameters are loaded by the lanes of the segment");
					t = io.t_start ? io.t_start[mem] : a.t0;
					dt = a.dt0;
					if (t + dt > a.t1) dt = a.t1 - t; // irk.hpp:514-520
					dts0 = dt;
					dts1 = dt;
					q_prev = tb.q_init;
					err = 0.0;
					alt = true;
					maxit = a.maxit0;
					step = 0;
					last_relax = 0;
					status = REHUEL_SUCCESS;
					if (io.rtol) {
						m_rtol = io.rtol[mem];
						m_atol = io.atol[mem];
						m_Rtol2 = io.newton_tol[mem] * io.newton_tol[mem];
					}
					c_attempt = c_rej_newton = c_rej_err = c_newton_ok = c_fun = c_jac = c_iters = 0;
					if (io.state_in) { // continue a stopped integration: the loop variables as they were at exit
						const double *si = io.state_in + mem;
						dt = si[REHUEL_STATE_DT * M];
						dts0 = si[REHUEL_STATE_DTS0 * M];
						dts1 = si[REHUEL_STATE_DTS1 * M];
						q_prev = si[REHUEL_STATE_ERRQ * M];
						alt = si[REHUEL_STATE_ALT * M] != 0.0;
						maxit = (int)si[REHUEL_STATE_MAXIT * M];
						step = (long long)si[REHUEL_STATE_STEP * M];
						last_relax = (long long)si[REHUEL_STATE_LAST_RELAX * M];
					}
					step0 = step;
					n_stored = a.store_samples ? 1u : 0u; // samples are only counted on this path
				}
			}
			__syncwarp();
		}
		if (!__any_sync(FULLM, has)) break;

		// ---------------- one attempt (irk.hpp:604-860) for every segment with an unfinished member
		bool on = has && (t < a.t1); // irk.hpp:604
		bool fin = has && !on;       // nothing (more) to integrate
		if (on) {
			if (t + dt > a.t1) dt = a.t1 - t; // irk.hpp:607-609
			++c_attempt;
			if (a.max_steps >= 0 && step > a.max_steps) { // irk.hpp:612-616
				status = REHUEL_ERROR_MAX_STEPS_EXCEEDED;
				fin = true;
				on = false;
			} else if (!(dt >= 2.2250738585072014e-308)) { // see irk_thread.cuh: the reference would never return
				status = REHUEL_GENERAL_ERROR | REHUEL_DT_TOO_SMALL;
				fin = true;
				on = false;
			}
		}
		// -------- Jacobian row wi, the rows of the matrices  mu I - J  (irk.hpp:422-436), factorisations
		const double rdt = 1.0 / dt;
		const double mur = (EST == EST_REUSE) ? 1.0 / (tb.gamma * dt) : tb.gamma_hat * rdt;
		const double mue = (EST == EST_OWN_LU) ? 1.0 / (tb.gamma * dt) : 0.0;
		{
			double jrow[N];
#pragma unroll
			for (int j = 0; j < N; ++j) jrow[j] = 0.0;
			if (row && on) F::jac_row(wi, t, y, par, jrow);
#pragma unroll
			for (int j = 0; j < N; ++j) {
				const double dg = (j == wi) ? 1.0 : 0.0;
				if (NR) ar[j] = dg * mur - jrow[j];
				if (EST == EST_OWN_LU) er[j] = dg * mue - jrow[j];
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					cr[c][j] = dg * (tb.alpha[c] * rdt) - jrow[j];
					ci[c][j] = dg * (tb.beta[c] * rdt);
				}
			}
		}
		if (on) ++c_jac;
		if (REHUEL_WD_ROLLED && S <= 3) { // rolled loops over rotating register rows (see wd_factor_rolled)
			if (NR) wd_factor_real_rolled<N, SEG>(ar, on, rs_r);
#pragma unroll
			for (int c = 0; c < NC; ++c) wd_factor_rolled<N, SEG>(cr[c], ci[c], on, rs_c[c]);
			if (EST == EST_OWN_LU) wd_factor_real_rolled<N, SEG>(er, on, rs_e);
		} else {
			if (NR) wd_factor_real<N, SEG>(ar, on, rs_r);
#pragma unroll
			for (int c = 0; c < NC; ++c) wd_factor<N, SEG>(cr[c], ci[c], on, rs_c[c]);
			if (EST == EST_OWN_LU) wd_factor_real<N, SEG>(er, on, rs_e);
		}

		// -------- simplified Newton (irk.hpp:398-493), the four members in lockstep
		double Yr[S], Rr[S];
		// residual R = Y - dt*(A (x) I) F(y+Y), returns |R|^2 of the segment's member   (irk.hpp:366-386)
		auto residual = [&]() -> double {
#pragma unroll
			for (int s = 0; s < S; ++s)
				if (row) work[s * N + wi] = yr + Yr[s]; // stage states, for the other lanes
			__syncwarp();
			double Fr[S];
#pragma unroll
			for (int s = 0; s < S; ++s) Fr[s] = row ? F::fun(wi, fma(tb.c[s], dt, t), work + s * N, par) : 0.0;
			__syncwarp(); // the work vectors are free again
			double r2 = 0.0;
#pragma unroll
			for (int s = 0; s < S; ++s) {
				double acc = 0.0;
#pragma unroll
				for (int k = 0; k < S; ++k) acc = fma(tb.A[s][k], Fr[k], acc);
				const double r = fma(-dt, acc, Yr[s]);
				Rr[s] = r;
				r2 = fma(r, r, r2);
			}
			return seg_sum<SEG>(r2);
		};
#pragma unroll
		for (int s = 0; s < S; ++s) Yr[s] = 0.0; // irk.hpp:415
		double Rnorm2 = residual();
		if (on) c_fun += S;
		int nstat = kNewtonMaxit;
		int it = 1;
		bool iter = on && it < maxit; // irk.hpp:458
		while (__any_sync(FULLM, iter)) {
#pragma unroll
			for (int s = 0; s < S; ++s) { // Z = (Ti (x) I) R
				double acc = 0.0;
#pragma unroll
				for (int k = 0; k < S; ++k) acc = fma(tb.Ti[s][k], Rr[k], acc);
				if (row) work[s * N + wi] = acc;
			}
			__syncwarp();
			if (NR) wd_solve_real<N, SEG>(ar, rs_r, work, iter);
#pragma unroll
			for (int c = 0; c < NC; ++c) wd_solve<N, SEG>(cr[c], ci[c], true, rs_c[c], work + (NR + 2 * c) * N, work + (NR + 2 * c + 1) * N, iter);
			__syncwarp();
			double z[S]; // W = -mu Z, componentwise
#pragma unroll
			for (int k = 0; k < S; ++k) z[k] = row ? work[k * N + wi] : 0.0;
			if (NR) z[0] *= -mur;
#pragma unroll
			for (int c = 0; c < NC; ++c) { // w = -(mr + i mi) w
				const double mr = tb.alpha[c] * rdt, mi = tb.beta[c] * rdt;
				const double wr = z[NR + 2 * c], wim = z[NR + 2 * c + 1];
				z[NR + 2 * c] = fma(-mr, wr, mi * wim);
				z[NR + 2 * c + 1] = fma(-mr, wim, -mi * wr);
			}
			double x2 = 0.0;
#pragma unroll
			for (int s = 0; s < S; ++s) { // dY = (T (x) I) W ; Y += dY
				double acc = 0.0;
#pragma unroll
				for (int k = 0; k < S; ++k) acc = fma(tb.T[s][k], z[k], acc);
				x2 = fma(acc, acc, x2);
				if (iter) Yr[s] += acc;
			}
			const double xnorm2 = seg_sum<SEG>(x2); // irk.hpp:466
			if (iter && REHUEL_SKIP_DEAD_RESIDUAL && xnorm2 < a.xtol2) { // the residual of irk.hpp:469-472 would never be looked at
				c_fun += S;                                               // (see irk_thread.cuh); its evaluations are booked
				nstat = kNewtonSuccess;
				iter = false;
			}
			if (__any_sync(FULLM, iter)) {
				const double r2 = residual();
				if (iter) {
					Rnorm2 = r2;
					c_fun += S;
				}
			}
			if (iter) {
				if (Rnorm2 < m_Rtol2 || xnorm2 < a.xtol2) {
					nstat = kNewtonSuccess;
					iter = false;
				} else {
					if (a.refresh_jac > 0 && it % a.refresh_jac == 0) ++c_jac;
					if (!(Rnorm2 <= 1.79769313486231570e308) && !(xnorm2 <= 1.79769313486231570e308)) { // non-finite iterate: see irk_thread.cuh
						const int rest = maxit - 1 - it;
						if (rest > 0) {
							c_fun += (unsigned)rest * S;
							if (a.refresh_jac > 0) c_jac += (unsigned)((maxit - 1) / a.refresh_jac - it / a.refresh_jac);
						}
						it = maxit;
						iter = false;
					} else {
						++it;
						if (!(it < maxit)) iter = false;
					}
				}
			}
		}

		const bool ok = on && nstat == kNewtonSuccess;
		if (on && !ok) { // irk.hpp:634-665
			if (wi == 0) trace_row(a, mem, step, t, dt, -1.0, it);
			if (!a.adaptive) {
				status = REHUEL_GENERAL_ERROR;
				fin = true;
			} else {
				dt *= 0.7;
				if (step - last_relax > 15) {
					maxit += a.maxit0;
					last_relax = step;
				}
				++c_rej_newton;
			}
		}
		if (ok) {
			++c_newton_ok;
			c_iters += (unsigned)it;
		}

		// -------- new solution + embedded error (irk.hpp:691-731), componentwise in registers
		const double gam = tb.gamma * dt;
		double dyv = 0.0, dalt = 0.0;
#pragma unroll
		for (int s = 0; s < S; ++s) {
			dyv = fma(Yr[s], tb.d[s], dyv);
			dalt = fma(Yr[s], tb.d2[s], dalt);
		}
		const double yn = yr + dyv;
		double ev = dalt - dyv;
		if (EST != EST_IDENTITY) ev = fma(gam, row ? F::fun(wi, t, y, par) : 0.0, dalt) - dyv; // irk.hpp:702
		if (ok) ++c_fun;
		const double esc = (EST == EST_IDENTITY) ? dt : 1.0 / tb.gamma; // dt*E^-1 = (1/gamma)(mu I - J)^-1
		auto apply_Einv_dt = [&](double v, bool active) -> double {
			if (EST != EST_IDENTITY) {
				if (row) work[wi] = v;
				__syncwarp();
				if (EST == EST_REUSE) wd_solve_real<N, SEG>(ar, rs_r, work, active);
				else wd_solve_real<N, SEG>(er, rs_e, work, active);
				__syncwarp();
				v = row ? work[wi] : 0.0;
				__syncwarp(); // before the next use of the work vectors
			}
			return v * esc;
		};
		ev = apply_Einv_dt(ev, ok); // 8.19
		const bool do_alt = ok && alt;
		if (__any_sync(FULLM, do_alt)) { // 8.20, irk.hpp:722-731
			double v2 = dalt - dyv;
			if (EST != EST_IDENTITY) {
				static_assert(EST == EST_IDENTITY || S >= 2, "the trial state of 8.20 uses the second work vector");
				if (row) work[N + wi] = yr + ev;
				__syncwarp();
				v2 = fma(gam, row ? F::fun(wi, t, work + N, par) : 0.0, dalt) - dyv;
				__syncwarp();
			}
			v2 = apply_Einv_dt(v2, do_alt);
			if (do_alt) {
				ev = v2;
				++c_fun;
			}
		}
		double tot = 0.0;
		{ // irk.hpp:733-746
			const double sc = fma(m_rtol, fmax(fabs(yr), fabs(yn)), m_atol);
			const double q = row ? ev / sc : 0.0;
			tot = seg_sum<SEG>(q * q);
		}
		if (ok) {
			err = sqrt(tot / (double)N);
			if (err < kMachinePrecision) err = kMachinePrecision;
			const bool rejected = a.adaptive && (err > 1.0); // irk.hpp:761
			if (rejected) {
				alt = true;
				++c_rej_err;
			}
			// step-size proposal (irk.hpp:770-801; shifts the error history, irk.hpp:756-758)
			const double new_dt = propose_dt(dt, err, q_prev, dts0, dts1, maxit, it, tb.min_order, tb.expo, a.max_dt);
			if (wi == 0) trace_row(a, mem, step, t, dt, err, it); // irk.hpp:805-810
			bool stuck = false;
			if (!rejected) { // irk.hpp:813-846
				yr = yn;
				if (row) y[wi] = yn;
				const double t_old = t;
				t += dt;
				++step;
				if (t == t_old) stuck = true;
				if (a.store_samples && (a.output_interval == 1ull || (unsigned long long)step % a.output_interval == 0ull)) ++n_stored;
				alt = false;
			}
			if (a.adaptive) dt = new_dt; // irk.hpp:850-852
			dts1 = dts0;
			dts0 = dt;
			if (stuck) {
				status = REHUEL_GENERAL_ERROR | REHUEL_DT_TOO_SMALL;
				fin = true;
			}
		}
		__syncwarp();
		if (has && !fin && !(t < a.t1)) fin = true;

		// ---------------- results of the members that are done
		if (fin) {
			if (row) io.y[wi * M + mem] = yr;
			if (wi == 0) {
				io.t_final[mem] = t;
				if (io.err_last) io.err_last[mem] = err;
				io.status[mem] = status;
				const unsigned steps_here = (unsigned)(step - step0);
				if (io.counters) {
					io.counters[REHUEL_CNT_ATTEMPT * M + mem] = c_attempt;
					io.counters[REHUEL_CNT_REJECT_NEWTON * M + mem] = c_rej_newton;
					io.counters[REHUEL_CNT_REJECT_ERR * M + mem] = c_rej_err;
					io.counters[REHUEL_CNT_NEWTON_SUCCESS * M + mem] = c_newton_ok;
					io.counters[REHUEL_CNT_NEWTON_MAXIT_EXCEED * M + mem] = c_rej_newton;
					io.counters[REHUEL_CNT_FUN_EVALS * M + mem] = c_fun;
					io.counters[REHUEL_CNT_JAC_EVALS * M + mem] = c_jac;
					io.counters[REHUEL_CNT_NEWTON_ITERS * M + mem] = c_iters;
					io.counters[REHUEL_CNT_STEPS * M + mem] = steps_here;
				}
				if (io.stats) { // ensemble sums
					atomicAdd(io.stats + REHUEL_CNT_ATTEMPT, (unsigned long long)c_attempt);
					atomicAdd(io.stats + REHUEL_CNT_REJECT_NEWTON, (unsigned long long)c_rej_newton);
					atomicAdd(io.stats + REHUEL_CNT_REJECT_ERR, (unsigned long long)c_rej_err);
					atomicAdd(io.stats + REHUEL_CNT_NEWTON_SUCCESS, (unsigned long long)c_newton_ok);
					atomicAdd(io.stats + REHUEL_CNT_NEWTON_MAXIT_EXCEED, (unsigned long long)c_rej_newton);
					atomicAdd(io.stats + REHUEL_CNT_FUN_EVALS, (unsigned long long)c_fun);
					atomicAdd(io.stats + REHUEL_CNT_JAC_EVALS, (unsigned long long)c_jac);
					atomicAdd(io.stats + REHUEL_CNT_NEWTON_ITERS, (unsigned long long)c_iters);
					atomicAdd(io.stats + REHUEL_CNT_STEPS, (unsigned long long)steps_here);
					if (status != 0) atomicAdd(io.stats + REHUEL_NCOUNTERS, 1ull);
					atomicMax(io.stats + REHUEL_NCOUNTERS + 1, (unsigned long long)c_attempt);
				}
				if (io.n_samples) io.n_samples[mem] = n_stored;
				if (io.state_out) {
					double *so = io.state_out + mem;
					so[REHUEL_STATE_DT * M] = dt;
					so[REHUEL_STATE_DTS0 * M] = dts0;
					so[REHUEL_STATE_DTS1 * M] = dts1;
					so[REHUEL_STATE_ERRQ * M] = q_prev;
					so[REHUEL_STATE_ALT * M] = alt ? 1.0 : 0.0;
					so[REHUEL_STATE_MAXIT * M] = (double)maxit;
					so[REHUEL_STATE_STEP * M] = (double)step;
					so[REHUEL_STATE_LAST_RELAX * M] = (double)last_relax;
				}
			}
			has = false;
		}
	}
}

} // namespace rehuel
