// Warp-per-system ensemble integrator for SMALL DENSE systems (4 < n <= 16, Jacobian not banded): one warp runs
// one member's adaptive integration, persistent warps pull members from the global work queue.
//
// Same algorithm, statement for statement, as the other kernels (irk_thread.cuh, irk_warp.cuh, irk_cta.cuh) and hence
// as the reference loop irk.hpp:604-860 / newton_solve_stages irk.hpp:398-493, and the same frame as the band kernel
// (irk_warp.cuh): stage increments, residual, transforms, update and estimator right-hand side in registers,
// component i on lane i; y and one set of work vectors in shared memory.  What differs is the linear algebra.  A
// dense n x n matrix with n <= 16 is small enough to live in REGISTERS, row i on lane i:
//   * LU with row partial pivoting (LAPACK's choice of pivot: first row of maximal |re|+|im|) is n unrolled steps of
//     a segmented shuffle butterfly (pivot search), a shuffle per entry of the pivot row (broadcast) and one FMA per
//     lane and entry; interchanges swap whole register rows between two lanes and are skipped when no matrix needs one;
//   * a solve is a gather of the right-hand side through the interchange list and 2n broadcast-and-FMA steps;
//   * a matrix only occupies a SEGMENT of 8 or 16 lanes, so the 2..4 matrices of an attempt (real block, complex
//     blocks, the estimator's own) sit side by side in the warp and are factorised / solved by the same
//     instructions (one routine for real and complex: a real matrix has a zero imaginary plane).
// No shared-memory matrix, no block barrier: an attempt costs a few thousand warp-instructions, where the
// CTA-per-system kernel (which these systems went through before) spends ~140 us of a whole SM.
//
// Functor concept: the componentwise one of irk_cta.cuh.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "../../include/rehuel_b200.h"
#include "irk_cta.cuh"
#include "irk_thread.cuh"
#include "irk_warp.cuh"
#include "tableau.hpp"

#ifndef REHUEL_WD_ROLLED
#define REHUEL_WD_ROLLED 1
#endif
#ifndef REHUEL_WD_MINB8
#define REHUEL_WD_MINB8 12
#endif
#ifndef REHUEL_WD_MINB16
#define REHUEL_WD_MINB16 12
#endif

namespace rehuel {

template <class F, int S, int NC, int EST>
struct WdSmem {
	static constexpr int N = F::N;
	static_assert(N > 4 && N <= 16, "the register-row LU keeps one matrix row per lane in segments of 8 or 16 lanes");
	static constexpr int SEG = (N <= 8) ? 8 : 16;                 // lanes per matrix
	static constexpr int NSEG = 32 / SEG;                         // matrices side by side
	static constexpr int NMAT = (S - 2 * NC) + NC + (EST == EST_OWN_LU ? 1 : 0);
	static constexpr int PASSES = (NMAT + NSEG - 1) / NSEG;
	static constexpr int PA = ((F::P > 0 ? F::P : 1) + 1) & ~1;
	static constexpr int o_y = 0;
	static constexpr int o_p = o_y + N;
	static constexpr int o_w = o_p + PA;                          // work vectors [S][N]
	static constexpr int total = o_w + (S > 1 ? S : 2) * N;
	static constexpr size_t bytes = (size_t)total * 8;
};

// ------------------------------------------------------------------ dense LU in registers, one lane per row
// ar/ai: this lane's row (lane wi = lane % SEG of the segment holds row wi; lanes wi >= N idle).  In place, LAPACK's
// packed result in LAPACK's row order: U(k, k+1..) on lane k, the multipliers L(i, k) on lane i at [k], the
// RECIPROCAL pivot on lane k at [k].  rsrc: the original index of the row that ended up on this lane.
template <int N, int SEG>
__device__ __forceinline__ void wd_factor(double (&ar)[N], double (&ai)[N], bool active, int &rsrc)
{
	const int wi = threadIdx.x % SEG;
	rsrc = wi;
#pragma unroll
	for (int k = 0; k < N; ++k) {
		// ---- pivot search among the rows >= k (izamax: first row of maximal |re|+|im|)
		double v = (active && wi >= k && wi < N) ? fabs(ar[k]) + fabs(ai[k]) : -1.0;
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
		if (__any_sync(0xffffffffu, p != k)) { // some matrix exchanges rows k and p: whole rows, as dgetrf does
			const int src = (wi == k) ? p : (wi == p ? k : wi);
#pragma unroll
			for (int j = 0; j < N; ++j) {
				ar[j] = __shfl_sync(0xffffffffu, ar[j], src, SEG);
				ai[j] = __shfl_sync(0xffffffffu, ai[j], src, SEG);
			}
			rsrc = __shfl_sync(0xffffffffu, rsrc, src, SEG);
		}
		// ---- reciprocal pivot 1/(a+ib) = (a-ib)/(a^2+b^2), multipliers, rank-1 update
		const double pr = __shfl_sync(0xffffffffu, ar[k], k, SEG), pi = __shfl_sync(0xffffffffu, ai[k], k, SEG);
		const double rden = fast_rcp(fma(pr, pr, pi * pi));
		const double rr = pr * rden, ri = -pi * rden;
		const bool below = wi > k && wi < N;
		const double lr = below ? fma(ar[k], rr, -ai[k] * ri) : 0.0;
		const double li = below ? fma(ar[k], ri, ai[k] * rr) : 0.0;
		if (below) {
			ar[k] = lr;
			ai[k] = li;
		}
		if (wi == k) {
			ar[k] = rr;
			ai[k] = ri;
		}
#pragma unroll
		for (int j = k + 1; j < N; ++j) {
			const double ur = __shfl_sync(0xffffffffu, ar[j], k, SEG), ui = __shfl_sync(0xffffffffu, ai[j], k, SEG);
			ar[j] = fma(-lr, ur, fma(li, ui, ar[j])); // lr = li = 0 on the rows that take no part
			ai[j] = fma(-lr, ui, fma(-li, ur, ai[j]));
		}
	}
}

// The same factorisation as a ROLLED loop.  Static register indexing forces wd_factor to unroll its N steps, and the
// kernels built on it are straight-line code that every warp streams through once per attempt: ncu shows them bound
// by instruction FETCH (`no_instruction` 2.3 - 3.7 stalls per issue at 32 - 40 % issue-slot use), not by arithmetic.
// Here the register row is ROTATED by one slot per step, so that the pivot column always sits in slot 0 and column
// k + j in slot j: one loop body serves every step (N times less code for the LU), at the price of 2N register moves
// per step and plane; after N steps the row is back in place, packed exactly like wd_factor's.  Same operations on
// the same numbers, so the factors are bit-identical.  Measured (n = 16, one member per warp): RADAU_IIA_53 40 -> 23.5 ms
// per 16 384 members; LOBATTO_IIIC_85, where four matrices make the LU the bulk of the executed instructions and the
// kernel spills, 77 -> 101 ms -- so the kernels roll the LU for the methods with up to three stages only.
template <int N, int SEG>
__device__ __forceinline__ void wd_factor_rolled(double (&ar)[N], double (&ai)[N], bool active, int &rsrc)
{
	const int wi = threadIdx.x % SEG;
	rsrc = wi;
#pragma unroll 1
	for (int k = 0; k < N; ++k) {
		double v = (active && wi >= k && wi < N) ? fabs(ar[0]) + fabs(ai[0]) : -1.0;
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
		if (__any_sync(0xffffffffu, p != k)) { // every lane holds the same rotation, so whole rows can still be exchanged
			const int src = (wi == k) ? p : (wi == p ? k : wi);
#pragma unroll
			for (int j = 0; j < N; ++j) {
				ar[j] = __shfl_sync(0xffffffffu, ar[j], src, SEG);
				ai[j] = __shfl_sync(0xffffffffu, ai[j], src, SEG);
			}
			rsrc = __shfl_sync(0xffffffffu, rsrc, src, SEG);
		}
		const double pr = __shfl_sync(0xffffffffu, ar[0], k, SEG), pi = __shfl_sync(0xffffffffu, ai[0], k, SEG);
		const double rden = fast_rcp(fma(pr, pr, pi * pi));
		const double rr = pr * rden, ri = -pi * rden;
		const bool below = wi > k && wi < N;
		const double lr = below ? fma(ar[0], rr, -ai[0] * ri) : 0.0;
		const double li = below ? fma(ar[0], ri, ai[0] * rr) : 0.0;
		const double kr = below ? lr : (wi == k ? rr : ar[0]), ki = below ? li : (wi == k ? ri : ai[0]); // column k as it is kept
#pragma unroll
		for (int j = 1; j < N; ++j) {
			if (j < N - k) { // warp-uniform: slot j holds column k + j
				const double ur = __shfl_sync(0xffffffffu, ar[j], k, SEG), ui = __shfl_sync(0xffffffffu, ai[j], k, SEG);
				ar[j] = fma(-lr, ur, fma(li, ui, ar[j]));
				ai[j] = fma(-lr, ui, fma(-li, ur, ai[j]));
			}
		}
#pragma unroll
		for (int j = 0; j + 1 < N; ++j) {
			ar[j] = ar[j + 1];
			ai[j] = ai[j + 1];
		}
		ar[N - 1] = kr;
		ai[N - 1] = ki;
	}
}

template <int N, int SEG>
__device__ __forceinline__ void wd_factor_real_rolled(double (&ar)[N], bool active, int &rsrc)
{
	const int wi = threadIdx.x % SEG;
	rsrc = wi;
#pragma unroll 1
	for (int k = 0; k < N; ++k) {
		double v = (active && wi >= k && wi < N) ? fabs(ar[0]) : -1.0;
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
		const double rr = fast_rcp(__shfl_sync(0xffffffffu, ar[0], k, SEG));
		const bool below = wi > k && wi < N;
		const double l = below ? ar[0] * rr : 0.0;
		const double kr = below ? l : (wi == k ? rr : ar[0]);
#pragma unroll
		for (int j = 1; j < N; ++j)
			if (j < N - k) ar[j] = fma(-l, __shfl_sync(0xffffffffu, ar[j], k, SEG), ar[j]);
#pragma unroll
		for (int j = 0; j + 1 < N; ++j) ar[j] = ar[j + 1];
		ar[N - 1] = kr;
	}
}

// x <- A^-1 x with the factors above; x in shared memory (xi only touched for a complex matrix).  Every lane of the
// warp runs the shuffles; `active` says whether this lane's segment has a right-hand side at all.
template <int N, int SEG>
__device__ __forceinline__ void wd_solve(const double (&ar)[N], const double (&ai)[N], bool is_c, int rsrc, double *xr, double *xi,
                                         bool active)
{
	const int wi = threadIdx.x % SEG;
	const bool mine = active && wi < N;
	double vr = mine ? xr[rsrc] : 0.0, vi = (mine && is_c) ? xi[rsrc] : 0.0; // P x
	__syncwarp(); // every lane has its entry before any lane stores
#pragma unroll
	for (int k = 0; k + 1 < N; ++k) { // forward, unit lower factor
		const double kr = __shfl_sync(0xffffffffu, vr, k, SEG), ki = __shfl_sync(0xffffffffu, vi, k, SEG);
		if (wi > k) {
			vr = fma(-ar[k], kr, fma(ai[k], ki, vr));
			vi = fma(-ar[k], ki, fma(-ai[k], kr, vi));
		}
	}
#pragma unroll
	for (int k = N - 1; k >= 0; --k) { // backward; lane k holds the reciprocal pivot at [k]
		const double tr = fma(vr, ar[k], -vi * ai[k]), ti = fma(vr, ai[k], vi * ar[k]);
		const double kr = __shfl_sync(0xffffffffu, tr, k, SEG), ki = __shfl_sync(0xffffffffu, ti, k, SEG);
		if (wi == k) {
			vr = kr;
			vi = ki;
		}
		if (wi < k) {
			vr = fma(-ar[k], kr, fma(ai[k], ki, vr));
			vi = fma(-ar[k], ki, fma(-ai[k], kr, vi));
		}
	}
	if (mine) {
		xr[wi] = vr;
		if (is_c) xi[wi] = vi;
	}
}

// ------------------------------------------------------------------ the kernel: one warp per block
template <class F, int S, int NR, int NC, int EST, bool DENSE = false>
__global__ void __launch_bounds__(32, (F::N <= 8) ? REHUEL_WD_MINB8 : REHUEL_WD_MINB16)
ensemble_irk_wdense(const __grid_constant__ KernelArgs a, const __grid_constant__ DevTableau<S, NC> tb)
{
	static_assert(S == NR + 2 * NC && NR <= 1, "eigen-structure");
	using L = WdSmem<F, S, NC, EST>;
	constexpr int N = F::N, SEG = L::SEG, NSEG = L::NSEG, NMAT = L::NMAT, PASSES = L::PASSES;
	constexpr int NPL = 1;        // components per lane: i = lane (N <= 16)
	constexpr bool FULL = false;  // lanes >= N own nothing
	static_assert(EST != EST_REUSE || NR == 1, "E's factors are the real Newton block's");
	extern __shared__ double sm[];
	double *y = sm + L::o_y, *par = sm + L::o_p, *work = sm + L::o_w;
	const int lane = threadIdx.x;
	const unsigned long long M = a.M;
	const rehuel_device_io &io = a.io;
	// ---- matrix roles: segment `seg` of pass `ps` holds matrix m = ps*NSEG + seg of the list
	//      [real Newton block (if NR)] [complex blocks 0..NC-1] [the estimator's own (if EST_OWN_LU)], row wi on lane wi of the segment
	const int seg = lane / SEG, wi = lane % SEG;
	constexpr int M_REAL = 0, M_CPLX = 1, M_EST = 2, M_NONE = 3;
	int m_type[PASSES], m_c[PASSES];
#pragma unroll
	for (int ps = 0; ps < PASSES; ++ps) {
		const int m = ps * NSEG + seg;
		m_type[ps] = (m >= NMAT) ? M_NONE : ((NR && m == 0) ? M_REAL : ((m - NR < NC) ? M_CPLX : M_EST));
		m_c[ps] = (m_type[ps] == M_CPLX) ? m - NR : 0;
	}
	double fr[PASSES][N], fi[PASSES][N]; // this lane's row of the factors of its matrices
	int rsrc[PASSES];                    // which row of the right-hand side ends up here (the interchanges)

	for (;;) {
		// ---------------- next member
		unsigned long long mem = 0;
		if (lane == 0) {
			mem = atomicAdd(io.queue, 1ull);
			if (mem < M && io.order) mem = io.order[mem];
		}
		mem = __shfl_sync(0xffffffffu, mem, 0);
		if (mem >= M) break;
		double yr[NPL]; // y, component lane + 32 j; the shared copy is for the neighbours (f and J couple components)
#pragma unroll
		for (int j = 0; j < NPL; ++j) {
			const int i = lane + 32 * j;
			yr[j] = 0.0;
			if (FULL || i < N) {
				yr[j] = io.y0_broadcast ? io.y0[i] : io.y0[i * M + mem];
				y[i] = yr[j];
			}
		}
		for (int k = lane; k < F::P; k += 32) par[k] = io.params[k * M + mem];
		__syncwarp();

		// ---------------- per-member scalar state (identical in every lane)
		double t = io.t_start ? io.t_start[mem] : a.t0, dt = a.dt0;
		if (t + dt > a.t1) dt = a.t1 - t; // irk.hpp:514-520
		double dts0 = dt, dts1 = dt, q_prev = tb.q_init, err = 0.0; // q_prev: see propose_dt (irk_thread.cuh)
		bool alt = true;
		int maxit = a.maxit0;
		long long step = 0, last_relax = 0;
		int status = REHUEL_SUCCESS;
		// per-member tolerances (io.rtol/atol/newton_tol, mixed-tolerance ensembles) or the uniform ones of the options
		const double m_rtol = io.rtol ? io.rtol[mem] : a.rtol, m_atol = io.rtol ? io.atol[mem] : a.atol;
		const double m_Rtol2 = io.rtol ? io.newton_tol[mem] * io.newton_tol[mem] : a.Rtol2;
		unsigned c_attempt = 0, c_rej_newton = 0, c_rej_err = 0, c_newton_ok = 0, c_fun = 0, c_jac = 0, c_iters = 0,
		         n_stored = 0;
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
		unsigned kd = 0; // dense output: next grid time (see irk_thread.cuh)
		if (DENSE) {
			while (kd < io.n_dense && fma((double)kd, io.dense_dt, io.dense_t0) < t) ++kd;
			if (kd < io.n_dense && fma((double)kd, io.dense_dt, io.dense_t0) == t) {
#pragma unroll
				for (int j = 0; j < NPL; ++j) {
					const int i = lane + 32 * j;
					if (FULL || i < N) io.y_dense[((size_t)kd * N + i) * M + mem] = yr[j];
				}
				++kd;
			}
		}
		if (a.store_samples) {
			if (io.n_samples_cap) {
				if (lane == 0) {
					io.t_samples[mem] = t;
					if (io.err_samples) io.err_samples[mem] = 0.0;
				}
#pragma unroll
				for (int j = 0; j < NPL; ++j) {
					const int i = lane + 32 * j;
					if (FULL || i < N) io.y_samples[i * M + mem] = yr[j];
				}
			}
			n_stored = 1;
		}

		double Yr[S][NPL], Rr[S][NPL]; // stage increments and residual, component lane + 32 j of every stage
		// residual R = Y - dt*(A (x) I) F(y+Y), returns |R|^2   (irk.hpp:366-386)
		auto residual = [&]() -> double {
#pragma unroll
			for (int s = 0; s < S; ++s)
#pragma unroll
				for (int j = 0; j < NPL; ++j) {
					const int i = lane + 32 * j;
					if (FULL || i < N) work[s * N + i] = yr[j] + Yr[s][j]; // stage states, for the neighbours
				}
			__syncwarp();
			double Fr[S][NPL];
#pragma unroll
			for (int s = 0; s < S; ++s)
#pragma unroll
				for (int j = 0; j < NPL; ++j) {
					const int i = lane + 32 * j;
					Fr[s][j] = (FULL || i < N) ? F::fun(i, fma(tb.c[s], dt, t), work + s * N, par) : 0.0;
				}
			__syncwarp(); // the work vectors are free again
			c_fun += S;
			double r2 = 0.0;
#pragma unroll
			for (int s = 0; s < S; ++s)
#pragma unroll
				for (int j = 0; j < NPL; ++j) {
					double acc = 0.0;
#pragma unroll
					for (int k = 0; k < S; ++k) acc = fma(tb.A[s][k], Fr[k][j], acc);
					const double r = fma(-dt, acc, Yr[s][j]);
					Rr[s][j] = r;
					r2 = fma(r, r, r2);
				}
			return warp_sum(r2);
		};

		while (t < a.t1) { // irk.hpp:604
			if (t + dt > a.t1) dt = a.t1 - t; // irk.hpp:607-609
			++c_attempt;
			if (a.max_steps >= 0 && step > a.max_steps) { // irk.hpp:612-616
				status = REHUEL_ERROR_MAX_STEPS_EXCEEDED;
				break;
			}
			if (!(dt >= 2.2250738585072014e-308)) { // see irk_thread.cuh: the reference would never return
				status = REHUEL_GENERAL_ERROR | REHUEL_DT_TOO_SMALL;
				break;
			}
			// -------- Jacobian row wi, the rows of this lane's matrices  mu I - J  (irk.hpp:422-436)
			const double rdt = 1.0 / dt;
			const double mur = (EST == EST_REUSE) ? 1.0 / (tb.gamma * dt) : tb.gamma_hat * rdt;
			const double mue = (EST == EST_OWN_LU) ? 1.0 / (tb.gamma * dt) : 0.0;
			{
				double row[N];
#pragma unroll
				for (int j = 0; j < N; ++j) row[j] = 0.0;
				if (wi < N) F::jac_row(wi, t, y, par, row);
#pragma unroll
				for (int ps = 0; ps < PASSES; ++ps) {
					double dr = 0.0, di = 0.0; // the shift mu of this matrix
					if (m_type[ps] == M_REAL) dr = mur;
					if (m_type[ps] == M_EST) dr = mue;
#pragma unroll
					for (int c = 0; c < NC; ++c)
						if (m_type[ps] == M_CPLX && m_c[ps] == c) {
							dr = tb.alpha[c] * rdt;
							di = tb.beta[c] * rdt;
						}
#pragma unroll
					for (int j = 0; j < N; ++j) {
						fr[ps][j] = (j == wi ? dr : 0.0) - row[j];
						fi[ps][j] = (j == wi ? di : 0.0);
					}
				}
			}
			++c_jac;
			// -------- all factorisations: NSEG matrices side by side per pass, row partial pivoting (LAPACK's choice)
#pragma unroll
			for (int ps = 0; ps < PASSES; ++ps) {
				if (REHUEL_WD_ROLLED && S <= 3) wd_factor_rolled<N, SEG>(fr[ps], fi[ps], m_type[ps] != M_NONE, rsrc[ps]);
				else wd_factor<N, SEG>(fr[ps], fi[ps], m_type[ps] != M_NONE, rsrc[ps]);
			}

			// -------- simplified Newton (irk.hpp:398-493)
#pragma unroll
			for (int s = 0; s < S; ++s)
#pragma unroll
				for (int j = 0; j < NPL; ++j) Yr[s][j] = 0.0; // irk.hpp:415
			double Rnorm2 = residual();
			int nstat = kNewtonMaxit;
			int it = 1;
			for (; it < maxit; ++it) { // irk.hpp:458
#pragma unroll
				for (int s = 0; s < S; ++s)
#pragma unroll
					for (int j = 0; j < NPL; ++j) { // Z = (Ti (x) I) R
						const int i = lane + 32 * j;
						double acc = 0.0;
#pragma unroll
						for (int k = 0; k < S; ++k) acc = fma(tb.Ti[s][k], Rr[k][j], acc);
						if (FULL || i < N) work[s * N + i] = acc;
					}
				__syncwarp();
#pragma unroll
				for (int ps = 0; ps < PASSES; ++ps) { // the right-hand sides of a pass side by side, like the factorisations
					double *xr = work + (m_type[ps] == M_CPLX ? (NR + 2 * m_c[ps]) * N : 0);
					wd_solve<N, SEG>(fr[ps], fi[ps], m_type[ps] == M_CPLX, rsrc[ps], xr, xr + N,
					                 m_type[ps] == M_REAL || m_type[ps] == M_CPLX);
				}
				__syncwarp();
				double z[NPL][S]; // W = -mu Z, componentwise
#pragma unroll
				for (int j = 0; j < NPL; ++j) {
					const int i = lane + 32 * j;
#pragma unroll
					for (int k = 0; k < S; ++k) z[j][k] = (FULL || i < N) ? work[k * N + i] : 0.0;
					if (NR) z[j][0] *= -mur;
#pragma unroll
					for (int c = 0; c < NC; ++c) { // w = -(mr + i mi) w
						const double mr = tb.alpha[c] * rdt, mi = tb.beta[c] * rdt;
						const double wr = z[j][NR + 2 * c], wi = z[j][NR + 2 * c + 1];
						z[j][NR + 2 * c] = fma(-mr, wr, mi * wi);
						z[j][NR + 2 * c + 1] = fma(-mr, wi, -mi * wr);
					}
				}
				double x2 = 0.0;
#pragma unroll
				for (int s = 0; s < S; ++s)
#pragma unroll
					for (int j = 0; j < NPL; ++j) { // dY = (T (x) I) W ; Y += dY
						double acc = 0.0;
#pragma unroll
						for (int k = 0; k < S; ++k) acc = fma(tb.T[s][k], z[j][k], acc);
						x2 = fma(acc, acc, x2);
						Yr[s][j] += acc;
					}
				const double xnorm2 = warp_sum(x2); // irk.hpp:466
				if (REHUEL_SKIP_DEAD_RESIDUAL && xnorm2 < a.xtol2) { // the residual of irk.hpp:469-472 would never be looked at
					c_fun += S;                                       // (see irk_thread.cuh); its evaluations are booked
					nstat = kNewtonSuccess;
					break;
				}
				Rnorm2 = residual();
				if (Rnorm2 < m_Rtol2) { nstat = kNewtonSuccess; break; }
				if (xnorm2 < a.xtol2) { nstat = kNewtonSuccess; break; }
				if (a.refresh_jac > 0 && it % a.refresh_jac == 0) ++c_jac;
				if (!(Rnorm2 <= 1.79769313486231570e308) && !(xnorm2 <= 1.79769313486231570e308)) { // non-finite iterate: see irk_thread.cuh
					const int rest = maxit - 1 - it;
					if (rest > 0) {
						c_fun += (unsigned)rest * S;
						if (a.refresh_jac > 0) c_jac += (unsigned)((maxit - 1) / a.refresh_jac - it / a.refresh_jac);
					}
					it = maxit;
					break;
				}
			}

			if (nstat != kNewtonSuccess) { // irk.hpp:634-665
				if (lane == 0) trace_row(a, mem, step, t, dt, -1.0, it);
				if (!a.adaptive) {
					status = REHUEL_GENERAL_ERROR;
					break;
				}
				dt *= 0.7;
				if (step - last_relax > 15) {
					maxit += a.maxit0;
					last_relax = step;
				}
				++c_rej_newton;
				continue;
			}
			++c_newton_ok;
			c_iters += (unsigned)it;

			// -------- new solution + embedded error (irk.hpp:691-731), componentwise in registers
			const double gam = tb.gamma * dt;
			double dyv[NPL], dalt[NPL], yn[NPL], ev[NPL];
#pragma unroll
			for (int j = 0; j < NPL; ++j) {
				const int i = lane + 32 * j;
				double s1 = 0.0, s2 = 0.0;
#pragma unroll
				for (int s = 0; s < S; ++s) {
					s1 = fma(Yr[s][j], tb.d[s], s1);
					s2 = fma(Yr[s][j], tb.d2[s], s2);
				}
				dyv[j] = s1;
				dalt[j] = s2;
				yn[j] = yr[j] + s1;
				double dyalt = s2;
				if (EST != EST_IDENTITY && (FULL || i < N)) dyalt = fma(gam, F::fun(i, t, y, par), s2); // irk.hpp:702
				ev[j] = dyalt - s1;
			}
			++c_fun;
			const double esc = (EST == EST_IDENTITY) ? dt : 1.0 / tb.gamma; // dt*E^-1 = (1/gamma)(mu I - J)^-1
			auto apply_Einv_dt = [&]() {
				if (EST != EST_IDENTITY) {
#pragma unroll
					for (int j = 0; j < NPL; ++j) {
						const int i = lane + 32 * j;
						if (FULL || i < N) work[i] = ev[j];
					}
					__syncwarp();
#pragma unroll
					for (int ps = 0; ps < PASSES; ++ps) {
						constexpr int mE = (EST == EST_REUSE) ? 0 : NMAT - 1; // E's factors: the real Newton block's or its own
						if (ps == mE / NSEG) wd_solve<N, SEG>(fr[ps], fi[ps], false, rsrc[ps], work, work, seg == mE % NSEG);
					}
					__syncwarp();
#pragma unroll
					for (int j = 0; j < NPL; ++j) {
						const int i = lane + 32 * j;
						ev[j] = (FULL || i < N) ? work[i] : 0.0;
					}
				}
#pragma unroll
				for (int j = 0; j < NPL; ++j) ev[j] *= esc;
			};
			apply_Einv_dt(); // 8.19
			if (alt) {       // 8.20, irk.hpp:722-731
				if (EST != EST_IDENTITY) {
					static_assert(EST == EST_IDENTITY || S >= 2, "the trial state of 8.20 uses the second work vector");
#pragma unroll
					for (int j = 0; j < NPL; ++j) {
						const int i = lane + 32 * j;
						if (FULL || i < N) work[N + i] = yr[j] + ev[j];
					}
					__syncwarp();
#pragma unroll
					for (int j = 0; j < NPL; ++j) {
						const int i = lane + 32 * j;
						const double fy = (FULL || i < N) ? F::fun(i, t, work + N, par) : 0.0;
						ev[j] = fma(gam, fy, dalt[j]) - dyv[j];
					}
				} else {
#pragma unroll
					for (int j = 0; j < NPL; ++j) ev[j] = dalt[j] - dyv[j];
				}
				++c_fun;
				apply_Einv_dt();
			}
			double tot = 0.0;
#pragma unroll
			for (int j = 0; j < NPL; ++j) { // irk.hpp:733-746
				const int i = lane + 32 * j;
				const double sc = fma(m_rtol, fmax(fabs(yr[j]), fabs(yn[j])), m_atol);
				const double q = (FULL || i < N) ? ev[j] / sc : 0.0;
				tot = fma(q, q, tot);
			}
			tot = warp_sum(tot);
			err = sqrt(tot / (double)N);
			if (err < kMachinePrecision) err = kMachinePrecision;
			const bool rejected = a.adaptive && (err > 1.0); // irk.hpp:761
			if (rejected) {
				alt = true;
				++c_rej_err;
			}
			// -------- step-size proposal (irk.hpp:770-801; shifts the error history, irk.hpp:756-758)
			const double new_dt = propose_dt(dt, err, q_prev, dts0, dts1, maxit, it, tb.min_order, tb.expo, a.max_dt);
			if (lane == 0) trace_row(a, mem, step, t, dt, err, it); // irk.hpp:805-810
			bool stuck = false;
			if (!rejected) { // irk.hpp:813-846
				if (DENSE) { // grid times inside (t, t + dt] from this step's collocation polynomial
					const double t_new = t + dt;
					while (kd < io.n_dense) {
						const double tk = fma((double)kd, io.dense_dt, io.dense_t0);
						if (tk > t_new) break;
						double w[S];
						dense_weights<S, NC>(tb, (tk - t) * rdt, w);
#pragma unroll
						for (int j = 0; j < NPL; ++j) {
							const int i = lane + 32 * j;
							double acc = yr[j];
#pragma unroll
							for (int s = 0; s < S; ++s) acc = fma(w[s], Yr[s][j], acc);
							if (FULL || i < N) io.y_dense[((size_t)kd * N + i) * M + mem] = (tk == t_new) ? yn[j] : acc;
						}
						++kd;
					}
				}
#pragma unroll
				for (int j = 0; j < NPL; ++j) {
					const int i = lane + 32 * j;
					yr[j] = yn[j];
					if (FULL || i < N) y[i] = yn[j];
				}
				const double t_old = t;
				t += dt;
				++step;
				if (t == t_old) stuck = true;
				if (a.store_samples && (a.output_interval == 1ull || (unsigned long long)step % a.output_interval == 0ull)) {
					if (n_stored < io.n_samples_cap) {
						const size_t row = n_stored;
						if (lane == 0) {
							io.t_samples[row * M + mem] = t;
							if (io.err_samples) io.err_samples[row * M + mem] = err;
						}
#pragma unroll
						for (int j = 0; j < NPL; ++j) {
							const int i = lane + 32 * j;
							if (FULL || i < N) io.y_samples[(row * N + i) * M + mem] = yn[j];
						}
					}
					++n_stored;
				}
				alt = false;
				__syncwarp();
			}
			if (a.adaptive) dt = new_dt; // irk.hpp:850-852
			dts1 = dts0;
			dts0 = dt;
			if (stuck) {
				status = REHUEL_GENERAL_ERROR | REHUEL_DT_TOO_SMALL;
				break;
			}
		}

		// ---------------- results
		__syncwarp();
#pragma unroll
		for (int j = 0; j < NPL; ++j) {
			const int i = lane + 32 * j;
			if (FULL || i < N) io.y[i * M + mem] = yr[j];
		}
		if (lane == 0) {
			io.t_final[mem] = t;
			if (io.err_last) io.err_last[mem] = err;
			io.status[mem] = status;
			const unsigned steps_here = (unsigned)(step - (io.state_in ? (long long)io.state_in[REHUEL_STATE_STEP * M + mem] : 0ll));
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
			if (io.stats) { // ensemble sums; members are few on this path, so straight to global memory
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
		__syncwarp();
	}
}

} // namespace rehuel
