// Warp-per-system ensemble integrator for systems with a BANDED Jacobian (4 < n <= 64, half
// bandwidth kBand <= 4): one warp runs one member's adaptive integration, persistent warps pull
// members from the global work queue.
//
// Same algorithm, statement for statement, as the other two kernels (irk_thread.cuh,
// irk_cta.cuh) and hence as the reference loop irk.hpp:604-860 / newton_solve_stages
// irk.hpp:398-493.  What the band buys: the decoupled Newton matrices  mu I - J  keep the band of
// J, their LU (row partial pivoting, dgbtf2's pattern: L within kBand below the diagonal, U within
// 2*kBand above it) costs O(n*kBand^2) instead of O(n^3), and a factorisation or a triangular
// solve is a short SEQUENTIAL recurrence -- no work for a whole thread block.  So:
//   * state vectors (y, Y, R, Z/F) and the band factors live in the warp's slice of shared memory
//     (~35 KB for n = 64, S = 5: six independent warps per SM instead of one 256-thread CTA);
//   * vector work (stage evaluations, the (Ti (x) I) / (T (x) I) transforms, norms) is spread over
//     the 32 lanes, reductions are shuffle butterflies (identical on every lane);
//   * every matrix is factorised by ONE lane, all matrices side by side (lane m <-> matrix m),
//     with a sliding window of kBand+1 rows in registers; every right-hand side is solved by ONE
//     lane, all of them side by side, with a sliding window of the solution in registers, so the
//     recurrence runs at register latency;
//   * scalar control (dt, controller history, counters) is kept redundantly by every lane.
// For the 64-equation Brusselator (kBand = 2) with LOBATTO_IIIC_85 an attempt costs ~10 us of one
// warp instead of ~120 us of one SM.
//
// Functor concept: the componentwise one of irk_cta.cuh plus `static constexpr int kBand`.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "../../include/rehuel_b200.h"
#include "irk_cta.cuh"
#include "irk_thread.cuh"
#include "tableau.hpp"

namespace rehuel {

template <class F, int S, int NC, int EST>
struct WarpSmem {
	static constexpr int N = F::N;
	static constexpr int KB = jac_band<F>::value;
	static constexpr int W = 3 * KB + 1; // band row: KB multipliers | diagonal | 2*KB of U
	static constexpr int PA = ((F::P > 0 ? F::P : 1) + 1) & ~1;
	static constexpr int NREAL = 1 + (EST == EST_OWN_LU ? 1 : 0); // slot 0: real Newton block, slot 1: estimator
	// offsets in doubles
	static constexpr int o_y = 0;
	static constexpr int o_yn = o_y + N;
	static constexpr int o_p = o_yn + N;
	static constexpr int o_Y = o_p + PA;
	static constexpr int o_R = o_Y + S * N;
	static constexpr int o_Z = o_R + S * N;     // Z of the Newton update; doubles as F of the residual
	static constexpr int o_e = o_Z + S * N;
	static constexpr int o_dy = o_e + N;
	static constexpr int o_dalt = o_dy + N;
	static constexpr int o_fy = o_dalt + N;
	static constexpr int PL = (N + KB) * W;                   // one band plane: N rows + KB zero rows of padding
	static constexpr int o_abr = o_fy + N;                    // real band factors [NREAL][N + KB][W]
	static constexpr int o_rdr = o_abr + NREAL * PL;          // reciprocal pivots [NREAL][N]
	static constexpr int o_abc = o_rdr + NREAL * N;           // complex band factors [NC][2][N + KB][W]
	static constexpr int o_rdc = o_abc + NC * 2 * PL;         // reciprocal pivots [NC][2][N]
	static constexpr int o_piv = o_rdc + NC * 2 * N;          // interchange offsets, bytes: [NREAL + NC][N]
	static constexpr size_t bytes = (size_t)o_piv * 8 + (((size_t)(NREAL + NC) * N + 15) & ~size_t(15));
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// ------------------------------------------------------------------ band LU, one lane per matrix
// ONE routine for real and complex matrices, so that the lanes that factorise the matrices of an
// attempt side by side execute the same instructions (a real matrix is a complex one whose
// imaginary plane is never loaded or stored: `is_c` predicates those accesses).
//
// In:  ab[i*W + KB + (j-i)] = A(i,j) for |i-j| <= KB, zero elsewhere; each plane carries KB extra
//      zero rows below row N-1, so no step of the recurrence needs a bounds check.
// Out: U(k,k..k+2KB) in ab[k*W + KB ..], multipliers L(k+r,k) in ab[(k+r)*W + KB - r], reciprocal
//      pivots in rd[k].
// PIVOT = false: plain elimination; returns true if row partial pivoting WOULD have exchanged rows
//                somewhere (the caller then redoes the factorisation with PIVOT = true -- where no
//                exchange is needed the two perform the same operations);
// PIVOT = true:  dgbtf2's pivoting: only the not yet eliminated part of the rows is exchanged, the
//                exchange offsets p-k go to piv[k] and the solves apply them step by step.
template <int N, int KB, bool PIVOT>
__device__ __forceinline__ bool band_lu(double *abr, double *abi, bool is_c, double *rdr, double *rdi, signed char *piv)
{
	// without interchanges U keeps the band of A (KB superdiagonals); with them it fills in up to 2*KB
	constexpr int W = 3 * KB + 1, WW = (PIVOT ? 2 * KB : KB) + 1;
	double wr[KB + 1][WW], wi[KB + 1][WW]; // rows k..k+KB, columns k..k+WW-1
#pragma unroll
	for (int r = 0; r <= KB; ++r)
#pragma unroll
		for (int c = 0; c < WW; ++c) {
			const bool in = (c - r <= KB && r - c <= KB);
			wr[r][c] = in ? abr[r * W + KB + c - r] : 0.0;
			wi[r][c] = (in && is_c) ? abi[r * W + KB + c - r] : 0.0;
		}
	bool would_swap = false;
#pragma unroll 1
	for (int k = 0; k < N; ++k) {
		int p = 0;
		double best = fabs(wr[0][0]) + fabs(wi[0][0]); // izamax's |re|+|im|; plain |re| for a real matrix
#pragma unroll
		for (int r = 1; r <= KB; ++r) {
			const double v = fabs(wr[r][0]) + fabs(wi[r][0]); // the padding rows hold zeros and never win
			if (v > best) {
				best = v;
				p = r;
			}
		}
		would_swap |= (p != 0);
		if (PIVOT) {
			piv[k] = (signed char)p;
#pragma unroll
			for (int r = 1; r <= KB; ++r) {
				const bool sw = (p == r);
#pragma unroll
				for (int c = 0; c < WW; ++c) {
					double u = wr[0][c], v = wr[r][c];
					wr[0][c] = sw ? v : u;
					wr[r][c] = sw ? u : v;
					u = wi[0][c];
					v = wi[r][c];
					wi[0][c] = sw ? v : u;
					wi[r][c] = sw ? u : v;
				}
			}
		}
		// reciprocal pivot 1/(a+ib) = (a-ib)/(a^2+b^2)
		const double pr = wr[0][0], pi = wi[0][0];
		const double rden = fast_rcp(fma(pr, pr, pi * pi));
		const double rr = pr * rden, ri = -pi * rden;
		rdr[k] = rr;
		if (is_c) rdi[k] = ri;
#pragma unroll
		for (int c = 0; c < WW; ++c) {
			abr[k * W + KB + c] = wr[0][c];
			if (is_c) abi[k * W + KB + c] = wi[0][c];
		}
#pragma unroll
		for (int r = 1; r <= KB; ++r) {
			const double lr = fma(wr[r][0], rr, -wi[r][0] * ri);
			const double li = fma(wr[r][0], ri, wi[r][0] * rr);
			abr[(k + r) * W + KB - r] = lr; // rows >= N are padding
			if (is_c) abi[(k + r) * W + KB - r] = li;
#pragma unroll
			for (int c = 1; c < WW; ++c) {
				const double ur = wr[0][c], ui = wi[0][c];
				wr[r][c] = fma(-lr, ur, fma(li, ui, wr[r][c]));
				wi[r][c] = fma(-lr, ui, fma(-li, ur, wi[r][c]));
			}
		}
		// slide: rows k+1..k+KB+1, columns k+1..k+WW (column k+1+c of the new row sits in slot c)
		const int in = k + 1 + KB;
#pragma unroll
		for (int r = 0; r < KB; ++r) {
#pragma unroll
			for (int c = 0; c + 1 < WW; ++c) {
				wr[r][c] = wr[r + 1][c + 1];
				wi[r][c] = wi[r + 1][c + 1];
			}
			// the column that enters the window: beyond every older row's band when U is 2*KB wide (fill-in only
			// comes from interchanges); with the narrow window it is the row's own, still untouched, last band entry
			wr[r][WW - 1] = PIVOT ? 0.0 : abr[(k + 1 + r) * W + 2 * KB - r];
			wi[r][WW - 1] = (PIVOT || !is_c) ? 0.0 : abi[(k + 1 + r) * W + 2 * KB - r];
		}
		if (in < N + KB) { // warp-uniform (k is)
#pragma unroll
			for (int c = 0; c < WW; ++c) {
				wr[KB][c] = abr[in * W + c];
				wi[KB][c] = is_c ? abi[in * W + c] : 0.0;
			}
		} else {
#pragma unroll
			for (int c = 0; c < WW; ++c) wr[KB][c] = wi[KB][c] = 0.0;
		}
	}
	return would_swap;
}

// ------------------------------------------------------------------ band solves, one lane per right-hand side
// x <- U^-1 L^-1 P x with the factors above (P only when PIVOT); x in shared memory, the recurrence in
// registers.  Real and complex systems share the routine like the factorisations do.  The forward
// window may read up to KB entries past x[N-1]: they only ever combine with the zero multipliers
// of the padding rows and are never stored.
template <int N, int KB, bool PIVOT>
__device__ __forceinline__ void band_solve(const double *abr, const double *abi, bool is_c, const double *rdr, const double *rdi,
                                           const signed char *piv, double *xr, double *xi)
{
	constexpr int W = 3 * KB + 1;
	double wr[KB + 1], wi[KB + 1];
#pragma unroll
	for (int r = 0; r <= KB; ++r) {
		wr[r] = xr[r];
		wi[r] = is_c ? xi[r] : 0.0;
	}
#pragma unroll 2
	for (int k = 0; k < N; ++k) { // forward: (interchange, then) eliminate with the unit lower factor
		if (PIVOT) {
			const int p = piv[k];
#pragma unroll
			for (int r = 1; r <= KB; ++r) {
				const bool sw = (p == r);
				double u = wr[0], v = wr[r];
				wr[0] = sw ? v : u;
				wr[r] = sw ? u : v;
				u = wi[0];
				v = wi[r];
				wi[0] = sw ? v : u;
				wi[r] = sw ? u : v;
			}
		}
#pragma unroll
		for (int r = 1; r <= KB; ++r) {
			const double lr = abr[(k + r) * W + KB - r], li = is_c ? abi[(k + r) * W + KB - r] : 0.0;
			const double a = wr[r], b = wi[r];
			wr[r] = fma(-lr, wr[0], fma(li, wi[0], a));
			wi[r] = fma(-lr, wi[0], fma(-li, wr[0], b));
		}
		xr[k] = wr[0];
		if (is_c) xi[k] = wi[0];
#pragma unroll
		for (int r = 0; r < KB; ++r) {
			wr[r] = wr[r + 1];
			wi[r] = wi[r + 1];
		}
		wr[KB] = xr[k + 1 + KB];
		wi[KB] = is_c ? xi[k + 1 + KB] : 0.0;
	}
	constexpr int UB = PIVOT ? 2 * KB : KB; // superdiagonals of U (see band_lu)
	double ur[UB + 1], ui[UB + 1]; // u[c] = x[k + c]
#pragma unroll
	for (int c = 0; c <= UB; ++c) ur[c] = ui[c] = 0.0;
#pragma unroll 2
	for (int k = N - 1; k >= 0; --k) { // backward; the newest unknown enters the chain last
		double ar = xr[k], ai = is_c ? xi[k] : 0.0;
#pragma unroll
		for (int c = UB; c >= 1; --c) {
			const double er = abr[k * W + KB + c], ei = is_c ? abi[k * W + KB + c] : 0.0; // beyond column N-1: stored zeros
			ar = fma(-er, ur[c], fma(ei, ui[c], ar));
			ai = fma(-er, ui[c], fma(-ei, ur[c], ai));
		}
		const double dr = rdr[k], di = is_c ? rdi[k] : 0.0;
		const double kr = fma(ar, dr, -ai * di);
		const double ki = fma(ar, di, ai * dr);
		xr[k] = kr;
		if (is_c) xi[k] = ki;
#pragma unroll
		for (int c = UB; c >= 2; --c) {
			ur[c] = ur[c - 1];
			ui[c] = ui[c - 1];
		}
		ur[1] = kr;
		ui[1] = ki;
	}
}

// ------------------------------------------------------------------ the kernel: one warp per block
template <class F, int S, int NR, int NC, int EST, bool DENSE = false>
__global__ void __launch_bounds__(32)
ensemble_irk_warp(const __grid_constant__ KernelArgs a, const __grid_constant__ DevTableau<S, NC> tb)
{
	static_assert(S == NR + 2 * NC && NR <= 1, "eigen-structure");
	using L = WarpSmem<F, S, NC, EST>;
	constexpr int N = F::N, KB = L::KB, W = L::W;
	static_assert(KB >= 1 && KB <= 4, "band kernels are built for half bandwidths 1..4");
	static_assert(1 + NC + 1 <= 32, "one lane per matrix");
	extern __shared__ double sm[];
	double *y = sm + L::o_y, *yn = sm + L::o_yn, *par = sm + L::o_p, *Y = sm + L::o_Y, *R = sm + L::o_R, *Z = sm + L::o_Z,
	       *Fv = sm + L::o_Z, *ev = sm + L::o_e, *dyv = sm + L::o_dy, *dalt = sm + L::o_dalt, *fy = sm + L::o_fy,
	       *abr = sm + L::o_abr, *rdr = sm + L::o_rdr, *abc = sm + L::o_abc, *rdc = sm + L::o_rdc;
	signed char *piv = reinterpret_cast<signed char *>(sm + L::o_piv); // [real slots][N] then [NC][N]
	signed char *pivc = piv + L::NREAL * N;
	const int lane = threadIdx.x;
	const unsigned long long M = a.M;
	const rehuel_device_io &io = a.io;
	constexpr int PL = L::PL;
	// lane -> matrix: lane 0 the real Newton block, lanes 1..NC the complex ones, lane NC+1 the estimator's own
	const bool m_cplx = lane >= 1 && lane <= NC;
	const bool m_est = (EST == EST_OWN_LU) && lane == NC + 1;
	const bool has_matrix = (NR && lane == 0) || m_cplx || m_est;
	const int m_c = m_cplx ? lane - 1 : 0;
	double *m_abr = m_cplx ? abc + (m_c * 2) * PL : (m_est ? abr + PL : abr);
	double *m_abi = m_cplx ? abc + (m_c * 2 + 1) * PL : m_abr;
	double *m_rdr = m_cplx ? rdc + (m_c * 2) * N : (m_est ? rdr + N : rdr);
	double *m_rdi = m_cplx ? rdc + (m_c * 2 + 1) * N : m_rdr;
	signed char *m_piv = m_cplx ? pivc + m_c * N : (m_est ? piv + N : piv);
	// lane -> right-hand side of a Newton update: the same lanes, estimator excluded
	const bool has_rhs = (NR && lane == 0) || m_cplx;
	double *m_xr = m_cplx ? Z + (NR + 2 * m_c) * N : Z;
	double *m_xi = m_cplx ? Z + (NR + 2 * m_c + 1) * N : Z;

	for (;;) {
		// ---------------- next member
		unsigned long long mem = 0;
		if (lane == 0) {
			mem = atomicAdd(io.queue, 1ull);
			if (mem < M && io.order) mem = io.order[mem];
		}
		mem = __shfl_sync(0xffffffffu, mem, 0);
		if (mem >= M) break;
		for (int k = lane; k < N; k += 32) y[k] = io.y0_broadcast ? io.y0[k] : io.y0[k * M + mem];
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
				for (int k = lane; k < N; k += 32) io.y_dense[((size_t)kd * N + k) * M + mem] = y[k];
				++kd;
			}
		}
		if (a.store_samples) {
			if (io.n_samples_cap) {
				if (lane == 0) {
					io.t_samples[mem] = t;
					if (io.err_samples) io.err_samples[mem] = 0.0;
				}
				for (int k = lane; k < N; k += 32) io.y_samples[k * M + mem] = y[k];
			}
			n_stored = 1;
		}

		// residual R = Y - dt*(A (x) I) F(y+Y), returns |R|^2   (irk.hpp:366-386)
		auto residual = [&]() -> double {
			for (int e = lane; e < S * N; e += 32) R[e] = y[e % N] + Y[e]; // stage states (R is free here)
			__syncwarp();
			for (int e = lane; e < S * N; e += 32) {
				const int s = e / N, i = e % N;
				Fv[e] = F::fun(i, fma(tb.c[s], dt, t), R + s * N, par);
			}
			__syncwarp();
			c_fun += S;
			double r2 = 0.0;
			for (int e = lane; e < S * N; e += 32) {
				const int s = e / N, i = e % N;
				double acc = 0.0;
#pragma unroll
				for (int j = 0; j < S; ++j) acc = fma(tb.A[s][j], Fv[j * N + i], acc);
				const double r = fma(-dt, acc, Y[e]);
				R[e] = r;
				r2 = fma(r, r, r2);
			}
			__syncwarp();
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
			// -------- Jacobian rows straight into the band storage of every matrix (irk.hpp:422-436)
			const double rdt = 1.0 / dt;
			const double mur = (EST == EST_REUSE) ? 1.0 / (tb.gamma * dt) : tb.gamma_hat * rdt;
			const double mue = (EST == EST_OWN_LU) ? 1.0 / (tb.gamma * dt) : 0.0;
			auto build_matrices = [&]() {
				for (int i = lane; i < N; i += 32) {
					double row[N];
#pragma unroll
					for (int j = 0; j < N; ++j) row[j] = 0.0;
					F::jac_row(i, t, y, par, row);
#pragma unroll
					for (int d = -KB; d <= 2 * KB; ++d) {
						const int j = i + d;
						const double jv = (d <= KB && j >= 0 && j < N) ? row[j] : 0.0;
						const double dg = (d == 0) ? 1.0 : 0.0;
						if (NR) abr[i * W + KB + d] = dg * mur - jv;
						if (EST == EST_OWN_LU) abr[PL + i * W + KB + d] = dg * mue - jv;
#pragma unroll
						for (int c = 0; c < NC; ++c) {
							abc[(c * 2 + 0) * PL + i * W + KB + d] = dg * (tb.alpha[c] * rdt) - jv;
							abc[(c * 2 + 1) * PL + i * W + KB + d] = dg * (tb.beta[c] * rdt);
						}
					}
				}
				for (int e = lane; e < KB * W; e += 32) { // the zero rows below row N-1 of every plane
					if (NR) abr[N * W + e] = 0.0;
					if (EST == EST_OWN_LU) abr[PL + N * W + e] = 0.0;
#pragma unroll
					for (int c = 0; c < 2 * NC; ++c) abc[c * PL + N * W + e] = 0.0;
				}
				__syncwarp();
			};
			++c_jac;
			// -------- all factorisations side by side, one lane per matrix, one routine for all of them:
			// plain elimination first; if pivoting would have exchanged rows anywhere, redo them pivoted
			build_matrices();
			bool flag = false;
			if (has_matrix) flag = band_lu<N, KB, false>(m_abr, m_abi, m_cplx, m_rdr, m_rdi, m_piv);
			__syncwarp();
			const bool piv_any = __any_sync(0xffffffffu, flag);
			if (piv_any) {
				build_matrices();
				if (has_matrix) band_lu<N, KB, true>(m_abr, m_abi, m_cplx, m_rdr, m_rdi, m_piv);
				__syncwarp();
			}

			// -------- simplified Newton (irk.hpp:398-493)
			for (int e = lane; e < S * N; e += 32) Y[e] = 0.0; // irk.hpp:415
			__syncwarp();
			double Rnorm2 = residual();
			int nstat = kNewtonMaxit;
			int it = 1;
			for (; it < maxit; ++it) { // irk.hpp:458
				for (int e = lane; e < S * N; e += 32) { // Z = (Ti (x) I) R
					const int s = e / N, i = e % N;
					double acc = 0.0;
#pragma unroll
					for (int j = 0; j < S; ++j) acc = fma(tb.Ti[s][j], R[j * N + i], acc);
					Z[e] = acc;
				}
				__syncwarp();
				if (has_rhs) { // all right-hand sides side by side, one lane each
					if (piv_any) band_solve<N, KB, true>(m_abr, m_abi, m_cplx, m_rdr, m_rdi, m_piv, m_xr, m_xi);
					else band_solve<N, KB, false>(m_abr, m_abi, m_cplx, m_rdr, m_rdi, m_piv, m_xr, m_xi);
				}
				__syncwarp();
				if (NR)
					for (int i = lane; i < N; i += 32) Z[i] *= -mur;
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					double *zr = Z + (NR + 2 * c) * N, *zi = zr + N;
					const double mr = tb.alpha[c] * rdt, mi = tb.beta[c] * rdt;
					for (int i = lane; i < N; i += 32) { // w = -(mr + i mi) w
						const double wr = zr[i], wi = zi[i];
						zr[i] = fma(-mr, wr, mi * wi);
						zi[i] = fma(-mr, wi, -mi * wr);
					}
				}
				__syncwarp();
				double x2 = 0.0;
				for (int e = lane; e < S * N; e += 32) { // dY = (T (x) I) W ; Y += dY
					const int s = e / N, i = e % N;
					double acc = 0.0;
#pragma unroll
					for (int j = 0; j < S; ++j) acc = fma(tb.T[s][j], Z[j * N + i], acc);
					x2 = fma(acc, acc, x2);
					Y[e] += acc;
				}
				__syncwarp();
				const double xnorm2 = warp_sum(x2); // irk.hpp:466
				if (REHUEL_SKIP_DEAD_RESIDUAL && xnorm2 < a.xtol2) { // the residual of irk.hpp:469-472 would never be looked at
					c_fun += S;                                       // (see irk_thread.cuh); its evaluations are booked
					nstat = kNewtonSuccess;
					break;
				}
				Rnorm2 = residual();
				if (Rnorm2 < a.Rtol2) { nstat = kNewtonSuccess; break; }
				if (xnorm2 < a.xtol2) { nstat = kNewtonSuccess; break; }
				if (a.refresh_jac > 0 && it % a.refresh_jac == 0) ++c_jac;
				if (!(Rnorm2 <= 1.79769313486231570e308)) { // NaN/Inf: see irk_thread.cuh
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

			// -------- new solution + embedded error (irk.hpp:691-731)
			const double gam = tb.gamma * dt;
			for (int i = lane; i < N; i += 32) {
				double s1 = 0.0, s2 = 0.0;
#pragma unroll
				for (int s = 0; s < S; ++s) {
					s1 = fma(Y[s * N + i], tb.d[s], s1);
					s2 = fma(Y[s * N + i], tb.d2[s], s2);
				}
				dyv[i] = s1;
				dalt[i] = s2;
				yn[i] = y[i] + s1;
				if (EST != EST_IDENTITY) fy[i] = F::fun(i, t, y, par); // irk.hpp:702
			}
			++c_fun;
			__syncwarp();
			for (int i = lane; i < N; i += 32) {
				const double dyalt = (EST != EST_IDENTITY) ? fma(gam, fy[i], dalt[i]) : dalt[i];
				ev[i] = dyalt - dyv[i];
			}
			__syncwarp();
			const double esc = (EST == EST_IDENTITY) ? dt : 1.0 / tb.gamma; // dt*E^-1 = (1/gamma)(mu I - J)^-1
			auto apply_Einv_dt = [&]() {
				if (EST != EST_IDENTITY) {
					if (lane == 0) {
						const int o = (EST == EST_REUSE) ? 0 : 1; // which real plane holds E's factors
						if (piv_any) band_solve<N, KB, true>(abr + o * PL, abr, false, rdr + o * N, rdr, piv + o * N, ev, ev);
						else band_solve<N, KB, false>(abr + o * PL, abr, false, rdr + o * N, rdr, piv + o * N, ev, ev);
					}
					__syncwarp();
				}
				for (int i = lane; i < N; i += 32) ev[i] *= esc;
				__syncwarp();
			};
			apply_Einv_dt(); // 8.19
			if (alt) {       // 8.20, irk.hpp:722-731
				if (EST != EST_IDENTITY) {
					for (int i = lane; i < N; i += 32) Z[i] = y[i] + ev[i];
					__syncwarp();
					for (int i = lane; i < N; i += 32) fy[i] = F::fun(i, t, Z, par);
					__syncwarp();
				}
				++c_fun;
				for (int i = lane; i < N; i += 32) {
					const double v = (EST != EST_IDENTITY) ? fma(gam, fy[i], dalt[i]) : dalt[i];
					ev[i] = v - dyv[i];
				}
				__syncwarp();
				apply_Einv_dt();
			}
			double tot = 0.0;
			for (int i = lane; i < N; i += 32) { // irk.hpp:733-746
				const double sc = fma(a.rtol, fmax(fabs(y[i]), fabs(yn[i])), a.atol);
				const double q = ev[i] / sc;
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
			bool stuck = false;
			if (!rejected) { // irk.hpp:813-846
				if (DENSE) { // grid times inside (t, t + dt] from this step's collocation polynomial
					const double t_new = t + dt;
					while (kd < io.n_dense) {
						const double tk = fma((double)kd, io.dense_dt, io.dense_t0);
						if (tk > t_new) break;
						double w[S];
						dense_weights<S, NC>(tb, (tk - t) * rdt, w);
						for (int i = lane; i < N; i += 32) {
							double acc = y[i];
#pragma unroll
							for (int s = 0; s < S; ++s) acc = fma(w[s], Y[s * N + i], acc);
							io.y_dense[((size_t)kd * N + i) * M + mem] = (tk == t_new) ? yn[i] : acc;
						}
						++kd;
					}
				}
				for (int i = lane; i < N; i += 32) y[i] = yn[i];
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
						for (int i = lane; i < N; i += 32) io.y_samples[(row * N + i) * M + mem] = yn[i];
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
		for (int k = lane; k < N; k += 32) io.y[k * M + mem] = y[k];
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
