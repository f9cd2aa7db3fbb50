// CTA-per-system ensemble integrator for larger systems (4 < n <= 64): one thread block runs one
// member's adaptive integration, persistent blocks pull members from the global work queue.
//
// Same algorithm, statement for statement, as the thread-per-system kernel (irk_thread.cuh) and
// hence as the reference loop irk.hpp:604-860 / newton_solve_stages irk.hpp:398-493; what changes
// is where the data lives and who does the work:
//   * state vectors (y, Y, R, Z, F) and the Jacobian live in shared memory;
//   * the decoupled Newton matrices  mu_r I - J  (real) and  mu_c I - J  (complex, one per
//     conjugate eigenvalue pair of inv(A)) are factorised IN shared memory by the whole block in
//     ONE fused right-looking loop (row partial pivoting like dgetrf/zgetrf): warp m picks the
//     pivot, swaps and scales for matrix m, then all threads apply the rank-1 updates of all
//     matrices -- two block barriers per elimination step for the whole set;
//   * the triangular solves of matrix m are done by warp m alone, warp-synchronously: each lane
//     keeps rows lane, lane+32 of the right-hand side in registers and the pivot entry is
//     broadcast with a shuffle, so a Newton iteration costs no block barrier inside the solves;
//   * scalar control (dt, controller history, counters) is kept redundantly by every thread; all
//     threads see identical reduction results (fixed summation order), so control flow is uniform.
// For the 64-equation Brusselator with LOBATTO_IIIC_85 this replaces the reference's dense
// 320x320 LU (21.8 Mflop per attempt) by one real and two complex 64x64 LUs (1.6 Mflop).
//
// Functor concept for this path (componentwise, so that the block can share the work):
//   static constexpr int N, P;
//   __device__ static double fun(int i, double t, const double *y, const double *p);          // f_i(t,y)
//   __device__ static void   jac_row(int i, double t, const double *y, const double *p, double *row); // row i, pre-zeroed
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "../../include/rehuel_b200.h"
#include "irk_thread.cuh"
#include "tableau.hpp"

namespace rehuel {

// Optional functor trait: half bandwidth of the Jacobian (J(i,j) == 0 for |i-j| > kBand).  The LU
// and the triangular solves then skip the entries that are structurally zero -- with row partial
// pivoting U fills in up to 2*kBand above the diagonal, L stays within kBand below it (dgbtrf's
// pattern).  Skipped operations act on exact zeros, so results are those of the dense path.
template <class F, typename = void>
struct jac_band { static constexpr int value = F::N - 1; };
template <class F>
struct jac_band<F, decltype((void)F::kBand)> { static constexpr int value = F::kBand < F::N - 1 ? F::kBand : F::N - 1; };

// Optional functor member for banded Jacobians: jac_band(i, t, y, p, double (&band)[2*kBand+1]) fills
// band[kBand + o] = J(i, i + o), o = -kBand .. kBand (entries beyond the matrix: 0), with static indices.
template <class F, class = void>
struct has_jac_band { static constexpr bool value = false; };
template <class F>
struct has_jac_band<F, decltype((void)&F::jac_band)> { static constexpr bool value = true; };

template <class F, int S, int NC>
struct CtaSmem {
	static constexpr int N = F::N;
	static constexpr int LD = N + 1; // odd leading dimension: column walks are bank-conflict free
	static constexpr int PA = ((F::P > 0 ? F::P : 1) + 1) & ~1;
	// offsets in doubles
	static constexpr int o_y = 0;
	static constexpr int o_yn = o_y + N;
	static constexpr int o_p = o_yn + N;
	static constexpr int o_Y = o_p + PA;
	static constexpr int o_R = o_Y + S * N;
	static constexpr int o_Z = o_R + S * N;
	static constexpr int o_F = o_Z + S * N;
	static constexpr int o_e = o_F + S * N;
	static constexpr int o_dy = o_e + N;
	static constexpr int o_dalt = o_dy + N;
	static constexpr int o_fy = o_dalt + N;
	static constexpr int o_red = o_fy + N;           // 64 doubles of reduction scratch
	static constexpr int o_rdr = o_red + 64;         // reciprocal pivots, real matrix
	static constexpr int o_rdc = o_rdr + N;          // reciprocal pivots, complex matrices (re, im)
	static constexpr int o_rde = o_rdc + 2 * N * (NC > 0 ? NC : 1); // reciprocal pivots, the estimator's own matrix
	static constexpr int o_J = o_rde + N;            // Jacobian, leading dimension LD: the estimator's own matrix
	static constexpr int o_LUr = o_J + N * LD;       // (1/gamma dt) I - J is built over it and factorised in place
	static constexpr int o_LUc = o_LUr + N * LD;
	static constexpr int o_int = o_LUc + 2 * N * LD * NC; // int region: perm arrays (real, complex, estimator)
	static constexpr int n_int = N * (2 + NC) + 1 + 8;
	static constexpr size_t bytes = (size_t)o_int * 8 + (size_t)n_int * 4;
};

// Sum over the block, identical on every thread (fixed order: butterfly inside a warp, then the
// warp partials in ascending order).  Two block barriers.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double *red)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0.0;
#pragma unroll
	for (int w = 0; w < THREADS / 32; ++w) s += red[w];
	__syncthreads();
	return s;
}

// Fused LU of NRm (0 or 1) + NEm (0 or 1) real and NCm complex N x N matrices in shared memory; matrix m is
// piloted by warp m (real Newton block, complex ones, then the estimator's own real matrix Ae).  perm[m][k] = row
// exchanged with row k at step k (ipiv); reciprocal pivots go to rdr / rdc / rde so that the solves multiply.
// A step is two block barriers and a chain (pivot search -> reciprocal -> multipliers -> rank-1 update) that
// nothing overlaps, so what counts is the NUMBER of steps: all matrices of an attempt share the N steps.
template <int N, int LD, int NRm, int NCm, int NEm, int THREADS, int KB>
__device__ __forceinline__ void lu_factor_set(double *Ar, double *Ac, double *Ae, double *rdr, double *rdc, double *rde, int *permr,
                                              int *permc, int *perme)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int NM = NRm + NCm + NEm;
	static_assert(NM <= THREADS / 32, "one pilot warp per matrix");
	for (int k = 0; k < N; ++k) {
		if (warp < NM) {
			const bool is_est = NEm && warp == NRm + NCm;
			const bool is_real = warp < NRm || is_est;
			const int c = is_real ? 0 : warp - NRm;
			double *are = is_est ? Ae : (is_real ? Ar : Ac + (size_t)c * 2 * N * LD);
			double *aim = are + N * LD; // only for complex
			// ---- pivot search: first row of maximal magnitude in column k (izamax: |re|+|im|)
			double best = -1.0;
			int bi = k;
			const int rmax = (k + KB + 1 < N) ? k + KB + 1 : N;         // rows that can be non-zero in column k
			const int cmax = (k + 2 * KB + 1 < N) ? k + 2 * KB + 1 : N; // columns that can be non-zero in those rows
			for (int r = k + lane; r < rmax; r += 32) {
				const double v = is_real ? fabs(are[r * LD + k]) : fabs(are[r * LD + k]) + fabs(aim[r * LD + k]);
				if (v > best) {
					best = v;
					bi = r;
				}
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				const double ob = __shfl_xor_sync(0xffffffffu, best, o);
				const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
				if (ob > best || (ob == best && oi < bi)) {
					best = ob;
					bi = oi;
				}
			}
			const int p = bi; // the butterfly leaves the same (value, first index) in every lane
			if (lane == 0) (is_est ? perme : (is_real ? permr : permc + c * N))[k] = p; // interchange list, as ipiv of dgetrf
			if (p != k) {
				// like dgbtrf, only the not yet eliminated part of the rows is exchanged; the
				// solves apply the interchanges step by step, so the multipliers stay put
				for (int j = k + lane; j < cmax; j += 32) {
					double u = are[k * LD + j];
					are[k * LD + j] = are[p * LD + j];
					are[p * LD + j] = u;
					if (!is_real) {
						u = aim[k * LD + j];
						aim[k * LD + j] = aim[p * LD + j];
						aim[p * LD + j] = u;
					}
				}
			}
			__syncwarp();
			// ---- reciprocal pivot (fast_rcp: see irk_thread.cuh), multipliers
			if (is_real) {
				const double rp = fast_rcp(are[k * LD + k]);
				if (lane == 0) (is_est ? rde : rdr)[k] = rp;
				for (int r = k + 1 + lane; r < rmax; r += 32) are[r * LD + k] *= rp;
			} else {
				const double pr = are[k * LD + k], pi = aim[k * LD + k];
				const double rden = fast_rcp(fma(pr, pr, pi * pi));
				const double rr = pr * rden, ri = -pi * rden;
				if (lane == 0) {
					rdc[(c * 2 + 0) * N + k] = rr;
					rdc[(c * 2 + 1) * N + k] = ri;
				}
				for (int r = k + 1 + lane; r < rmax; r += 32) {
					const double xr = are[r * LD + k], xi = aim[r * LD + k];
					are[r * LD + k] = fma(xr, rr, -xi * ri);
					aim[r * LD + k] = fma(xr, ri, xi * rr);
				}
			}
		}
		__syncthreads();
		// ---- rank-1 update of every trailing block: 16 x (THREADS/16) thread grid, strided; a thread keeps its (up to CB)
		// pivot-row entries of every matrix in registers while it walks down its rows
		if (k + 1 < N) {
			constexpr int TX = 16, TY = THREADS / 16, CB = (N + TX - 1) / TX;
			const int tx = tid % TX, ty = tid / TX;
			const int rmax = (k + KB + 1 < N) ? k + KB + 1 : N;
			const int cmax = (k + 2 * KB + 1 < N) ? k + 2 * KB + 1 : N;
			const int j0 = k + 1 + tx;
			if (NRm) {
				double u[CB];
#pragma unroll
				for (int b = 0; b < CB; ++b) u[b] = (j0 + TX * b < cmax) ? Ar[k * LD + j0 + TX * b] : 0.0;
				for (int i = k + 1 + ty; i < rmax; i += TY) {
					const double l = Ar[i * LD + k];
#pragma unroll
					for (int b = 0; b < CB; ++b)
						if (j0 + TX * b < cmax) Ar[i * LD + j0 + TX * b] = fma(-l, u[b], Ar[i * LD + j0 + TX * b]);
				}
			}
			if (NEm) {
				double u[CB];
#pragma unroll
				for (int b = 0; b < CB; ++b) u[b] = (j0 + TX * b < cmax) ? Ae[k * LD + j0 + TX * b] : 0.0;
				for (int i = k + 1 + ty; i < rmax; i += TY) {
					const double l = Ae[i * LD + k];
#pragma unroll
					for (int b = 0; b < CB; ++b)
						if (j0 + TX * b < cmax) Ae[i * LD + j0 + TX * b] = fma(-l, u[b], Ae[i * LD + j0 + TX * b]);
				}
			}
#pragma unroll
			for (int c = 0; c < NCm; ++c) {
				double *are = Ac + (size_t)c * 2 * N * LD;
				double *aim = are + N * LD;
				double ur[CB], ui[CB];
#pragma unroll
				for (int b = 0; b < CB; ++b) {
					const bool in = j0 + TX * b < cmax;
					ur[b] = in ? are[k * LD + j0 + TX * b] : 0.0;
					ui[b] = in ? aim[k * LD + j0 + TX * b] : 0.0;
				}
				for (int i = k + 1 + ty; i < rmax; i += TY) {
					const double lr = are[i * LD + k], li = aim[i * LD + k];
#pragma unroll
					for (int b = 0; b < CB; ++b) {
						const int j = j0 + TX * b;
						if (j < cmax) {
							are[i * LD + j] = fma(-lr, ur[b], fma(li, ui[b], are[i * LD + j]));
							aim[i * LD + j] = fma(-lr, ui[b], fma(-li, ur[b], aim[i * LD + j]));
						}
					}
				}
			}
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// PANEL-BLOCKED LU for dense matrices (N a multiple of kPanel): the same eliminations in the same order as
// lu_factor_set -- every entry sees the identical sequence of FMAs, so the factors are bit-identical -- but
//   A. a panel of kPanel columns is factorised by ONE warp per matrix in registers (two rows per lane, pivot search by
//      three REDUX, row exchanges and the pivot-row broadcast by shuffles): no block barrier inside the panel;
//   B. the panel's row exchanges and the unit-lower triangular solve are applied to the columns right of it, one thread
//      per column and matrix;
//   C. the trailing block takes the rank-kPanel update A22 -= L21 U12 in register tiles (4 x 4 real, 2 x 4 complex):
//      per FMA 0.5 (real) / 0.4 (complex) shared-memory accesses instead of 2.3 for a rank-1 update.
// Three block barriers per panel (24 per set of factorisations for N = 64) instead of two per column (128).
// Row exchanges follow LAPACK inside a panel (whole panel rows are exchanged, so the multipliers of the earlier panel
// columns move with their rows) and leave the columns LEFT of the panel alone: the forward solve applies the kPanel
// exchanges of a panel before its kPanel elimination steps (PANEL flag of the solves).
#ifndef REHUEL_CTA_PANEL
#define REHUEL_CTA_PANEL 1
#endif
constexpr int kPanel = 8;

// warp-wide argmax of a non-negative double (one candidate per lane, `row` its row index; best < 0: no candidate):
// the smallest row among the maxima, N if there is none (NaN column)
template <int N>
__device__ __forceinline__ int warp_pivot_row(double best, int row)
{
	const int hi = __double2hiint(best); // non-negative doubles order like their bit patterns
	const int mh = __reduce_max_sync(0xffffffffu, hi);
	const unsigned lo = (hi == mh) ? (unsigned)__double2loint(best) : 0u;
	const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
	const unsigned cand = (mh >= 0 && hi == mh && lo == ml) ? (unsigned)row : (unsigned)N;
	return (int)__reduce_min_sync(0xffffffffu, cand);
}

// One warp factorises a panel: the loop over its kPanel columns is ROLLED -- straight-line code that one warp runs once
// is bound by instruction fetch (the unrolled version spent half its time without an instruction) -- so the registers
// hold a window of the panel whose column 0 is the current column: a step stores its finished column (multipliers)
// and pivot row to shared memory and shifts the window left while it updates it.
template <int N, int LD>
__device__ __forceinline__ void lu_panel_real(double *A, double *rd, int *perm, int c0)
{
	constexpr int PW = kPanel, Q = (N + 31) / 32;
	const int lane = threadIdx.x & 31;
	double a[Q][PW];
#pragma unroll
	for (int q = 0; q < Q; ++q) {
		const int r = lane + 32 * q;
#pragma unroll
		for (int j = 0; j < PW; ++j) a[q][j] = (r >= c0 && r < N) ? A[r * LD + c0 + j] : 0.0;
	}
#pragma unroll 1
	for (int j = 0; j < PW; ++j) {
		const int k = c0 + j;
		double best = -1.0;
		int brow = N;
#pragma unroll
		for (int q = 0; q < Q; ++q) {
			const int r = lane + 32 * q;
			const double v = fabs(a[q][0]);
			if (r >= k && r < N && v > best) {
				best = v;
				brow = r;
			}
		}
		int p = warp_pivot_row<N>(best, brow);
		if (p >= N) p = k;
		if (lane == 0) perm[k] = p;
		const int lk = k & 31, lp = p & 31;
		const bool qk = Q > 1 && k >= 32, qp = Q > 1 && p >= 32;
		if (p != k) { // warp-uniform: exchange the panel rows k and p -- the window, and the finished columns in memory
#pragma unroll
			for (int jj = 0; jj < PW; ++jj) {
				const double vk = __shfl_sync(0xffffffffu, qk ? a[Q - 1][jj] : a[0][jj], lk);
				const double vp = __shfl_sync(0xffffffffu, qp ? a[Q - 1][jj] : a[0][jj], lp);
#pragma unroll
				for (int q = 0; q < Q; ++q) {
					const int r = lane + 32 * q;
					if (r == k) a[q][jj] = vp;
					if (r == p) a[q][jj] = vk;
				}
			}
			if (lane < j) {
				const double v = A[k * LD + c0 + lane];
				A[k * LD + c0 + lane] = A[p * LD + c0 + lane];
				A[p * LD + c0 + lane] = v;
			}
		}
		double u[PW];
#pragma unroll
		for (int jj = 0; jj < PW; ++jj) u[jj] = __shfl_sync(0xffffffffu, qk ? a[Q - 1][jj] : a[0][jj], lk);
		const double rp = fast_rcp(u[0]);
		if (lane == 0) rd[k] = rp;
		// the pivot row is final: lane jj stores U(k, k + jj)
		{
			double mine = u[0];
#pragma unroll
			for (int jj = 1; jj < PW; ++jj)
				if (lane == jj) mine = u[jj];
			if (lane < PW - j) A[k * LD + k + lane] = mine;
		}
#pragma unroll
		for (int q = 0; q < Q; ++q) {
			const int r = lane + 32 * q;
			if (r > k && r < N) {
				const double l = a[q][0] * rp;
				A[r * LD + k] = l;
#pragma unroll
				for (int jj = 1; jj < PW; ++jj) a[q][jj - 1] = fma(-l, u[jj], a[q][jj]);
				a[q][PW - 1] = 0.0;
			}
		}
		__syncwarp(); // the stored columns are exchanged by other lanes in later steps
	}
}

template <int N, int LD>
__device__ __forceinline__ void lu_panel_cplx(double *Are, double *rdre, double *rdim, int *perm, int c0)
{
	constexpr int PW = kPanel, Q = (N + 31) / 32;
	const int lane = threadIdx.x & 31;
	double *Aim = Are + N * LD;
	double ar[Q][PW], ai[Q][PW];
#pragma unroll
	for (int q = 0; q < Q; ++q) {
		const int r = lane + 32 * q;
#pragma unroll
		for (int j = 0; j < PW; ++j) {
			const bool in = r >= c0 && r < N;
			ar[q][j] = in ? Are[r * LD + c0 + j] : 0.0;
			ai[q][j] = in ? Aim[r * LD + c0 + j] : 0.0;
		}
	}
#pragma unroll 1
	for (int j = 0; j < PW; ++j) {
		const int k = c0 + j;
		double best = -1.0;
		int brow = N;
#pragma unroll
		for (int q = 0; q < Q; ++q) {
			const int r = lane + 32 * q;
			const double v = fabs(ar[q][0]) + fabs(ai[q][0]); // izamax
			if (r >= k && r < N && v > best) {
				best = v;
				brow = r;
			}
		}
		int p = warp_pivot_row<N>(best, brow);
		if (p >= N) p = k;
		if (lane == 0) perm[k] = p;
		const int lk = k & 31, lp = p & 31;
		const bool qk = Q > 1 && k >= 32, qp = Q > 1 && p >= 32;
		if (p != k) {
#pragma unroll
			for (int jj = 0; jj < PW; ++jj) {
				const double vkr = __shfl_sync(0xffffffffu, qk ? ar[Q - 1][jj] : ar[0][jj], lk);
				const double vki = __shfl_sync(0xffffffffu, qk ? ai[Q - 1][jj] : ai[0][jj], lk);
				const double vpr = __shfl_sync(0xffffffffu, qp ? ar[Q - 1][jj] : ar[0][jj], lp);
				const double vpi = __shfl_sync(0xffffffffu, qp ? ai[Q - 1][jj] : ai[0][jj], lp);
#pragma unroll
				for (int q = 0; q < Q; ++q) {
					const int r = lane + 32 * q;
					if (r == k) { ar[q][jj] = vpr; ai[q][jj] = vpi; }
					if (r == p) { ar[q][jj] = vkr; ai[q][jj] = vki; }
				}
			}
			if (lane < j) {
				double v = Are[k * LD + c0 + lane];
				Are[k * LD + c0 + lane] = Are[p * LD + c0 + lane];
				Are[p * LD + c0 + lane] = v;
				v = Aim[k * LD + c0 + lane];
				Aim[k * LD + c0 + lane] = Aim[p * LD + c0 + lane];
				Aim[p * LD + c0 + lane] = v;
			}
		}
		double ur[PW], ui[PW];
#pragma unroll
		for (int jj = 0; jj < PW; ++jj) {
			ur[jj] = __shfl_sync(0xffffffffu, qk ? ar[Q - 1][jj] : ar[0][jj], lk);
			ui[jj] = __shfl_sync(0xffffffffu, qk ? ai[Q - 1][jj] : ai[0][jj], lk);
		}
		const double rden = fast_rcp(fma(ur[0], ur[0], ui[0] * ui[0]));
		const double rr = ur[0] * rden, ri = -ui[0] * rden;
		if (lane == 0) {
			rdre[k] = rr;
			rdim[k] = ri;
		}
		{
			double mr = ur[0], mi = ui[0];
#pragma unroll
			for (int jj = 1; jj < PW; ++jj)
				if (lane == jj) {
					mr = ur[jj];
					mi = ui[jj];
				}
			if (lane < PW - j) {
				Are[k * LD + k + lane] = mr;
				Aim[k * LD + k + lane] = mi;
			}
		}
#pragma unroll
		for (int q = 0; q < Q; ++q) {
			const int r = lane + 32 * q;
			if (r > k && r < N) {
				const double xr = ar[q][0], xi = ai[q][0];
				const double lr = fma(xr, rr, -xi * ri), li = fma(xr, ri, xi * rr);
				Are[r * LD + k] = lr;
				Aim[r * LD + k] = li;
#pragma unroll
				for (int jj = 1; jj < PW; ++jj) {
					ar[q][jj - 1] = fma(-lr, ur[jj], fma(li, ui[jj], ar[q][jj]));
					ai[q][jj - 1] = fma(-lr, ui[jj], fma(-li, ur[jj], ai[q][jj]));
				}
				ar[q][PW - 1] = 0.0;
				ai[q][PW - 1] = 0.0;
			}
		}
		__syncwarp();
	}
}

template <int N, int LD, int NRm, int NCm, int NEm, int THREADS>
__device__ __forceinline__ void lu_factor_set_panel(double *Ar, double *Ac, double *Ae, double *rdr, double *rdc, double *rde,
                                                    int *permr, int *permc, int *perme, PhaseClock &pc)
{
	(void)pc;
	constexpr int PW = kPanel, NM = NRm + NCm + NEm, NREAL = NRm + NEm;
	static_assert(N % PW == 0 && N <= 64 && NM <= THREADS / 32, "one pilot warp per matrix, two rows per lane");
	const int tid = threadIdx.x, warp = tid >> 5;
	// matrix m: real Newton block, estimator (the real ones first), then the complex blocks
	auto real_mat = [&](int m) -> double * { return (NRm && m == 0) ? Ar : Ae; };
	auto real_perm = [&](int m) -> int * { return (NRm && m == 0) ? permr : perme; };
	for (int c0 = 0; c0 < N; c0 += PW) {
		// ---- A: the panel, one warp per matrix
		if (warp < NREAL) {
			lu_panel_real<N, LD>(real_mat(warp), (NRm && warp == 0) ? rdr : rde, real_perm(warp), c0);
		} else if (warp < NM) {
			const int c = warp - NREAL;
			lu_panel_cplx<N, LD>(Ac + (size_t)c * 2 * N * LD, rdc + (c * 2) * N, rdc + (c * 2 + 1) * N, permc + c * N, c0);
		}
		__syncthreads();
		REHUEL_PHASE_MARK(pc, 1);
		const int n2 = N - c0 - PW, cb = c0 + PW;
		if (n2 <= 0) break;
		// ---- B: exchanges + L11^-1 on the columns right of the panel: one (matrix, column) per thread
		for (int task = tid; task < NM * N; task += THREADS) {
			const int m = task / N, jc = task % N;
			if (jc < cb) continue;
			// (every load first: issued one by one between the dependent FMAs they would set the pace)
			if (m < NREAL) {
				double *A = real_mat(m);
				const int *perm = real_perm(m);
				int pv[PW];
#pragma unroll
				for (int j = 0; j < PW; ++j) pv[j] = perm[c0 + j];
				double l11[PW][PW];
#pragma unroll
				for (int j = 0; j < PW; ++j)
#pragma unroll
					for (int i = j + 1; i < PW; ++i) l11[i][j] = A[(c0 + i) * LD + c0 + j];
#pragma unroll
				for (int j = 0; j < PW; ++j) {
					const int k = c0 + j, p = pv[j];
					if (p != k) {
						const double v = A[k * LD + jc];
						A[k * LD + jc] = A[p * LD + jc];
						A[p * LD + jc] = v;
					}
				}
				double u[PW];
#pragma unroll
				for (int j = 0; j < PW; ++j) u[j] = A[(c0 + j) * LD + jc];
#pragma unroll
				for (int j = 0; j < PW; ++j)
#pragma unroll
					for (int i = j + 1; i < PW; ++i) u[i] = fma(-l11[i][j], u[j], u[i]);
#pragma unroll
				for (int j = 1; j < PW; ++j) A[(c0 + j) * LD + jc] = u[j];
			} else {
				const int c = m - NREAL;
				double *Are = Ac + (size_t)c * 2 * N * LD, *Aim = Are + N * LD;
				const int *perm = permc + c * N;
				int pv[PW];
#pragma unroll
				for (int j = 0; j < PW; ++j) pv[j] = perm[c0 + j];
				double l11r[PW][PW], l11i[PW][PW];
#pragma unroll
				for (int j = 0; j < PW; ++j)
#pragma unroll
					for (int i = j + 1; i < PW; ++i) {
						l11r[i][j] = Are[(c0 + i) * LD + c0 + j];
						l11i[i][j] = Aim[(c0 + i) * LD + c0 + j];
					}
#pragma unroll
				for (int j = 0; j < PW; ++j) {
					const int k = c0 + j, p = pv[j];
					if (p != k) {
						double v = Are[k * LD + jc];
						Are[k * LD + jc] = Are[p * LD + jc];
						Are[p * LD + jc] = v;
						v = Aim[k * LD + jc];
						Aim[k * LD + jc] = Aim[p * LD + jc];
						Aim[p * LD + jc] = v;
					}
				}
				double ur[PW], ui[PW];
#pragma unroll
				for (int j = 0; j < PW; ++j) {
					ur[j] = Are[(c0 + j) * LD + jc];
					ui[j] = Aim[(c0 + j) * LD + jc];
				}
#pragma unroll
				for (int j = 0; j < PW; ++j)
#pragma unroll
					for (int i = j + 1; i < PW; ++i) {
						const double lr = l11r[i][j], li = l11i[i][j];
						ur[i] = fma(-lr, ur[j], fma(li, ui[j], ur[i]));
						ui[i] = fma(-lr, ui[j], fma(-li, ur[j], ui[i]));
					}
#pragma unroll
				for (int j = 1; j < PW; ++j) {
					Are[(c0 + j) * LD + jc] = ur[j];
					Aim[(c0 + j) * LD + jc] = ui[j];
				}
			}
		}
		__syncthreads();
		REHUEL_PHASE_MARK(pc, 2);
		// ---- C: A22 -= L21 U12 in register tiles; the rows (columns) of a tile are n2/4 (n2/2 for the two rows of a
		// complex tile) apart, so that neighbouring threads read neighbouring columns
		const int TJ = n2 / 4, TI2 = n2 / 2;
		const int n_real = NREAL * TJ * TJ, n_all = n_real + NCm * TI2 * TJ;
		// x / d for x < 2048, d <= 512 as a multiplication (exact: x (d - 2^20 mod d) < 2^20)
		auto magic = [](int d) -> unsigned { return (1u << 20) / (unsigned)d + 1u; };
		const unsigned m_tj = magic(TJ), m_tt = magic(TJ * TJ), m_t2 = magic(TI2 * TJ);
		auto divm = [](int x, unsigned m) -> int { return (int)(((unsigned)x * m) >> 20); };
		static_assert(N <= 64, "unit indices stay below 2048");
		for (int unit = tid; unit < n_all; unit += THREADS) {
			if (unit < n_real) {
				const int m = divm(unit, m_tt), w = unit - m * TJ * TJ;
				double *A = real_mat(m);
				const int wi = divm(w, m_tj);
				const int r0 = cb + wi, cc = cb + w - wi * TJ;
				double acc[4][4];
#pragma unroll
				for (int x = 0; x < 4; ++x)
#pragma unroll
					for (int y = 0; y < 4; ++y) acc[x][y] = A[(r0 + TJ * x) * LD + cc + TJ * y];
#pragma unroll
				for (int kk = 0; kk < PW; ++kk) {
					double l[4], u[4];
#pragma unroll
					for (int x = 0; x < 4; ++x) l[x] = A[(r0 + TJ * x) * LD + c0 + kk];
#pragma unroll
					for (int y = 0; y < 4; ++y) u[y] = A[(c0 + kk) * LD + cc + TJ * y];
#pragma unroll
					for (int x = 0; x < 4; ++x)
#pragma unroll
						for (int y = 0; y < 4; ++y) acc[x][y] = fma(-l[x], u[y], acc[x][y]);
				}
#pragma unroll
				for (int x = 0; x < 4; ++x)
#pragma unroll
					for (int y = 0; y < 4; ++y) A[(r0 + TJ * x) * LD + cc + TJ * y] = acc[x][y];
			} else {
				const int v = unit - n_real;
				const int c = divm(v, m_t2), w = v - c * TI2 * TJ;
				double *Are = Ac + (size_t)c * 2 * N * LD, *Aim = Are + N * LD;
				const int wi = divm(w, m_tj);
				const int r0 = cb + wi, cc = cb + w - wi * TJ;
				double accr[2][4], acci[2][4];
#pragma unroll
				for (int x = 0; x < 2; ++x)
#pragma unroll
					for (int y = 0; y < 4; ++y) {
						accr[x][y] = Are[(r0 + TI2 * x) * LD + cc + TJ * y];
						acci[x][y] = Aim[(r0 + TI2 * x) * LD + cc + TJ * y];
					}
#pragma unroll
				for (int kk = 0; kk < PW; ++kk) {
					double lr[2], li[2], ur[4], ui[4];
#pragma unroll
					for (int x = 0; x < 2; ++x) {
						lr[x] = Are[(r0 + TI2 * x) * LD + c0 + kk];
						li[x] = Aim[(r0 + TI2 * x) * LD + c0 + kk];
					}
#pragma unroll
					for (int y = 0; y < 4; ++y) {
						ur[y] = Are[(c0 + kk) * LD + cc + TJ * y];
						ui[y] = Aim[(c0 + kk) * LD + cc + TJ * y];
					}
#pragma unroll
					for (int x = 0; x < 2; ++x)
#pragma unroll
						for (int y = 0; y < 4; ++y) {
							accr[x][y] = fma(-lr[x], ur[y], fma(li[x], ui[y], accr[x][y]));
							acci[x][y] = fma(-lr[x], ui[y], fma(-li[x], ur[y], acci[x][y]));
						}
				}
#pragma unroll
				for (int x = 0; x < 2; ++x)
#pragma unroll
					for (int y = 0; y < 4; ++y) {
						Are[(r0 + TI2 * x) * LD + cc + TJ * y] = accr[x][y];
						Aim[(r0 + TI2 * x) * LD + cc + TJ * y] = acci[x][y];
					}
			}
		}
		__syncthreads();
		REHUEL_PHASE_MARK(pc, 3);
	}
}

// Warp-synchronous solve with the packed real LU: x <- U^-1 L^-1 P x.  x in shared memory.
// PANEL: factors of lu_factor_set_panel -- the kPanel row exchanges of a panel come before its elimination steps.
template <int N, int LD, int KB, bool PANEL = false>
__device__ __forceinline__ void lu_solve_real_warp(const double *A, const double *rd, const int *perm, double *x)
{
	constexpr int Q = (N + 31) / 32, GROUP = PANEL ? kPanel : 1;
	const int lane = threadIdx.x & 31;
	double v[Q];
#pragma unroll
	for (int q = 0; q < Q; ++q) {
		const int i = lane + 32 * q;
		v[q] = (i < N) ? x[i] : 0.0;
	}
#pragma unroll
	for (int kq = 0; kq < Q; ++kq) // forward, unit lower (kq static: v[] stays in registers)
		for (int kb = 0; kb < 32 && kq * 32 + kb < N; kb += GROUP) {
			int pv[GROUP]; // the group's interchanges first: a load per step would stall the warp every time
#pragma unroll
			for (int g = 0; g < GROUP; ++g) pv[g] = perm[kq * 32 + kb + g];
#pragma unroll
			for (int kk = kb; kk < kb + GROUP; ++kk) {
				const int k = kq * 32 + kk;
				const int p = pv[kk - kb]; // warp-uniform
				if (p != k) { // exchange entries k and p (p > k)
					const double xk = __shfl_sync(0xffffffffu, v[kq], kk);
					const double xp = __shfl_sync(0xffffffffu, (Q > 1 && p >= 32) ? v[Q - 1] : v[0], p & 31);
#pragma unroll
					for (int q = 0; q < Q; ++q) {
						const int i = lane + 32 * q;
						if (i == k) v[q] = xp;
						if (i == p) v[q] = xk;
					}
				}
			}
			// the group's columns first (independent loads), so that the chain of a step is one shuffle and one FMA
			double l[GROUP][Q];
#pragma unroll
			for (int g = 0; g < GROUP; ++g)
#pragma unroll
				for (int q = 0; q < Q; ++q) {
					const int i = lane + 32 * q, k = kq * 32 + kb + g;
					l[g][q] = (q >= kq && i > k && i < N && i <= k + KB) ? A[i * LD + k] : 0.0;
				}
#pragma unroll
			for (int g = 0; g < GROUP; ++g) {
				const int k = kq * 32 + kb + g;
				const double xk = __shfl_sync(0xffffffffu, v[kq], kb + g);
#pragma unroll
				for (int q = kq; q < Q; ++q) {
					const int i = lane + 32 * q;
					if (i > k && i < N && i <= k + KB) v[q] = fma(-l[g][q], xk, v[q]);
				}
			}
		}
#pragma unroll
	for (int kq = Q - 1; kq >= 0; --kq) // backward
		for (int kb = 32 - GROUP; kb >= 0; kb -= GROUP) {
			if (kq * 32 + kb >= N) continue; // N is a multiple of GROUP
			double uu[GROUP][Q], rdk[GROUP];
#pragma unroll
			for (int g = 0; g < GROUP; ++g) {
				const int k = kq * 32 + kb + g;
				rdk[g] = rd[k];
#pragma unroll
				for (int q = 0; q < Q; ++q) {
					const int i = lane + 32 * q;
					uu[g][q] = (q <= kq && i < k && i + 2 * KB >= k) ? A[i * LD + k] : 0.0;
				}
			}
#pragma unroll
			for (int g = GROUP - 1; g >= 0; --g) {
				const int k = kq * 32 + kb + g;
				const double xk = __shfl_sync(0xffffffffu, v[kq], kb + g) * rdk[g];
#pragma unroll
				for (int q = 0; q <= kq; ++q) {
					const int i = lane + 32 * q;
					if (i == k) v[q] = xk;
					if (i < k && i + 2 * KB >= k) v[q] = fma(-uu[g][q], xk, v[q]);
				}
			}
		}
	__syncwarp();
#pragma unroll
	for (int q = 0; q < Q; ++q) {
		const int i = lane + 32 * q;
		if (i < N) x[i] = v[q];
	}
	__syncwarp();
}

template <int N, int LD, int KB, bool PANEL = false>
__device__ __forceinline__ void lu_solve_cplx_warp(const double *Are, const double *rdre, const double *rdim,
                                                   const int *perm, double *xr, double *xi)
{
	constexpr int Q = (N + 31) / 32, GROUP = PANEL ? kPanel : 1;
	const int lane = threadIdx.x & 31;
	const double *Aim = Are + N * LD;
	double vr[Q], vi[Q];
#pragma unroll
	for (int q = 0; q < Q; ++q) {
		const int i = lane + 32 * q;
		vr[q] = (i < N) ? xr[i] : 0.0;
		vi[q] = (i < N) ? xi[i] : 0.0;
	}
#pragma unroll
	for (int kq = 0; kq < Q; ++kq)
		for (int kb = 0; kb < 32 && kq * 32 + kb < N; kb += GROUP) {
			int pv[GROUP];
#pragma unroll
			for (int g = 0; g < GROUP; ++g) pv[g] = perm[kq * 32 + kb + g];
#pragma unroll
			for (int kk = kb; kk < kb + GROUP; ++kk) {
				const int k = kq * 32 + kk;
				const int p = pv[kk - kb];
				if (p != k) {
					const double kr = __shfl_sync(0xffffffffu, vr[kq], kk);
					const double ki = __shfl_sync(0xffffffffu, vi[kq], kk);
					const double pr = __shfl_sync(0xffffffffu, (Q > 1 && p >= 32) ? vr[Q - 1] : vr[0], p & 31);
					const double pi = __shfl_sync(0xffffffffu, (Q > 1 && p >= 32) ? vi[Q - 1] : vi[0], p & 31);
#pragma unroll
					for (int q = 0; q < Q; ++q) {
						const int i = lane + 32 * q;
						if (i == k) { vr[q] = pr; vi[q] = pi; }
						if (i == p) { vr[q] = kr; vi[q] = ki; }
					}
				}
			}
			double lr[GROUP][Q], li[GROUP][Q];
#pragma unroll
			for (int g = 0; g < GROUP; ++g)
#pragma unroll
				for (int q = 0; q < Q; ++q) {
					const int i = lane + 32 * q, k = kq * 32 + kb + g;
					const bool on = q >= kq && i > k && i < N && i <= k + KB;
					lr[g][q] = on ? Are[i * LD + k] : 0.0;
					li[g][q] = on ? Aim[i * LD + k] : 0.0;
				}
#pragma unroll
			for (int g = 0; g < GROUP; ++g) {
				const int k = kq * 32 + kb + g;
				const double kr = __shfl_sync(0xffffffffu, vr[kq], kb + g);
				const double ki = __shfl_sync(0xffffffffu, vi[kq], kb + g);
#pragma unroll
				for (int q = kq; q < Q; ++q) {
					const int i = lane + 32 * q;
					if (i > k && i < N && i <= k + KB) {
						vr[q] = fma(-lr[g][q], kr, fma(li[g][q], ki, vr[q]));
						vi[q] = fma(-lr[g][q], ki, fma(-li[g][q], kr, vi[q]));
					}
				}
			}
		}
#pragma unroll
	for (int kq = Q - 1; kq >= 0; --kq)
		for (int kb = 32 - GROUP; kb >= 0; kb -= GROUP) {
			if (kq * 32 + kb >= N) continue; // N is a multiple of GROUP
			double ur[GROUP][Q], ui[GROUP][Q], dre[GROUP], dim[GROUP];
#pragma unroll
			for (int g = 0; g < GROUP; ++g) {
				const int k = kq * 32 + kb + g;
				dre[g] = rdre[k];
				dim[g] = rdim[k];
#pragma unroll
				for (int q = 0; q < Q; ++q) {
					const int i = lane + 32 * q;
					const bool on = q <= kq && i < k && i + 2 * KB >= k;
					ur[g][q] = on ? Are[i * LD + k] : 0.0;
					ui[g][q] = on ? Aim[i * LD + k] : 0.0;
				}
			}
#pragma unroll
			for (int g = GROUP - 1; g >= 0; --g) {
				const int k = kq * 32 + kb + g;
				const double tr = __shfl_sync(0xffffffffu, vr[kq], kb + g);
				const double ti = __shfl_sync(0xffffffffu, vi[kq], kb + g);
				const double kr = fma(tr, dre[g], -ti * dim[g]);
				const double ki = fma(tr, dim[g], ti * dre[g]);
#pragma unroll
				for (int q = 0; q <= kq; ++q) {
					const int i = lane + 32 * q;
					if (i == k) {
						vr[q] = kr;
						vi[q] = ki;
					}
					if (i < k && i + 2 * KB >= k) {
						vr[q] = fma(-ur[g][q], kr, fma(ui[g][q], ki, vr[q]));
						vi[q] = fma(-ur[g][q], ki, fma(-ui[g][q], kr, vi[q]));
					}
				}
			}
		}
	__syncwarp();
#pragma unroll
	for (int q = 0; q < Q; ++q) {
		const int i = lane + 32 * q;
		if (i < N) {
			xr[i] = vr[q];
			xi[i] = vi[q];
		}
	}
	__syncwarp();
}

template <class F, int S, int NR, int NC, int EST, int THREADS, bool DENSE = false>
__global__ void __launch_bounds__(THREADS, THREADS <= 128 ? 4 : 1)
ensemble_irk_cta(const __grid_constant__ KernelArgs a, const __grid_constant__ DevTableau<S, NC> tb)
{
	static_assert(S == NR + 2 * NC && NR <= 1, "eigen-structure");
	using L = CtaSmem<F, S, NC>;
	constexpr int N = F::N, LD = L::LD, KB = jac_band<F>::value;
	// dense Jacobians in whole panels take the blocked factorisation (REHUEL_CTA_PANEL=0: the column-by-column one)
	constexpr bool PANEL = REHUEL_CTA_PANEL && N % kPanel == 0 && 2 * KB >= N;
	static_assert(N <= 64, "the warp-synchronous solves keep at most two rows per lane");
	extern __shared__ double sm[];
	double *y = sm + L::o_y, *yn = sm + L::o_yn, *par = sm + L::o_p, *Y = sm + L::o_Y, *R = sm + L::o_R,
	       *Z = sm + L::o_Z, *Fv = sm + L::o_F, *ev = sm + L::o_e, *dyv = sm + L::o_dy, *dalt = sm + L::o_dalt,
	       *fy = sm + L::o_fy, *red = sm + L::o_red, *rdr = sm + L::o_rdr, *rdc = sm + L::o_rdc, *rde = sm + L::o_rde,
	       *J = sm + L::o_J, *LUr = sm + L::o_LUr, *LUc = sm + L::o_LUc;
	int *permr = reinterpret_cast<int *>(sm + L::o_int);
	int *permc = permr + N;
	int *perme = permc + N * NC;
	int *s_ctl = permr + ((N * (2 + NC) + 1) & ~1); // [0..1]: member index (64 bit, hence the even offset)
	const int tid = threadIdx.x, warp = tid >> 5;
	const unsigned long long M = a.M;
	const rehuel_device_io &io = a.io;
	PhaseClock pc;
#ifdef REHUEL_PHASE_CLOCKS
	pc.on = tid == 0 && blockIdx.x == 0;
	for (int i = 0; i < 9; ++i) pc.acc[i] = 0;
	pc.start();
#endif

	for (;;) {
		// ---------------- next member
		if (tid == 0) {
			unsigned long long q = atomicAdd(io.queue, 1ull);
			if (q < M && io.order) q = io.order[q];
			*reinterpret_cast<unsigned long long *>(s_ctl) = q;
		}
		__syncthreads();
		const unsigned long long mem = *reinterpret_cast<unsigned long long *>(s_ctl);
		if (mem >= M) break;
		for (int k = tid; k < N; k += THREADS) y[k] = io.y0_broadcast ? io.y0[k] : io.y0[k * M + mem];
		for (int k = tid; k < F::P; k += THREADS) par[k] = io.params[k * M + mem];
		__syncthreads();

		// ---------------- per-member scalar state (identical in every thread)
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
				for (int k = tid; k < N; k += THREADS) io.y_dense[((size_t)kd * N + k) * M + mem] = y[k];
				++kd;
			}
		}
		if (a.store_samples) {
			if (io.n_samples_cap) {
				if (tid == 0) {
					io.t_samples[mem] = t;
					if (io.err_samples) io.err_samples[mem] = 0.0;
				}
				for (int k = tid; k < N; k += THREADS) io.y_samples[k * M + mem] = y[k];
			}
			n_stored = 1;
		}

		// residual R = Y - dt*(A (x) I) F(y+Y), returns |R|^2   (irk.hpp:366-386)
		auto residual = [&]() -> double {
			for (int e = tid; e < S * N; e += THREADS) Z[e] = y[e % N] + Y[e]; // stage states (Z is free here)
			__syncthreads();
			for (int e = tid; e < S * N; e += THREADS) {
				const int s = e / N, i = e % N;
				Fv[e] = F::fun(i, fma(tb.c[s], dt, t), Z + s * N, par);
			}
			__syncthreads();
			c_fun += S;
			double r2 = 0.0;
			for (int e = tid; e < S * N; e += THREADS) {
				const int s = e / N, i = e % N;
				double acc = 0.0;
#pragma unroll
				for (int j = 0; j < S; ++j) acc = fma(tb.A[s][j], Fv[j * N + i], acc);
				const double r = fma(-dt, acc, Y[e]);
				R[e] = r;
				r2 = fma(r, r, r2);
			}
			return block_sum<THREADS>(r2, red);
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
			// -------- Jacobian and the decoupled Newton matrices (irk.hpp:422-436)
			const double rdt = 1.0 / dt;
			const double mur = (EST == EST_REUSE) ? 1.0 / (tb.gamma * dt) : tb.gamma_hat * rdt;
			const double mue = (EST == EST_OWN_LU) ? 1.0 / (tb.gamma * dt) : 0.0;
			for (int e = tid; e < N * LD; e += THREADS) J[e] = 0.0;
			__syncthreads();
			for (int i = tid; i < N; i += THREADS) F::jac_row(i, t, y, par, J + i * LD);
			++c_jac;
			__syncthreads();
			for (int e = tid; e < N * N; e += THREADS) {
				const int i = e / N, j = e % N;
				const double jv = J[i * LD + j];
				if (NR) LUr[i * LD + j] = (i == j ? mur : 0.0) - jv;
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					LUc[(size_t)c * 2 * N * LD + i * LD + j] = (i == j ? tb.alpha[c] * rdt : 0.0) - jv;
					LUc[(size_t)c * 2 * N * LD + N * LD + i * LD + j] = (i == j ? tb.beta[c] * rdt : 0.0);
				}
				// E/gam = (1/gam) I - J of the error estimate gets its own factors, in place of J (read above by this thread)
				if (EST == EST_OWN_LU) J[i * LD + j] = (i == j ? mue : 0.0) - jv;
			}
			__syncthreads();
			// all factorisations of the attempt in the same N steps, the estimator's included
			REHUEL_PHASE_MARK(pc, 0);
			if constexpr (PANEL) lu_factor_set_panel<N, LD, NR, NC, (EST == EST_OWN_LU ? 1 : 0), THREADS>(LUr, LUc, J, rdr, rdc, rde, permr, permc, perme, pc);
			else lu_factor_set<N, LD, NR, NC, (EST == EST_OWN_LU ? 1 : 0), THREADS, KB>(LUr, LUc, J, rdr, rdc, rde, permr, permc, perme);

			// -------- simplified Newton (irk.hpp:398-493)
			for (int e = tid; e < S * N; e += THREADS) Y[e] = 0.0; // irk.hpp:415
			__syncthreads();
			REHUEL_PHASE_MARK(pc, 8);
			double Rnorm2 = residual();
			REHUEL_PHASE_MARK(pc, 4);
			int nstat = kNewtonMaxit;
			int it = 1;
			for (; it < maxit; ++it) { // irk.hpp:458
				for (int e = tid; e < S * N; e += THREADS) { // Z = (Ti (x) I) R
					const int s = e / N, i = e % N;
					double acc = 0.0;
#pragma unroll
					for (int j = 0; j < S; ++j) acc = fma(tb.Ti[s][j], R[j * N + i], acc);
					Z[e] = acc;
				}
				__syncthreads();
				REHUEL_PHASE_MARK(pc, 5);
				if (NR && warp == 0) {
					lu_solve_real_warp<N, LD, KB, PANEL>(LUr, rdr, permr, Z);
					for (int i = tid & 31; i < N; i += 32) Z[i] *= -mur;
				}
				if (warp >= NR && warp < NR + NC) {
					const int c = warp - NR;
					double *zr = Z + (NR + 2 * c) * N, *zi = zr + N;
					lu_solve_cplx_warp<N, LD, KB, PANEL>(LUc + (size_t)c * 2 * N * LD, rdc + (c * 2) * N, rdc + (c * 2 + 1) * N,
					                          permc + c * N, zr, zi);
					const double mr = tb.alpha[c] * rdt, mi = tb.beta[c] * rdt;
					for (int i = tid & 31; i < N; i += 32) { // w = -(mr + i mi) w
						const double wr = zr[i], wi = zi[i];
						zr[i] = fma(-mr, wr, mi * wi);
						zi[i] = fma(-mr, wi, -mi * wr);
					}
				}
				__syncthreads();
				REHUEL_PHASE_MARK(pc, 6);
				double x2 = 0.0;
				for (int e = tid; e < S * N; e += THREADS) { // dY = (T (x) I) W ; Y += dY
					const int s = e / N, i = e % N;
					double acc = 0.0;
#pragma unroll
					for (int j = 0; j < S; ++j) acc = fma(tb.T[s][j], Z[j * N + i], acc);
					x2 = fma(acc, acc, x2);
					Y[e] += acc;
				}
				const double xnorm2 = block_sum<THREADS>(x2, red); // irk.hpp:466 (also orders Z reads before reuse)
				REHUEL_PHASE_MARK(pc, 5);
				if (REHUEL_SKIP_DEAD_RESIDUAL && xnorm2 < a.xtol2) { // the residual of irk.hpp:469-472 would never be looked at
					c_fun += S;                                       // (see irk_thread.cuh); its evaluations are booked
					nstat = kNewtonSuccess;
					break;
				}
				Rnorm2 = residual();
				REHUEL_PHASE_MARK(pc, 4);
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
				if (threadIdx.x == 0) trace_row(a, mem, step, t, dt, -1.0, it);
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
			REHUEL_PHASE_MARK(pc, 8);

			// -------- new solution + embedded error (irk.hpp:691-731)
			const double gam = tb.gamma * dt;
			for (int i = tid; i < N; i += THREADS) {
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
			__syncthreads();
			for (int i = tid; i < N; i += THREADS) {
				const double dyalt = (EST != EST_IDENTITY) ? fma(gam, fy[i], dalt[i]) : dalt[i];
				ev[i] = dyalt - dyv[i];
			}
			__syncthreads();
			const double esc = (EST == EST_IDENTITY) ? dt : 1.0 / tb.gamma; // dt*E^-1 = (1/gamma)(mu I - J)^-1
			auto apply_Einv_dt = [&]() {
				if (EST != EST_IDENTITY) {
					if (warp == 0) {
						if (EST == EST_OWN_LU) lu_solve_real_warp<N, LD, KB, PANEL>(J, rde, perme, ev);
						else lu_solve_real_warp<N, LD, KB, PANEL>(LUr, rdr, permr, ev);
					}
					__syncthreads();
				}
				for (int i = tid; i < N; i += THREADS) ev[i] *= esc;
				__syncthreads();
			};
			apply_Einv_dt(); // 8.19
			if (alt) {       // 8.20, irk.hpp:722-731
				if (EST != EST_IDENTITY) {
					for (int i = tid; i < N; i += THREADS) Z[i] = y[i] + ev[i];
					__syncthreads();
					for (int i = tid; i < N; i += THREADS) fy[i] = F::fun(i, t, Z, par);
					__syncthreads();
				}
				++c_fun;
				for (int i = tid; i < N; i += THREADS) {
					const double v = (EST != EST_IDENTITY) ? fma(gam, fy[i], dalt[i]) : dalt[i];
					ev[i] = v - dyv[i];
				}
				__syncthreads();
				apply_Einv_dt();
			}
			double tot = 0.0;
			for (int i = tid; i < N; i += THREADS) { // irk.hpp:733-746
				const double sc = fma(m_rtol, fmax(fabs(y[i]), fabs(yn[i])), m_atol);
				const double q = ev[i] / sc;
				tot = fma(q, q, tot);
			}
			tot = block_sum<THREADS>(tot, red);
			REHUEL_PHASE_MARK(pc, 7);
			err = sqrt(tot / (double)N);
			if (err < kMachinePrecision) err = kMachinePrecision;
			const bool rejected = a.adaptive && (err > 1.0); // irk.hpp:761
			if (rejected) {
				alt = true;
				++c_rej_err;
			}
			// -------- step-size proposal (irk.hpp:770-801; shifts the error history, irk.hpp:756-758)
			const double new_dt = propose_dt(dt, err, q_prev, dts0, dts1, maxit, it, tb.min_order, tb.expo, a.max_dt);
			if (threadIdx.x == 0) trace_row(a, mem, step, t, dt, err, it); // irk.hpp:805-810
			bool stuck = false;
			if (!rejected) { // irk.hpp:813-846
				if (DENSE) { // grid times inside (t, t + dt] from this step's collocation polynomial
					const double t_new = t + dt;
					while (kd < io.n_dense) {
						const double tk = fma((double)kd, io.dense_dt, io.dense_t0);
						if (tk > t_new) break;
						double w[S];
						dense_weights<S, NC>(tb, (tk - t) * rdt, w);
						for (int i = tid; i < N; i += THREADS) {
							double acc = y[i];
#pragma unroll
							for (int s = 0; s < S; ++s) acc = fma(w[s], Y[s * N + i], acc);
							io.y_dense[((size_t)kd * N + i) * M + mem] = (tk == t_new) ? yn[i] : acc;
						}
						++kd;
					}
				}
				for (int i = tid; i < N; i += THREADS) y[i] = yn[i];
				const double t_old = t;
				t += dt;
				++step;
				if (t == t_old) stuck = true;
				if (a.store_samples && (a.output_interval == 1ull || (unsigned long long)step % a.output_interval == 0ull)) {
					if (n_stored < io.n_samples_cap) {
						const size_t row = n_stored;
						if (tid == 0) {
							io.t_samples[row * M + mem] = t;
							if (io.err_samples) io.err_samples[row * M + mem] = err;
						}
						for (int i = tid; i < N; i += THREADS) io.y_samples[(row * N + i) * M + mem] = yn[i];
					}
					++n_stored;
				}
				alt = false;
				__syncthreads();
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
		__syncthreads();
		for (int k = tid; k < N; k += THREADS) io.y[k * M + mem] = y[k];
		if (tid == 0) {
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
		__syncthreads();
	}
#ifdef REHUEL_PHASE_CLOCKS
	if (pc.on) {
		long long tot = 0;
		for (int i = 0; i < 9; ++i) tot += pc.acc[i];
		printf("cta phase clocks (block 0, %lld cycles): setup %.1f%% panel %.1f%% exch+L11 %.1f%% update %.1f%% residual %.1f%% "
		       "transforms %.1f%% newton-solves %.1f%% estimate %.1f%% rest %.1f%%\n", tot, 100.0 * pc.acc[0] / tot, 100.0 * pc.acc[1] / tot,
		       100.0 * pc.acc[2] / tot, 100.0 * pc.acc[3] / tot, 100.0 * pc.acc[4] / tot, 100.0 * pc.acc[5] / tot, 100.0 * pc.acc[6] / tot,
		       100.0 * pc.acc[7] / tot, 100.0 * pc.acc[8] / tot);
	}
#endif
}

} // namespace rehuel
