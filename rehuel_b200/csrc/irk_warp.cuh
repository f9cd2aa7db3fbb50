// Warp-per-system ensemble integrator for systems with a BANDED Jacobian (4 < n <= 64, half
// bandwidth kBand <= 4): one warp runs one member's adaptive integration, persistent warps pull
// members from the global work queue.
//
// Same algorithm, statement for statement, as the other two kernels (irk_thread.cuh,
// irk_cta.cuh) and hence as the reference loop irk.hpp:604-860 / newton_solve_stages
// irk.hpp:398-493.  What the band buys: the decoupled Newton matrices  mu I - J  keep the band of
// J, their LU costs O(n*kBand^2) instead of O(n^3), and a factorisation or a triangular solve is a
// short SEQUENTIAL recurrence -- no work for a whole thread block.  What bounds the kernel is the
// LATENCY of those recurrences times the number of members an SM can hold, so:
//   * the recurrences are halved: every matrix is factorised, every right-hand side solved, by a
//     PAIR of lanes with the twisted factorisation of band_twisted.cuh (one lane top-down, its
//     neighbour bottom-up, meeting on kBand separator rows); all matrices of an attempt side by
//     side (lanes 2m, 2m+1 <-> matrix m), one routine for real and complex ones;
//   * shared memory per member is what caps the resident warps, so it only holds what lanes
//     must see of each other: y, one set of work vectors (stage states -> right-hand sides of the
//     solves) and the band factors in the narrow twisted layout (~20 KB for n = 64, S = 5: eleven
//     warps per SM).  The stage increments Y, the residual R and everything of the update /
//     estimator phase live in REGISTERS, component i on lane i mod 32: the (Ti (x) I), (T (x) I)
//     and (A (x) I) transforms never leave the lane;
//   * the twisted routines do not interchange rows; they report whether partial pivoting would
//     have, and then the attempt's matrices are rebuilt in the wide pivoted layout (dgbtf2's
//     pattern: L within kBand below the diagonal, U within 2*kBand above it) in a per-warp slab of
//     GLOBAL memory and factorised / solved by the one-lane-per-matrix routines below.  mu I - J
//     practically never needs it (the 1-D Brusselator never does), so its cost does not matter,
//     but it keeps the kernel correct for any banded functor;
//   * scalar control (dt, controller history, counters) is kept redundantly by every lane.
//
// Functor concept: the componentwise one of irk_cta.cuh plus `static constexpr int kBand`.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "../../include/rehuel_b200.h"
#include "band_twisted.cuh"
#include "irk_cta.cuh"
#include "irk_thread.cuh"
#include "tableau.hpp"

namespace rehuel {

template <class F, int S, int NC, int EST>
struct WarpSmem {
	static constexpr int N = F::N;
	static constexpr int KB = jac_band<F>::value;
	using G = Tw<N, KB>;
	static constexpr int PA = ((F::P > 0 ? F::P : 1) + 1) & ~1;
	static constexpr int NREAL = 1 + (EST == EST_OWN_LU ? 1 : 0); // slot 0: real Newton block, slot 1: estimator
	static constexpr int NMAT = NREAL + NC;
	// ---- shared memory, offsets in doubles
	static constexpr int o_y = 0;
	static constexpr int o_p = o_y + N;
	static constexpr int o_w = o_p + PA;                          // work vectors [S][N] (+ KB entries the pivoted forward sweep may look at)
	static constexpr int o_fr = o_w + S * N + KB;                 // real twisted factors    [NREAL][half][HP]
	static constexpr int o_fc = o_fr + NREAL * 2 * G::HP;         // complex twisted factors [NC][half][re|im][HP]
	static constexpr int o_q = o_fc + NC * 2 * 2 * G::HP;         // inverse separator blocks [NMAT][half][re|im][KB*KB]
	static constexpr int total = o_q + NMAT * 2 * 2 * KB * KB;
	static constexpr size_t bytes = (size_t)total * 8;
	// ---- per-warp slab in global memory for the pivoted fallback, offsets in doubles
	static constexpr int W = 3 * KB + 1;                          // band row: KB multipliers | diagonal | 2*KB of U
	static constexpr int PL = (N + KB) * W;                       // one band plane: N rows + KB zero rows of padding
	static constexpr int g_abr = 0;                               // real band factors [NREAL][N + KB][W]
	static constexpr int g_rdr = g_abr + NREAL * PL;              // reciprocal pivots [NREAL][N]
	static constexpr int g_abc = g_rdr + NREAL * N;               // complex band factors [NC][2][N + KB][W]
	static constexpr int g_rdc = g_abc + NC * 2 * PL;             // reciprocal pivots [NC][2][N]
	static constexpr int g_piv = g_rdc + NC * 2 * N;              // interchange offsets, bytes: [NMAT][N]
	static constexpr size_t slab_doubles = (size_t)g_piv + ((size_t)NMAT * N + 7) / 8;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// Row i of the Jacobian as its band, band[KB + o] = J(i, i + o): straight from the functor's jac_band if it has one
// (registers), else cut out of a dense row (local memory: measurably slower, but every componentwise functor works).
template <class F, int KB>
__device__ __forceinline__ void jacobian_band(int i, double t, const double *y, const double *par, double (&band)[2 * KB + 1])
{
	if constexpr (has_jac_band<F>::value) {
		F::jac_band(i, t, y, par, band);
	} else {
		constexpr int N = F::N;
		double row[N];
#pragma unroll
		for (int j = 0; j < N; ++j) row[j] = 0.0;
		F::jac_row(i, t, y, par, row);
#pragma unroll
		for (int o = -KB; o <= KB; ++o) {
			const int j = i + o;
			band[KB + o] = (j >= 0 && j < N) ? row[j] : 0.0;
		}
	}
}

// ------------------------------------------------------------------ twisted band LU / solve, one lane PAIR per matrix
// `mask`: the lanes that execute the call together (whole pairs).  The sweeps are band_twisted.cuh's; here the two
// halves exchange their separator data with one shuffle per number.
template <int N, int KB>
__device__ __forceinline__ bool tw_factor_pair(unsigned mask, double *pr, double *pi, bool is_c, int np, double *qr, double *qi)
{
	double sr[KB][KB], si[KB][KB], tr[KB][KB], ti[KB][KB];
	bool flag = tw_lu_sweep<N, KB>(pr, pi, is_c, np, sr, si);
	__syncwarp(mask);
#pragma unroll
	for (int j = 0; j < KB; ++j)
#pragma unroll
		for (int c = 0; c < KB; ++c) { // the partner's block, turned into this half's orientation
			tr[j][c] = __shfl_xor_sync(mask, sr[KB - 1 - j][KB - 1 - c], 1);
			ti[j][c] = __shfl_xor_sync(mask, si[KB - 1 - j][KB - 1 - c], 1);
		}
	flag |= tw_separator<N, KB>(pr, pi, is_c, np, sr, si, tr, ti, qr, qi);
	return flag;
}

// x <- A^-1 x.  xr/xi point at this half's first unknown, xs = +1 / -1 (see tw_forward).
template <int N, int KB>
__device__ __forceinline__ void tw_solve_pair(unsigned mask, const double *pr, const double *pi, bool is_c, int np, const double *qr,
                                              const double *qi, double *xr, double *xi, int xs)
{
	double gr[KB], gi[KB], hr[KB], hi[KB], sr[KB], si[KB];
	tw_forward<N, KB>(pr, pi, is_c, np, xr, xi, xs, gr, gi);
	__syncwarp(mask);
#pragma unroll
	for (int j = 0; j < KB; ++j) {
		hr[j] = __shfl_xor_sync(mask, gr[KB - 1 - j], 1);
		hi[j] = __shfl_xor_sync(mask, gi[KB - 1 - j], 1);
	}
	tw_separator_solve<N, KB>(qr, qi, is_c, np, xr, xi, xs, gr, gi, hr, hi, sr, si);
	tw_backward<N, KB>(pr, pi, is_c, np, xr, xi, xs, sr, si);
	__syncwarp(mask); // both halves have read the separator's right-hand side
	if (xs > 0) {
#pragma unroll
		for (int j = 0; j < KB; ++j) {
			xr[np + j] = sr[j];
			if (is_c) xi[np + j] = si[j];
		}
	}
}

// ------------------------------------------------------------------ pivoted band LU, one lane per matrix (fallback)
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
__global__ void __launch_bounds__(32, 11)
ensemble_irk_warp(const __grid_constant__ KernelArgs a, const __grid_constant__ DevTableau<S, NC> tb)
{
	static_assert(S == NR + 2 * NC && NR <= 1, "eigen-structure");
	using L = WarpSmem<F, S, NC, EST>;
	using G = typename L::G;
	constexpr int N = F::N, KB = L::KB, W = L::W, HP = G::HP;
	constexpr int NPL = (N + 31) / 32;     // components per lane: i = lane + 32*j
	constexpr bool FULL = (N % 32 == 0);   // every lane owns NPL components
	static_assert(KB >= 1 && KB <= 4, "band kernels are built for half bandwidths 1..4");
	static_assert(2 * (1 + NC + 1) <= 32, "one lane pair per matrix");
	static_assert(EST != EST_REUSE || NR == 1, "E's factors are the real Newton block's");
	extern __shared__ double sm[];
	double *y = sm + L::o_y, *par = sm + L::o_p, *work = sm + L::o_w;
	const int lane = threadIdx.x;
	const unsigned long long M = a.M;
	const rehuel_device_io &io = a.io;
	// ---- twisted roles: lanes 2m, 2m+1 <-> matrix m (0: real Newton block, 1..NC: complex ones, NC+1: estimator's own),
	//      even lane = top half, odd lane = bottom half
	const int tm = lane >> 1, th = lane & 1;
	const bool t_real = NR && tm == 0;
	const bool t_cplx = tm >= 1 && tm <= NC;
	const bool t_est = (EST == EST_OWN_LU) && tm == NC + 1;
	const bool t_has = t_real || t_cplx || t_est;
	const int t_c = t_cplx ? tm - 1 : 0;
	double *t_pr = t_cplx ? sm + L::o_fc + ((t_c * 2 + th) * 2) * HP : sm + L::o_fr + ((t_est ? 1 : 0) * 2 + th) * HP;
	double *t_pi = t_cplx ? t_pr + HP : t_pr;
	const int t_slot = t_cplx ? L::NREAL + t_c : (t_est ? 1 : 0);
	double *t_qr = sm + L::o_q + ((t_slot * 2 + th) * 2) * KB * KB, *t_qi = t_qr + KB * KB;
	const int t_np = th ? G::NP1 : G::NP0, t_xs = th ? -1 : 1;
	const bool t_rhs = t_real || t_cplx;                       // right-hand sides of a Newton update
	double *t_xr = (t_cplx ? work + (NR + 2 * t_c) * N : work) + (th ? N - 1 : 0);
	double *t_xi = t_cplx ? t_xr + N : t_xr;
	const bool t_E = (EST == EST_REUSE) ? t_real : t_est;      // the pair that holds the factors of E = I - gamma dt J
	double *t_er = work + (th ? N - 1 : 0);                    // E's right-hand side is work[0..N)
	const unsigned m_has = __ballot_sync(0xffffffffu, t_has), m_rhs = __ballot_sync(0xffffffffu, t_rhs),
	               m_E = __ballot_sync(0xffffffffu, t_E);
	// ---- pivoted fallback roles: lane m <-> matrix m, factors in this warp's global slab
	double *slab = a.pivot_scratch ? a.pivot_scratch + (size_t)blockIdx.x * L::slab_doubles : nullptr;
	double *abr = slab + L::g_abr, *rdr = slab + L::g_rdr, *abc = slab + L::g_abc, *rdc = slab + L::g_rdc;
	signed char *piv = reinterpret_cast<signed char *>(slab + L::g_piv); // [real slots][N] then [NC][N]
	signed char *pivc = piv + L::NREAL * N;
	constexpr int PL = L::PL;
	const bool p_cplx = lane >= 1 && lane <= NC;
	const bool p_est = (EST == EST_OWN_LU) && lane == NC + 1;
	const bool p_has = (NR && lane == 0) || p_cplx || p_est;
	const int p_c = p_cplx ? lane - 1 : 0;
	double *p_abr = p_cplx ? abc + (p_c * 2) * PL : (p_est ? abr + PL : abr);
	double *p_abi = p_cplx ? abc + (p_c * 2 + 1) * PL : p_abr;
	double *p_rdr = p_cplx ? rdc + (p_c * 2) * N : (p_est ? rdr + N : rdr);
	double *p_rdi = p_cplx ? rdc + (p_c * 2 + 1) * N : p_rdr;
	signed char *p_piv = p_cplx ? pivc + p_c * N : (p_est ? piv + N : piv);
	const bool p_rhs = (NR && lane == 0) || p_cplx;
	double *p_xr = p_cplx ? work + (NR + 2 * p_c) * N : work;
	double *p_xi = p_cplx ? work + (NR + 2 * p_c + 1) * N : work;

	PhaseClock pc;
#ifdef REHUEL_PHASE_CLOCKS
	pc.on = threadIdx.x == 0 && blockIdx.x == 0;
	for (int i = 0; i < 9; ++i) pc.acc[i] = 0;
	pc.start();
#endif
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
			// -------- Jacobian rows straight into the band storage of every matrix (irk.hpp:422-436)
			const double rdt = 1.0 / dt;
			const double mur = (EST == EST_REUSE) ? 1.0 / (tb.gamma * dt) : tb.gamma_hat * rdt;
			const double mue = (EST == EST_OWN_LU) ? 1.0 / (tb.gamma * dt) : 0.0;
			auto build_twisted = [&]() {
#pragma unroll
				for (int jj = 0; jj < NPL; ++jj) {
					const int i = lane + 32 * jj;
					if (FULL || i < N) {
						double band[2 * KB + 1]; // J(i, i-KB .. i+KB)
						jacobian_band<F, KB>(i, t, y, par, band);
#pragma unroll
						for (int d = -KB; d <= KB; ++d) {
							const double jv = band[KB + d];
							const double dg = (d == 0) ? 1.0 : 0.0;
							if (NR) tw_put<N, KB>(sm + L::o_fr, sm + L::o_fr + HP, i, d, dg * mur - jv);
							if (EST == EST_OWN_LU) tw_put<N, KB>(sm + L::o_fr + 2 * HP, sm + L::o_fr + 3 * HP, i, d, dg * mue - jv);
#pragma unroll
							for (int c = 0; c < NC; ++c) {
								double *top = sm + L::o_fc + (c * 4) * HP, *bot = top + 2 * HP;
								tw_put<N, KB>(top, bot, i, d, dg * (tb.alpha[c] * rdt) - jv);
								tw_put<N, KB>(top + HP, bot + HP, i, d, dg * (tb.beta[c] * rdt));
							}
						}
					}
				}
				__syncwarp();
			};
			auto build_pivoted = [&]() {
				for (int i = lane; i < N; i += 32) {
					double band[2 * KB + 1];
					jacobian_band<F, KB>(i, t, y, par, band);
#pragma unroll
					for (int d = -KB; d <= 2 * KB; ++d) {
						const double jv = (d <= KB) ? band[KB + (d <= KB ? d : 0)] : 0.0;
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
			// -------- all factorisations side by side, one lane pair per matrix, one routine for all of them;
			// if partial pivoting would have exchanged rows anywhere, redo them pivoted (one lane per matrix)
			REHUEL_PHASE_MARK(pc, 8);
			build_twisted();
			REHUEL_PHASE_MARK(pc, 0);
			bool flag = false;
			if (t_has) flag = tw_factor_pair<N, KB>(m_has, t_pr, t_pi, t_cplx, t_np, t_qr, t_qi);
			__syncwarp();
			const bool piv_any = __any_sync(0xffffffffu, flag) || a.force_pivot;
			if (piv_any) {
				if (!slab) { // launch_warp always provides one
					status = REHUEL_GENERAL_ERROR;
					break;
				}
				build_pivoted();
				if (p_has) band_lu<N, KB, true>(p_abr, p_abi, p_cplx, p_rdr, p_rdi, p_piv);
				__syncwarp();
			}

			// -------- simplified Newton (irk.hpp:398-493)
#pragma unroll
			for (int s = 0; s < S; ++s)
#pragma unroll
				for (int j = 0; j < NPL; ++j) Yr[s][j] = 0.0; // irk.hpp:415
			REHUEL_PHASE_MARK(pc, 1);
			double Rnorm2 = residual();
			REHUEL_PHASE_MARK(pc, 2);
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
				REHUEL_PHASE_MARK(pc, 3);
				if (!piv_any) { // all right-hand sides side by side, one lane pair each
					if (t_rhs) tw_solve_pair<N, KB>(m_rhs, t_pr, t_pi, t_cplx, t_np, t_qr, t_qi, t_xr, t_xi, t_xs);
				} else {
					if (p_rhs) band_solve<N, KB, true>(p_abr, p_abi, p_cplx, p_rdr, p_rdi, p_piv, p_xr, p_xi);
				}
				__syncwarp();
				REHUEL_PHASE_MARK(pc, 4);
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
				REHUEL_PHASE_MARK(pc, 3);
				if (REHUEL_SKIP_DEAD_RESIDUAL && xnorm2 < a.xtol2) { // the residual of irk.hpp:469-472 would never be looked at
					c_fun += S;                                       // (see irk_thread.cuh); its evaluations are booked
					nstat = kNewtonSuccess;
					break;
				}
				Rnorm2 = residual();
				REHUEL_PHASE_MARK(pc, 2);
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
			REHUEL_PHASE_MARK(pc, 8);

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
					REHUEL_PHASE_MARK(pc, 5);
					if (!piv_any) {
						if (t_E) tw_solve_pair<N, KB>(m_E, t_pr, t_pr, false, t_np, t_qr, t_qi, t_er, t_er, t_xs);
					} else if (lane == 0) {
						const int o = (EST == EST_REUSE) ? 0 : 1; // which real plane holds E's factors
						band_solve<N, KB, true>(abr + o * PL, abr, false, rdr + o * N, rdr, piv + o * N, work, work);
					}
					__syncwarp();
					REHUEL_PHASE_MARK(pc, 6);
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
			REHUEL_PHASE_MARK(pc, 5);
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
#ifdef REHUEL_PHASE_CLOCKS
	if (pc.on) {
		long long tot = 0;
		for (int i = 0; i < 9; ++i) tot += pc.acc[i];
		printf("warp phase clocks (warp 0, %lld cycles): build %.1f%% factor %.1f%% residual %.1f%% transforms %.1f%% newton-solves %.1f%% "
		       "estimate(other) %.1f%% estimate-solve %.1f%% rest %.1f%%\n", tot, 100.0 * pc.acc[0] / tot, 100.0 * pc.acc[1] / tot,
		       100.0 * pc.acc[2] / tot, 100.0 * pc.acc[3] / tot, 100.0 * pc.acc[4] / tot, 100.0 * pc.acc[5] / tot, 100.0 * pc.acc[6] / tot,
		       100.0 * pc.acc[8] / tot);
	}
#endif
}

} // namespace rehuel
