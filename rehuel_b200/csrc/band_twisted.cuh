// Twisted ("burn at both ends") LU and solve for banded matrices, two lanes per matrix.
//
// The warp-per-system kernel (irk_warp.cuh) solves with the decoupled Newton matrices  mu I - J
// (irk.hpp:427-436 after the eigen-decoupling of DESIGN.md 5.1), J banded with half bandwidth KB.
// A band LU or band solve is a SEQUENTIAL recurrence over the N rows; its length, not its flop
// count, is what an attempt of a 64-equation member costs.  The twisted factorisation halves it:
//
//   unknowns   T = {0 .. NP0-1} | separator S = {NP0 .. NP0+KB-1} | B = {NP0+KB .. N-1}
//
// T and B are not coupled directly (band), so one lane eliminates T top-down while its neighbour
// eliminates B bottom-up -- which is the SAME top-down routine on the matrix with rows and
// columns reversed.  Both sweeps end on the KB separator rows; their Schur complements add up,
//   S_c = A_SS - A_ST A_TT^-1 A_TS - A_SB A_BB^-1 A_BS ,
// a dense KB x KB block that is inverted once per factorisation.  A solve runs the two forward
// sweeps side by side, exchanges KB numbers (one shuffle each), applies S_c^-1, and runs the two
// back substitutions side by side, outwards from the separator.
//
// No row interchanges here.  Every sweep reports whether partial pivoting WOULD have exchanged
// rows somewhere (a pivot smaller than an entry it eliminates); the caller then redoes the
// attempt's factorisations with the pivoted one-lane-per-matrix routines of irk_warp.cuh.
//
// Storage of one half (`Tw<N,KB>::HP` doubles per plane, real and imaginary planes separate):
// rows in the half's own order (bottom half: reversed), row r holds A(r, r-KB .. r+KB) at
// [r*W + 0 .. 2KB], W = 2KB+1, for the NP pivot rows, the KB separator rows and one row of
// look-ahead.  After tw_lu_sweep: reciprocal pivot at [k*W+KB], U(k,k+c) at [k*W+KB+c],
// multipliers L(k+r,k) at [(k+r)*W+KB-r]; in the separator rows the multipliers end left of
// column NP, so A_SS is still there when the separator block is put together.
//
// The sweeps are __host__ __device__: tests/band_twisted_host.cu runs the very same code on the
// CPU against a dense solve.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace rehuel {

template <int N, int KB>
struct Tw {
	static_assert(KB >= 1 && N - KB >= 2, "twisted band factorisation needs at least one pivot row on either side of the separator");
	static constexpr int NP0 = (N - KB) / 2; // pivot rows of the top half
	static constexpr int NP1 = N - KB - NP0; // pivot rows of the bottom half (>= NP0)
	static constexpr int W = 2 * KB + 1;
	static constexpr int HR = NP1 + KB + 1;  // stored rows per half: pivot rows, separator rows, one row of look-ahead
	static constexpr int HP = HR * W;        // doubles per half-plane
	static_assert(HR <= N, "half planes hold rows of the matrix only");
};

__host__ __device__ __forceinline__ double tw_rcp(double x)
{
#ifdef __CUDA_ARCH__
	double r; // MUFU.RCP64H seed + one cubic step, see fast_rcp (irk_thread.cuh)
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	const double e = fma(-x, r, 1.0); // |e| <= 2^-20: the seed carries the upper 20 mantissa bits
	const double t = fma(e, e, e);    // r (1 + e + e^2): the neglected e^3 is below 2^-60
	return fma(r, t, r);
#else
	return 1.0 / x;
#endif
}

// One half's elimination sweep over its `np` pivot rows.  pr/pi: the half's planes (pi is never
// touched for a real matrix).  Out: the separator block as reduced by THIS half in sr/si (own
// orientation).  Returns true if partial pivoting would have exchanged rows.
template <int N, int KB>
__host__ __device__ __forceinline__ bool tw_lu_sweep(double *pr, double *pi, bool is_c, int np, double (&sr)[KB][KB],
                                                     double (&si)[KB][KB])
{
	constexpr int W = Tw<N, KB>::W, WW = KB + 1;
	double wr[KB + 1][WW], wi[KB + 1][WW]; // rows k..k+KB, columns k..k+KB
#pragma unroll
	for (int r = 0; r <= KB; ++r)
#pragma unroll
		for (int c = 0; c < WW; ++c) {
			wr[r][c] = pr[r * W + KB + c - r];
			wi[r][c] = is_c ? pi[r * W + KB + c - r] : 0.0;
		}
	bool would_swap = false;
#pragma unroll 1
	for (int k = 0; k < np; ++k) {
		const double best = fabs(wr[0][0]) + fabs(wi[0][0]); // izamax's |re|+|im|
#pragma unroll
		for (int r = 1; r <= KB; ++r) would_swap |= (fabs(wr[r][0]) + fabs(wi[r][0]) > best);
		// reciprocal pivot 1/(a+ib) = (a-ib)/(a^2+b^2); it replaces the diagonal
		const double ar = wr[0][0], ai = wi[0][0];
		const double rden = tw_rcp(fma(ar, ar, ai * ai));
		const double rr = ar * rden, ri = -ai * rden;
		pr[k * W + KB] = rr;
		if (is_c) pi[k * W + KB] = ri;
#pragma unroll
		for (int c = 1; c < WW; ++c) {
			pr[k * W + KB + c] = wr[0][c];
			if (is_c) pi[k * W + KB + c] = wi[0][c];
		}
#pragma unroll
		for (int r = 1; r <= KB; ++r) {
			const double lr = fma(wr[r][0], rr, -wi[r][0] * ri);
			const double li = fma(wr[r][0], ri, wi[r][0] * rr);
			pr[(k + r) * W + KB - r] = lr;
			if (is_c) pi[(k + r) * W + KB - r] = li;
#pragma unroll
			for (int c = 1; c < WW; ++c) {
				const double ur = wr[0][c], ui = wi[0][c];
				wr[r][c] = fma(-lr, ur, fma(li, ui, wr[r][c]));
				wi[r][c] = fma(-lr, ui, fma(-li, ur, wi[r][c]));
			}
		}
		// slide to rows k+1..k+1+KB, columns k+1..k+1+KB
#pragma unroll
		for (int r = 0; r < KB; ++r) {
#pragma unroll
			for (int c = 0; c + 1 < WW; ++c) {
				wr[r][c] = wr[r + 1][c + 1];
				wi[r][c] = wi[r + 1][c + 1];
			}
			// the column that enters is the row's own last band entry, not touched by any earlier pivot row
			wr[r][WW - 1] = pr[(k + 1 + r) * W + 2 * KB - r];
			wi[r][WW - 1] = is_c ? pi[(k + 1 + r) * W + 2 * KB - r] : 0.0;
		}
#pragma unroll
		for (int c = 0; c < WW; ++c) { // the row that enters (the look-ahead row at the last step; never used then)
			wr[KB][c] = pr[(k + 1 + KB) * W + c];
			wi[KB][c] = is_c ? pi[(k + 1 + KB) * W + c] : 0.0;
		}
	}
#pragma unroll
	for (int j = 0; j < KB; ++j)
#pragma unroll
		for (int c = 0; c < KB; ++c) {
			sr[j][c] = wr[j][c];
			si[j][c] = wi[j][c];
		}
	return would_swap;
}

// S_c = own + partner's (already turned into this half's orientation by the caller) - A_SS, inverted in place
// (Gauss-Jordan, no interchanges) and stored at qr/qi[KB*KB].  A_SS is still in the half's separator rows: the
// multipliers written there sit left of column np.  Returns the would-have-pivoted flag.
template <int N, int KB>
__host__ __device__ __forceinline__ bool tw_separator(const double *pr, const double *pi, bool is_c, int np, double (&sr)[KB][KB],
                                                      double (&si)[KB][KB], const double (&or_)[KB][KB], const double (&oi)[KB][KB],
                                                      double *qr, double *qi)
{
	constexpr int W = Tw<N, KB>::W;
#pragma unroll
	for (int j = 0; j < KB; ++j)
#pragma unroll
		for (int c = 0; c < KB; ++c) {
			sr[j][c] += or_[j][c] - pr[(np + j) * W + KB + c - j];
			si[j][c] += oi[j][c] - (is_c ? pi[(np + j) * W + KB + c - j] : 0.0);
		}
	bool would_swap = false;
#pragma unroll
	for (int k = 0; k < KB; ++k) {
		const double best = fabs(sr[k][k]) + fabs(si[k][k]);
#pragma unroll
		for (int r = k + 1; r < KB; ++r) would_swap |= (fabs(sr[r][k]) + fabs(si[r][k]) > best);
		const double ar = sr[k][k], ai = si[k][k];
		const double rden = tw_rcp(fma(ar, ar, ai * ai));
		const double rr = ar * rden, ri = -ai * rden;
		sr[k][k] = 1.0;
		si[k][k] = 0.0;
#pragma unroll
		for (int c = 0; c < KB; ++c) {
			const double xr = sr[k][c], xi = si[k][c];
			sr[k][c] = fma(xr, rr, -xi * ri);
			si[k][c] = fma(xr, ri, xi * rr);
		}
#pragma unroll
		for (int r = 0; r < KB; ++r) {
			if (r == k) continue;
			const double fr = sr[r][k], fi = si[r][k];
			sr[r][k] = 0.0;
			si[r][k] = 0.0;
#pragma unroll
			for (int c = 0; c < KB; ++c) {
				sr[r][c] = fma(-fr, sr[k][c], fma(fi, si[k][c], sr[r][c]));
				si[r][c] = fma(-fr, si[k][c], fma(-fi, sr[k][c], si[r][c]));
			}
		}
	}
#pragma unroll
	for (int j = 0; j < KB; ++j)
#pragma unroll
		for (int c = 0; c < KB; ++c) {
			qr[j * KB + c] = sr[j][c];
			if (is_c) qi[j * KB + c] = si[j][c];
		}
	return would_swap;
}

// Entry A(i, i+d) of a matrix goes to the top half (row i) and/or the bottom half (row N-1-i, band offset -d).
template <int N, int KB>
__host__ __device__ __forceinline__ void tw_put(double *top, double *bot, int i, int d, double v)
{
	constexpr int W = Tw<N, KB>::W, HR = Tw<N, KB>::HR;
	if (i < HR) top[i * W + KB + d] = v;
	if (N - 1 - i < HR) bot[(N - 1 - i) * W + KB - d] = v;
}

// Forward sweep of one half.  xr/xi point at the half's FIRST unknown, xs = +1 (top) or -1 (bottom) walks
// through its unknowns in the half's order.  Out: the separator right-hand side as reduced by this half.
template <int N, int KB>
__host__ __device__ __forceinline__ void tw_forward(const double *pr, const double *pi, bool is_c, int np, double *xr, double *xi, int xs,
                                                    double (&gr)[KB], double (&gi)[KB])
{
	constexpr int W = Tw<N, KB>::W;
	double wr[KB + 1], wi[KB + 1];
#pragma unroll
	for (int r = 0; r <= KB; ++r) {
		wr[r] = xr[r * xs];
		wi[r] = is_c ? xi[r * xs] : 0.0;
	}
#pragma unroll 2
	for (int k = 0; k < np; ++k) {
#pragma unroll
		for (int r = 1; r <= KB; ++r) {
			const double lr = pr[(k + r) * W + KB - r], li = is_c ? pi[(k + r) * W + KB - r] : 0.0;
			const double a = wr[r], b = wi[r];
			wr[r] = fma(-lr, wr[0], fma(li, wi[0], a));
			wi[r] = fma(-lr, wi[0], fma(-li, wr[0], b));
		}
		xr[k * xs] = wr[0];
		if (is_c) xi[k * xs] = wi[0];
#pragma unroll
		for (int r = 0; r < KB; ++r) {
			wr[r] = wr[r + 1];
			wi[r] = wi[r + 1];
		}
		const bool more = k + 1 < np; // the last step would look at the other half's first unknown: not needed
		wr[KB] = more ? xr[(k + 1 + KB) * xs] : 0.0;
		wi[KB] = (more && is_c) ? xi[(k + 1 + KB) * xs] : 0.0;
	}
#pragma unroll
	for (int j = 0; j < KB; ++j) {
		gr[j] = wr[j];
		gi[j] = wi[j];
	}
}

// Separator unknowns  x_S = S_c^-1 (g_own + g_partner - b_S), everything in this half's orientation; b_S is still
// in x (no sweep writes the separator entries).
template <int N, int KB>
__host__ __device__ __forceinline__ void tw_separator_solve(const double *qr, const double *qi, bool is_c, int np, const double *xr,
                                                            const double *xi, int xs, const double (&gr)[KB], const double (&gi)[KB],
                                                            const double (&hr)[KB], const double (&hi)[KB], double (&sr)[KB],
                                                            double (&si)[KB])
{
	double tr[KB], ti[KB];
#pragma unroll
	for (int j = 0; j < KB; ++j) {
		tr[j] = gr[j] + hr[j] - xr[(np + j) * xs];
		ti[j] = is_c ? gi[j] + hi[j] - xi[(np + j) * xs] : 0.0;
	}
#pragma unroll
	for (int j = 0; j < KB; ++j) {
		double ar = 0.0, ai = 0.0;
#pragma unroll
		for (int c = 0; c < KB; ++c) {
			const double er = qr[j * KB + c], ei = is_c ? qi[j * KB + c] : 0.0;
			ar = fma(er, tr[c], fma(-ei, ti[c], ar));
			ai = fma(er, ti[c], fma(ei, tr[c], ai));
		}
		sr[j] = ar;
		si[j] = ai;
	}
}

// Back substitution of one half, outwards from the separator (sr/si: its solution, own orientation).
template <int N, int KB>
__host__ __device__ __forceinline__ void tw_backward(const double *pr, const double *pi, bool is_c, int np, double *xr, double *xi, int xs,
                                                     const double (&sr)[KB], const double (&si)[KB])
{
	constexpr int W = Tw<N, KB>::W;
	double ur[KB + 1], ui[KB + 1]; // u[c] = x[k + c]
#pragma unroll
	for (int c = 1; c <= KB; ++c) {
		ur[c] = sr[c - 1];
		ui[c] = si[c - 1];
	}
#pragma unroll 2
	for (int k = np - 1; k >= 0; --k) { // the newest unknown enters the chain last
		double ar = xr[k * xs], ai = is_c ? xi[k * xs] : 0.0;
#pragma unroll
		for (int c = KB; c >= 1; --c) {
			const double er = pr[k * W + KB + c], ei = is_c ? pi[k * W + KB + c] : 0.0;
			ar = fma(-er, ur[c], fma(ei, ui[c], ar));
			ai = fma(-er, ui[c], fma(-ei, ur[c], ai));
		}
		const double dr = pr[k * W + KB], di = is_c ? pi[k * W + KB] : 0.0;
		const double kr = fma(ar, dr, -ai * di);
		const double ki = fma(ar, di, ai * dr);
		xr[k * xs] = kr;
		if (is_c) xi[k * xs] = ki;
#pragma unroll
		for (int c = KB; c >= 2; --c) {
			ur[c] = ur[c - 1];
			ui[c] = ui[c - 1];
		}
		ur[1] = kr;
		ui[1] = ki;
	}
}

} // namespace rehuel
