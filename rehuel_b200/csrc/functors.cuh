// Device functors: the reference's functor concept (functor.hpp:35-73: fun(t,y), jac(t,y) with
// J(i,j) = d f_i / d y_j) as templated __device__ callables.  Parameters the reference keeps as
// data members (vdpol::mu, bruss::a/b, exponential::l) or hard-codes (rober: 0.04/1e4/3e7,
// test_equations.hpp:155-157) are per-member parameter vectors p[P] here.
//
// Expression shapes follow test_equations.hpp so that the CPU oracle and the device agree to
// rounding; --fmad contraction is allowed.
#pragma once

namespace rehuel {

// Robertson chemical kinetics; test_equations.hpp:150-176.  p = (k1, k2, k3).
struct RobertsonF {
	static constexpr int N = 3;
	static constexpr int P = 3;
	static constexpr bool kAutonomous = true;
	static constexpr int kFunFlops = 11, kJacFlops = 8; // SURVEY.md 8(d) accounting
	static constexpr unsigned kJacZeroMask = (1u << (2 * 3 + 0)) | (1u << (2 * 3 + 2)); // J(2,0) = J(2,2) = 0
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		f[0] = -p[0] * y[0] + p[1] * y[1] * y[2];
		f[1] = p[0] * y[0] - p[1] * y[1] * y[2] - p[2] * y[1] * y[1];
		f[2] = p[2] * y[1] * y[1];
	}
	__device__ __forceinline__ static void jac(double, const double (&y)[N], const double (&p)[P], double (&J)[N][N])
	{
		J[0][0] = -p[0];
		J[0][1] = p[1] * y[2];
		J[0][2] = p[1] * y[1];
		J[1][0] = p[0];
		J[1][1] = -p[1] * y[2] - (2.0 * p[2]) * y[1];
		J[1][2] = -p[1] * y[1];
		J[2][0] = 0.0;
		J[2][1] = (2.0 * p[2]) * y[1];
		J[2][2] = 0.0;
	}
};

// Van der Pol in the reference's form (mu is the SMALL parameter); test_equations.hpp:91-115.
struct VanDerPolF {
	static constexpr int N = 2;
	static constexpr int P = 1;
	static constexpr bool kAutonomous = true;
	static constexpr int kFunFlops = 6, kJacFlops = 7;
	// The reference divides by mu (test_equations.hpp:101-108); here the member's 1/mu is formed once (loop-invariant:
	// the compiler hoists it out of the Newton iteration) and multiplied with -- the same value to one rounding.  A
	// failing Newton iteration of the stiff members spins for thousands of iterations on a single lane, and the IEEE
	// division was a fifth of its instructions (REHUEL_VDP_DIVIDE=1 restores it).
#ifndef REHUEL_VDP_DIVIDE
#define REHUEL_VDP_DIVIDE 0
#endif
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		f[0] = y[1];
		if (REHUEL_VDP_DIVIDE) f[1] = ((1 - y[0] * y[0]) * y[1] - y[0]) / p[0];
		else f[1] = ((1 - y[0] * y[0]) * y[1] - y[0]) * (1.0 / p[0]);
	}
	__device__ __forceinline__ static void jac(double, const double (&y)[N], const double (&p)[P], double (&J)[N][N])
	{
		J[0][0] = 0.0;
		J[0][1] = 1.0;
		if (REHUEL_VDP_DIVIDE) {
			J[1][0] = -(2.0 * y[0] * y[1] + 1.0) / p[0];
			J[1][1] = (1.0 - y[0] * y[0]) / p[0];
		} else {
			const double rmu = 1.0 / p[0];
			J[1][0] = -(2.0 * y[0] * y[1] + 1.0) * rmu;
			J[1][1] = (1.0 - y[0] * y[0]) * rmu;
		}
	}
};

// y' = l*y; test_equations.hpp:37-59.  p = (l).
struct ExponentialF {
	static constexpr int N = 1;
	static constexpr int P = 1;
	static constexpr bool kAutonomous = true;
	static constexpr int kFunFlops = 1, kJacFlops = 0;
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		f[0] = p[0] * y[0];
	}
	__device__ __forceinline__ static void jac(double, const double (&)[N], const double (&p)[P], double (&J)[N][N])
	{
		J[0][0] = p[0];
	}
};

// y'' + 1001 y' + 1000 y = 0 as a 2x2 linear system; test_equations.hpp:207-250.  No parameters
// (P is 1 only so that arrays have a size; the value is ignored).
struct StiffEqF {
	static constexpr int N = 2;
	static constexpr int P = 0;
	static constexpr bool kAutonomous = true;
	static constexpr int kFunFlops = 3, kJacFlops = 0;
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&)[1], double (&f)[N])
	{
		f[0] = y[1];
		f[1] = -1000.0 * y[0] + -1001.0 * y[1];
	}
	__device__ __forceinline__ static void jac(double, const double (&)[N], const double (&)[1], double (&J)[N][N])
	{
		J[0][0] = 0.0;
		J[0][1] = 1.0;
		J[1][0] = -1000.0;
		J[1][1] = -1001.0;
	}
};

// Two-variable Brusselator; test_equations.hpp:120-145.  p = (a, b).
struct BrusselatorF {
	static constexpr int N = 2;
	static constexpr int P = 2;
	static constexpr bool kAutonomous = true;
	static constexpr int kFunFlops = 9, kJacFlops = 8;
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		f[0] = p[0] + y[0] * y[0] * y[1] - p[1] * y[0] - y[0];
		f[1] = p[1] * y[0] - y[0] * y[0] * y[1];
	}
	__device__ __forceinline__ static void jac(double, const double (&y)[N], const double (&p)[P], double (&J)[N][N])
	{
		J[0][0] = 2 * y[0] * y[1] - p[1] - 1;
		J[0][1] = y[0] * y[0];
		J[1][0] = p[1] - 2 * y[0] * y[1];
		J[1][1] = -y[0] * y[0];
	}
};

// Lorenz system; test_equations.hpp:519-543.  p = (sigma, rho, beta).
struct LorenzF {
	static constexpr int N = 3;
	static constexpr int P = 3;
	static constexpr bool kAutonomous = true;
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		f[0] = p[0] * (y[1] - y[0]);
		f[1] = y[0] * (p[1] - y[2]) - y[1];
		f[2] = y[0] * y[1] - p[2] * y[2];
	}
	__device__ __forceinline__ static void jac(double, const double (&y)[N], const double (&p)[P], double (&J)[N][N])
	{
		J[0][0] = -p[0];
		J[0][1] = p[0];
		J[0][2] = 0.0;
		J[1][0] = p[1] - y[2];
		J[1][1] = -1.0;
		J[1][2] = -y[0];
		J[2][0] = y[1];
		J[2][1] = y[0];
		J[2][2] = -p[2];
	}
};

// Harmonic oscillator; test_equations.hpp:63-84.  p = (w).
struct HarmonicF {
	static constexpr int N = 2;
	static constexpr int P = 1;
	static constexpr bool kAutonomous = true;
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		f[0] = p[0] * y[1];
		f[1] = -p[0] * y[0];
	}
	__device__ __forceinline__ static void jac(double, const double (&)[N], const double (&p)[P], double (&J)[N][N])
	{
		J[0][0] = 0.0;
		J[0][1] = p[0];
		J[1][0] = -p[0];
		J[1][1] = 0.0;
	}
};

// Dimerisation reaction; test_equations.hpp:180-203.  p = (rate); irate = 1/rate like the reference's constructor.
struct DimerF {
	static constexpr int N = 2;
	static constexpr int P = 1;
	static constexpr bool kAutonomous = true;
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		const double rate = p[0], irate = 1.0 / p[0];
		f[0] = -2 * rate * y[0] * y[0] + 2 * irate * y[1];
		f[1] = rate * y[0] * y[0] - irate * y[1];
	}
	__device__ __forceinline__ static void jac(double, const double (&y)[N], const double (&p)[P], double (&J)[N][N])
	{
		const double rate = p[0], irate = 1.0 / p[0];
		J[0][0] = -4 * rate * y[0];
		J[0][1] = 2 * irate;
		J[1][0] = 2 * rate * y[0];
		J[1][1] = -irate;
	}
};

// Four-species reversible polymerisation kinetics; test_equations.hpp:484-517.  p = (b2, b3, b4).
// The one built-in system with n = 4, the largest the thread-per-system kernels take.
struct Kinetic4F {
	static constexpr int N = 4;
	static constexpr int P = 3;
	static constexpr bool kAutonomous = true;
	__device__ __forceinline__ static void fun(double, const double (&y)[N], const double (&p)[P], double (&f)[N])
	{
		const double b2 = p[0], b3 = p[1], b4 = p[2];
		f[0] = (-2 * y[0] * y[0] - y[0] * (y[1] + y[2])) + (2 * b2 * y[1] + b3 * y[2] + b4 * y[3]);
		f[1] = y[0] * y[0] - y[0] * y[1] + b3 * y[2] - b2 * y[1];
		f[2] = y[0] * y[1] - y[0] * y[2] + b4 * y[3] - b3 * y[2];
		f[3] = y[0] * y[2] - b4 * y[3];
	}
	__device__ __forceinline__ static void jac(double, const double (&y)[N], const double (&p)[P], double (&J)[N][N])
	{
		const double b2 = p[0], b3 = p[1], b4 = p[2];
		J[0][0] = -4 * y[0] - y[1] - y[2];
		J[0][1] = -y[0] + 2 * b2;
		J[0][2] = -y[0] + b3;
		J[0][3] = b4;
		J[1][0] = 2 * y[0] - y[1];
		J[1][1] = -y[0] - b2;
		J[1][2] = b3;
		J[1][3] = 0.0;
		J[2][0] = y[1] - y[2];
		J[2][1] = y[0];
		J[2][2] = -y[0] - b3;
		J[2][3] = b4;
		J[3][0] = y[2];
		J[3][1] = 0.0;
		J[3][2] = y[0];
		J[3][3] = -b4;
	}
};

// test_equations.hpp:254-421 (`three_body`): planar three-body problem, y = (x0,y0,x1,y1,x2,y2, p0..p5), p = (m1,m2,m3).
// Reproduced AS WRITTEN in the reference (parity, not repair): dydt[9] uses x3mx2 where y3my2 is meant
// (:296), and jac() returns MINUS a matrix that puts the force derivatives in the momentum-momentum block and 1/m on
// the position diagonal (:329-420, with r13_3 for an m1*m2 term at :344 and r13_2 at :363) -- it is not the derivative
// of fun(), the simplified Newton iteration merely converges linearly with it.  Componentwise (n = 12): dense Jacobian,
// one warp per system with the matrices in registers (irk_wdense.cuh).
struct ThreeBodyF {
	static constexpr int N = 12;
	static constexpr int P = 3;
	static constexpr bool kAutonomous = true;
	__device__ __forceinline__ static double fun(int i, double, const double *y, const double *p)
	{
		const double m1 = p[0], m2 = p[1], m3 = p[2];
		if (i < 6) return y[6 + i] / (i < 2 ? m1 : (i < 4 ? m2 : m3));
		const double x2mx1 = y[2] - y[0], y2my1 = y[3] - y[1];
		const double x3mx1 = y[4] - y[0], y3my1 = y[5] - y[1];
		const double x3mx2 = y[4] - y[2], y3my2 = y[5] - y[3];
		const double r12_2 = x2mx1 * x2mx1 + y2my1 * y2my1;
		const double r13_2 = x3mx1 * x3mx1 + y3my1 * y3my1;
		const double r23_2 = x3mx2 * x3mx2 + y3my2 * y3my2;
		const double r12_3 = r12_2 * sqrt(r12_2), r13_3 = r13_2 * sqrt(r13_2), r23_3 = r23_2 * sqrt(r23_2);
		switch (i) {
		case 6: return m1 * m3 * x3mx1 / r13_3 + m1 * m2 * x2mx1 / r12_3;
		case 7: return m1 * m3 * y3my1 / r13_3 + m1 * m2 * y2my1 / r12_3;
		case 8: return -m1 * m2 * x2mx1 / r12_3 + m2 * m3 * x3mx2 / r23_3;
		case 9: return -m1 * m2 * y2my1 / r12_3 + m2 * m3 * x3mx2 / r23_3; // sic (:296)
		case 10: return -m2 * m3 * x3mx2 / r23_3 - m1 * m3 * x3mx1 / r13_3;
		default: return -m2 * m3 * y3my2 / r23_3 - m1 * m3 * y3my1 / r13_3;
		}
	}
	// entry (a, b), a >= b, of the reference's symmetric 6 x 6 block J(6+a, 6+b) BEFORE the final negation
	__device__ __forceinline__ static double block(int a, int b, const double *y, const double *p)
	{
		const double m1 = p[0], m2 = p[1], m3 = p[2];
		const double x2mx1 = y[2] - y[0], y2my1 = y[3] - y[1];
		const double x3mx1 = y[4] - y[0], y3my1 = y[5] - y[1];
		const double x3mx2 = y[4] - y[2], y3my2 = y[5] - y[3];
		const double x2mx1_2 = x2mx1 * x2mx1, y2my1_2 = y2my1 * y2my1;
		const double x3mx1_2 = x3mx1 * x3mx1, y3my1_2 = y3my1 * y3my1;
		const double x3mx2_2 = x3mx2 * x3mx2, y3my2_2 = y3my2 * y3my2;
		const double r12_2 = x2mx1_2 + y2my1_2, r13_2 = x3mx1_2 + y3my1_2, r23_2 = x3mx2_2 + y3my2_2;
		const double r12_3 = r12_2 * sqrt(r12_2), r13_3 = r13_2 * sqrt(r13_2), r23_3 = r23_2 * sqrt(r23_2);
		const double r12_5 = r12_3 * r12_2, r13_5 = r13_3 * r13_2, r23_5 = r23_3 * r23_2;
		switch (a * 6 + b) {
		case 0 * 6 + 0: return -m1 * m3 / r13_3 + 3 * m1 * m3 * x3mx1_2 / r13_5 - m1 * m2 / r12_3 + 3 * m1 * m2 * x2mx1_2 / r12_5;
		case 1 * 6 + 1: return -m1 * m3 / r13_3 + 3 * m1 * m3 * y3my1_2 / r13_5 - m1 * m2 / r12_3 + 3 * m1 * m2 * y2my1_2 / r12_5;
		case 2 * 6 + 2: return -m2 * m3 / r23_3 + 3 * m2 * m3 * x3mx2_2 / r23_5 - m1 * m2 / r12_3 + 3 * m1 * m2 * x2mx1_2 / r12_5;
		case 3 * 6 + 3: return -m2 * m3 / r23_3 + 3 * m2 * m3 * y3my2_2 / r23_5 - m1 * m2 / r13_3 + 3 * m1 * m2 * y2my1_2 / r12_5; // sic (:344)
		case 4 * 6 + 4: return -m1 * m3 / r13_3 + 3 * m1 * m3 * x3mx1_2 / r13_5 - m2 * m3 / r23_3 + 3 * m2 * m3 * x3mx2_2 / r23_5;
		case 5 * 6 + 5: return -m1 * m3 / r13_3 + 3 * m1 * m3 * y3my1_2 / r13_5 - m2 * m3 / r23_3 + 3 * m2 * m3 * y3my2_2 / r23_5;
		case 1 * 6 + 0: return 3 * m1 * m3 * x3mx1 * y3my1 / r13_5 + 3 * m1 * m2 * x2mx1 * y2my1 / r12_5;
		case 2 * 6 + 0: return m1 * m2 / r12_3 - 3 * m1 * m2 * x2mx1_2 / r12_5;
		case 3 * 6 + 0: return -3 * m1 * m2 * x2mx1 * y2my1 / r12_5;
		case 4 * 6 + 0: return m1 * m3 / r13_3 - 3 * m1 * m3 * x3mx1_2 / r13_5;
		case 5 * 6 + 0: return -3 * m1 * m3 * x3mx1 * y3my1 / r13_5;
		case 2 * 6 + 1: return -3 * m1 * m2 * x2mx1 * y2my1 / r12_5;
		case 3 * 6 + 1: return m1 * m2 / r12_3 - 3 * m1 * m2 * y2my1_2 / r12_5;
		case 4 * 6 + 1: return -3 * m1 * m3 * x3mx1 * y3my1 / r13_5;
		case 5 * 6 + 1: return m1 * m3 / r13_2 - 3 * m1 * m3 * y3my1_2 / r13_5; // sic (:363)
		case 3 * 6 + 2: return 3 * m2 * m3 * x3mx2 * y3my2 / r23_5 + 3 * m1 * m2 * x2mx1 * y2my1 / r12_5;
		case 4 * 6 + 2: return m2 * m3 / r23_3 - 3 * m2 * m3 * x3mx2_2 / r23_5;
		case 5 * 6 + 2: return -3 * m2 * m3 * x3mx2 * y3my2 / r23_5;
		case 4 * 6 + 3: return -3 * m2 * m3 * x3mx2 * y3my2 / r23_5;
		case 5 * 6 + 3: return m2 * m3 / r23_3 - 3 * m2 * m3 * y3my2_2 / r23_5;
		default: return 3 * m2 * m3 * x3mx2 * y3my2 / r23_5 + 3 * m1 * m3 * x3mx1 * y3my1 / r13_5; // (5, 4)
		}
	}
	__device__ __forceinline__ static void jac_row(int i, double, const double *y, const double *p, double *row)
	{
		if (i < 6) {
			row[i] = -(1.0 / (i < 2 ? p[0] : (i < 4 ? p[1] : p[2])));
			return;
		}
		const int a = i - 6;
#pragma unroll
		for (int b = 0; b < 6; ++b) row[6 + b] = -(a >= b ? block(a, b, y, p) : block(b, a, y, p));
	}
};

// 1-D Brusselator, method of lines on NG interior grid points, interleaved (u_i, v_i), boundary
// values u = 1, v = 3 (the reference only has a "TODO: Discretization of a
// PDE", test_equations.hpp:546).  p = (alpha, A, B).  Componentwise functor for the CTA-per-system
// path (irk_cta.cuh); expression shapes are those of the oracle twin (oracle/ref_driver.cpp).
template <int NG>
struct Brusselator1dF {
	static constexpr int N = 2 * NG;
	static constexpr int P = 3;
	static constexpr int kBand = 2; // interleaved (u_i,v_i): couplings reach i +- 2
	__device__ __forceinline__ static double fun(int i, double, const double *y, const double *p)
	{
		const int g = i >> 1;
		const double alpha = p[0], A = p[1], B = p[2];
		const double d = alpha * (NG + 1.0) * (NG + 1.0);
		const double u = y[2 * g], v = y[2 * g + 1];
		if ((i & 1) == 0) {
			const double ul = (g == 0) ? 1.0 : y[2 * (g - 1)];
			const double ur = (g == NG - 1) ? 1.0 : y[2 * (g + 1)];
			return A + u * u * v - (B + 1.0) * u + d * (ul - 2.0 * u + ur);
		}
		const double vl = (g == 0) ? 3.0 : y[2 * (g - 1) + 1];
		const double vr = (g == NG - 1) ? 3.0 : y[2 * (g + 1) + 1];
		return B * u - u * u * v + d * (vl - 2.0 * v + vr);
	}
	__device__ __forceinline__ static void jac_row(int i, double, const double *y, const double *p, double *row)
	{
		const int g = i >> 1;
		const double alpha = p[0], B = p[2];
		const double d = alpha * (NG + 1.0) * (NG + 1.0);
		const double u = y[2 * g], v = y[2 * g + 1];
		if ((i & 1) == 0) {
			row[2 * g] = 2.0 * u * v - (B + 1.0) - 2.0 * d;
			row[2 * g + 1] = u * u;
			if (g > 0) row[2 * (g - 1)] = d;
			if (g < NG - 1) row[2 * (g + 1)] = d;
		} else {
			row[2 * g] = B - 2.0 * u * v;
			row[2 * g + 1] = -u * u - 2.0 * d;
			if (g > 0) row[2 * (g - 1) + 1] = d;
			if (g < NG - 1) row[2 * (g + 1) + 1] = d;
		}
	}
	// Optional for banded functors: row i of J as its band, band[kBand + o] = J(i, i + o) (entries beyond the matrix: 0).
	// Statically indexed, so the warp-per-system kernel keeps it in registers instead of a dense row in local memory.
	// Same expressions as jac_row.
	__device__ __forceinline__ static void jac_band(int i, double, const double *y, const double *p, double (&band)[2 * kBand + 1])
	{
		const int g = i >> 1;
		const double alpha = p[0], B = p[2];
		const double d = alpha * (NG + 1.0) * (NG + 1.0);
		const double u = y[2 * g], v = y[2 * g + 1];
		const bool is_u = (i & 1) == 0;
		band[0] = (g > 0) ? d : 0.0;
		band[1] = is_u ? 0.0 : B - 2.0 * u * v;
		band[2] = is_u ? 2.0 * u * v - (B + 1.0) - 2.0 * d : -u * u - 2.0 * d;
		band[3] = is_u ? u * u : 0.0;
		band[4] = (g < NG - 1) ? d : 0.0;
	}
};

// The same equations with the Jacobian treated as DENSE (no kBand trait): takes the general
// CTA-per-system path with the shared-memory block LU (irk_cta.cuh).
template <int NG>
struct Brusselator1dDenseF {
	static constexpr int N = 2 * NG;
	static constexpr int P = 3;
	__device__ __forceinline__ static double fun(int i, double t, const double *y, const double *p)
	{
		return Brusselator1dF<NG>::fun(i, t, y, p);
	}
	__device__ __forceinline__ static void jac_row(int i, double t, const double *y, const double *p, double *row)
	{
		Brusselator1dF<NG>::jac_row(i, t, y, p, row);
	}
};

} // namespace rehuel
