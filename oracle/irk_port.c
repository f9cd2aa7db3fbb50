/* TEST INFRASTRUCTURE ONLY (oracle).  Never linked into or called from the product library.
 *
 * Plain-C restatement of the reference's implicit-RK hot path, one ensemble member at a time:
 *
 *   irk::get_coefficients      /root/reference/irk.cpp:215-689  -> get_tableau()
 *   irk::construct_R           /root/reference/irk.hpp:366-386  -> construct_R()
 *   irk::newton_solve_stages   /root/reference/irk.hpp:398-493  -> newton_solve_stages()
 *   irk::irk_guts              /root/reference/irk.hpp:509-869  -> irk_guts()
 *   irk::odeint                /root/reference/irk.hpp:884-898  -> odeint()
 *
 * and of the dense kernels the reference takes from Armadillo/LAPACK (third-party, not vendored,
 * version unpinned by the reference): partial-pivoted LU (dgetrf), triangular solves (dtrtrs),
 * general solve (dgesv), inverse (dgetri).
 *
 * Parity status: PINNED.  This port is checked bit for bit against oracle/_ref (the unmodified
 * reference sources compiled against oracle/arma_shim) in tests/test_oracle.py, and against the
 * reference's own known-answer test test/irk.cpp:250-296 (irk_calc_stages).  To make the bitwise
 * comparison meaningful the floating-point operation ORDER below deliberately mirrors the
 * reference expressions (e.g. the stage residual is formed with the dense dt*kron(A,I) matrix,
 * zeros included, exactly like irk.hpp:384); both are compiled with -ffp-contract=off.
 *
 * Exports the same C entry points as oracle/ref_driver.cpp, so oracle/refsolver.py drives either.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_S 7

/* ------------------------------------------------------------------ method ids, enums.hpp:38-58 */
enum {
	IMPLICIT_EULER = 200, RADAU_IIA_32 = 210, RADAU_IIA_53 = 211, RADAU_IIA_95 = 212,
	RADAU_IIA_137 = 213, LOBATTO_IIIA_43 = 220, LOBATTO_IIIA_85 = 221, LOBATTO_IIIA_127 = 222,
	LOBATTO_IIIC_43 = 230, LOBATTO_IIIC_64 = 231, LOBATTO_IIIC_85 = 232, LOBATTO_IIIC_127 = 233,
	GAUSS_LEGENDRE_42 = 240, GAUSS_LEGENDRE_63 = 241, GAUSS_LEGENDRE_105 = 242,
	GAUSS_LEGENDRE_147 = 243
};
static const struct { int id; const char *name; } METHOD_NAMES[] = {
	{200, "IMPLICIT_EULER"}, {210, "RADAU_IIA_32"}, {211, "RADAU_IIA_53"}, {212, "RADAU_IIA_95"},
	{213, "RADAU_IIA_137"}, {220, "LOBATTO_IIIA_43"}, {221, "LOBATTO_IIIA_85"},
	{222, "LOBATTO_IIIA_127"}, {230, "LOBATTO_IIIC_43"}, {231, "LOBATTO_IIIC_64"},
	{232, "LOBATTO_IIIC_85"}, {233, "LOBATTO_IIIC_127"}, {240, "GAUSS_LEGENDRE_42"},
	{241, "GAUSS_LEGENDRE_63"}, {242, "GAUSS_LEGENDRE_105"}, {243, "GAUSS_LEGENDRE_147"},
};
/* odeint_status_codes, enums.hpp:103-127; newton ret codes, newton.hpp:45-51 */
enum { ST_SUCCESS = 0, ST_GENERAL_ERROR = 4, ST_ERROR_MAX_STEPS_EXCEEDED = 128 };
enum { NEWTON_SUCCESS = 0, NEWTON_MAXIT_EXCEEDED = 3 };
static const double MACHINE_PRECISION = 1e-17; /* enums.hpp:129 */

/* ------------------------------------------------------------------ tableaux */
typedef struct {
	int s, has_b2, order, order2;
	double A[MAX_S * MAX_S]; /* row-major s x s */
	double b[MAX_S], c[MAX_S], b2[MAX_S], gamma;
} tableau;

#define SETA(i, j, v) (tb->A[(i) * tb->s + (j)] = (v))

#include "tableau_dec.inc" /* RADAU_IIA_95 / RADAU_IIA_137 as the reference prints them (generated) */

/* irk.cpp:393-437, 439-533: A, b, c by decimals;  b2(i) = b(i) -+ |coef| * gamma */
static void set_decimal(tableau *tb, int s, const double *A, const double *c, const double *b,
                        const double *b2coef, double gamma, int order, int order2)
{
	tb->s = s; tb->has_b2 = 1; tb->gamma = gamma; tb->order = order; tb->order2 = order2;
	for (int i = 0; i < s; ++i) {
		for (int j = 0; j < s; ++j) tb->A[i * s + j] = A[i * s + j];
		tb->c[i] = c[i];
		tb->b[i] = b[i];
		double cg = b2coef[i] * gamma;
		tb->b2[i] = b[i] + cg;
	}
}

/* irk.cpp:215-689.  Only methods whose A is invertible (irk.hpp:599 inverts A) and that carry
 * an embedded pair are restated; unknown ids return 0 like verify_solver_coeffs failing. */
static int get_tableau(int method, tableau *tb)
{
	const double sqrt5 = sqrt(5.0), sqrt6 = sqrt(6.0), sqrt21 = sqrt(21.0);
	memset(tb, 0, sizeof *tb);
	tb->gamma = 0.0; /* irk.cpp:231: stays 0 unless the case sets it */
	switch (method) {
	case IMPLICIT_EULER: /* irk.cpp:238-244: no b2 => odeint switches adaptivity off (irk.hpp:890-895) */
		tb->s = 1; tb->has_b2 = 0;
		SETA(0,0, 1.0); tb->b[0] = 1.0; tb->c[0] = 1.0;
		tb->order = 1; tb->order2 = 0;
		return 1;
	case RADAU_IIA_95: /* irk.cpp:393-437 */
		set_decimal(tb, 5, radau_iia_95_A, radau_iia_95_c, radau_iia_95_b, radau_iia_95_b2_gamma_coef,
		            radau_iia_95_gamma, radau_iia_95_order, radau_iia_95_order2);
		return 1;
	case RADAU_IIA_137: /* irk.cpp:439-533 */
		set_decimal(tb, 7, radau_iia_137_A, radau_iia_137_c, radau_iia_137_b, radau_iia_137_b2_gamma_coef,
		            radau_iia_137_gamma, radau_iia_137_order, radau_iia_137_order2);
		return 1;
	case RADAU_IIA_32: /* irk.cpp:342-360 */
		tb->s = 2; tb->has_b2 = 1;
		SETA(0,0, 5.0/12.0); SETA(0,1, -1.0/12.0);
		SETA(1,0, 3.0/4.0);  SETA(1,1, 1.0/4.0);
		tb->gamma = 1.0/3.0;
		tb->c[0] = 1.0/3.0; tb->c[1] = 1.0;
		tb->b[0] = 3.0/4.0; tb->b[1] = 1.0/4.0;
		tb->b2[0] = 2.0/3.0; tb->b2[1] = 1.0/3.0;
		tb->order = 3; tb->order2 = 2;
		return 1;
	case RADAU_IIA_53: /* irk.cpp:363-391 */
		tb->s = 3; tb->has_b2 = 1;
		SETA(0,0, (88 - 7*sqrt6)/360.0); SETA(0,1, (296 - 169*sqrt6)/1800.0); SETA(0,2, (-2+3*sqrt6)/225.0);
		SETA(1,0, (296 + 169*sqrt6)/1800.0); SETA(1,1, (88 + 7*sqrt6)/360.0); SETA(1,2, (-2-3*sqrt6)/225.0);
		SETA(2,0, (16.0 - sqrt6)/36.0); SETA(2,1, (16 + sqrt6)/36.0); SETA(2,2, 1.0 / 9.0);
		tb->gamma = 2.74888829595677e-01;
		tb->c[0] = (4.0-sqrt6)/10.0; tb->c[1] = (4.0+sqrt6) / 10.0; tb->c[2] = 1.0;
		tb->b[0] = (16 - sqrt6)/36.0; tb->b[1] = (16 + sqrt6)/36.0; tb->b[2] = 1.0 / 9.0;
		tb->b2[0] = -((18*sqrt6 + 12)*tb->gamma - 16 + sqrt6)/36.0;
		tb->b2[1] = ((18*sqrt6 - 12)*tb->gamma + 16 + sqrt6)/36.0;
		tb->b2[2] = -(3*tb->gamma - 1) / 9.0;
		tb->order = 5; tb->order2 = 3;
		return 1;
	case LOBATTO_IIIC_43: /* irk.cpp:261-272; gamma stays 0 */
		tb->s = 3; tb->has_b2 = 1;
		SETA(0,0, 1.0/6.0); SETA(0,1, -1.0/3.0);  SETA(0,2, 1.0/6.0);
		SETA(1,0, 1.0/6.0); SETA(1,1, 5.0/12.0);  SETA(1,2, -1.0/12.0);
		SETA(2,0, 1.0/6.0); SETA(2,1, 2.0/3.0);   SETA(2,2, 1.0/6.0);
		tb->b[0] = 1.0/6.0; tb->b[1] = 2.0/3.0; tb->b[2] = 1.0/6.0;
		tb->b2[0] = 1.0/6.0; tb->b2[1] = 1.0/6.0; tb->b2[2] = 2.0/3.0;
		tb->c[0] = 0.0; tb->c[1] = 0.5; tb->c[2] = 1.0;
		tb->order = 4; tb->order2 = 2;
		return 1;
	case LOBATTO_IIIC_64: /* irk.cpp:274-286; gamma stays 0 */
		tb->s = 4; tb->has_b2 = 1;
		SETA(0,0, 1.0/12.0); SETA(0,1, -sqrt5/12.); SETA(0,2, sqrt5/12.); SETA(0,3, -1.0/12.0);
		SETA(1,0, 1.0/12.0); SETA(1,1, 0.25); SETA(1,2, (10 - 7*sqrt5)/60.); SETA(1,3, sqrt5/60.);
		SETA(2,0, 1.0/12.0); SETA(2,1, (10 + 7*sqrt5)/60.); SETA(2,2, 0.25); SETA(2,3, -sqrt5/60.);
		SETA(3,0, 1.0/12.0); SETA(3,1, 5.0/12.0); SETA(3,2, 5.0/12.0); SETA(3,3, 1.0/12.0);
		tb->b[0] = 1.0/12.0; tb->b[1] = 5.0/12.0; tb->b[2] = 5.0/12.0; tb->b[3] = 1.0/12.0;
		tb->b2[0] = 1.0/12.0; tb->b2[1] = 1.0/12.0; tb->b2[2] = 5.0/12.0; tb->b2[3] = 5.0/12.0;
		tb->c[0] = 0.0; tb->c[1] = 0.5 - sqrt5/10.; tb->c[2] = 0.5 + sqrt5/10.; tb->c[3] = 1.0;
		tb->order = 6; tb->order2 = 4;
		return 1;
	case LOBATTO_IIIC_85: /* irk.cpp:288-327 */
		tb->s = 5; tb->has_b2 = 1;
		SETA(0,0, 1.0/20.0); SETA(0,1, -7.0/60.0); SETA(0,2, 2.0/15.0); SETA(0,3, -7.0/60.0); SETA(0,4, 1.0/20.0);
		SETA(1,0, 1.0/20.0); SETA(1,1, 29.0/180.0); SETA(1,2, (47.0 - 15*sqrt21)/315.0);
		SETA(1,3, (203.0 - 30*sqrt21)/1260.0); SETA(1,4, -3.0/140.0);
		SETA(2,0, 1.0/20.0); SETA(2,1, (329.0 + 105.0*sqrt21)/2880.0); SETA(2,2, 73.0/360.0);
		SETA(2,3, (329.0 - 105.0*sqrt21)/2880.0); SETA(2,4, 3.0/160.0);
		SETA(3,0, 1.0/20.0); SETA(3,1, (203.0 + 30*sqrt21)/1260.0); SETA(3,2, (47+15*sqrt21)/315.0);
		SETA(3,3, 29/180.0); SETA(3,4, -3/140.0);
		SETA(4,0, 1.0/20.0); SETA(4,1, 49/180.0); SETA(4,2, 16/45.0); SETA(4,3, 49/180.0); SETA(4,4, 1.0/20.0);
		tb->c[0] = 0.0; tb->c[1] = 0.5 - sqrt21/14.0; tb->c[2] = 0.5; tb->c[3] = 0.5 + sqrt21/14.0; tb->c[4] = 1.0;
		tb->b[0] = 0.05; tb->b[1] = 49/180.0; tb->b[2] = 16/45.0; tb->b[3] = 49/180.0; tb->b[4] = 0.05;
		tb->b2[0] = 0.05; tb->b2[1] = 0.05; tb->b2[2] = 49/180.0; tb->b2[3] = 16/45.0; tb->b2[4] = 49/180.0;
		tb->gamma = 0.1894964436645163;
		tb->order = 8; tb->order2 = 5;
		return 1;
	default:
		return 0;
	}
}

/* ------------------------------------------------------------------ dense kernels (column-major) */
#define AT(a, n, i, j) ((a)[(size_t)(i) + (size_t)(j) * (size_t)(n)])

/* Unblocked right-looking LU, row partial pivoting, first maximal |a| wins (dgetf2). */
static int getrf(double *a, int n, int *piv)
{
	int ok = 1;
	for (int k = 0; k < n; ++k) {
		int p = k;
		double best = fabs(AT(a, n, k, k));
		for (int i = k + 1; i < n; ++i) {
			double v = fabs(AT(a, n, i, k));
			if (v > best) { best = v; p = i; }
		}
		piv[k] = p;
		if (AT(a, n, p, k) == 0.0) { ok = 0; continue; }
		if (p != k)
			for (int j = 0; j < n; ++j) {
				double tmp = AT(a, n, k, j); AT(a, n, k, j) = AT(a, n, p, j); AT(a, n, p, j) = tmp;
			}
		const double rp = 1.0 / AT(a, n, k, k);
		for (int i = k + 1; i < n; ++i) AT(a, n, i, k) *= rp;
		for (int j = k + 1; j < n; ++j) {
			const double akj = AT(a, n, k, j);
			for (int i = k + 1; i < n; ++i) AT(a, n, i, j) -= AT(a, n, i, k) * akj;
		}
	}
	return ok;
}

/* x <- (unit-lower L)^-1 x, x <- U^-1 x on the packed LU. */
static void lu_forward_unit(const double *lu, int n, double *x)
{
	for (int i = 0; i < n; ++i) {
		double acc = x[i];
		for (int k = 0; k < i; ++k) acc -= AT(lu, n, i, k) * x[k];
		x[i] = acc;
	}
}
static void lu_backward(const double *lu, int n, double *x)
{
	for (int i = n - 1; i >= 0; --i) {
		double acc = x[i];
		for (int k = i + 1; k < n; ++k) acc -= AT(lu, n, i, k) * x[k];
		x[i] = acc / AT(lu, n, i, i);
	}
}

/* arma::solve(A,b): LU of a copy, apply the row interchanges, two substitutions. */
static int dense_solve(const double *A, int n, double *x /* in: b, out: x */, double *work, int *piv)
{
	memcpy(work, A, sizeof(double) * (size_t)n * (size_t)n);
	if (!getrf(work, n, piv)) return 0;
	for (int k = 0; k < n; ++k)
		if (piv[k] != k) { double tmp = x[k]; x[k] = x[piv[k]]; x[piv[k]] = tmp; }
	lu_forward_unit(work, n, x);
	lu_backward(work, n, x);
	return 1;
}

/* d = (A^-1)^T w, with A^-1 = solve(A, I) column by column (irk.hpp:599-601). A row-major in tb. */
static void inv_transpose_times(const tableau *tb, const double *w, double *d)
{
	int s = tb->s, piv[MAX_S];
	double Acm[MAX_S * MAX_S], work[MAX_S * MAX_S], Ai[MAX_S * MAX_S], col[MAX_S];
	for (int i = 0; i < s; ++i)
		for (int j = 0; j < s; ++j) AT(Acm, s, i, j) = tb->A[i * s + j];
	for (int j = 0; j < s; ++j) {
		for (int i = 0; i < s; ++i) col[i] = (i == j) ? 1.0 : 0.0;
		dense_solve(Acm, s, col, work, piv);
		for (int i = 0; i < s; ++i) AT(Ai, s, i, j) = col[i];
	}
	/* (Ai^T * w)_i = sum_k Ai(k,i) w_k, ascending k from 0.0 */
	for (int i = 0; i < s; ++i) {
		double acc = 0.0;
		for (int k = 0; k < s; ++k) acc += AT(Ai, s, k, i) * w[k];
		d[i] = acc;
	}
}

/* ------------------------------------------------------------------ problems (functor concept) */
enum { REF_ROBERTSON = 1, REF_VDPOL = 2, REF_BRUSS1D = 3, REF_ROBER_HARDCODED = 4,
       REF_EXPONENTIAL = 5, REF_STIFF_EQ = 6, REF_BRUSS2 = 7, REF_LORENZ = 8, REF_HARMONIC = 9, REF_DIMER = 10,
       REF_KINETIC4 = 11, REF_THREE_BODY = 12 };

typedef struct {
	int id, n;
	double p[4];
} problem;

/* f(t,y) -> out[n].  Expression shapes follow test_equations.hpp. */
static void fun(const problem *pr, double t, const double *y, double *f)
{
	(void)t;
	switch (pr->id) {
	case REF_ROBER_HARDCODED: /* test_equations.hpp:153-158 */
		f[0] = -0.04*y[0] + 1e4 * y[1]*y[2];
		f[1] = 0.04*y[0] - 1e4 * y[1]*y[2] - 3e7*y[1]*y[1];
		f[2] = 3e7*y[1]*y[1];
		break;
	case REF_ROBERTSON: {
		const double k1 = pr->p[0], k2 = pr->p[1], k3 = pr->p[2];
		f[0] = -k1*y[0] + k2 * y[1]*y[2];
		f[1] = k1*y[0] - k2 * y[1]*y[2] - k3*y[1]*y[1];
		f[2] = k3*y[1]*y[1];
		break;
	}
	case REF_VDPOL: /* test_equations.hpp:97-100 */
		f[0] = y[1];
		f[1] = ((1 - y[0]*y[0]) * y[1] - y[0]) / pr->p[0];
		break;
	case REF_EXPONENTIAL: /* test_equations.hpp:48-51 */
		f[0] = pr->p[0]*y[0];
		break;
	case REF_STIFF_EQ: { /* test_equations.hpp:237-240: A*y, dense product from 0.0 */
		double a0 = 0.0, a1 = 0.0;
		a0 += 0.0 * y[0]; a0 += 1.0 * y[1];
		a1 += -1000.0 * y[0]; a1 += -1001.0 * y[1];
		f[0] = a0; f[1] = a1;
		break;
	}
	case REF_BRUSS2: { /* test_equations.hpp:126-130 */
		const double a = pr->p[0], b = pr->p[1];
		f[0] = a + y[0]*y[0]*y[1] - b*y[0] - y[0];
		f[1] = b*y[0] - y[0]*y[0]*y[1];
		break;
	}
	case REF_LORENZ: { /* test_equations.hpp:526-532 */
		const double s = pr->p[0], r = pr->p[1], b = pr->p[2];
		f[0] = s*(y[1] - y[0]);
		f[1] = y[0]*(r - y[2]) - y[1];
		f[2] = y[0]*y[1] - b*y[2];
		break;
	}
	case REF_HARMONIC: /* test_equations.hpp:73-76 */
		f[0] = pr->p[0]*y[1];
		f[1] = -pr->p[0]*y[0];
		break;
	case REF_DIMER: { /* test_equations.hpp:185-189; irate = 1.0/rate as in the constructor */
		const double rate = pr->p[0], irate = 1.0/rate;
		f[0] = -2*rate*y[0]*y[0] + 2*irate*y[1];
		f[1] = rate*y[0]*y[0] - irate*y[1];
		break;
	}
	case REF_KINETIC4: { /* test_equations.hpp:492-502 */
		const double b2 = pr->p[0], b3 = pr->p[1], b4 = pr->p[2];
		double r0 = -2*y[0]*y[0] - y[0]*(y[1] + y[2]);
		r0 += 2*b2*y[1] + b3*y[2] + b4*y[3];
		f[0] = r0;
		f[1] = y[0]*y[0] - y[0]*y[1] + b3*y[2] - b2*y[1];
		f[2] = y[0]*y[1] - y[0]*y[2] + b4*y[3] - b3*y[2];
		f[3] = y[0]*y[2] - b4*y[3];
		break;
	}
	case REF_THREE_BODY: { /* test_equations.hpp:263-302, as written (dydt[9] uses x3mx2) */
		const double m1 = pr->p[0], m2 = pr->p[1], m3 = pr->p[2];
		f[0] = y[6] / m1; f[1] = y[7] / m1;
		f[2] = y[8] / m2; f[3] = y[9] / m2;
		f[4] = y[10] / m3; f[5] = y[11] / m3;
		double x2mx1 = y[2] - y[0], y2my1 = y[3] - y[1];
		double x3mx1 = y[4] - y[0], y3my1 = y[5] - y[1];
		double x3mx2 = y[4] - y[2], y3my2 = y[5] - y[3];
		double r12_2 = x2mx1*x2mx1 + y2my1*y2my1;
		double r13_2 = x3mx1*x3mx1 + y3my1*y3my1;
		double r23_2 = x3mx2*x3mx2 + y3my2*y3my2;
		double r12_3 = r12_2*sqrt(r12_2), r13_3 = r13_2*sqrt(r13_2), r23_3 = r23_2*sqrt(r23_2);
		f[6]  =  m1*m3*x3mx1 / r13_3 + m1*m2*x2mx1 / r12_3;
		f[7]  =  m1*m3*y3my1 / r13_3 + m1*m2*y2my1 / r12_3;
		f[8]  = -m1*m2*x2mx1 / r12_3 + m2*m3*x3mx2 / r23_3;
		f[9]  = -m1*m2*y2my1 / r12_3 + m2*m3*x3mx2 / r23_3;
		f[10] = -m2*m3*x3mx2 / r23_3 - m1*m3*x3mx1 / r13_3;
		f[11] = -m2*m3*y3my2 / r23_3 - m1*m3*y3my1 / r13_3;
		break;
	}
	case REF_BRUSS1D: { /* ours; twin of bruss1d in ref_driver.cpp */
		const int Ng = pr->n / 2;
		const double alpha = pr->p[0], A = pr->p[1], B = pr->p[2];
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
		break;
	}
	}
}

/* J(i,j) = d f_i / d y_j, column-major n x n. */
static void jac(const problem *pr, double t, const double *y, double *J)
{
	const int n = pr->n;
	(void)t;
	memset(J, 0, sizeof(double) * (size_t)n * (size_t)n);
	switch (pr->id) {
	case REF_ROBER_HARDCODED: /* test_equations.hpp:161-175 */
		AT(J,n,0,0) = -0.04; AT(J,n,0,1) = 1e4*y[2]; AT(J,n,0,2) = 1e4*y[1];
		AT(J,n,1,0) = 0.04; AT(J,n,1,1) = -1e4*y[2] - 6e7*y[1]; AT(J,n,1,2) = -1e4*y[1];
		AT(J,n,2,0) = 0.0; AT(J,n,2,2) = 0.0; AT(J,n,2,1) = 6e7*y[1];
		break;
	case REF_ROBERTSON: {
		const double k1 = pr->p[0], k2 = pr->p[1], k3 = pr->p[2];
		AT(J,n,0,0) = -k1; AT(J,n,0,1) = k2*y[2]; AT(J,n,0,2) = k2*y[1];
		AT(J,n,1,0) = k1; AT(J,n,1,1) = -k2*y[2] - (2.0*k3)*y[1]; AT(J,n,1,2) = -k2*y[1];
		AT(J,n,2,0) = 0.0; AT(J,n,2,2) = 0.0; AT(J,n,2,1) = (2.0*k3)*y[1];
		break;
	}
	case REF_VDPOL: { /* test_equations.hpp:102-112 */
		const double mu = pr->p[0];
		AT(J,n,0,0) = 0.0; AT(J,n,0,1) = 1.0;
		AT(J,n,1,0) = -(2.0 * y[0] * y[1] + 1.0) / mu;
		AT(J,n,1,1) = ( 1.0 - y[0]*y[0] ) / mu;
		break;
	}
	case REF_EXPONENTIAL:
		J[0] = pr->p[0];
		break;
	case REF_STIFF_EQ:
		AT(J,n,0,0) = 0.0; AT(J,n,0,1) = 1.0; AT(J,n,1,0) = -1000.0; AT(J,n,1,1) = -1001.0;
		break;
	case REF_BRUSS2: { /* test_equations.hpp:132-142 */
		const double b = pr->p[1];
		AT(J,n,0,0) = 2*y[0]*y[1] - b - 1; AT(J,n,0,1) = y[0]*y[0];
		AT(J,n,1,0) = b - 2*y[0]*y[1]; AT(J,n,1,1) = -y[0]*y[0];
		break;
	}
	case REF_LORENZ: { /* test_equations.hpp:534-540 */
		const double s = pr->p[0], r = pr->p[1], b = pr->p[2];
		AT(J,n,0,0) = -s; AT(J,n,0,1) = s; AT(J,n,0,2) = 0;
		AT(J,n,1,0) = r - y[2]; AT(J,n,1,1) = -1.0; AT(J,n,1,2) = -y[0];
		AT(J,n,2,0) = y[1]; AT(J,n,2,1) = y[0]; AT(J,n,2,2) = -b;
		break;
	}
	case REF_HARMONIC: /* test_equations.hpp:78-81 */
		AT(J,n,0,0) = 0; AT(J,n,0,1) = pr->p[0]; AT(J,n,1,0) = -pr->p[0]; AT(J,n,1,1) = 0;
		break;
	case REF_DIMER: { /* test_equations.hpp:191-200 */
		const double rate = pr->p[0], irate = 1.0/rate;
		AT(J,n,0,0) = -4*rate*y[0]; AT(J,n,0,1) = 2*irate;
		AT(J,n,1,0) = 2*rate*y[0]; AT(J,n,1,1) = -irate;
		break;
	}
	case REF_KINETIC4: { /* test_equations.hpp:508-512 */
		const double b2 = pr->p[0], b3 = pr->p[1], b4 = pr->p[2];
		AT(J,n,0,0) = -4*y[0] - y[1] - y[2]; AT(J,n,0,1) = -y[0] + 2*b2; AT(J,n,0,2) = -y[0] + b3; AT(J,n,0,3) = b4;
		AT(J,n,1,0) = 2*y[0] - y[1]; AT(J,n,1,1) = -y[0] - b2; AT(J,n,1,2) = b3; AT(J,n,1,3) = 0.0;
		AT(J,n,2,0) = y[1] - y[2]; AT(J,n,2,1) = y[0]; AT(J,n,2,2) = -y[0] - b3; AT(J,n,2,3) = b4;
		AT(J,n,3,0) = y[2]; AT(J,n,3,1) = 0; AT(J,n,3,2) = y[0]; AT(J,n,3,3) = -b4;
		break;
	}
	case REF_THREE_BODY: { /* test_equations.hpp:304-421, as written: the function returns -J of this matrix */
		const double m1 = pr->p[0], m2 = pr->p[1], m3 = pr->p[2];
		double x2mx1 = y[2] - y[0], y2my1 = y[3] - y[1];
		double x3mx1 = y[4] - y[0], y3my1 = y[5] - y[1];
		double x3mx2 = y[4] - y[2], y3my2 = y[5] - y[3];
		double x2mx1_2 = x2mx1*x2mx1, y2my1_2 = y2my1*y2my1;
		double x3mx1_2 = x3mx1*x3mx1, y3my1_2 = y3my1*y3my1;
		double x3mx2_2 = x3mx2*x3mx2, y3my2_2 = y3my2*y3my2;
		double r12_2 = x2mx1_2 + y2my1_2, r13_2 = x3mx1_2 + y3my1_2, r23_2 = x3mx2_2 + y3my2_2;
		double r12_3 = r12_2*sqrt(r12_2), r13_3 = r13_2*sqrt(r13_2), r23_3 = r23_2*sqrt(r23_2);
		double r12_5 = r12_3*r12_2, r13_5 = r13_3*r13_2, r23_5 = r23_3*r23_2;
		AT(J,n, 6, 6) = -m1*m3 / r13_3 + 3*m1*m3*x3mx1_2 / r13_5 - m1*m2 / r12_3 + 3*m1*m2*x2mx1_2 / r12_5;
		AT(J,n, 8, 8) = -m2*m3 / r23_3 + 3*m2*m3*x3mx2_2 / r23_5 - m1*m2 / r12_3 + 3*m1*m2*x2mx1_2 / r12_5;
		AT(J,n,10,10) = -m1*m3 / r13_3 + 3*m1*m3*x3mx1_2 / r13_5 - m2*m3 / r23_3 + 3*m2*m3*x3mx2_2 / r23_5;
		AT(J,n, 7, 7) = -m1*m3 / r13_3 + 3*m1*m3*y3my1_2 / r13_5 - m1*m2 / r12_3 + 3*m1*m2*y2my1_2 / r12_5;
		AT(J,n, 9, 9) = -m2*m3 / r23_3 + 3*m2*m3*y3my2_2 / r23_5 - m1*m2 / r13_3 + 3*m1*m2*y2my1_2 / r12_5;
		AT(J,n,11,11) = -m1*m3 / r13_3 + 3*m1*m3*y3my1_2 / r13_5 - m2*m3 / r23_3 + 3*m2*m3*y3my2_2 / r23_5;
		AT(J,n, 7, 6) = 3*m1*m3*x3mx1*y3my1 / r13_5 +3*m1*m2*x2mx1*y2my1 / r12_5;
		AT(J,n, 8, 6) = m1*m2/r12_3 - 3*m1*m2*x2mx1_2 / r12_5;
		AT(J,n, 9, 6) = -3*m1*m2*x2mx1*y2my1 / r12_5;
		AT(J,n,10, 6) = m1*m3/r13_3 - 3*m1*m3*x3mx1_2 / r13_5;
		AT(J,n,11, 6) = -3*m1*m3*x3mx1*y3my1 / r13_5;
		AT(J,n, 8, 7) = -3*m1*m2*x2mx1*y2my1 / r12_5;
		AT(J,n, 9, 7) = m1*m2/r12_3 - 3*m1*m2*y2my1_2 / r12_5;
		AT(J,n,10, 7) = -3*m1*m3*x3mx1*y3my1 / r13_5;
		AT(J,n,11, 7) = m1*m3/r13_2 - 3*m1*m3*y3my1_2 / r13_5;
		AT(J,n, 9, 8) = 3*m2*m3*x3mx2*y3my2 / r23_5 + 3*m1*m2*x2mx1*y2my1 / r12_5;
		AT(J,n,10, 8) = m2*m3 / r23_3 - 3*m2*m3*x3mx2_2 / r23_5;
		AT(J,n,11, 8) = -3*m2*m3*x3mx2*y3my2 / r23_5;
		AT(J,n,10, 9) = -3*m2*m3*x3mx2*y3my2 / r23_5;
		AT(J,n,11, 9) = m2*m3/r23_3 - 3*m2*m3*y3my2_2 / r23_5;
		AT(J,n,11,10) = 3*m2*m3*x3mx2*y3my2 / r23_5 + 3*m1*m3*x3mx1*y3my1 / r13_5;
		for (int a = 6; a < 12; ++a)
			for (int b = a + 1; b < 12; ++b) AT(J,n,a,b) = AT(J,n,b,a);
		AT(J,n,0,0) = 1.0 / m1; AT(J,n,1,1) = 1.0 / m1;
		AT(J,n,2,2) = 1.0 / m2; AT(J,n,3,3) = 1.0 / m2;
		AT(J,n,4,4) = 1.0 / m3; AT(J,n,5,5) = 1.0 / m3;
		for (int k = 0; k < n*n; ++k) J[k] = -J[k];
		break;
	}
	case REF_BRUSS1D: {
		const int Ng = n / 2;
		const double alpha = pr->p[0], B = pr->p[2];
		const double d = alpha * (Ng + 1.0) * (Ng + 1.0);
		for (int i = 0; i < Ng; ++i) {
			double u = y[2*i], v = y[2*i+1];
			AT(J,n,2*i,2*i)     = 2.0*u*v - (B + 1.0) - 2.0*d;
			AT(J,n,2*i,2*i+1)   = u*u;
			AT(J,n,2*i+1,2*i)   = B - 2.0*u*v;
			AT(J,n,2*i+1,2*i+1) = -u*u - 2.0*d;
			if (i > 0)      { AT(J,n,2*i,2*(i-1)) = d; AT(J,n,2*i+1,2*(i-1)+1) = d; }
			if (i < Ng - 1) { AT(J,n,2*i,2*(i+1)) = d; AT(J,n,2*i+1,2*(i+1)+1) = d; }
		}
		break;
	}
	}
}

int ref_problem_dims(int id, int n_hint, int *n, int *p)
{
	switch (id) {
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

/* ------------------------------------------------------------------ per-member workspace */
typedef struct {
	int n, s, N;
	double *Y, *R, *dY, *F, *tmp, *ytmp; /* N, N, N, N, N, n */
	double *J, *JY, *LU, *K;             /* n*n, N*N, N*N, N*N */
	int *piv;                            /* N: LAPACK-style interchange list */
	double *E, *Ework, *ev;              /* n*n, n*n, n */
	int *epiv;
} workspace;

static int ws_alloc(workspace *w, int n, int s)
{
	size_t N = (size_t)n * (size_t)s;
	memset(w, 0, sizeof *w);
	w->n = n; w->s = s; w->N = (int)N;
	w->Y = calloc(N, 8); w->R = calloc(N, 8); w->dY = calloc(N, 8); w->F = calloc(N, 8);
	w->tmp = calloc(N, 8); w->ytmp = calloc((size_t)n, 8);
	w->J = calloc((size_t)n * n, 8); w->JY = calloc(N * N, 8); w->LU = calloc(N * N, 8);
	w->K = calloc(N * N, 8);
	w->piv = calloc(N, sizeof(int));
	w->E = calloc((size_t)n * n, 8); w->Ework = calloc((size_t)n * n, 8); w->ev = calloc((size_t)n, 8);
	w->epiv = calloc((size_t)n, sizeof(int));
	return w->Y && w->R && w->dY && w->F && w->tmp && w->ytmp && w->J && w->JY && w->LU && w->K &&
	       w->piv && w->E && w->Ework && w->ev && w->epiv;
}
static void ws_free(workspace *w)
{
	free(w->Y); free(w->R); free(w->dY); free(w->F); free(w->tmp); free(w->ytmp); free(w->J);
	free(w->JY); free(w->LU); free(w->K); free(w->piv); free(w->E); free(w->Ework); free(w->ev);
	free(w->epiv);
}

typedef struct {
	unsigned long long attempt, reject_newton, reject_err, newton_success, newton_incr_diverge,
	    newton_iter_error_too_large, newton_maxit_exceed, fun_evals, jac_evals;
} counters;

/* irk.hpp:366-386.  R = Y - (dt*kron(A,I_n)) * F with the dense N x N product, ascending column
 * index, accumulator started at 0.0 -- the order the Armadillo stand-in uses. */
static void construct_R(const problem *pr, const tableau *tb, workspace *w, const double *y,
                        double t, double dt)
{
	const int n = w->n, s = w->s, N = w->N;
	for (int i = 0; i < s; ++i) {
		for (int k = 0; k < n; ++k) w->ytmp[k] = y[k] + w->Y[i*n + k];
		fun(pr, t + tb->c[i]*dt, w->ytmp, w->F + (size_t)i*n);
	}
	/* K = dt * kron(A, I): entry ((i,k),(j,l)) = dt * (A_ij * delta_kl) */
	for (int j = 0; j < s; ++j)
		for (int i = 0; i < s; ++i)
			for (int l = 0; l < n; ++l)
				for (int k = 0; k < n; ++k)
					AT(w->K, N, i*n + k, j*n + l) = dt * (tb->A[i*s + j] * (k == l ? 1.0 : 0.0));
	for (int r = 0; r < N; ++r) {
		double acc = 0.0;
		for (int q = 0; q < N; ++q) acc += AT(w->K, N, r, q) * w->F[q];
		w->R[r] = w->Y[r] - acc;
	}
}

static double dot(const double *a, const double *b, int n)
{
	double acc = 0.0;
	for (int i = 0; i < n; ++i) acc += a[i] * b[i];
	return acc;
}

/* J = jac(t,y); J_Y = I_N - dt*kron(A,J); P*J_Y = L*U   (irk.hpp:422-436) */
static void refresh_jacobi_matrix(const problem *pr, const tableau *tb, workspace *w,
                                  const double *y, double t, double dt, int plu, counters *cnt)
{
	const int n = w->n, s = w->s, N = w->N;
	jac(pr, t, y, w->J);
	for (int j = 0; j < s; ++j)
		for (int i = 0; i < s; ++i)
			for (int l = 0; l < n; ++l)
				for (int k = 0; k < n; ++k) {
					double eye = (i*n + k == j*n + l) ? 1.0 : 0.0;
					AT(w->JY, N, i*n + k, j*n + l) = eye - dt * (tb->A[i*s + j] * AT(w->J, n, k, l));
				}
	if (plu) {
		memcpy(w->LU, w->JY, sizeof(double) * (size_t)N * (size_t)N);
		getrf(w->LU, N, w->piv);
	}
	cnt->jac_evals++;
}

/* irk.hpp:398-493.  adaptive_step / plu are the two template flags. */
static int newton_solve_stages(const problem *pr, const tableau *tb, workspace *w, const double *y,
                               double t, double dt, int maxit, int refresh_jac, double xtol,
                               double Rtol, int adaptive_step, int plu, int *iters_out,
                               double *res_out, counters *cnt)
{
	const int N = w->N, s = w->s;
	memset(w->Y, 0, sizeof(double) * (size_t)N);                       /* 415 */
	refresh_jacobi_matrix(pr, tb, w, y, t, dt, plu, cnt);               /* 438 */
	const double xtol2 = xtol*xtol, Rtol2 = Rtol*Rtol;                  /* 441-442 */
	construct_R(pr, tb, w, y, t, dt);                                   /* 443 */
	cnt->fun_evals += (unsigned long long)s;
	double step = 1.0;
	double Rnorm2 = dot(w->R, w->R, N);
	if (adaptive_step) step /= sqrt(1.0 + Rnorm2);
	double xnorm2 = 0;
	int status = NEWTON_MAXIT_EXCEEDED;
	int iters = 1;
	for (; iters < maxit; ++iters) {                                    /* 458 */
		if (plu) {
			/* tmp = solve(trimatl(-L), P*R); dY = solve(trimatu(U), tmp)   (461-462) */
			for (int i = 0; i < N; ++i) w->tmp[i] = w->R[i];
			for (int k = 0; k < N; ++k)
				if (w->piv[k] != k) {
					double sw = w->tmp[k]; w->tmp[k] = w->tmp[w->piv[k]]; w->tmp[w->piv[k]] = sw;
				}
			for (int i = 0; i < N; ++i) { /* non-unit forward substitution with diag(-L) = -1 */
				double acc = w->tmp[i];
				for (int k = 0; k < i; ++k) acc -= (-AT(w->LU, N, i, k)) * w->tmp[k];
				w->tmp[i] = acc / -1.0;
			}
			for (int i = 0; i < N; ++i) w->dY[i] = w->tmp[i];
			lu_backward(w->LU, N, w->dY);
		} else {
			/* dY = -solve(J_Y, R): refactorises every iteration (464) */
			for (int i = 0; i < N; ++i) w->dY[i] = w->R[i];
			dense_solve(w->JY, N, w->dY, w->LU, w->piv);
			for (int i = 0; i < N; ++i) w->dY[i] = -w->dY[i];
		}
		xnorm2 = dot(w->dY, w->dY, N);                                  /* 466 */
		for (int i = 0; i < N; ++i) w->Y[i] += step * w->dY[i];         /* 468 */
		construct_R(pr, tb, w, y, t, dt);                               /* 469 */
		cnt->fun_evals += (unsigned long long)s;
		Rnorm2 = dot(w->R, w->R, N);
		if (Rnorm2 < Rtol2) { status = NEWTON_SUCCESS; break; }         /* 473-476 */
		if (xnorm2 < xtol2) { status = NEWTON_SUCCESS; break; }         /* 477-480 */
		if (adaptive_step) step = 1.0 / sqrt(1.0 + Rnorm2);
		if (iters % refresh_jac == 0)                                   /* 485-487 */
			refresh_jacobi_matrix(pr, tb, w, y, t, dt, plu, cnt);
	}
	*iters_out = iters;
	*res_out = Rnorm2;
	return status;
}

typedef struct {
	double rel_tol, abs_tol, max_dt;
	long long max_steps;
	int adaptive_step_size;
	int out_interval;
	double newton_tol, newton_dx_delta;
	int newton_maxit, newton_refresh_jac;
	int output_mode;
	unsigned long long output_interval;
} ref_options;

typedef struct {
	size_t cap, len; /* len counts every stored sample, even beyond cap */
	double *t, *y, *err;
} trajectory;

static void traj_push(trajectory *tr, int n, double t, const double *y, double err)
{
	if (!tr) return;
	if (tr->len < tr->cap) {
		tr->t[tr->len] = t;
		tr->err[tr->len] = err;
		for (int k = 0; k < n; ++k) tr->y[tr->len * (size_t)n + k] = y[k];
	}
	tr->len++;
}

/* irk.hpp:509-869 (+ the wrapper 884-898).  Final state = last stored sample, like the driver
 * around the reference reads sol.y_vals.back(). */
static int irk_guts(const problem *pr, const tableau *tb, workspace *w, double t0, double t1,
                    const double *y0, const ref_options *o, double dt, double *y_out,
                    double *t_out, double *err_out, counters *cnt, double *accept_frac,
                    trajectory *tr)
{
	const int n = w->n, s = w->s;
	int adaptive = o->adaptive_step_size != 0;
	if (adaptive && !tb->has_b2) adaptive = 0;                          /* 890-895 */
	if (t0 + dt > t1) dt = t1 - t0;                                     /* 514-520 */
	double t = t0;
	int status = ST_SUCCESS;
	const int newton_maxit0 = o->newton_maxit;
	int newton_maxit = newton_maxit0;
	long long last_maxit_relax_step = 0, step = 0;
	double y[n], y_n[n], delta_y[n], delta_alt[n], dy_alt[n], delta_delta[n], err_est[n], fy[n],
	    ytmp[n], rhs[n];
	double d_weights[MAX_S], d2_weights[MAX_S];
	for (int k = 0; k < n; ++k) { y[k] = y0[k]; err_est[k] = 0.0; }
	double dts[3] = {dt, dt, dt}, errs[3] = {0.9, 0.9, 0.9};            /* 568-573 */
	double err = 0.0, new_dt = dt;
	memset(cnt, 0, sizeof *cnt);
	/* stored-sample bookkeeping: sample 0 is (t0, y0) with err 0 (581-585) */
	double last_t = t, last_err = 0.0;
	for (int k = 0; k < n; ++k) y_out[k] = y[k];
	traj_push(tr, n, t, y, 0.0);
	int alternative_error_formula = 1;                                  /* 587 */
	const int min_order = tb->order < tb->order2 ? tb->order : tb->order2; /* 588 */
	const double xtol = o->newton_dx_delta, Rtol = o->newton_tol;       /* 594-595 */
	inv_transpose_times(tb, tb->b, d_weights);                          /* 599-600 */
	if (tb->has_b2) inv_transpose_times(tb, tb->b2, d2_weights);        /* 601 */

	while (t < t1) {                                                    /* 604 */
		if (t + dt > t1) dt = t1 - t;                                   /* 607-609 */
		cnt->attempt++;
		if (o->max_steps >= 0 && step > o->max_steps) {                 /* 612-616 */
			status = ST_ERROR_MAX_STEPS_EXCEEDED;
			goto done;
		}
		int integrator_status = 0, iters = 0;
		double res = 0.0;
		int nstat = newton_solve_stages(pr, tb, w, y, t, dt, newton_maxit, o->newton_refresh_jac,
		                                xtol, Rtol, 0, 1, &iters, &res, cnt); /* 621-629 */
		if (nstat != NEWTON_SUCCESS) {                                  /* 634 */
			if (!adaptive) { status = ST_GENERAL_ERROR; goto done; }    /* 635-642 */
			dt *= 0.7;                                                  /* 644 */
			if (step - last_maxit_relax_step > 15) {                    /* 645-648 */
				newton_maxit += newton_maxit0;
				last_maxit_relax_step = step;
			}
			cnt->reject_newton++;
			cnt->newton_maxit_exceed++;                                 /* 662-664 */
			continue;                                                   /* 665 */
		}
		cnt->newton_success++;                                          /* 667 */

		const double gam = tb->gamma*dt;                                /* 692 */
		/* YYs = reshape(Y,n,s): column i = stage i; delta_y = YYs*d_weights (695-696) */
		for (int k = 0; k < n; ++k) {
			double acc = 0.0;
			for (int i = 0; i < s; ++i) acc += w->Y[i*n + k] * d_weights[i];
			delta_y[k] = acc;
		}
		if (adaptive)
			for (int k = 0; k < n; ++k) {                               /* 698-700 */
				double acc = 0.0;
				for (int i = 0; i < s; ++i) acc += w->Y[i*n + k] * d2_weights[i];
				delta_alt[k] = acc;
			}
		else {
			/* The reference adds an EMPTY vector here and Armadillo throws:
			 * fixed-step irk_guts is unusable.  Mirror that as a hard error. */
			status = -1000;
			goto done;
		}
		fun(pr, t, y, fy);                                              /* 702 */
		cnt->fun_evals++;
		for (int k = 0; k < n; ++k) dy_alt[k] = gam * fy[k] + delta_alt[k];
		for (int k = 0; k < n; ++k) y_n[k] = y[k] + delta_y[k];         /* 705 */
		for (int k = 0; k < n; ++k) delta_delta[k] = dy_alt[k] - delta_y[k]; /* 707 */

		/* solve_tmp = I - gam*J ; err = dt*solve(solve_tmp, delta_delta)   (717-719) */
		for (int j = 0; j < n; ++j)
			for (int i = 0; i < n; ++i)
				AT(w->E, n, i, j) = (i == j ? 1.0 : 0.0) - gam * AT(w->J, n, i, j);
		for (int k = 0; k < n; ++k) rhs[k] = delta_delta[k];
		dense_solve(w->E, n, rhs, w->Ework, w->epiv);
		for (int k = 0; k < n; ++k) err_est[k] = dt * rhs[k];
		if (alternative_error_formula) {                                /* 722-731 */
			for (int k = 0; k < n; ++k) ytmp[k] = y[k] + err_est[k];
			fun(pr, t, ytmp, fy);
			cnt->fun_evals++;
			for (int k = 0; k < n; ++k) rhs[k] = gam * fy[k];
			for (int k = 0; k < n; ++k) rhs[k] += delta_alt[k];
			for (int k = 0; k < n; ++k) rhs[k] = rhs[k] - delta_y[k];
			dense_solve(w->E, n, rhs, w->Ework, w->epiv);
			for (int k = 0; k < n; ++k) err_est[k] = dt * rhs[k];
		}
		double err_tot = 0.0, nn = 0.0;                                 /* 733-746 */
		for (int i = 0; i < n; ++i) {
			double erri = err_est[i];
			double y0i = fabs(y[i]), y1i = fabs(y_n[i]);
			double sci = o->abs_tol + o->rel_tol * (y0i < y1i ? y1i : y0i);
			double add = erri / sci;
			err_tot += add * add;
			nn += 1.0;
		}
		err = sqrt(err_tot / nn);                                       /* 749 */
		if (err < MACHINE_PRECISION) err = MACHINE_PRECISION;           /* 752-754 */
		errs[2] = errs[1]; errs[1] = errs[0]; errs[0] = err;            /* 756-758 */
		if (adaptive && (err > 1.0)) {                                  /* 761-766 */
			alternative_error_formula = 1;
			integrator_status = 1;
			cnt->reject_err++;
		}
		double fac = 0.9 * (newton_maxit + 1.0);                        /* 771-772 */
		fac /= (newton_maxit + iters);
		double expt = 1.0 / (1.0 + min_order);                          /* 774 */
		double err_inv = 1.0 / err;
		double scale_27 = pow(err_inv, expt);
		double dt_rat = dts[0] / dts[1];
		double err_frac = errs[1] / errs[0];
		if (errs[1] == 0 || errs[0] == 0) err_frac = 1.0;               /* 779-781 */
		double err_rat = pow(err_frac, expt);
		double scale_28 = scale_27 * dt_rat * err_rat;
		double min_scales = scale_28 < scale_27 ? scale_28 : scale_27;  /* 796: std::min(a,b) */
		new_dt = fac * dt * (min_scales < 8.0 ? min_scales : 8.0);      /* 798: std::min(8.0,m) */
		if (o->max_dt > 0) new_dt = new_dt < o->max_dt ? new_dt : o->max_dt; /* 799-801 */

		if (!adaptive || integrator_status == 0) {                      /* 813 */
			for (int k = 0; k < n; ++k) y[k] = y_n[k];
			t += dt;
			++step;
			if (step % (long long)o->output_interval == 0) {            /* 825 */
				if (o->output_mode & 1) {                               /* store_in_vectors */
					last_t = t; last_err = err;
					for (int k = 0; k < n; ++k) y_out[k] = y_n[k];
					traj_push(tr, n, t, y_n, err);
				}
			}
			alternative_error_formula = 0;                              /* 845 */
		}
		if (adaptive) dt = new_dt;                                      /* 850-852 */
		dts[2] = dts[1]; dts[1] = dts[0]; dts[0] = dt;                  /* 853-855 */
	}
done:
	*t_out = last_t;
	*err_out = last_err;
	*accept_frac = (double)step / (double)cnt->attempt;                 /* 864 */
	return status;
}

/* ------------------------------------------------------------------ exported entry points */
static _Thread_local char g_err[256];
const char *ref_last_error(void) { return g_err; }

int ref_name_to_method(const char *name)
{
	for (size_t i = 0; i < sizeof METHOD_NAMES / sizeof METHOD_NAMES[0]; ++i)
		if (!strcmp(METHOD_NAMES[i].name, name)) return METHOD_NAMES[i].id;
	return 0; /* irk.cpp:715-718: std::map::operator[] default */
}
const char *ref_method_to_name(int m)
{
	for (size_t i = 0; i < sizeof METHOD_NAMES / sizeof METHOD_NAMES[0]; ++i)
		if (METHOD_NAMES[i].id == m) return METHOD_NAMES[i].name;
	return "";
}

int ref_get_coefficients(int method, double *A, double *b, double *c, double *b2, double *gamma,
                         int *order, int *order2, double *d, double *d2)
{
	tableau tb;
	if (!get_tableau(method, &tb)) return 0;
	for (int i = 0; i < tb.s; ++i) {
		for (int j = 0; j < tb.s; ++j) A[i*tb.s + j] = tb.A[i*tb.s + j];
		b[i] = tb.b[i]; c[i] = tb.c[i]; b2[i] = tb.b2[i];
	}
	*gamma = tb.gamma; *order = tb.order; *order2 = tb.order2;
	inv_transpose_times(&tb, tb.b, d);
	inv_transpose_times(&tb, tb.b2, d2);
	return tb.s;
}

int ref_project_b(int method, double theta, double *out) { (void)method; (void)theta; (void)out; return 0; }

int ref_irk_ensemble(int problem_id, int method, double t0, double t1, double dt0, size_t M,
                     int n_hint, const double *y0, int y0_broadcast, const double *params,
                     const ref_options *opts, size_t start, size_t stride, double *y,
                     double *t_final, double *err_last, int *status, unsigned long long *cnts,
                     double *accept_frac, double *elapsed_ms, size_t traj_cap, size_t *traj_len,
                     double *traj_t, double *traj_y, double *traj_err, char *log_buf, size_t log_cap)
{
	int n = 0, P = 0;
	tableau tb;
	workspace w;
	if (ref_problem_dims(problem_id, n_hint, &n, &P) || !opts || stride == 0) return -1;
	if (!get_tableau(method, &tb)) { snprintf(g_err, sizeof g_err, "method %d not in the port", method); return -1; }
	if (!ws_alloc(&w, n, tb.s)) { ws_free(&w); return -4; }
	struct timespec a, b;
	clock_gettime(CLOCK_MONOTONIC, &a);
	int rc = 0;
	for (size_t i = start; i < M; i += stride) {
		problem pr = { problem_id, n, {0, 0, 0, 0} };
		double Y0[n], yo[n];
		for (int k = 0; k < P; ++k) pr.p[k] = params[(size_t)k*M + i];
		for (int k = 0; k < n; ++k) Y0[k] = y0_broadcast ? y0[k] : y0[(size_t)k*M + i];
		counters c;
		trajectory tr = { traj_cap, 0, traj_t, traj_y, traj_err };
		int st = irk_guts(&pr, &tb, &w, t0, t1, Y0, opts, dt0, yo, &t_final[i], &err_last[i], &c,
		                  &accept_frac[i], traj_cap ? &tr : NULL);
		if (st == -1000) {
			snprintf(g_err, sizeof g_err, "addition: incompatible matrix dimensions: %dx1 and 0x1", n);
			rc = -2;
			break;
		}
		status[i] = st;
		for (int k = 0; k < n; ++k) y[(size_t)k*M + i] = yo[k];
		cnts[0*M+i] = c.attempt; cnts[1*M+i] = c.reject_newton; cnts[2*M+i] = c.reject_err;
		cnts[3*M+i] = c.newton_success; cnts[4*M+i] = c.newton_incr_diverge;
		cnts[5*M+i] = c.newton_iter_error_too_large; cnts[6*M+i] = c.newton_maxit_exceed;
		cnts[7*M+i] = c.fun_evals; cnts[8*M+i] = c.jac_evals;
		if (traj_cap && traj_len) *traj_len = tr.len;
	}
	clock_gettime(CLOCK_MONOTONIC, &b);
	if (elapsed_ms) *elapsed_ms = (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6;
	if (log_buf && log_cap) log_buf[0] = 0;
	ws_free(&w);
	return rc;
}

int ref_newton_solve_stages(int problem_id, int method, int variant, int n_hint,
                            const double *params, const double *y_in, double t, double dt,
                            int maxit, int refresh_jac, double xtol, double Rtol, double *Y_out,
                            double *y1_out, int *iters, double *res, unsigned long long *fun_evals,
                            unsigned long long *jac_evals)
{
	int n = 0, P = 0;
	tableau tb;
	workspace w;
	if (ref_problem_dims(problem_id, n_hint, &n, &P)) return -1;
	if (!get_tableau(method, &tb)) return -1;
	if (!ws_alloc(&w, n, tb.s)) { ws_free(&w); return -4; }
	problem pr = { problem_id, n, {0, 0, 0, 0} };
	for (int k = 0; k < P; ++k) pr.p[k] = params[k];
	counters c;
	memset(&c, 0, sizeof c);
	int st = newton_solve_stages(&pr, &tb, &w, y_in, t, dt, maxit, refresh_jac, xtol, Rtol,
	                             variant == 0, variant != 0, iters, res, &c);
	double d[MAX_S];
	inv_transpose_times(&tb, tb.b, d);
	for (int k = 0; k < w.N; ++k) Y_out[k] = w.Y[k];
	for (int k = 0; k < n; ++k) {
		double acc = 0.0;
		for (int i = 0; i < tb.s; ++i) acc += w.Y[i*n + k] * d[i];
		y1_out[k] = y_in[k] + acc;
	}
	*fun_evals = c.fun_evals;
	*jac_evals = c.jac_evals;
	ws_free(&w);
	return st;
}

/* Fixed-step oracle: loop of newton_solve_stages<F,false,true> + y += YYs*d_weights
 * (test/irk.cpp:271-286; see ref_driver.cpp ref_fixed_step). */
int ref_fixed_step(int problem_id, int method, int n_hint, const double *params, const double *y_in,
                   double t0, double dt, long long nsteps, int maxit, int refresh_jac, double xtol,
                   double Rtol, double *y_out, unsigned long long *newton_iters)
{
	int n = 0, P = 0;
	tableau tb;
	workspace w;
	if (ref_problem_dims(problem_id, n_hint, &n, &P)) return -1;
	if (!get_tableau(method, &tb)) return -1;
	if (!ws_alloc(&w, n, tb.s)) { ws_free(&w); return -4; }
	problem pr = { problem_id, n, {0, 0, 0, 0} };
	for (int k = 0; k < P; ++k) pr.p[k] = params[k];
	double d[MAX_S], y[n];
	inv_transpose_times(&tb, tb.b, d);
	for (int k = 0; k < n; ++k) y[k] = y_in[k];
	int failed = 0;
	double t = t0;
	*newton_iters = 0;
	for (long long st = 0; st < nsteps; ++st) {
		counters c;
		int iters = 0;
		double res = 0.0;
		memset(&c, 0, sizeof c);
		int status = newton_solve_stages(&pr, &tb, &w, y, t, dt, maxit, refresh_jac, xtol, Rtol, 0, 1,
		                                 &iters, &res, &c);
		if (status != NEWTON_SUCCESS) ++failed;
		*newton_iters += (unsigned long long)iters;
		for (int k = 0; k < n; ++k) {
			double acc = 0.0;
			for (int i = 0; i < tb.s; ++i) acc += w.Y[i*n + k] * d[i];
			y[k] = y[k] + acc;
		}
		t += dt;
	}
	for (int k = 0; k < n; ++k) y_out[k] = y[k];
	ws_free(&w);
	return failed;
}

/* fun/jac of the restated functors, J row-major (see ref_driver.cpp ref_eval_functor). */
int ref_eval_functor(int problem_id, int n_hint, const double *params, double t, const double *y, double *f_out,
                     double *J_out)
{
	int n = 0, P = 0;
	if (ref_problem_dims(problem_id, n_hint, &n, &P)) return -1;
	problem pr;
	pr.id = problem_id;
	pr.n = n;
	for (int k = 0; k < 4; ++k) pr.p[k] = (params && k < P) ? params[k] : 0.0;
	double *J = (double *)malloc(sizeof(double) * (size_t)n * (size_t)n);
	if (!J) return -2;
	fun(&pr, t, y, f_out);
	jac(&pr, t, y, J);
	for (int i = 0; i < n; ++i)
		for (int j = 0; j < n; ++j) J_out[i*n + j] = AT(J, n, i, j);
	free(J);
	return 0;
}
