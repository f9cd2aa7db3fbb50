// Host-side tableaux.  Replaces irk::get_coefficients (reference irk.cpp:215-689), the
// name<->id maps (irk.cpp:709-728, enums.hpp:38-58), project_b (irk.cpp:731-747) and
// collocation_interpolate_coeffs (irk.cpp:111-210).
#include "tableau.hpp"

#include <cmath>
#include <cstring>
#include <initializer_list>

namespace rehuel {

namespace {

struct MethodName {
	int id;
	const char *name;
};
// enums.hpp:38-58 -- the complete id space of the reference, including ids it declares but never
// implements (LOBATTO_IIIA_85, LOBATTO_IIIC_127, GAUSS_LEGENDRE_63/105), so that name<->id round
// trips behave the same.
const MethodName kMethodNames[] = {
    {200, "IMPLICIT_EULER"},   {210, "RADAU_IIA_32"},      {211, "RADAU_IIA_53"},
    {212, "RADAU_IIA_95"},     {213, "RADAU_IIA_137"},     {220, "LOBATTO_IIIA_43"},
    {221, "LOBATTO_IIIA_85"},  {222, "LOBATTO_IIIA_127"},  {230, "LOBATTO_IIIC_43"},
    {231, "LOBATTO_IIIC_64"},  {232, "LOBATTO_IIIC_85"},   {233, "LOBATTO_IIIC_127"},
    {240, "GAUSS_LEGENDRE_42"}, {241, "GAUSS_LEGENDRE_63"}, {242, "GAUSS_LEGENDRE_105"},
    {243, "GAUSS_LEGENDRE_147"},
};

struct EigEntry {
	const char *name;
	int s, n_real, n_cplx;
	double gamma_hat[1];
	double alpha[3];
	double beta[3];
	double T[MAX_S * MAX_S];
	double Ti[MAX_S * MAX_S];
};
const EigEntry kEig[] = {
#include "tableau_eig.inc"
};

#include "tableau_dec.inc" // RADAU_IIA_95 / RADAU_IIA_137 as the reference prints them

// x = A^-1 rhs for a small dense system, Gaussian elimination with row partial pivoting.
bool small_solve(int s, const double A[MAX_S][MAX_S], const double *rhs, double *x)
{
	double a[MAX_S][MAX_S + 1];
	for (int i = 0; i < s; ++i) {
		for (int j = 0; j < s; ++j) a[i][j] = A[i][j];
		a[i][s] = rhs[i];
	}
	for (int k = 0; k < s; ++k) {
		int p = k;
		for (int i = k + 1; i < s; ++i)
			if (std::fabs(a[i][k]) > std::fabs(a[p][k])) p = i;
		if (a[p][k] == 0.0) return false;
		if (p != k)
			for (int j = 0; j <= s; ++j) {
				double tmp = a[k][j];
				a[k][j] = a[p][j];
				a[p][j] = tmp;
			}
		for (int i = k + 1; i < s; ++i) {
			double l = a[i][k] / a[k][k];
			for (int j = k; j <= s; ++j) a[i][j] -= l * a[k][j];
		}
	}
	for (int i = s - 1; i >= 0; --i) {
		double acc = a[i][s];
		for (int j = i + 1; j < s; ++j) acc -= a[i][j] * x[j];
		x[i] = acc / a[i][i];
	}
	return true;
}

// d = A^-T w  <=>  A^T d = w
bool solve_transposed(const Tableau &tb, const double *w, double *d)
{
	double At[MAX_S][MAX_S];
	for (int i = 0; i < tb.s; ++i)
		for (int j = 0; j < tb.s; ++j) At[i][j] = tb.A[j][i];
	return small_solve(tb.s, At, w, d);
}

// Integrated Lagrange basis on the nodes c: b_j(theta) = int_0^theta l_j(x) dx, stored so that
// b_j(theta) = sum_k b_interp[j][k] * theta^(s-k)   (layout of irk.cpp:196-206 / 731-747).
void build_interpolant(Tableau &tb)
{
	const int s = tb.s;
	for (int j = 0; j < s; ++j) {
		// numerator polynomial prod_{m != j} (x - c_m), coefficients ascending in x
		double poly[MAX_S + 1] = {1.0};
		int deg = 0;
		double denom = 1.0;
		for (int m = 0; m < s; ++m) {
			if (m == j) continue;
			denom *= tb.c[j] - tb.c[m];
			for (int q = deg + 1; q >= 1; --q) poly[q] = poly[q - 1] - tb.c[m] * poly[q];
			poly[0] = -tb.c[m] * poly[0];
			++deg;
		}
		// integrate: x^q -> x^(q+1)/(q+1); power q+1 sits in column s-(q+1)
		for (int q = 0; q <= deg; ++q) tb.b_interp[j][s - (q + 1)] = poly[q] / denom / (q + 1.0);
	}
	tb.has_interp = true;
}

// Tableau given by decimal arrays (tableau_dec.inc): A row-major, b2 = b + coef*gamma.
void set_decimal(Tableau &tb, int s, const double *A, const double *c, const double *b, const double *b2coef,
                 double gamma, int order, int order2)
{
	tb.s = s;
	tb.gamma = gamma;
	tb.order = order;
	tb.order2 = order2;
	for (int i = 0; i < s; ++i) {
		for (int j = 0; j < s; ++j) tb.A[i][j] = A[i * s + j];
		tb.c[i] = c[i];
		tb.b[i] = b[i];
		const double cg = b2coef[i] * gamma; // the reference writes  b(i) -+ |coef|*gamma
		tb.b2[i] = b[i] + cg;
	}
}

void set_row(Tableau &tb, int i, std::initializer_list<double> row)
{
	int j = 0;
	for (double v : row) tb.A[i][j++] = v;
}
void set_vec(double *dst, std::initializer_list<double> v)
{
	int j = 0;
	for (double x : v) dst[j++] = x;
}

} // namespace

int name_to_method(const char *name)
{
	if (!name) return 0;
	for (const MethodName &m : kMethodNames)
		if (!std::strcmp(m.name, name)) return m.id;
	return 0;
}

const char *method_to_name(int method)
{
	for (const MethodName &m : kMethodNames)
		if (m.id == method) return m.name;
	return "";
}

bool get_coefficients(int method, Tableau &tb)
{
	tb = Tableau();
	tb.method = method;
	tb.name = method_to_name(method);
	const double r5 = std::sqrt(5.0), r6 = std::sqrt(6.0), r21 = std::sqrt(21.0);
	bool interp = false;
	tb.has_b2 = true;

	switch (method) {
	case 200: // IMPLICIT_EULER, irk.cpp:238-244: no embedded weights => odeint runs it with a fixed step (irk.hpp:890-895)
		tb.s = 1;
		tb.A[0][0] = 1.0;
		tb.b[0] = 1.0;
		tb.c[0] = 1.0;
		tb.order = 1;
		tb.order2 = 0;
		tb.has_b2 = false;
		break;
	case 212: // RADAU_IIA_95, irk.cpp:393-437
		set_decimal(tb, 5, radau_iia_95_A, radau_iia_95_c, radau_iia_95_b, radau_iia_95_b2_gamma_coef, radau_iia_95_gamma,
		            radau_iia_95_order, radau_iia_95_order2);
		interp = true;
		break;
	case 213: // RADAU_IIA_137, irk.cpp:439-533
		set_decimal(tb, 7, radau_iia_137_A, radau_iia_137_c, radau_iia_137_b, radau_iia_137_b2_gamma_coef, radau_iia_137_gamma,
		            radau_iia_137_order, radau_iia_137_order2);
		interp = true;
		break;
	case 210: // RADAU_IIA_32, irk.cpp:342-360
		tb.s = 2;
		set_row(tb, 0, {5.0 / 12.0, -1.0 / 12.0});
		set_row(tb, 1, {3.0 / 4.0, 1.0 / 4.0});
		set_vec(tb.c, {1.0 / 3.0, 1.0});
		set_vec(tb.b, {3.0 / 4.0, 1.0 / 4.0});
		set_vec(tb.b2, {2.0 / 3.0, 1.0 / 3.0});
		tb.gamma = 1.0 / 3.0;
		tb.order = 3;
		tb.order2 = 2;
		interp = true;
		break;
	case 211: // RADAU_IIA_53, irk.cpp:363-391
		tb.s = 3;
		set_row(tb, 0, {(88 - 7 * r6) / 360.0, (296 - 169 * r6) / 1800.0, (-2 + 3 * r6) / 225.0});
		set_row(tb, 1, {(296 + 169 * r6) / 1800.0, (88 + 7 * r6) / 360.0, (-2 - 3 * r6) / 225.0});
		set_row(tb, 2, {(16.0 - r6) / 36.0, (16 + r6) / 36.0, 1.0 / 9.0});
		set_vec(tb.c, {(4.0 - r6) / 10.0, (4.0 + r6) / 10.0, 1.0});
		set_vec(tb.b, {(16 - r6) / 36.0, (16 + r6) / 36.0, 1.0 / 9.0});
		tb.gamma = 2.74888829595677e-01; // the reference's 15-digit value, kept verbatim for parity
		set_vec(tb.b2, {-((18 * r6 + 12) * tb.gamma - 16 + r6) / 36.0,
		                ((18 * r6 - 12) * tb.gamma + 16 + r6) / 36.0, -(3 * tb.gamma - 1) / 9.0});
		tb.order = 5;
		tb.order2 = 3;
		interp = true;
		break;
	case 230: // LOBATTO_IIIC_43, irk.cpp:261-272 (gamma deliberately left 0, irk.cpp:231)
		tb.s = 3;
		set_row(tb, 0, {1.0 / 6.0, -1.0 / 3.0, 1.0 / 6.0});
		set_row(tb, 1, {1.0 / 6.0, 5.0 / 12.0, -1.0 / 12.0});
		set_row(tb, 2, {1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0});
		set_vec(tb.c, {0.0, 0.5, 1.0});
		set_vec(tb.b, {1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0});
		set_vec(tb.b2, {1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0});
		tb.order = 4;
		tb.order2 = 2;
		break;
	case 231: // LOBATTO_IIIC_64, irk.cpp:274-286 (gamma left 0)
		tb.s = 4;
		set_row(tb, 0, {1.0 / 12.0, -r5 / 12., r5 / 12., -1.0 / 12.0});
		set_row(tb, 1, {1.0 / 12.0, 0.25, (10 - 7 * r5) / 60., r5 / 60.});
		set_row(tb, 2, {1.0 / 12.0, (10 + 7 * r5) / 60., 0.25, -r5 / 60.});
		set_row(tb, 3, {1.0 / 12.0, 5.0 / 12.0, 5.0 / 12.0, 1.0 / 12.0});
		set_vec(tb.c, {0.0, 0.5 - r5 / 10., 0.5 + r5 / 10., 1.0});
		set_vec(tb.b, {1.0 / 12.0, 5.0 / 12.0, 5.0 / 12.0, 1.0 / 12.0});
		set_vec(tb.b2, {1.0 / 12.0, 1.0 / 12.0, 5.0 / 12.0, 5.0 / 12.0});
		tb.order = 6;
		tb.order2 = 4;
		break;
	case 232: // LOBATTO_IIIC_85, irk.cpp:288-327
		tb.s = 5;
		set_row(tb, 0, {1.0 / 20.0, -7.0 / 60.0, 2.0 / 15.0, -7.0 / 60.0, 1.0 / 20.0});
		set_row(tb, 1, {1.0 / 20.0, 29.0 / 180.0, (47.0 - 15 * r21) / 315.0, (203.0 - 30 * r21) / 1260.0, -3.0 / 140.0});
		set_row(tb, 2, {1.0 / 20.0, (329.0 + 105.0 * r21) / 2880.0, 73.0 / 360.0, (329.0 - 105.0 * r21) / 2880.0, 3.0 / 160.0});
		set_row(tb, 3, {1.0 / 20.0, (203.0 + 30 * r21) / 1260.0, (47 + 15 * r21) / 315.0, 29 / 180.0, -3 / 140.0});
		set_row(tb, 4, {1.0 / 20.0, 49 / 180.0, 16 / 45.0, 49 / 180.0, 1.0 / 20.0});
		set_vec(tb.c, {0.0, 0.5 - r21 / 14.0, 0.5, 0.5 + r21 / 14.0, 1.0});
		set_vec(tb.b, {0.05, 49 / 180.0, 16 / 45.0, 49 / 180.0, 0.05});
		set_vec(tb.b2, {0.05, 0.05, 49 / 180.0, 16 / 45.0, 49 / 180.0}); // b rotated by one
		tb.gamma = 0.1894964436645163; // verbatim; NOT 1/gamma_hat (differs by 4e-6) => EST_OWN_LU
		tb.order = 8;
		tb.order2 = 5;
		interp = true;
		break;
	default:
		return false;
	}

	if (!solve_transposed(tb, tb.b, tb.d)) return false;
	if (tb.has_b2 && !solve_transposed(tb, tb.b2, tb.d2)) return false;

	for (const EigEntry &e : kEig) {
		if (std::strcmp(e.name, tb.name)) continue;
		if (e.s != tb.s) return false;
		tb.n_real = e.n_real;
		tb.n_cplx = e.n_cplx;
		tb.gamma_hat = e.n_real ? e.gamma_hat[0] : 0.0;
		for (int k = 0; k < e.n_cplx; ++k) {
			tb.alpha[k] = e.alpha[k];
			tb.beta[k] = e.beta[k];
		}
		for (int i = 0; i < tb.s; ++i)
			for (int j = 0; j < tb.s; ++j) {
				tb.T[i][j] = e.T[i * tb.s + j];
				tb.Ti[i][j] = e.Ti[i * tb.s + j];
			}
		if (tb.gamma == 0.0) tb.est_mode = EST_IDENTITY;
		else if (tb.n_real && std::fabs(tb.gamma * tb.gamma_hat - 1.0) < 1e-13) tb.est_mode = EST_REUSE;
		else tb.est_mode = EST_OWN_LU;
		if (interp) {
			build_interpolant(tb);
			// w_interp[:, k] = A^-T b_interp[:, k]
			for (int k = 0; k < tb.s; ++k) {
				double col[MAX_S], w[MAX_S];
				for (int j = 0; j < tb.s; ++j) col[j] = tb.b_interp[j][k];
				if (!solve_transposed(tb, col, w)) return false;
				for (int i = 0; i < tb.s; ++i) tb.w_interp[i][k] = w[i];
			}
		}
		// the structure the thread-per-system kernels rely on (tableau.hpp: Tableau::canonical)
		tb.canonical = true;
		for (int i = 0; i < tb.s; ++i) tb.canonical = tb.canonical && tb.d[i] == (i == tb.s - 1 ? 1.0 : 0.0);
		for (int j = 0; j < tb.s; ++j) {
			const bool unit = j < tb.n_real || ((j - tb.n_real) % 2 == 0);
			tb.canonical = tb.canonical && tb.T[tb.s - 1][j] == (unit ? 1.0 : 0.0);
		}
		return true;
	}
	return false;
}

bool project_b(const Tableau &tb, double theta, double *out)
{
	if (!tb.has_interp) return false;
	const int s = tb.s;
	double ts[MAX_S];
	double tt = theta;
	for (int i = 0; i < s; ++i) { // ts = { theta^s, ..., theta^2, theta }
		ts[s - i - 1] = tt;
		tt *= theta;
	}
	for (int j = 0; j < s; ++j) {
		double acc = 0.0;
		for (int k = 0; k < s; ++k) acc += tb.b_interp[j][k] * ts[k];
		out[j] = acc;
	}
	return true;
}

} // namespace rehuel
