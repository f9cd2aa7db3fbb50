// Butcher tableaux of the implicit methods + everything the kernels derive from them.
// Replaces irk::solver_coeffs (reference irk.hpp:68-89) and irk::get_coefficients
// (irk.cpp:215-689); the derived weights d = A^-T b, d2 = A^-T b2 are irk.hpp:599-601.
#pragma once

#include <cmath>

namespace rehuel {

constexpr int MAX_S = 7;

// How the kernels obtain E^-1 for the embedded error estimate, E = I - gamma*dt*J (irk.hpp:717):
enum EstMode {
	EST_IDENTITY = 0, // gamma == 0 (LOBATTO_IIIC_43/64 leave it at 0, irk.cpp:231) => E = I
	EST_REUSE = 1,    // gamma == 1/gamma_hat: E is a multiple of the real Newton block, reuse its LU
	EST_OWN_LU = 2    // anything else: factorise E separately
};

// Host-side description of a method.
struct Tableau {
	const char *name = "";
	int method = 0;
	int s = 0;
	double A[MAX_S][MAX_S] = {};
	double b[MAX_S] = {}, c[MAX_S] = {}, b2[MAX_S] = {};
	bool has_b2 = false;
	double gamma = 0.0;
	int order = 0, order2 = 0;
	double d[MAX_S] = {}, d2[MAX_S] = {}; // A^-T b, A^-T b2
	// inv(A) = T Lambda Ti with Lambda = diag(gamma_hat, [[alpha,-beta],[beta,alpha]], ...)
	int n_real = 0, n_cplx = 0;
	double gamma_hat = 0.0, alpha[3] = {}, beta[3] = {};
	double T[MAX_S][MAX_S] = {}, Ti[MAX_S][MAX_S] = {};
	int est_mode = EST_IDENTITY;
	// dense output: b_j(theta) = sum_k b_interp[j][k] theta^(s-k)  (irk.cpp:111-210, 731-747)
	bool has_interp = false;
	double b_interp[MAX_S][MAX_S] = {};
	// dense output in terms of the stage increments Y_i the kernels hold:  y(t + theta dt) = y + sum_i w_i(theta) Y_i,
	// w(theta) = A^-T b(theta)  =>  w_i(theta) = sum_k w_interp[i][k] theta^(s-k)   (w(1) = d)
	double w_interp[MAX_S][MAX_S] = {};
	// The thread-per-system kernels skip multiplications by 1 and 0 that every built tableau has (checked by
	// get_coefficients, which refuses a tableau without them):
	//   d = A^-T b = e_s exactly (all methods here are stiffly accurate: b is the last row of A), and
	//   the last row of T is 1 for the real block and (1, 0) for every complex pair (tools/gen_tableaux.py scales
	//   the eigenvectors that way).
	bool canonical = false;
};

// Returns false if the method id is unknown / unsupported by the GPU path.
bool get_coefficients(int method, Tableau &tb);

int name_to_method(const char *name);   // irk.cpp:715-718 (0 if unknown)
const char *method_to_name(int method); // irk.cpp:709-712 ("" if unknown)

// irk::project_b (irk.cpp:731-747): { b_1(theta), ..., b_s(theta) }.
bool project_b(const Tableau &tb, double theta, double *out);

// Device-side constant block, passed by value as a __grid_constant__ kernel parameter so that
// every coefficient is a constant-bank operand of the FP64 instructions.
template <int S, int NC>
struct DevTableau {
	double A[S][S];
	double Asum[S];   // row sums of A: the stage residual at Y = 0 for an autonomous f (irk_thread.cuh)
	double TiAsum[S]; // Ti * Asum: the same residual in the transformed variables
	double c[S];
	double d[S], d2[S];
	double T[S][S], Ti[S][S];
	double gamma;     // tableau gamma (error estimator)
	double rgamma;    // 1/gamma (0 if gamma == 0)
	double gamma_hat; // real eigenvalue of inv(A) (0 if none)
	double alpha[NC > 0 ? NC : 1], beta[NC > 0 ? NC : 1];
	double Wd[S][S];  // dense output: w_i(theta) = sum_k Wd[i][k] theta^(S-k), y(t + theta dt) = y + sum_i w_i Y_i
	double expo;      // 1/(1+min(order,order2))  (irk.hpp:588, 774)
	double q_init;    // 0.9^expo: the controller's initial error history (irk.hpp:573) as propose_dt carries it
	int min_order;
};

template <int S, int NC>
inline DevTableau<S, NC> make_dev_tableau(const Tableau &tb)
{
	DevTableau<S, NC> dt{};
	for (int i = 0; i < S; ++i) {
		for (int j = 0; j < S; ++j) {
			dt.A[i][j] = tb.A[i][j];
			dt.T[i][j] = tb.T[i][j];
			dt.Ti[i][j] = tb.Ti[i][j];
			dt.Wd[i][j] = tb.w_interp[i][j];
		}
		for (int j = 0; j < S; ++j) dt.Asum[i] += tb.A[i][j];
		dt.c[i] = tb.c[i];
		dt.d[i] = tb.d[i];
		dt.d2[i] = tb.d2[i];
	}
	for (int i = 0; i < S; ++i)
		for (int j = 0; j < S; ++j) dt.TiAsum[i] += tb.Ti[i][j] * dt.Asum[j];
	dt.gamma = tb.gamma;
	dt.rgamma = tb.gamma != 0.0 ? 1.0 / tb.gamma : 0.0;
	dt.gamma_hat = tb.gamma_hat;
	for (int k = 0; k < NC; ++k) {
		dt.alpha[k] = tb.alpha[k];
		dt.beta[k] = tb.beta[k];
	}
	dt.min_order = tb.order < tb.order2 ? tb.order : tb.order2;
	dt.expo = 1.0 / (1.0 + dt.min_order);
	dt.q_init = std::pow(0.9, dt.expo);
	return dt;
}

} // namespace rehuel
