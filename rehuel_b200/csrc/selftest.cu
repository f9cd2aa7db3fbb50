// Diagnostic entry point: runs the warp kernel's band LU / band solve (irk_warp.cuh, band_twisted.cuh)
// on caller-given band matrices, one warp per problem, the matrices side by side in the lanes exactly
// as the integrator uses them.  Lets the tests check both factorisation variants (twisted elimination
// by lane pairs + "would have pivoted" flag; dgbtf2-style pivoting by single lanes) against a dense
// LAPACK solve on matrices the ODE problems never produce (the Brusselator's never need an interchange).
#include <cuda_runtime.h>

#include "../../include/rehuel_b200.h"
#include "irk_warp.cuh"
#include "launch.cuh"

namespace rehuel {

constexpr int kSelfN = 64, kSelfKB = 2, kSelfW = 3 * kSelfKB + 1, kSelfPL = (kSelfN + kSelfKB) * kSelfW;

// pivoted: per problem and lane (4 lanes): planes re/im [N+KB][W], rhs re/im [N]
__global__ void __launch_bounds__(32) band_selftest_pivoted_kernel(const double *__restrict__ ab_in, const double *__restrict__ rhs_in,
                                                                  const int *__restrict__ is_cplx, double *__restrict__ x_out,
                                                                  int *__restrict__ flags, double *__restrict__ scratch)
{
	// the factors live in GLOBAL memory like the integrator's fallback slab
	double *ab = scratch + (size_t)blockIdx.x * (4 * 2 * kSelfPL + 4 * 2 * kSelfN);
	double *rd = ab + 4 * 2 * kSelfPL;
	__shared__ double x[4][2][kSelfN + 2 * kSelfKB];
	__shared__ signed char piv[4][kSelfN];
	const int lane = threadIdx.x, prob = blockIdx.x;
	for (int e = lane; e < 4 * 2 * kSelfPL; e += 32) ab[e] = ab_in[(size_t)prob * 4 * 2 * kSelfPL + e];
	for (int e = lane; e < 4 * 2 * (kSelfN + 2 * kSelfKB); e += 32) (&x[0][0][0])[e] = 0.0;
	__syncwarp();
	for (int e = lane; e < 4 * 2 * kSelfN; e += 32) {
		const int m = e / (2 * kSelfN), part = (e / kSelfN) % 2, i = e % kSelfN;
		x[m][part][i] = rhs_in[(size_t)prob * 4 * 2 * kSelfN + e];
	}
	__syncwarp();
	if (lane < 4) {
		const bool c = is_cplx[prob * 4 + lane] != 0;
		double *pr = ab + (lane * 2) * kSelfPL, *pi = pr + kSelfPL, *dr = rd + (lane * 2) * kSelfN, *di = dr + kSelfN;
		const bool f = band_lu<kSelfN, kSelfKB, true>(pr, pi, c, dr, di, piv[lane]);
		flags[prob * 4 + lane] = f ? 1 : 0;
		band_solve<kSelfN, kSelfKB, true>(pr, pi, c, dr, di, piv[lane], x[lane][0], x[lane][1]);
	}
	__syncwarp();
	for (int e = lane; e < 4 * 2 * kSelfN; e += 32) {
		const int m = e / (2 * kSelfN), part = (e / kSelfN) % 2, i = e % kSelfN;
		x_out[(size_t)prob * 4 * 2 * kSelfN + e] = x[m][part][i];
	}
}

// twisted: lanes 2m, 2m+1 <-> matrix m; half planes, separator blocks and right-hand sides in shared memory
__global__ void __launch_bounds__(32) band_selftest_twisted_kernel(const double *__restrict__ ab_in, const double *__restrict__ rhs_in,
                                                                  const int *__restrict__ is_cplx, double *__restrict__ x_out,
                                                                  int *__restrict__ flags)
{
	using G = Tw<kSelfN, kSelfKB>;
	__shared__ double hp[4][2][2][G::HP]; // [matrix][half][re|im]
	__shared__ double q[4][2][2][kSelfKB * kSelfKB];
	__shared__ double x[4][2][kSelfN];
	const int lane = threadIdx.x, prob = blockIdx.x;
	constexpr int D = 2 * kSelfKB + 1;
	for (int e = lane; e < 4 * 2 * kSelfN * D; e += 32) {
		const int d = e % D - kSelfKB, i = (e / D) % kSelfN, part = (e / (D * kSelfN)) % 2, m = e / (D * kSelfN * 2);
		const double v = ab_in[(size_t)prob * 4 * 2 * kSelfPL + (size_t)(m * 2 + part) * kSelfPL + i * kSelfW + kSelfKB + d];
		tw_put<kSelfN, kSelfKB>(hp[m][0][part], hp[m][1][part], i, d, v);
	}
	for (int e = lane; e < 4 * 2 * kSelfN; e += 32) {
		const int m = e / (2 * kSelfN), part = (e / kSelfN) % 2, i = e % kSelfN;
		x[m][part][i] = rhs_in[(size_t)prob * 4 * 2 * kSelfN + e];
	}
	__syncwarp();
	if (lane < 8) {
		const int m = lane >> 1, h = lane & 1;
		const bool c = is_cplx[prob * 4 + m] != 0;
		const int np = h ? G::NP1 : G::NP0;
		bool f = tw_factor_pair<kSelfN, kSelfKB>(0xffu, hp[m][h][0], hp[m][h][1], c, np, q[m][h][0], q[m][h][1]);
		f |= __shfl_xor_sync(0xffu, f ? 1 : 0, 1) != 0;
		if (h == 0) flags[prob * 4 + m] = f ? 1 : 0;
		double *xr = x[m][0] + (h ? kSelfN - 1 : 0), *xi = x[m][1] + (h ? kSelfN - 1 : 0);
		tw_solve_pair<kSelfN, kSelfKB>(0xffu, hp[m][h][0], hp[m][h][1], c, np, q[m][h][0], q[m][h][1], xr, xi, h ? -1 : 1);
	}
	__syncwarp();
	for (int e = lane; e < 4 * 2 * kSelfN; e += 32) {
		const int m = e / (2 * kSelfN), part = (e / kSelfN) % 2, i = e % kSelfN;
		x_out[(size_t)prob * 4 * 2 * kSelfN + e] = x[m][part][i];
	}
}

} // namespace rehuel

extern "C" int rehuel_selftest_band(int n_problems, int pivot, const double *ab, const double *rhs, const int *is_cplx,
                                    double *x, int *flags)
{
	using namespace rehuel;
	if (n_problems <= 0 || !ab || !rhs || !is_cplx || !x || !flags) return set_error(REHUEL_EINVAL, "selftest: bad arguments");
	const size_t nab = (size_t)n_problems * 4 * 2 * kSelfPL, nx = (size_t)n_problems * 4 * 2 * kSelfN;
	double *d_ab = nullptr, *d_rhs = nullptr, *d_x = nullptr, *d_scratch = nullptr;
	int *d_c = nullptr, *d_f = nullptr;
	REHUEL_CUDA_TRY(cudaMalloc(&d_ab, nab * 8));
	REHUEL_CUDA_TRY(cudaMalloc(&d_rhs, nx * 8));
	REHUEL_CUDA_TRY(cudaMalloc(&d_x, nx * 8));
	REHUEL_CUDA_TRY(cudaMalloc(&d_c, (size_t)n_problems * 4 * sizeof(int)));
	REHUEL_CUDA_TRY(cudaMalloc(&d_f, (size_t)n_problems * 4 * sizeof(int)));
	REHUEL_CUDA_TRY(cudaMemcpy(d_ab, ab, nab * 8, cudaMemcpyHostToDevice));
	REHUEL_CUDA_TRY(cudaMemcpy(d_rhs, rhs, nx * 8, cudaMemcpyHostToDevice));
	REHUEL_CUDA_TRY(cudaMemcpy(d_c, is_cplx, (size_t)n_problems * 4 * sizeof(int), cudaMemcpyHostToDevice));
	if (pivot) {
		REHUEL_CUDA_TRY(cudaMalloc(&d_scratch, (size_t)n_problems * (4 * 2 * kSelfPL + 4 * 2 * kSelfN) * 8));
		band_selftest_pivoted_kernel<<<n_problems, 32>>>(d_ab, d_rhs, d_c, d_x, d_f, d_scratch);
	} else {
		band_selftest_twisted_kernel<<<n_problems, 32>>>(d_ab, d_rhs, d_c, d_x, d_f);
	}
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	REHUEL_CUDA_TRY(cudaMemcpy(x, d_x, nx * 8, cudaMemcpyDeviceToHost));
	REHUEL_CUDA_TRY(cudaMemcpy(flags, d_f, (size_t)n_problems * 4 * sizeof(int), cudaMemcpyDeviceToHost));
	cudaFree(d_ab);
	cudaFree(d_rhs);
	cudaFree(d_x);
	cudaFree(d_c);
	cudaFree(d_f);
	cudaFree(d_scratch);
	return REHUEL_OK;
}
