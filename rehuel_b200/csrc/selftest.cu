// Diagnostic entry point: runs the warp kernel's band LU / band solve (irk_warp.cuh) on caller-given
// band matrices, one warp per problem, lane m <-> matrix m exactly as the integrator uses them.
// Lets the tests check both factorisation variants (plain elimination + "would have pivoted"
// flag, dgbtf2-style pivoting) against a dense LAPACK solve on matrices the ODE problems never
// produce (the Brusselator's matrices never need an interchange).
#include <cuda_runtime.h>

#include "../../include/rehuel_b200.h"
#include "irk_warp.cuh"
#include "launch.cuh"

namespace rehuel {

constexpr int kSelfN = 64, kSelfKB = 2, kSelfW = 3 * kSelfKB + 1, kSelfPL = (kSelfN + kSelfKB) * kSelfW;

// per problem and lane (4 lanes): planes re/im [N+KB][W], rhs re/im [N]
template <bool PIVOT>
__global__ void __launch_bounds__(32) band_selftest_kernel(const double *__restrict__ ab_in, const double *__restrict__ rhs_in,
                                                          const int *__restrict__ is_cplx, double *__restrict__ x_out,
                                                          int *__restrict__ flags)
{
	__shared__ double ab[4][2][kSelfPL];
	__shared__ double rd[4][2][kSelfN];
	__shared__ double x[4][2][kSelfN + 2 * kSelfKB];
	__shared__ signed char piv[4][kSelfN];
	const int lane = threadIdx.x, prob = blockIdx.x;
	for (int e = lane; e < 4 * 2 * kSelfPL; e += 32) (&ab[0][0][0])[e] = ab_in[(size_t)prob * 4 * 2 * kSelfPL + e];
	for (int e = lane; e < 4 * 2 * (kSelfN + 2 * kSelfKB); e += 32) (&x[0][0][0])[e] = 0.0;
	__syncwarp();
	for (int e = lane; e < 4 * 2 * kSelfN; e += 32) {
		const int m = e / (2 * kSelfN), part = (e / kSelfN) % 2, i = e % kSelfN;
		x[m][part][i] = rhs_in[(size_t)prob * 4 * 2 * kSelfN + e];
	}
	__syncwarp();
	if (lane < 4) {
		const bool c = is_cplx[prob * 4 + lane] != 0;
		const bool f = band_lu<kSelfN, kSelfKB, PIVOT>(ab[lane][0], ab[lane][1], c, rd[lane][0], rd[lane][1], piv[lane]);
		flags[prob * 4 + lane] = f ? 1 : 0;
		band_solve<kSelfN, kSelfKB, PIVOT>(ab[lane][0], ab[lane][1], c, rd[lane][0], rd[lane][1], piv[lane], x[lane][0], x[lane][1]);
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
	double *d_ab = nullptr, *d_rhs = nullptr, *d_x = nullptr;
	int *d_c = nullptr, *d_f = nullptr;
	REHUEL_CUDA_TRY(cudaMalloc(&d_ab, nab * 8));
	REHUEL_CUDA_TRY(cudaMalloc(&d_rhs, nx * 8));
	REHUEL_CUDA_TRY(cudaMalloc(&d_x, nx * 8));
	REHUEL_CUDA_TRY(cudaMalloc(&d_c, (size_t)n_problems * 4 * sizeof(int)));
	REHUEL_CUDA_TRY(cudaMalloc(&d_f, (size_t)n_problems * 4 * sizeof(int)));
	REHUEL_CUDA_TRY(cudaMemcpy(d_ab, ab, nab * 8, cudaMemcpyHostToDevice));
	REHUEL_CUDA_TRY(cudaMemcpy(d_rhs, rhs, nx * 8, cudaMemcpyHostToDevice));
	REHUEL_CUDA_TRY(cudaMemcpy(d_c, is_cplx, (size_t)n_problems * 4 * sizeof(int), cudaMemcpyHostToDevice));
	if (pivot) band_selftest_kernel<true><<<n_problems, 32>>>(d_ab, d_rhs, d_c, d_x, d_f);
	else band_selftest_kernel<false><<<n_problems, 32>>>(d_ab, d_rhs, d_c, d_x, d_f);
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	REHUEL_CUDA_TRY(cudaMemcpy(x, d_x, nx * 8, cudaMemcpyDeviceToHost));
	REHUEL_CUDA_TRY(cudaMemcpy(flags, d_f, (size_t)n_problems * 4 * sizeof(int), cudaMemcpyDeviceToHost));
	cudaFree(d_ab);
	cudaFree(d_rhs);
	cudaFree(d_x);
	cudaFree(d_c);
	cudaFree(d_f);
	return REHUEL_OK;
}
