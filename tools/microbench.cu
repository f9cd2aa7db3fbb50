// Stand-alone sm_100a micro-benchmarks behind two design decisions (DESIGN.md sections 5.5 and 6):
//
//   1. issue model of the FP64 pipe: DFMA chains alone, DFMA interleaved 1:1 with independent integer instructions,
//      and 1:1 with FP32 FMAs.  If the mixed kernels take as long as the DFMA-only one, non-FP64 instructions hide
//      behind the 2-cycle FP64 issue; if they take longer, every instruction of the integration kernels costs time.
//   2. FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64) against the FP64 FMA pipe: raw rate with register-resident
//      accumulators, and the rank-4 trailing update C -= A B (C 64 x 64, A 64 x 4, B 4 x 64 in shared memory) of a
//      blocked LU as irk_cta.cuh would use it -- north_star: tensor cores "only if ncu shows them beating the FP64
//      FMA pipe".
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/microbench.cu -o build/microbench
//   build/microbench            -> one JSON line per experiment
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                   \
	do {                                                                                        \
		cudaError_t e_ = (x);                                                                   \
		if (e_ != cudaSuccess) {                                                                \
			std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                       \
			std::exit(1);                                                                       \
		}                                                                                       \
	} while (0)

// ---------------------------------------------------------------- 1. issue model
template <int MODE> // 0: DFMA only; 1: + 2-3 ALU integer ops per DFMA; 2: + one FFMA; 3: + one IMAD; 4: + one LOP3; 5: + one shared-memory load
__global__ void __launch_bounds__(256) issue_kernel(double *sink, int iters, double a, double b, unsigned m, float fa)
{
	double x[8];
	unsigned u[8];
	float f[8];
	__shared__ unsigned smem[1024];
	if (MODE == 5) {
		for (int e = threadIdx.x; e < 1024; e += blockDim.x) smem[e] = e * m;
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		x[i] = threadIdx.x * 1e-9 + i;
		u[i] = threadIdx.x * 2654435761u + i;
		f[i] = threadIdx.x * 1e-3f + i;
	}
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				x[i] = fma(x[i], a, b);
				if (MODE == 1) u[i] = (u[i] ^ m) + (u[i] >> 3); // LOP3 + shift/add: 2-3 integer instructions
				if (MODE == 2) f[i] = fmaf(f[i], fa, 1.0f);
				if (MODE == 3) u[i] = u[i] * m + 12345u;            // IMAD (FMA pipe)
				if (MODE == 4) u[i] = u[i] ^ (m + (unsigned)r);     // one LOP3 (ALU pipe)
				if (MODE == 5) u[i] += smem[(threadIdx.x + i * 32 + r) & 1023]; // LDS + IADD
			}
		}
	}
	double r = 0;
	unsigned ur = 0;
	float fr = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		r += x[i];
		ur += u[i];
		fr += f[i];
	}
	if (r == 123.456 || ur == 0x12345u || fr == 77.5f) sink[0] = r + ur + fr;
}

// ---------------------------------------------------------------- 2a. raw DMMA rate
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_kernel(double *sink, int iters, double a, double b)
{
	double c[8][2];
#pragma unroll
	for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9 + i;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int r = 0; r < 8; ++r)
#pragma unroll
			for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
	}
	double r = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) r += c[i][0] + c[i][1];
	if (r == 123.456) sink[0] = r;
}

// ---------------------------------------------------------------- 2b. rank-4 trailing update of a 64 x 64 block in shared memory
constexpr int NB = 64, KB = 4, LD = NB + 1; // padded rows: no bank conflicts on column walks

// FMA version, the access pattern of irk_cta.cuh: 256 threads, thread (ty, tx) owns C(ty + 16 p, tx + 16 q), p, q < 4
__global__ void __launch_bounds__(256) update_fma_kernel(double *out, int reps)
{
	__shared__ double C[NB * LD], A[NB * KB], B[KB * LD];
	const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
	for (int e = tid; e < NB * LD; e += 256) C[e] = 1.0 + 1e-3 * e;
	for (int e = tid; e < NB * KB; e += 256) A[e] = 1e-3 * (e % 7);
	for (int e = tid; e < KB * LD; e += 256) B[e] = 1e-3 * (e % 5);
	__syncthreads();
	for (int r = 0; r < reps; ++r) {
		double a[4][KB], b[KB][4];
#pragma unroll
		for (int p = 0; p < 4; ++p)
#pragma unroll
			for (int k = 0; k < KB; ++k) a[p][k] = A[(ty + 16 * p) * KB + k];
#pragma unroll
		for (int k = 0; k < KB; ++k)
#pragma unroll
			for (int q = 0; q < 4; ++q) b[k][q] = B[k * LD + tx + 16 * q];
#pragma unroll
		for (int p = 0; p < 4; ++p)
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				double c = C[(ty + 16 * p) * LD + tx + 16 * q];
#pragma unroll
				for (int k = 0; k < KB; ++k) c = fma(-a[p][k], b[k][q], c);
				C[(ty + 16 * p) * LD + tx + 16 * q] = c;
			}
		__syncthreads();
	}
	if (tid == 0) out[blockIdx.x] = C[5 * LD + 7];
}

// DMMA version: 8 warps, warp w owns the 8-row strip w of C (8 tiles of 8 x 8); per tile one m8n8k4
__global__ void __launch_bounds__(256) update_dmma_kernel(double *out, int reps)
{
	__shared__ double C[NB * LD], A[NB * KB], B[KB * LD];
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	for (int e = tid; e < NB * LD; e += 256) C[e] = 1.0 + 1e-3 * e;
	for (int e = tid; e < NB * KB; e += 256) A[e] = 1e-3 * (e % 7);
	for (int e = tid; e < KB * LD; e += 256) B[e] = 1e-3 * (e % 5);
	__syncthreads();
	const int g = lane >> 2, q = lane & 3; // fragment coordinates of m8n8k4
	for (int r = 0; r < reps; ++r) {
		const double a = -A[(8 * w + g) * KB + q]; // A fragment: row g, column q of the strip
#pragma unroll
		for (int tile = 0; tile < 8; ++tile) {
			const double b = B[q * LD + 8 * tile + g]; // B fragment: row q, column g of the tile
			double c0 = C[(8 * w + g) * LD + 8 * tile + 2 * q], c1 = C[(8 * w + g) * LD + 8 * tile + 2 * q + 1];
			dmma(c0, c1, a, b);
			C[(8 * w + g) * LD + 8 * tile + 2 * q] = c0;
			C[(8 * w + g) * LD + 8 * tile + 2 * q + 1] = c1;
		}
		__syncthreads();
	}
	if (tid == 0) out[blockIdx.x] = C[5 * LD + 7];
}

template <class L>
static float best_ms(L launch, int reps = 5)
{
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	launch();
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < reps; ++r) {
		CK(cudaEventRecord(e0));
		launch();
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (ms < best) best = ms;
	}
	CK(cudaGetLastError());
	return best;
}

int main(int argc, char **argv)
{
	const bool only_updates = argc > 1; // any argument: just the two update kernels (for ncu)
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	double *sink;
	CK(cudaMalloc(&sink, sizeof(double) * 4096));
	const int block = 256, grid = sms * 8;
	if (!only_updates) {
		const int iters = 512;
		const double fmas = 64.0 * iters * (double)block * grid;
		float t0 = best_ms([&] { issue_kernel<0><<<grid, block>>>(sink, iters, 1.0000001, 1e-9, 0x9e3779b9u, 1.0001f); });
		float t1 = best_ms([&] { issue_kernel<1><<<grid, block>>>(sink, iters, 1.0000001, 1e-9, 0x9e3779b9u, 1.0001f); });
		float t2 = best_ms([&] { issue_kernel<2><<<grid, block>>>(sink, iters, 1.0000001, 1e-9, 0x9e3779b9u, 1.0001f); });
		float t3 = best_ms([&] { issue_kernel<3><<<grid, block>>>(sink, iters, 1.0000001, 1e-9, 0x9e3779b9u, 1.0001f); });
		float t4 = best_ms([&] { issue_kernel<4><<<grid, block>>>(sink, iters, 1.0000001, 1e-9, 0x9e3779b9u, 1.0001f); });
		float t5 = best_ms([&] { issue_kernel<5><<<grid, block>>>(sink, iters, 1.0000001, 1e-9, 0x9e3779b9u, 1.0001f); });
		std::printf("{\"experiment\": \"issue_model\", \"dfma_only_ms\": %.4f, \"dfma_plus_int_ms\": %.4f, \"dfma_plus_ffma_ms\": %.4f, "
		            "\"dfma_only_tflops\": %.2f, \"int_over_dfma\": %.3f, \"ffma_over_dfma\": %.3f, \"imad_over_dfma\": %.3f, "
		            "\"lop3_over_dfma\": %.3f, \"lds_iadd_over_dfma\": %.3f}\n",
		            t0, t1, t2, 2 * fmas / (t0 * 1e-3) / 1e12, t1 / t0, t2 / t0, t3 / t0, t4 / t0, t5 / t0);
		float td = best_ms([&] { dmma_kernel<<<grid, block>>>(sink, iters, 1.0000001, 1e-9); });
		const double dflops = 2.0 * 256.0 * 64.0 * iters * (block / 32.0) * grid; // 8 x 8 x 4 FMAs per warp-instruction
		std::printf("{\"experiment\": \"dmma_rate\", \"dmma_ms\": %.4f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"dmma_over_dfma\": %.3f}\n", td,
		            dflops / (td * 1e-3) / 1e12, 2 * fmas / (t0 * 1e-3) / 1e12, (dflops / td) / (2 * fmas / t0));
	}
	const int reps = 20000, ugrid = sms;
	float tf = best_ms([&] { update_fma_kernel<<<ugrid, 256>>>(sink, reps); }, 3);
	float tm = best_ms([&] { update_dmma_kernel<<<ugrid, 256>>>(sink, reps); }, 3);
	const double uflops = 2.0 * NB * NB * KB * (double)reps * ugrid;
	std::printf("{\"experiment\": \"rank4_update_64x64_smem\", \"fma_us_per_update\": %.4f, \"dmma_us_per_update\": %.4f, \"fma_tflops\": %.2f, "
	            "\"dmma_tflops\": %.2f, \"dmma_over_fma_time\": %.3f}\n",
	            tf * 1e3 / reps, tm * 1e3 / reps, uflops / (tf * 1e-3) / 1e12, uflops / (tm * 1e-3) / 1e12, tm / tf);
	CK(cudaFree(sink));
	return 0;
}
