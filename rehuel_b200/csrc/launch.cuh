// Kernel launch templates + the host-side entry points that a C++ user functor needs.
//
// The library instantiates these for its built-in functors (capi.cu).  A user-written functor
// (the reference's functor.hpp:35-73 concept as a __device__ callable, see functors.cuh for the
// shape) is compiled by nvcc in the USER's translation unit through include/rehuel_b200/irk.cuh,
// which instantiates dispatch_method<F> there and hands its address to ensemble_host() exported
// by librehuel_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

#include "../../include/rehuel_b200.h"
#include "irk_cta.cuh"
#include "irk_thread.cuh"
#include "irk_warp.cuh"
#include "irk_wdense.cuh"
#include "irk_wseg.cuh"
#include "tableau.hpp"

namespace rehuel {

// ---- exported by librehuel_b200.so (capi.cu)
int set_error(int code, const char *fmt, ...); // records the thread's last error text, returns code
void count_launch();                            // bumps rehuel_kernel_launches()
int debug_force_pivot();                        // rehuel_debug_force_pivot()'s switch
int retain_async_pool(int dev);                 // once per device: the stream-ordered pool keeps freed blocks

struct KernelInfo {
	int regs, local_bytes, smem_bytes, ctas_per_sm, block;
};
// Launches (info == nullptr) or describes (info != nullptr) the kernel for one tableau.
using LaunchFn = int (*)(const KernelArgs &, const Tableau &, cudaStream_t, KernelInfo *);

// Batched irk::odeint with HOST buffers for an arbitrary functor (n equations, p parameters);
// same semantics as rehuel_irk_ensemble (include/rehuel_b200.h).
int ensemble_host(LaunchFn launch, int n, int p, int method, double t0, double t1, double dt0, size_t M,
                  const double *y0, int y0_broadcast, const double *params, const rehuel_options *opts,
                  size_t n_samples_cap, int n_gpus, rehuel_result *out);
// The same, continuing from per-member start times and controller state (host arrays [M], [REHUEL_NSTATE][M]; NULL = fresh start).
int ensemble_host_from(LaunchFn launch, int n, int p, int method, double t0, double t1, double dt0, size_t M,
                       const double *y0, int y0_broadcast, const double *params, const rehuel_options *opts,
                       size_t n_samples_cap, int n_gpus, const double *t_start, const double *state_in, rehuel_result *out);
// Device-resident variant; same semantics as rehuel_irk_ensemble_device.
int ensemble_device(LaunchFn launch, int n, int p, int method, double t0, double t1, double dt0, size_t M,
                    const rehuel_options *opts, const rehuel_device_io *io, cudaStream_t stream);

// Parameter-locality hand-out order (order.cu); same semantics as rehuel_locality_order.
size_t locality_order_workspace(size_t M);
int locality_order(const double *params, int p, size_t M, int descending, unsigned *order, void *workspace,
                   size_t workspace_bytes, cudaStream_t stream);

#define REHUEL_CUDA_TRY(expr)                                                                              \
	do {                                                                                                   \
		cudaError_t e_ = (expr);                                                                           \
		if (e_ != cudaSuccess)                                                                             \
			return ::rehuel::set_error(e_ == cudaErrorMemoryAllocation ? REHUEL_ENOMEM : REHUEL_ECUDA,     \
			                           "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

// work-queue cursor and trace length start at zero for every launch
inline cudaError_t reset_cursors(const KernelArgs &ka, cudaStream_t stream)
{
	cudaError_t e = cudaMemsetAsync(ka.io.queue, 0, sizeof(unsigned long long), stream);
	if (e == cudaSuccess && ka.io.trace_len) e = cudaMemsetAsync(ka.io.trace_len, 0, sizeof(unsigned), stream);
	return e;
}

#ifndef REHUEL_CTA_SMALL_THREADS
#define REHUEL_CTA_SMALL_THREADS 128
#endif
#ifndef REHUEL_BLOCK
#define REHUEL_BLOCK 128
#endif
#ifndef REHUEL_MINB
#define REHUEL_MINB 4
#endif

// Launch (or, when `info` is given, only describe) the thread-per-system kernel.
template <class F, int S, int NR, int NC, int EST>
int launch_thread(const KernelArgs &ka, const Tableau &tb, cudaStream_t stream, KernelInfo *info)
{
	// 128 registers/thread (4 CTAs of 128 per SM) is the measured optimum for stage systems up to
	// 3x3 (a few spills, 16 warps/SM); larger stage systems keep 255 registers.
	constexpr int BLOCK = REHUEL_BLOCK, MIN_BLOCKS = (S * F::N <= 9) ? REHUEL_MINB : 2;
	void (*kern)(const KernelArgs, const DevTableau<S, NC>) = ensemble_irk_thread<F, S, NR, NC, EST, BLOCK, MIN_BLOCKS, false>;
	if (ka.io.rtol) { // per-member tolerances: separate instantiation so the uniform kernels stay lean
		kern = ensemble_irk_thread<F, S, NR, NC, EST, BLOCK, MIN_BLOCKS, true>;
	}
	if (ka.io.n_dense) { // dense output: separate instantiation (the stage values stay live to the end of the attempt)
		if (ka.io.rtol) return set_error(REHUEL_EINVAL, "dense output and per-member tolerances cannot be combined");
		if constexpr (EST != EST_IDENTITY) { // exactly the tableaux that carry b_interp
			kern = ensemble_irk_thread<F, S, NR, NC, EST, BLOCK, MIN_BLOCKS, false, true>;
		} else {
			return set_error(REHUEL_EINVAL, "Chosen method does not have dense output!"); // irk.cpp:733
		}
	}
	constexpr int MO = static_min_order<S, NR, NC, EST>::value;
	if (tb.s != S || tb.n_real != NR || tb.n_cplx != NC || tb.est_mode != EST || (REHUEL_CANONICAL_TABLEAU && !tb.canonical) ||
	    (MO != 0 && tb.has_b2 && MO != (tb.order < tb.order2 ? tb.order : tb.order2)))
		return set_error(REHUEL_EINVAL, "internal: tableau %s does not have the structure its kernel was built for", tb.name);
	int dev = 0, sms = 0, per_sm = 0;
	REHUEL_CUDA_TRY(cudaGetDevice(&dev));
	{
		// SM count and occupancy per (kernel instantiation, device), queried once: three driver calls per launch
		// are noticeable when a host call issues 8 chunk launches back to back.  Benign race: same values.
		static int cache_sms[3][64], cache_per_sm[3][64];
		const int v = ka.io.rtol ? 1 : (ka.io.n_dense ? 2 : 0);
		if (dev < 64 && cache_per_sm[v][dev] > 0) {
			sms = cache_sms[v][dev];
			per_sm = cache_per_sm[v][dev];
		} else {
			REHUEL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
			REHUEL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, 0));
			if (dev < 64 && per_sm > 0) {
				cache_sms[v][dev] = sms;
				cache_per_sm[v][dev] = per_sm;
			}
		}
	}
	if (per_sm < 1) return set_error(REHUEL_ECUDA, "kernel does not fit on an SM");
	if (info) {
		cudaFuncAttributes fa;
		REHUEL_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
		info->regs = fa.numRegs;
		info->local_bytes = (int)fa.localSizeBytes;
		info->smem_bytes = (int)fa.sharedSizeBytes;
		info->ctas_per_sm = per_sm;
		info->block = BLOCK;
		return REHUEL_OK;
	}
	if (ka.M == 0) return REHUEL_OK;
	// persistent grid: one wave of resident CTAs, lanes pull members from io.queue
	unsigned long long want = (ka.M + BLOCK - 1) / BLOCK;
	unsigned long long cap = (unsigned long long)sms * per_sm;
	unsigned grid = (unsigned)(want < cap ? want : cap);
	DevTableau<S, NC> dtb = make_dev_tableau<S, NC>(tb);
	REHUEL_CUDA_TRY(reset_cursors(ka, stream));
	KernelArgs k = ka;
	if (k.batch_lanes <= 0) { // gangs of 32, unless the launch is so small that whole gangs quantise it (irk_thread.cuh: kBatchLanes)
		const unsigned long long lanes = cap * BLOCK;
		const bool whole_gangs = ka.M % lanes == 0; // the host layer's chunks: every lane gets the same number of members
		k.batch_lanes = (whole_gangs || ka.M >= (unsigned long long)kGangMinPerLane * lanes) ? kBatchLanes : kSmallBatchLanes;
	}
	kern<<<grid, BLOCK, 0, stream>>>(k, dtb);
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	return REHUEL_OK;
}

// ---- diagnostic: evaluate a functor's f(t,y) and J(t,y) on the device (rehuel_eval_functor)
using EvalFn = int (*)(double t, const double *y, const double *p, double *f, double *J /* row-major n x n */);

template <class F>
__global__ void eval_functor_kernel(double t, const double *y, const double *p, double *f, double *J)
{
	constexpr int N = F::N, PA = F::P > 0 ? F::P : 1;
	double yy[N], pp[PA], ff[N], JJ[N][N];
	for (int k = 0; k < N; ++k) yy[k] = y[k];
	for (int k = 0; k < PA; ++k) pp[k] = (k < F::P) ? p[k] : 0.0;
	F::fun(t, yy, pp, ff);
	F::jac(t, yy, pp, JJ);
	for (int i = 0; i < N; ++i) {
		f[i] = ff[i];
		for (int j = 0; j < N; ++j) J[i * N + j] = JJ[i][j];
	}
}

template <class F>
__global__ void eval_functor_rows_kernel(double t, const double *y, const double *p, double *f, double *J)
{
	constexpr int N = F::N;
	for (int i = threadIdx.x; i < N; i += blockDim.x) {
		f[i] = F::fun(i, t, y, p);
		for (int j = 0; j < N; ++j) J[i * N + j] = 0.0;
		if constexpr (has_jac_band<F>::value) { // what the warp-per-system kernel calls
			constexpr int KB = jac_band<F>::value;
			double band[2 * KB + 1];
			F::jac_band(i, t, y, p, band);
			for (int o = -KB; o <= KB; ++o)
				if (i + o >= 0 && i + o < N) J[i * N + i + o] = band[KB + o];
		} else {
			F::jac_row(i, t, y, p, J + i * N);
		}
	}
}

// host pointers in, host pointers out (a test hook, not a fast path)
template <class F, bool COMPONENTWISE>
int eval_functor(double t, const double *y, const double *p, double *f, double *J)
{
	constexpr int N = F::N, PA = F::P > 0 ? F::P : 1;
	struct Scratch { // freed on every exit path
		double *d = nullptr;
		~Scratch() { cudaFree(d); }
	} sc;
	REHUEL_CUDA_TRY(cudaMalloc(&sc.d, sizeof(double) * (N + PA + N + N * N)));
	double *dy = sc.d, *dp = sc.d + N, *df = dp + PA, *dJ = df + N;
	REHUEL_CUDA_TRY(cudaMemcpy(dy, y, sizeof(double) * N, cudaMemcpyHostToDevice));
	if (F::P > 0) REHUEL_CUDA_TRY(cudaMemcpy(dp, p, sizeof(double) * F::P, cudaMemcpyHostToDevice));
	if constexpr (COMPONENTWISE) eval_functor_rows_kernel<F><<<1, 64>>>(t, dy, dp, df, dJ);
	else eval_functor_kernel<F><<<1, 1>>>(t, dy, dp, df, dJ);
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	REHUEL_CUDA_TRY(cudaMemcpy(f, df, sizeof(double) * N, cudaMemcpyDeviceToHost));
	REHUEL_CUDA_TRY(cudaMemcpy(J, dJ, sizeof(double) * N * N, cudaMemcpyDeviceToHost));
	return REHUEL_OK;
}

// Method id -> kernel instantiation (stage count and eigen-structure are compile-time).
template <class F>
int dispatch_method(const KernelArgs &ka, const Tableau &tb, cudaStream_t stream, KernelInfo *info)
{
	static_assert(F::N >= 1 && F::N <= 4, "thread-per-system kernels cover n <= 4 (larger systems: CTA-per-system path)");
	switch (tb.method) {
	case REHUEL_IMPLICIT_EULER: return launch_thread<F, 1, 1, 0, EST_IDENTITY>(ka, tb, stream, info);
	case REHUEL_RADAU_IIA_95: return launch_thread<F, 5, 1, 2, EST_REUSE>(ka, tb, stream, info);
	case REHUEL_RADAU_IIA_137: return launch_thread<F, 7, 1, 3, EST_REUSE>(ka, tb, stream, info);
	case REHUEL_RADAU_IIA_32: return launch_thread<F, 2, 0, 1, EST_OWN_LU>(ka, tb, stream, info);
	case REHUEL_RADAU_IIA_53: return launch_thread<F, 3, 1, 1, EST_REUSE>(ka, tb, stream, info);
	case REHUEL_LOBATTO_IIIC_43: return launch_thread<F, 3, 1, 1, EST_IDENTITY>(ka, tb, stream, info);
	case REHUEL_LOBATTO_IIIC_64: return launch_thread<F, 4, 0, 2, EST_IDENTITY>(ka, tb, stream, info);
	case REHUEL_LOBATTO_IIIC_85: return launch_thread<F, 5, 1, 2, EST_OWN_LU>(ka, tb, stream, info);
	}
	return set_error(REHUEL_EINVAL, "method %d (%s) has no GPU kernel", tb.method, tb.name);
}

// ---- CTA-per-system path (4 < n <= 64), irk_cta.cuh
template <class F, int S, int NR, int NC, int EST>
int launch_cta(const KernelArgs &ka, const Tableau &tb, cudaStream_t stream, KernelInfo *info)
{
	// up to n = 32 a member's matrices take a quarter of an SM's shared memory: smaller blocks, several per SM
	constexpr int THREADS = (F::N <= 32) ? REHUEL_CTA_SMALL_THREADS : 256;
	void (*kern)(const KernelArgs, const DevTableau<S, NC>) = ensemble_irk_cta<F, S, NR, NC, EST, THREADS, false>;
	if (ka.io.n_dense) {
		if constexpr (EST != EST_IDENTITY) {
			kern = ensemble_irk_cta<F, S, NR, NC, EST, THREADS, true>;
		} else {
			return set_error(REHUEL_EINVAL, "Chosen method does not have dense output!"); // irk.cpp:733
		}
	}
	const size_t smem = CtaSmem<F, S, NC>::bytes;
	int dev = 0, sms = 0, per_sm = 0;
	REHUEL_CUDA_TRY(cudaGetDevice(&dev));
	REHUEL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	int smem_max = 0;
	REHUEL_CUDA_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
	if (smem > (size_t)smem_max) // e.g. RADAU_IIA_137 at n = 64 with a dense Jacobian: one real + three complex 64 x 65 matrices
		return set_error(REHUEL_EINVAL, "the CTA-per-system kernel would need %zu bytes of shared memory for this method and system size "
		                                "(%d stages, n = %d); a block can have %d", smem, S, (int)F::N, smem_max);
	REHUEL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	REHUEL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
	if (per_sm < 1) return set_error(REHUEL_ECUDA, "CTA-per-system kernel needs %zu bytes of shared memory", smem);
	if (info) {
		cudaFuncAttributes fa;
		REHUEL_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
		info->regs = fa.numRegs;
		info->local_bytes = (int)fa.localSizeBytes;
		info->smem_bytes = (int)smem;
		info->ctas_per_sm = per_sm;
		info->block = THREADS;
		return REHUEL_OK;
	}
	if (ka.M == 0) return REHUEL_OK;
	unsigned long long cap = (unsigned long long)sms * per_sm;
	unsigned grid = (unsigned)(ka.M < cap ? ka.M : cap);
	DevTableau<S, NC> dtb = make_dev_tableau<S, NC>(tb);
	REHUEL_CUDA_TRY(reset_cursors(ka, stream));
	kern<<<grid, THREADS, smem, stream>>>(ka, dtb);
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	return REHUEL_OK;
}

// ---- warp-per-system path for banded Jacobians (4 < n <= 64, kBand <= 4), irk_warp.cuh
template <class F, int S, int NR, int NC, int EST>
int launch_warp(const KernelArgs &ka, const Tableau &tb, cudaStream_t stream, KernelInfo *info)
{
	void (*kern)(const KernelArgs, const DevTableau<S, NC>) = ensemble_irk_warp<F, S, NR, NC, EST, false>;
	if (ka.io.n_dense) {
		if constexpr (EST != EST_IDENTITY) {
			kern = ensemble_irk_warp<F, S, NR, NC, EST, true>;
		} else {
			return set_error(REHUEL_EINVAL, "Chosen method does not have dense output!"); // irk.cpp:733
		}
	}
	const size_t smem = WarpSmem<F, S, NC, EST>::bytes;
	if (tb.s != S || tb.n_real != NR || tb.n_cplx != NC || tb.est_mode != EST)
		return set_error(REHUEL_EINVAL, "internal: tableau %s does not have the structure its kernel was built for", tb.name);
	int dev = 0, sms = 0, per_sm = 0;
	REHUEL_CUDA_TRY(cudaGetDevice(&dev));
	REHUEL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	REHUEL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	REHUEL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
	if (per_sm < 1) return set_error(REHUEL_ECUDA, "warp-per-system kernel needs %zu bytes of shared memory", smem);
	if (info) {
		cudaFuncAttributes fa;
		REHUEL_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
		info->regs = fa.numRegs;
		info->local_bytes = (int)fa.localSizeBytes;
		info->smem_bytes = (int)smem;
		info->ctas_per_sm = per_sm;
		info->block = 32;
		return REHUEL_OK;
	}
	if (ka.M == 0) return REHUEL_OK;
	unsigned long long cap = (unsigned long long)sms * per_sm;
	unsigned grid = (unsigned)(ka.M < cap ? ka.M : cap);
	DevTableau<S, NC> dtb = make_dev_tableau<S, NC>(tb);
	REHUEL_CUDA_TRY(reset_cursors(ka, stream));
	// one slab per warp for the pivoted band fallback (irk_warp.cuh); stream-ordered, so concurrent launches on other
	// streams get their own and the block goes back to the (retained) pool as soon as this kernel is done
	KernelArgs k = ka;
	int rp = retain_async_pool(dev);
	if (rp) return rp;
	REHUEL_CUDA_TRY(cudaMallocAsync((void **)&k.pivot_scratch, WarpSmem<F, S, NC, EST>::slab_doubles * sizeof(double) * grid, stream));
	k.force_pivot = debug_force_pivot();
	kern<<<grid, 32, smem, stream>>>(k, dtb);
	const cudaError_t launched = cudaGetLastError(); // the slab goes back to the pool whether or not the launch worked
	REHUEL_CUDA_TRY(cudaFreeAsync(k.pivot_scratch, stream));
	REHUEL_CUDA_TRY(launched);
	count_launch();
	return REHUEL_OK;
}

// ---- warp-per-system path for small dense systems (4 < n <= 16), irk_wdense.cuh
template <class F, int S, int NR, int NC, int EST>
int launch_wdense(const KernelArgs &ka, const Tableau &tb, cudaStream_t stream, KernelInfo *info)
{
	void (*kern)(const KernelArgs, const DevTableau<S, NC>) = ensemble_irk_wdense<F, S, NR, NC, EST, false>;
	if (ka.io.n_dense) {
		if constexpr (EST != EST_IDENTITY) {
			kern = ensemble_irk_wdense<F, S, NR, NC, EST, true>;
		} else {
			return set_error(REHUEL_EINVAL, "Chosen method does not have dense output!"); // irk.cpp:733
		}
	}
	const size_t smem = WdSmem<F, S, NC, EST>::bytes;
	if (tb.s != S || tb.n_real != NR || tb.n_cplx != NC || tb.est_mode != EST)
		return set_error(REHUEL_EINVAL, "internal: tableau %s does not have the structure its kernel was built for", tb.name);
	int dev = 0, sms = 0, per_sm = 0;
	REHUEL_CUDA_TRY(cudaGetDevice(&dev));
	REHUEL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	REHUEL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
	if (per_sm < 1) return set_error(REHUEL_ECUDA, "warp-per-system kernel does not fit on an SM");
	if (info) {
		cudaFuncAttributes fa;
		REHUEL_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
		info->regs = fa.numRegs;
		info->local_bytes = (int)fa.localSizeBytes;
		info->smem_bytes = (int)smem;
		info->ctas_per_sm = per_sm;
		info->block = 32;
		return REHUEL_OK;
	}
	if (ka.M == 0) return REHUEL_OK;
	unsigned long long cap = (unsigned long long)sms * per_sm;
	unsigned grid = (unsigned)(ka.M < cap ? ka.M : cap);
	DevTableau<S, NC> dtb = make_dev_tableau<S, NC>(tb);
	REHUEL_CUDA_TRY(reset_cursors(ka, stream));
	kern<<<grid, 32, smem, stream>>>(ka, dtb);
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	return REHUEL_OK;
}

// ---- segment-per-system path for the smallest dense systems (4 < n <= 8, four members per warp), irk_wseg.cuh
template <class F, int S, int NR, int NC, int EST>
int launch_wseg(const KernelArgs &ka, const Tableau &tb, cudaStream_t stream, KernelInfo *info)
{
	void (*kern)(const KernelArgs, const DevTableau<S, NC>) = ensemble_irk_wseg<F, S, NR, NC, EST>;
	const size_t smem = WsSmem<F, S>::bytes;
	if (tb.s != S || tb.n_real != NR || tb.n_cplx != NC || tb.est_mode != EST)
		return set_error(REHUEL_EINVAL, "internal: tableau %s does not have the structure its kernel was built for", tb.name);
	int dev = 0, sms = 0, per_sm = 0;
	REHUEL_CUDA_TRY(cudaGetDevice(&dev));
	REHUEL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	REHUEL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
	if (per_sm < 1) return set_error(REHUEL_ECUDA, "segment-per-system kernel does not fit on an SM");
	if (info) {
		cudaFuncAttributes fa;
		REHUEL_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
		info->regs = fa.numRegs;
		info->local_bytes = (int)fa.localSizeBytes;
		info->smem_bytes = (int)smem;
		info->ctas_per_sm = per_sm;
		info->block = 32;
		return REHUEL_OK;
	}
	if (ka.M == 0) return REHUEL_OK;
	constexpr unsigned long long MPW = WsSmem<F, S>::MPW;
	const unsigned long long want = (ka.M + MPW - 1) / MPW, cap = (unsigned long long)sms * per_sm;
	unsigned grid = (unsigned)(want < cap ? want : cap);
	DevTableau<S, NC> dtb = make_dev_tableau<S, NC>(tb);
	REHUEL_CUDA_TRY(reset_cursors(ka, stream));
	kern<<<grid, 32, smem, stream>>>(ka, dtb);
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	return REHUEL_OK;
}

// 4 < n <= 64: banded functors (kBand <= 4) go warp-per-system with band factors in shared memory, small dense
// systems (n <= 16) warp-per-system with the matrices in registers (n <= 8: four members per warp), everything else
// CTA-per-system
template <class F, int S, int NR, int NC, int EST>
int launch_block(const KernelArgs &ka, const Tableau &tb, cudaStream_t stream, KernelInfo *info)
{
	constexpr int KB = jac_band<F>::value;
	if constexpr (KB <= 4 && F::N >= 2 * KB + 2) {
		return launch_warp<F, S, NR, NC, EST>(ka, tb, stream, info);
	} else if constexpr (F::N <= 8) {
		// four members per warp (two per warp at 8 < n <= 16 was measured and buys nothing: there the matrices of ONE
		// member already fill the warp in irk_wdense.cuh); dense output and stored samples are only built into the
		// one-member-per-warp kernel
		if (ka.io.n_dense || (ka.store_samples && ka.io.n_samples_cap)) return launch_wdense<F, S, NR, NC, EST>(ka, tb, stream, info);
		return launch_wseg<F, S, NR, NC, EST>(ka, tb, stream, info);
	} else if constexpr (F::N <= 16) {
		return launch_wdense<F, S, NR, NC, EST>(ka, tb, stream, info);
	} else {
		return launch_cta<F, S, NR, NC, EST>(ka, tb, stream, info);
	}
}

template <class F>
int dispatch_method_cta(const KernelArgs &ka, const Tableau &tb, cudaStream_t stream, KernelInfo *info)
{
	static_assert(F::N > 4 && F::N <= 64, "CTA-per-system kernels cover 4 < n <= 64");
	switch (tb.method) {
	case REHUEL_IMPLICIT_EULER: return launch_block<F, 1, 1, 0, EST_IDENTITY>(ka, tb, stream, info);
	case REHUEL_RADAU_IIA_32: return launch_block<F, 2, 0, 1, EST_OWN_LU>(ka, tb, stream, info);
	case REHUEL_RADAU_IIA_53: return launch_block<F, 3, 1, 1, EST_REUSE>(ka, tb, stream, info);
	case REHUEL_RADAU_IIA_95: return launch_block<F, 5, 1, 2, EST_REUSE>(ka, tb, stream, info);
	case REHUEL_RADAU_IIA_137: return launch_block<F, 7, 1, 3, EST_REUSE>(ka, tb, stream, info);
	case REHUEL_LOBATTO_IIIC_43: return launch_block<F, 3, 1, 1, EST_IDENTITY>(ka, tb, stream, info);
	case REHUEL_LOBATTO_IIIC_64: return launch_block<F, 4, 0, 2, EST_IDENTITY>(ka, tb, stream, info);
	case REHUEL_LOBATTO_IIIC_85: return launch_block<F, 5, 1, 2, EST_OWN_LU>(ka, tb, stream, info);
	}
	return set_error(REHUEL_EINVAL, "method %d (%s) has no CTA-per-system kernel", tb.method, tb.name);
}

} // namespace rehuel
