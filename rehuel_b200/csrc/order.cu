// Parameter-locality hand-out order for ensembles.
//
// The thread-per-system kernel (irk_thread.cuh) runs 32 members side by side in a warp; every
// trip of the warp costs as much as its slowest lane (most Newton iterations), and lanes whose
// members finish at different times refill one by one.  Members with nearby parameters take
// nearly the same steps, so handing the queue out in a space-filling-curve order of the
// parameter vectors keeps the lanes of a warp in lockstep (measured on config 2: 3.00 -> 2.44 ms,
// results bit-identical -- the order only changes WHICH lane integrates a member).
//
// The reference has nothing comparable (one system per call, irk.hpp:884-898); this is scheduling
// of the batched overload.  Three small kernels + one radix sort of 32-bit (key, index) pairs:
//   1. per-parameter min / max over the members (order-preserving integer image of the doubles)
//   2. Morton key: D = min(p, 3) coordinates, floor(21 / D) bits each, interleaved.  A strictly
//      positive parameter is mapped through its IEEE bit pattern (monotone, piecewise linear in
//      log2), so decade sweeps (rate constants, stiffness parameters) spread evenly; a parameter
//      that changes sign is mapped linearly.
//   3. cub::DeviceRadixSort::SortPairs over the key bits actually used (plumbing, like NCCL).
#include <cuda_runtime.h>

#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/rehuel_b200.h"
#include "launch.cuh"

namespace rehuel {

constexpr int kOrderMaxDims = 3;
constexpr int kOrderKeyBits = 21;

// order-preserving map double -> u64 (negative values below positive ones)
__device__ __forceinline__ unsigned long long sortable_bits(double x)
{
	const unsigned long long b = (unsigned long long)__double_as_longlong(x);
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_sortable_bits(unsigned long long s)
{
	const unsigned long long b = (s >> 63) ? (s & 0x7fffffffffffffffull) : ~s;
	return __longlong_as_double((long long)b);
}

// minmax[d] = min, minmax[kOrderMaxDims + d] = max of parameter d, as sortable bits
__global__ void __launch_bounds__(256) order_minmax_kernel(const double *__restrict__ params, size_t M, int dims,
                                                          unsigned long long *__restrict__ minmax)
{
	unsigned long long lo[kOrderMaxDims], hi[kOrderMaxDims];
#pragma unroll
	for (int d = 0; d < kOrderMaxDims; ++d) {
		lo[d] = ~0ull;
		hi[d] = 0ull;
	}
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
		for (int d = 0; d < kOrderMaxDims; ++d)
			if (d < dims) {
				const double v = params[d * M + i];
				if (v == v) { // NaN parameters do not take part
					const unsigned long long s = sortable_bits(v);
					lo[d] = s < lo[d] ? s : lo[d];
					hi[d] = s > hi[d] ? s : hi[d];
				}
			}
	}
	// warp -> block -> one atomic per block and slot (thousands of same-address atomics would serialise)
	__shared__ unsigned long long s_lo[8][kOrderMaxDims], s_hi[8][kOrderMaxDims];
#pragma unroll
	for (int d = 0; d < kOrderMaxDims; ++d) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const unsigned long long a = __shfl_down_sync(0xffffffffu, lo[d], o);
			const unsigned long long b = __shfl_down_sync(0xffffffffu, hi[d], o);
			lo[d] = a < lo[d] ? a : lo[d];
			hi[d] = b > hi[d] ? b : hi[d];
		}
		if ((threadIdx.x & 31) == 0) {
			s_lo[threadIdx.x >> 5][d] = lo[d];
			s_hi[threadIdx.x >> 5][d] = hi[d];
		}
	}
	__syncthreads();
	if (threadIdx.x < (unsigned)dims) {
		const int d = threadIdx.x;
		unsigned long long a = s_lo[0][d], b = s_hi[0][d];
		for (int w = 1; w < 8; ++w) {
			a = s_lo[w][d] < a ? s_lo[w][d] : a;
			b = s_hi[w][d] > b ? s_hi[w][d] : b;
		}
		atomicMin(minmax + d, a);
		atomicMax(minmax + kOrderMaxDims + d, b);
	}
}

__global__ void __launch_bounds__(256) order_key_kernel(const double *__restrict__ params, size_t M, int dims, int bits,
                                                       int descending, const unsigned long long *__restrict__ minmax,
                                                       unsigned *__restrict__ keys, unsigned *__restrict__ idx)
{
	double lo[kOrderMaxDims], scale[kOrderMaxDims];
	bool logmap[kOrderMaxDims];
	const double cells = (double)(1u << bits);
#pragma unroll
	for (int d = 0; d < kOrderMaxDims; ++d) {
		lo[d] = 0.0;
		scale[d] = 0.0;
		logmap[d] = false;
		if (d < dims) {
			const double a = from_sortable_bits(minmax[d]), b = from_sortable_bits(minmax[kOrderMaxDims + d]);
			logmap[d] = a > 0.0; // strictly positive sweep: use the bit pattern (~ log2)
			const double fa = logmap[d] ? (double)__double_as_longlong(a) : a;
			const double fb = logmap[d] ? (double)__double_as_longlong(b) : b;
			lo[d] = fa;
			scale[d] = (fb > fa) ? cells / (fb - fa) : 0.0;
		}
	}
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < M; i += (size_t)gridDim.x * blockDim.x) {
		unsigned q[kOrderMaxDims];
#pragma unroll
		for (int d = 0; d < kOrderMaxDims; ++d) {
			q[d] = 0u;
			if (d < dims) {
				const double v = params[d * M + i];
				const double f = logmap[d] ? (double)__double_as_longlong(v) : v;
				double x = (f - lo[d]) * scale[d];
				x = (x == x) ? x : 0.0;
				x = x < 0.0 ? 0.0 : (x > cells - 1.0 ? cells - 1.0 : x);
				q[d] = (unsigned)x;
			}
		}
		unsigned key = 0u;
		for (int b = 0; b < bits; ++b)
#pragma unroll
			for (int d = 0; d < kOrderMaxDims; ++d)
				if (d < dims) key |= ((q[d] >> b) & 1u) << (b * dims + d);
		if (descending) key = ((1u << (bits * dims)) - 1u) - key;
		keys[i] = key;
		idx[i] = (unsigned)i;
	}
}

static size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

// cub's own scratch need (asks the current device; 0 when there is none)
static size_t sort_temp_bytes(size_t M)
{
	size_t tmp = 0;
	if (cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr,
	                                    (unsigned *)nullptr, (int)M, 0, kOrderKeyBits) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return tmp ? tmp : 1;
}

size_t locality_order_workspace(size_t M)
{
	if (M == 0 || M > 0x7fffffffull) return 0;
	const size_t tmp = sort_temp_bytes(M);
	if (!tmp) {
		set_error(REHUEL_ECUDA, "locality order: no CUDA device to size the sort for");
		return 0;
	}
	// minmax | keys in | keys out | idx in | cub temp
	return 256 + 3 * align256(sizeof(unsigned) * M) + align256(tmp) + 256;
}

int locality_order(const double *params, int p, size_t M, int descending, unsigned *order, void *workspace,
                   size_t workspace_bytes, cudaStream_t stream)
{
	if (!params || p < 1 || !order || !workspace)
		return set_error(REHUEL_EINVAL, "locality order: params (p >= 1), order and workspace are required");
	if (M == 0) return REHUEL_OK;
	if (M > 0x7fffffffull) return set_error(REHUEL_EINVAL, "locality order: at most 2^31-1 members per call");
	const size_t need = locality_order_workspace(M);
	if (!need) return REHUEL_ECUDA;
	if (workspace_bytes < need)
		return set_error(REHUEL_EINVAL, "locality order: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
	const int dims = p < kOrderMaxDims ? p : kOrderMaxDims;
	int bits = kOrderKeyBits / dims;
	if (const char *e = std::getenv("REHUEL_ORDER_BITS")) { // tuning aid: total key bits
		const int v = std::atoi(e) / dims;
		if (v >= 1 && v <= bits) bits = v;
	}
	char *w = static_cast<char *>(workspace);
	unsigned long long *minmax = reinterpret_cast<unsigned long long *>(w);
	w += 256;
	unsigned *keys_in = reinterpret_cast<unsigned *>(w);
	w += align256(sizeof(unsigned) * M);
	unsigned *keys_out = reinterpret_cast<unsigned *>(w);
	w += align256(sizeof(unsigned) * M);
	unsigned *idx_in = reinterpret_cast<unsigned *>(w);
	w += align256(sizeof(unsigned) * M);
	size_t tmp = sort_temp_bytes(M);
	// min slots start at all-ones, max slots at zero
	REHUEL_CUDA_TRY(cudaMemsetAsync(minmax, 0xff, sizeof(unsigned long long) * kOrderMaxDims, stream));
	REHUEL_CUDA_TRY(cudaMemsetAsync(minmax + kOrderMaxDims, 0, sizeof(unsigned long long) * kOrderMaxDims, stream));
	int dev = 0, sms = 0;
	REHUEL_CUDA_TRY(cudaGetDevice(&dev));
	REHUEL_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	const unsigned long long want = (M + 255) / 256;
	const unsigned grid = (unsigned)(want < (unsigned long long)sms * 8 ? want : (unsigned long long)sms * 8);
	const unsigned grid_mm = grid; // one atomic pair per block and parameter: a few thousand in all
	order_minmax_kernel<<<grid_mm, 256, 0, stream>>>(params, M, dims, minmax);
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	order_key_kernel<<<grid, 256, 0, stream>>>(params, M, dims, bits, descending, minmax, keys_in, idx_in);
	REHUEL_CUDA_TRY(cudaGetLastError());
	count_launch();
	REHUEL_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w, tmp, (const unsigned *)keys_in, keys_out, (const unsigned *)idx_in, order,
	                                                (int)M, 0, bits * dims, stream));
	return REHUEL_OK;
}

} // namespace rehuel

extern "C" {

size_t rehuel_locality_order_workspace(size_t M) { return rehuel::locality_order_workspace(M); }

int rehuel_locality_order(const double *params, int p, size_t M, int descending, unsigned *order, void *workspace,
                          size_t workspace_bytes, void *cuda_stream)
{
	return rehuel::locality_order(params, p, M, descending, order, workspace, workspace_bytes,
	                              static_cast<cudaStream_t>(cuda_stream));
}
}
