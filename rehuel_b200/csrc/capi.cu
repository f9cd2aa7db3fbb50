// Host layer + C ABI (include/rehuel_b200.h) of the batched ensemble integrator.
//
// Maps the reference's public surface for the implicit-RK path onto kernel launches:
//   irk::odeint (irk.hpp:884-898)            -> rehuel_irk_ensemble / rehuel_irk_single
//   irk::get_coefficients (irk.cpp:215-689)  -> rehuel_get_coefficients
//   irk::merge_rk_output (irk.cpp:750-783)   -> rehuel_result_merge
//   default option constructors              -> rehuel_default_options
// There is NO CPU fallback: every solve runs the sm_100a kernels or fails with REHUEL_ECUDA.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/rehuel_b200.h"
#include "irk_thread.cuh"
#include "launch.cuh"
#include "tableau.hpp"

namespace rehuel {

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
int set_error(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}
#define fail set_error
#define CUDA_TRY REHUEL_CUDA_TRY

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1); }

static std::atomic<int> g_force_pivot{0};
int debug_force_pivot() { return g_force_pivot.load(); }

// cudaMallocAsync's default pool gives freed blocks back to the driver at the next synchronisation unless told to keep
// them; launch_warp allocates its fallback slabs from it on every launch.
int retain_async_pool(int dev)
{
	static std::atomic<unsigned long long> done{0};
	if (dev >= 0 && dev < 64 && (done.load() >> dev & 1ull)) return REHUEL_OK;
	cudaMemPool_t pool;
	REHUEL_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
	unsigned long long keep = ~0ull;
	REHUEL_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
	if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev);
	return REHUEL_OK;
}

// ------------------------------------------------------------------ FP64 FMA peak probe
// Register-resident DFMA chains: 8 independent accumulators per thread, no memory traffic.
// Gives the achievable FP64 FMA rate of THIS device at its current clocks = the roofline
// denominator for the (compute-bound) integration kernels.
__global__ void __launch_bounds__(256) fp64_fma_peak_kernel(double *sink, int iters, double a, double b)
{
	double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
	       x7 = x0 + 7;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 8; ++u) {
			x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
			x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
		}
	}
	const double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
	if (r == 123.456) sink[0] = r; // never true; keeps the chains alive
}

// ------------------------------------------------------------------ problem registry
struct ProblemInfo {
	int id;
	const char *name;
	int n, p; // n == 0: size-generic (n = n_hint)
};
static const ProblemInfo kProblems[] = {
    {REHUEL_PROBLEM_ROBERTSON, "robertson", 3, 3},
    {REHUEL_PROBLEM_VAN_DER_POL, "van-der-pol", 2, 1},
    {REHUEL_PROBLEM_BRUSSELATOR_1D, "brusselator-1d", 0, 3},
    {REHUEL_PROBLEM_BRUSSELATOR_1D_DENSE, "brusselator-1d-dense", 0, 3},
    {REHUEL_PROBLEM_EXPONENTIAL, "exponential", 1, 1},
    {REHUEL_PROBLEM_STIFF_EQ, "stiff-equation", 2, 0},
    {REHUEL_PROBLEM_BRUSSELATOR, "brusselator", 2, 2},
    {REHUEL_PROBLEM_LORENZ, "lorenz", 3, 3},
    {REHUEL_PROBLEM_HARMONIC, "harmonic", 2, 1},
    {REHUEL_PROBLEM_DIMER, "dimer", 2, 1},
    {REHUEL_PROBLEM_KINETIC_4, "kinetic-4", 4, 3},
    {REHUEL_PROBLEM_THREE_BODY, "three-body", 12, 3},
};
static const ProblemInfo *find_problem(int id)
{
	for (const ProblemInfo &p : kProblems)
		if (p.id == id) return &p;
	return nullptr;
}

// ------------------------------------------------------------------ kernel dispatch
// The kernel templates (launch.cuh) are instantiated per functor in instances/inst_*.cu.
LaunchFn launcher_robertson();
LaunchFn launcher_van_der_pol();
LaunchFn launcher_exponential();
LaunchFn launcher_stiff_eq();
LaunchFn launcher_brusselator();
LaunchFn launcher_lorenz();
LaunchFn launcher_harmonic();
LaunchFn launcher_dimer();
LaunchFn launcher_kinetic4();
LaunchFn launcher_three_body();
LaunchFn launcher_bruss1d_64();
LaunchFn launcher_bruss1d_8();
LaunchFn launcher_bruss1d_32();
LaunchFn launcher_bruss1d_16();
LaunchFn launcher_bruss1d_12();
LaunchFn launcher_bruss1d_dense_64();
LaunchFn launcher_bruss1d_dense_8();
LaunchFn launcher_bruss1d_dense_32();
LaunchFn launcher_bruss1d_dense_16();
LaunchFn launcher_bruss1d_dense_12();

EvalFn evaluator_robertson();
EvalFn evaluator_van_der_pol();
EvalFn evaluator_exponential();
EvalFn evaluator_stiff_eq();
EvalFn evaluator_brusselator();
EvalFn evaluator_lorenz();
EvalFn evaluator_harmonic();
EvalFn evaluator_dimer();
EvalFn evaluator_kinetic4();
EvalFn evaluator_three_body();
EvalFn evaluator_bruss1d_64();
EvalFn evaluator_bruss1d_8();
EvalFn evaluator_bruss1d_32();
EvalFn evaluator_bruss1d_16();
EvalFn evaluator_bruss1d_12();
EvalFn evaluator_bruss1d_dense_64();
EvalFn evaluator_bruss1d_dense_8();
EvalFn evaluator_bruss1d_dense_32();
EvalFn evaluator_bruss1d_dense_16();
EvalFn evaluator_bruss1d_dense_12();

static LaunchFn find_launcher(int problem, int n)
{
	if (problem == REHUEL_PROBLEM_BRUSSELATOR_1D || problem == REHUEL_PROBLEM_BRUSSELATOR_1D_DENSE) {
		const bool dense = problem == REHUEL_PROBLEM_BRUSSELATOR_1D_DENSE;
		if (n == 64) return dense ? launcher_bruss1d_dense_64() : launcher_bruss1d_64();
		if (n == 32) return dense ? launcher_bruss1d_dense_32() : launcher_bruss1d_32();
		if (n == 16) return dense ? launcher_bruss1d_dense_16() : launcher_bruss1d_16();
		if (n == 12) return dense ? launcher_bruss1d_dense_12() : launcher_bruss1d_12();
		if (n == 8) return dense ? launcher_bruss1d_dense_8() : launcher_bruss1d_8();
		set_error(REHUEL_EINVAL, "brusselator-1d is built for n = 64, 32, 16, 12 and 8 (Ng = 32, 16, 8, 6, 4), got n = %d", n);
		return nullptr;
	}
	switch (problem) {
	case REHUEL_PROBLEM_ROBERTSON: return launcher_robertson();
	case REHUEL_PROBLEM_VAN_DER_POL: return launcher_van_der_pol();
	case REHUEL_PROBLEM_EXPONENTIAL: return launcher_exponential();
	case REHUEL_PROBLEM_STIFF_EQ: return launcher_stiff_eq();
	case REHUEL_PROBLEM_BRUSSELATOR: return launcher_brusselator();
	case REHUEL_PROBLEM_LORENZ: return launcher_lorenz();
	case REHUEL_PROBLEM_HARMONIC: return launcher_harmonic();
	case REHUEL_PROBLEM_DIMER: return launcher_dimer();
	case REHUEL_PROBLEM_KINETIC_4: return launcher_kinetic4();
	case REHUEL_PROBLEM_THREE_BODY: return launcher_three_body();
	}
	set_error(REHUEL_EINVAL, "problem %d has no GPU kernel", problem);
	return nullptr;
}

static EvalFn find_evaluator(int problem, int n)
{
	switch (problem) {
	case REHUEL_PROBLEM_ROBERTSON: return evaluator_robertson();
	case REHUEL_PROBLEM_VAN_DER_POL: return evaluator_van_der_pol();
	case REHUEL_PROBLEM_EXPONENTIAL: return evaluator_exponential();
	case REHUEL_PROBLEM_STIFF_EQ: return evaluator_stiff_eq();
	case REHUEL_PROBLEM_BRUSSELATOR: return evaluator_brusselator();
	case REHUEL_PROBLEM_LORENZ: return evaluator_lorenz();
	case REHUEL_PROBLEM_HARMONIC: return evaluator_harmonic();
	case REHUEL_PROBLEM_DIMER: return evaluator_dimer();
	case REHUEL_PROBLEM_KINETIC_4: return evaluator_kinetic4();
	case REHUEL_PROBLEM_THREE_BODY: return evaluator_three_body();
	case REHUEL_PROBLEM_BRUSSELATOR_1D: return n == 64 ? evaluator_bruss1d_64() : (n == 32 ? evaluator_bruss1d_32() : (n == 16 ? evaluator_bruss1d_16() : (n == 12 ? evaluator_bruss1d_12() : (n == 8 ? evaluator_bruss1d_8() : nullptr))));
	case REHUEL_PROBLEM_BRUSSELATOR_1D_DENSE:
		return n == 64 ? evaluator_bruss1d_dense_64()
		               : (n == 32 ? evaluator_bruss1d_dense_32() : (n == 16 ? evaluator_bruss1d_dense_16() : (n == 12 ? evaluator_bruss1d_dense_12() : (n == 8 ? evaluator_bruss1d_dense_8() : nullptr))));
	}
	return nullptr;
}

// Shared argument checking; fills the tableau and the scalar part of KernelArgs.
static int prepare(int method, double t0, double t1, double dt0, size_t M, const rehuel_options *o, Tableau &tb,
                   KernelArgs &ka)
{
	if (!o) return fail(REHUEL_EINVAL, "options must not be NULL (the reference asserts newton_opts, irk.hpp:548)");
	if (!get_coefficients(method, tb)) return fail(REHUEL_EINVAL, "Method %d not supported!", method); // irk.cpp:234-236
	if (!(dt0 > 0.0)) return fail(REHUEL_EINVAL, "Cannot use time step size <= 0!"); // irk.hpp:549
	if (o->newton_maxit < 1) return fail(REHUEL_EINVAL, "newton_maxit must be >= 1");
	if (o->newton_refresh_jac <= 0) return fail(REHUEL_EINVAL, "newton_refresh_jac must be > 0 (the reference divides by it, irk.hpp:485)");
	if (o->output_interval == 0) return fail(REHUEL_EINVAL, "output_interval must be > 0 (irk.hpp:825 takes step %% output_interval)");
	if (o->shard_mode != REHUEL_SHARD_CONTIGUOUS && o->shard_mode != REHUEL_SHARD_INTERLEAVED)
		return fail(REHUEL_EINVAL, "shard_mode must be REHUEL_SHARD_CONTIGUOUS or REHUEL_SHARD_INTERLEAVED");
	if (o->handout_order < REHUEL_ORDER_INDEX || o->handout_order > REHUEL_ORDER_LOCALITY_DESC)
		return fail(REHUEL_EINVAL, "handout_order must be REHUEL_ORDER_INDEX, _LOCALITY or _LOCALITY_DESC");
	std::memset(&ka, 0, sizeof ka);
	ka.t0 = t0;
	ka.t1 = t1;
	ka.dt0 = dt0;
	ka.rtol = o->rel_tol;
	ka.atol = o->abs_tol;
	ka.max_dt = o->max_dt;
	ka.max_steps = o->max_steps;
	// irk.hpp:890-895: adaptivity is switched off for tableaux without an embedded pair
	ka.adaptive = (o->adaptive_step_size != 0 && tb.has_b2) ? 1 : 0;
	ka.xtol2 = o->newton_dx_delta * o->newton_dx_delta;
	ka.Rtol2 = o->newton_tol * o->newton_tol;
	ka.maxit0 = o->newton_maxit;
	ka.refresh_jac = o->newton_refresh_jac;
	ka.store_samples = (o->output_mode & 1) ? 1 : 0;
	ka.output_interval = o->output_interval;
	ka.M = M;
	return REHUEL_OK;
}

// ------------------------------------------------------------------ host-buffer path
// Block cache for page-locked host memory and device memory: cudaMallocHost / cudaMalloc /
// cudaFree cost milliseconds (and synchronise the device), far more than the solve itself, so
// blocks are recycled across calls.  rehuel_release_cached_memory() returns them to the driver.
class BlockCache {
public:
	// kind < 0: page-locked host memory; kind >= 0: device memory of that device
	void *get(int kind, size_t bytes)
	{
		if (bytes == 0) bytes = 1;
		{
			std::lock_guard<std::mutex> g(mu_);
			size_t best = SIZE_MAX;
			for (size_t i = 0; i < blocks_.size(); ++i) {
				const Block &b = blocks_[i];
				if (b.used || b.kind != kind || b.bytes < bytes || b.bytes > 2 * bytes + (1u << 20)) continue;
				if (best == SIZE_MAX || b.bytes < blocks_[best].bytes) best = i;
			}
			if (best != SIZE_MAX) {
				blocks_[best].used = true;
				return blocks_[best].ptr;
			}
		}
		void *q = nullptr;
		cudaError_t e;
		if (kind < 0) {
			e = cudaMallocHost(&q, bytes);
		} else {
			int cur = 0;
			cudaGetDevice(&cur);
			if (cur != kind) cudaSetDevice(kind);
			e = cudaMalloc(&q, bytes);
			if (cur != kind) cudaSetDevice(cur);
		}
		if (e != cudaSuccess) {
			cudaGetLastError();
			trim(); // give cached blocks back and retry once
			e = (kind < 0) ? cudaMallocHost(&q, bytes) : cudaErrorMemoryAllocation;
			if (kind >= 0) {
				int cur = 0;
				cudaGetDevice(&cur);
				if (cur != kind) cudaSetDevice(kind);
				e = cudaMalloc(&q, bytes);
				if (cur != kind) cudaSetDevice(cur);
			}
			if (e != cudaSuccess) {
				cudaGetLastError();
				fail(REHUEL_ENOMEM, "%s(%zu bytes): %s", kind < 0 ? "cudaMallocHost" : "cudaMalloc", bytes, cudaGetErrorString(e));
				return nullptr;
			}
		}
		std::lock_guard<std::mutex> g(mu_);
		blocks_.push_back(Block{q, bytes, kind, true});
		return q;
	}
	void put(void *ptr)
	{
		if (!ptr) return;
		std::lock_guard<std::mutex> g(mu_);
		for (Block &b : blocks_)
			if (b.ptr == ptr) {
				b.used = false;
				return;
			}
	}
	void trim()
	{
		std::lock_guard<std::mutex> g(mu_);
		std::vector<Block> keep;
		for (Block &b : blocks_) {
			if (b.used) {
				keep.push_back(b);
				continue;
			}
			if (b.kind < 0) {
				cudaFreeHost(b.ptr);
			} else {
				int cur = 0;
				cudaGetDevice(&cur);
				cudaSetDevice(b.kind);
				cudaFree(b.ptr);
				cudaSetDevice(cur);
			}
		}
		blocks_.swap(keep);
	}

private:
	struct Block {
		void *ptr;
		size_t bytes;
		int kind;
		bool used;
	};
	std::mutex mu_;
	std::vector<Block> blocks_;
};
static BlockCache g_cache;

// Bump allocator over one cached slab (256-byte aligned sub-arrays).
struct Slab {
	char *base = nullptr;
	size_t size = 0, used = 0;
	template <class T>
	T *take(size_t count)
	{
		if (count == 0) return nullptr;
		size_t off = (used + 255) & ~size_t(255);
		used = off + count * sizeof(T);
		return (base && used <= size) ? reinterpret_cast<T *>(base + off) : nullptr;
	}
	static size_t need(size_t count, size_t elem) { return count ? ((count * elem + 255) & ~size_t(255)) + 256 : 0; }
};

struct ResultImpl {
	std::vector<void *> pinned; // cached page-locked slabs owned by the result
};

template <class T>
static int pinned_alloc(ResultImpl *impl, T **ptr, size_t count)
{
	*ptr = nullptr;
	if (count == 0) return REHUEL_OK;
	void *q = g_cache.get(-1, count * sizeof(T));
	if (!q) return REHUEL_ENOMEM;
	impl->pinned.push_back(q);
	*ptr = static_cast<T *>(q);
	return REHUEL_OK;
}

// A stream with its three timing events; recycled across calls like the memory blocks (creating and
// destroying 6-8 streams and 18-24 events per call costs more host time than issuing the work).
struct StreamSet {
	int dev = -1;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
};
class StreamPool {
public:
	int get(int dev, StreamSet &out) // the caller has made `dev` current
	{
		{
			std::lock_guard<std::mutex> g(mu_);
			for (size_t i = 0; i < free_.size(); ++i)
				if (free_[i].dev == dev) {
					out = free_[i];
					free_.erase(free_.begin() + (long)i);
					return REHUEL_OK;
				}
		}
		out = StreamSet();
		out.dev = dev;
		CUDA_TRY(cudaStreamCreateWithFlags(&out.stream, cudaStreamNonBlocking));
		CUDA_TRY(cudaEventCreate(&out.ev0));
		CUDA_TRY(cudaEventCreate(&out.ev1));
		CUDA_TRY(cudaEventCreate(&out.ev2));
		return REHUEL_OK;
	}
	void put(const StreamSet &s) // only after the stream has been synchronised
	{
		if (!s.stream) return;
		std::lock_guard<std::mutex> g(mu_);
		free_.push_back(s);
	}
	void trim()
	{
		std::lock_guard<std::mutex> g(mu_);
		int cur = 0;
		cudaGetDevice(&cur);
		for (StreamSet &s : free_) {
			cudaSetDevice(s.dev);
			if (s.ev0) cudaEventDestroy(s.ev0);
			if (s.ev1) cudaEventDestroy(s.ev1);
			if (s.ev2) cudaEventDestroy(s.ev2);
			if (s.stream) cudaStreamDestroy(s.stream);
		}
		free_.clear();
		cudaSetDevice(cur);
	}

private:
	std::mutex mu_;
	std::vector<StreamSet> free_;
};
static StreamPool g_streams;

// One unit of device work: a contiguous member range on one device with its own stream, so that
// the H2D copy, the kernel and the D2H copies of different chunks overlap.
struct DeviceShard {
	int dev = 0;
	size_t off = 0, m = 0;
	StreamSet ss;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
	void *slab = nullptr;
	rehuel_device_io io{};
	unsigned long long *d_stats = nullptr;
	bool synced = false;
	DeviceShard() = default;
	DeviceShard(const DeviceShard &) = delete;
	DeviceShard &operator=(const DeviceShard &) = delete;
	~DeviceShard()
	{
		if (stream && !synced) { // error path: nothing may still be using the slab or the stream
			cudaSetDevice(dev);
			cudaStreamSynchronize(stream);
		}
		g_cache.put(slab);
		g_streams.put(ss);
	}
};

static void free_result(rehuel_result *r)
{
	if (!r) return;
	ResultImpl *impl = static_cast<ResultImpl *>(r->impl);
	if (impl) {
		for (void *q : impl->pinned) g_cache.put(q);
		delete impl;
	}
	std::memset(r, 0, sizeof *r);
}

// Ensemble sums of the per-member counters (the role merge_rk_output's counter adds play in the reference,
// irk.cpp:773-780) are accumulated by the integration kernels themselves (rehuel_device_io.stats).
constexpr int kStatWords = 13;

int ensemble_device(LaunchFn launch, int n, int p, int method, double t0, double t1, double dt0, size_t M,
                    const rehuel_options *opts, const rehuel_device_io *io, cudaStream_t stream)
{
	(void)n;
	Tableau tb;
	KernelArgs ka;
	int rc = prepare(method, t0, t1, dt0, M, opts, tb, ka);
	if (rc) return rc;
	if (!io || !io->y0 || (p && !io->params) || !io->y || !io->t_final || !io->status || !io->queue)
		return fail(REHUEL_EINVAL, "rehuel_device_io: y0, params, y, t_final, status and queue are required");
	ka.io = *io;
	if (!ka.io.t_samples || !ka.io.y_samples || ka.io.n_samples_cap == 0) {
		ka.io.n_samples_cap = 0; // samples are counted but not stored
	}
	if (!ka.io.trace || !ka.io.trace_len) {
		ka.io.trace = nullptr;
		ka.io.trace_len = nullptr;
	}
	if (ka.io.n_samples_cap == 0 && !ka.io.n_samples) ka.store_samples = 0; // nobody looks at the sample count
	if (!ka.io.y_dense || ka.io.n_dense == 0) {
		ka.io.n_dense = 0;
		ka.io.y_dense = nullptr;
	} else if (!tb.has_interp) {
		return fail(REHUEL_EINVAL, "Chosen method does not have dense output!"); // irk.cpp:733
	} else if (!(ka.io.dense_dt > 0.0)) {
		return fail(REHUEL_EINVAL, "dense_dt must be > 0");
	}
	if ((ka.io.rtol || ka.io.atol || ka.io.newton_tol) && !(ka.io.rtol && ka.io.atol && ka.io.newton_tol))
		return fail(REHUEL_EINVAL, "per-member tolerances: rtol, atol and newton_tol must all be given");
	return launch(ka, tb, stream, nullptr);
}

int ensemble_host(LaunchFn launch, int n, int p, int method, double t0, double t1, double dt0, size_t M,
                  const double *y0, int y0_broadcast, const double *params, const rehuel_options *opts,
                  size_t n_samples_cap, int n_gpus, rehuel_result *out)
{
	return ensemble_host_from(launch, n, p, method, t0, t1, dt0, M, y0, y0_broadcast, params, opts, n_samples_cap, n_gpus,
	                          nullptr, nullptr, out);
}

// NaN fill for the dense-output block: grid times a member never reaches stay NaN
__global__ void __launch_bounds__(256) fill_nan_kernel(double *x, size_t count)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
		x[i] = __longlong_as_double(0x7ff8000000000000ll);
}

int ensemble_host_from(LaunchFn launch, int n, int p, int method, double t0, double t1, double dt0, size_t M,
                       const double *y0, int y0_broadcast, const double *params, const rehuel_options *opts,
                       size_t n_samples_cap, int n_gpus, const double *t_start, const double *state_in, rehuel_result *out)
{
	auto tic = std::chrono::steady_clock::now();
	if (!out) return fail(REHUEL_EINVAL, "out must not be NULL");
	std::memset(out, 0, sizeof *out);
	Tableau tb;
	KernelArgs ka;
	int rc = prepare(method, t0, t1, dt0, M, opts, tb, ka);
	if (rc) return rc;
	if (!y0 || (p && !params)) return fail(REHUEL_EINVAL, "y0/params must not be NULL");
	const size_t n_dense = (size_t)opts->dense_samples;
	const bool keep_state = opts->keep_state != 0;
	if (n_dense && !tb.has_interp) return fail(REHUEL_EINVAL, "Chosen method does not have dense output!"); // irk.cpp:733
	if (n_dense && !(opts->dense_dt > 0.0)) return fail(REHUEL_EINVAL, "dense_dt must be > 0");
	int ndev = 0;
	CUDA_TRY(cudaGetDeviceCount(&ndev));
	if (n_gpus < 1) n_gpus = 1;
	if (n_gpus > ndev) return fail(REHUEL_EINVAL, "n_gpus=%d but only %d CUDA devices are visible", n_gpus, ndev);
	if (!ka.store_samples) n_samples_cap = 0;
	int cur_dev = 0;
	CUDA_TRY(cudaGetDevice(&cur_dev));
	const int dev0 = (n_gpus == 1) ? cur_dev : 0;

	// ---- result storage: ONE cached page-locked slab (so that the D2H copies are asynchronous)
	ResultImpl *impl = new (std::nothrow) ResultImpl();
	if (!impl) return fail(REHUEL_ENOMEM, "out of host memory");
	out->impl = impl;
	out->M = M;
	out->n = n;
	out->n_samples_cap = n_samples_cap;
	const size_t max_shards = (size_t)n_gpus * 16;
	size_t hbytes = Slab::need(M, 8) + Slab::need((size_t)n * M, 8) + Slab::need(M, 8) + Slab::need(M, 4) +
	                Slab::need((size_t)REHUEL_NCOUNTERS * M, 4) + Slab::need(M, 4) +
	                Slab::need(max_shards * kStatWords, 8) + 3 * Slab::need(n_samples_cap * M, 8) * (size_t)(n > 1 ? n : 1) +
	                (keep_state ? Slab::need((size_t)REHUEL_NSTATE * M, 8) : 0) + Slab::need(n_dense * n * M, 8);
	Slab hs;
	hs.base = static_cast<char *>(g_cache.get(-1, hbytes));
	hs.size = hbytes;
	if (!hs.base) {
		free_result(out);
		return REHUEL_ENOMEM;
	}
	impl->pinned.push_back(hs.base);
	// one block of doubles [n + 2][M] (y rows, t_final, err_last) and one of 32-bit words [1 + 9][M] (status,
	// counters): each chunk comes back with TWO pitched copies instead of five
	double *hdbl = hs.take<double>((size_t)(n + 2) * M);
	int *hint = hs.take<int>((size_t)(1 + REHUEL_NCOUNTERS) * M);
	if (!hdbl || !hint) {
		free_result(out);
		return fail(REHUEL_ENOMEM, "internal: result slab too small");
	}
	out->y = hdbl;
	out->t_final = hdbl + (size_t)n * M;
	out->err_last = hdbl + (size_t)(n + 1) * M;
	out->status = hint;
	unsigned *counters = reinterpret_cast<unsigned *>(hint + M); // [NCOUNTERS][M]
	out->n_samples = n_samples_cap ? hs.take<unsigned>(M) : nullptr; // NULL: no samples were asked for
	unsigned long long *h_stats = hs.take<unsigned long long>(max_shards * kStatWords);
	std::memset(h_stats, 0, sizeof(unsigned long long) * max_shards * kStatWords);
	if (keep_state) out->state = hs.take<double>((size_t)REHUEL_NSTATE * M);
	if (n_dense) {
		out->y_dense = hs.take<double>(n_dense * n * M);
		out->n_dense = n_dense;
	}
	if (n_samples_cap) {
		out->t_samples = hs.take<double>(n_samples_cap * M);
		out->y_samples = hs.take<double>(n_samples_cap * n * M);
		out->err_samples = hs.take<double>(n_samples_cap * M);
	}
	out->attempt = counters + (size_t)REHUEL_CNT_ATTEMPT * M;
	out->reject_newton = counters + (size_t)REHUEL_CNT_REJECT_NEWTON * M;
	out->reject_err = counters + (size_t)REHUEL_CNT_REJECT_ERR * M;
	out->newton_success = counters + (size_t)REHUEL_CNT_NEWTON_SUCCESS * M;
	out->newton_maxit_exceed = counters + (size_t)REHUEL_CNT_NEWTON_MAXIT_EXCEED * M;
	out->fun_evals = counters + (size_t)REHUEL_CNT_FUN_EVALS * M;
	out->jac_evals = counters + (size_t)REHUEL_CNT_JAC_EVALS * M;
	out->newton_iters = counters + (size_t)REHUEL_CNT_NEWTON_ITERS * M;
	out->steps = counters + (size_t)REHUEL_CNT_STEPS * M;

	// ---- work units: GPU g owns the contiguous members [M*g/G, M*(g+1)/G); each GPU's share is cut
	// into up to 8 chunks on separate streams so that the H2D copy of chunk i+1 and the D2H copies
	// of chunk i-1 overlap the kernel of chunk i.  Chunk sizes taper off at the end (see `weight`).
	std::deque<DeviceShard> shards; // deque: elements are never moved (DeviceShard owns CUDA handles)
	// tuning aids: REHUEL_CHUNKS=k (chunks per GPU), REHUEL_CHUNK_WEIGHTS="1,2,4,4,2,1" (relative chunk sizes)
	long env_chunks = 0;
	if (const char *e = std::getenv("REHUEL_CHUNKS")) env_chunks = std::atol(e);
	size_t wts[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1}, n_wts = 0;
	if (const char *e = std::getenv("REHUEL_CHUNK_WEIGHTS")) {
		for (const char *q = e; *q && n_wts < 16;) {
			char *end = nullptr;
			const long v = std::strtol(q, &end, 10);
			if (end == q) break;
			wts[n_wts++] = v > 0 ? (size_t)v : 1;
			q = (*end == ',') ? end + 1 : end;
		}
	}
	// members [lo, hi) cut into chunks that are dealt to `owners` devices in turn (1 owner: all to dev_first)
	auto add_chunks = [&](size_t lo, size_t hi, int dev_first, int owners) {
		size_t chunks = (hi - lo) / owners / (128u << 10);
		chunks = chunks < 1 ? 1 : (chunks > 8 ? 8 : chunks); // per device; 8 equal chunks measured best at 2^20 members
		if (env_chunks >= 1 && env_chunks <= 16) chunks = (size_t)env_chunks;
		if (n_wts >= 1) chunks = n_wts;
		if (n_samples_cap) chunks = 1;
		const bool weighted = n_wts >= 1 && !n_samples_cap && owners == 1;
		chunks *= (size_t)owners;
		size_t wsum = 0, wacc = 0;
		for (size_t c = 0; c < chunks; ++c) wsum += weighted ? wts[c] : 1;
		// boundaries on multiples of 1024 members: every row of every chunk then starts 4 KiB-aligned on both
		// sides of the pitched copies (unaligned rows cost ~8 % of the whole call)
		auto cut = [&](size_t w) -> size_t {
			if (w == 0) return lo;
			if (w >= wsum) return hi;
			const size_t x = lo + (hi - lo) * w / wsum;
			const size_t r = (x + 512) & ~size_t(1023);
			return r < lo ? lo : (r > hi ? hi : r);
		};
		for (size_t c = 0; c < chunks; ++c) {
			shards.emplace_back();
			DeviceShard &sh = shards.back();
			sh.dev = dev_first + (int)(c % (size_t)owners);
			sh.off = cut(wacc);
			wacc += weighted ? wts[c] : 1;
			sh.m = cut(wacc) - sh.off;
		}
	};
	if (opts->shard_mode == REHUEL_SHARD_INTERLEAVED && n_gpus > 1) {
		add_chunks(0, M, dev0, n_gpus); // chunk k -> GPU k mod G: balances ensembles whose cost grows with the index
	} else {
		for (int g = 0; g < n_gpus; ++g) add_chunks(M * (size_t)g / (size_t)n_gpus, M * (size_t)(g + 1) / (size_t)n_gpus, dev0 + g, 1);
	}
	{ // the work units must tile [0, M) exactly: a gap would leave members silently unsolved
		size_t next = 0;
		for (const DeviceShard &sh : shards) {
			if (sh.off != next) {
				free_result(out);
				return fail(REHUEL_EINVAL, "internal: chunk layout leaves a gap at member %zu", next);
			}
			next += sh.m;
		}
		if (next != M) {
			free_result(out);
			return fail(REHUEL_EINVAL, "internal: chunk layout covers %zu of %zu members", next, M);
		}
	}
	std::vector<double> t_mark(shards.size() + 1, 0.0); // host time at which each shard's work was issued (REHUEL_TIMING)
	auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tic).count(); };
	auto body = [&]() -> int {
		t_mark[0] = now_ms();
		for (size_t si = 0; si < shards.size(); ++si) {
			DeviceShard &sh = shards[si];
			const size_t m = sh.m;
			t_mark[si + 1] = t_mark[si];
			if (m == 0) continue;
			CUDA_TRY(cudaSetDevice(sh.dev));
			if (int rs = g_streams.get(sh.dev, sh.ss)) return rs;
			sh.stream = sh.ss.stream;
			sh.ev0 = sh.ss.ev0;
			sh.ev1 = sh.ss.ev1;
			sh.ev2 = sh.ss.ev2;
			const size_t dbytes = Slab::need(y0_broadcast ? (size_t)n : (size_t)n * m, 8) + Slab::need((size_t)p * m, 8) +
			                      Slab::need((size_t)(n + 2) * m, 8) + Slab::need((size_t)(1 + REHUEL_NCOUNTERS) * m, 4) +
			                      Slab::need(m, 4) + Slab::need(1, 8) + Slab::need(kStatWords, 8) +
			                      3 * Slab::need(n_samples_cap * m, 8) * (size_t)(n > 1 ? n : 1) +
			                      (t_start ? Slab::need(m, 8) : 0) + (state_in ? Slab::need((size_t)REHUEL_NSTATE * m, 8) : 0) +
			                      (keep_state ? Slab::need((size_t)REHUEL_NSTATE * m, 8) : 0) + Slab::need(n_dense * n * m, 8);
			// hand-out order (order.cu): only pays where the lanes of a warp integrate different members
			const bool reorder = opts->handout_order != REHUEL_ORDER_INDEX && p > 0 && n <= 4 && m >= 4096;
			const size_t order_ws = reorder ? locality_order_workspace(m) : 0;
			if (reorder && !order_ws) return REHUEL_ECUDA;
			const size_t dbytes_all = dbytes + (reorder ? Slab::need(m, 4) + Slab::need(order_ws, 1) : 0);
			Slab ds;
			sh.slab = g_cache.get(sh.dev, dbytes_all);
			if (!sh.slab) return REHUEL_ENOMEM;
			ds.base = static_cast<char *>(sh.slab);
			ds.size = dbytes_all;
			double *d_y0 = ds.take<double>(y0_broadcast ? (size_t)n : (size_t)n * m);
			double *d_par = ds.take<double>((size_t)p * m);
			double *ddbl = ds.take<double>((size_t)(n + 2) * m); // same row order as the host block
			int *dint = ds.take<int>((size_t)(1 + REHUEL_NCOUNTERS) * m);
			sh.io.y = ddbl;
			sh.io.t_final = ddbl + (size_t)n * m;
			sh.io.err_last = ddbl + (size_t)(n + 1) * m;
			sh.io.status = dint;
			sh.io.counters = reinterpret_cast<unsigned *>(dint + m);
			sh.io.n_samples = ds.take<unsigned>(m);
			sh.io.queue = ds.take<unsigned long long>(1);
			sh.d_stats = ds.take<unsigned long long>(kStatWords);
			double *d_tstart = t_start ? ds.take<double>(m) : nullptr;
			double *d_state_in = state_in ? ds.take<double>((size_t)REHUEL_NSTATE * m) : nullptr;
			sh.io.state_out = keep_state ? ds.take<double>((size_t)REHUEL_NSTATE * m) : nullptr;
			sh.io.t_start = d_tstart;
			sh.io.state_in = d_state_in;
			if (n_dense) {
				sh.io.n_dense = n_dense;
				sh.io.dense_t0 = opts->dense_t0;
				sh.io.dense_dt = opts->dense_dt;
				sh.io.y_dense = ds.take<double>(n_dense * n * m);
			}
			if (n_samples_cap) {
				sh.io.t_samples = ds.take<double>(n_samples_cap * m);
				sh.io.y_samples = ds.take<double>(n_samples_cap * n * m);
				sh.io.err_samples = ds.take<double>(n_samples_cap * m);
			}
			sh.io.y0 = d_y0;
			sh.io.params = d_par;
			sh.io.y0_broadcast = y0_broadcast;
			sh.io.n_samples_cap = n_samples_cap;
			// H2D: SoA rows are strided by M on the host and by m on the device
			if (y0_broadcast) {
				CUDA_TRY(cudaMemcpyAsync(d_y0, y0, sizeof(double) * n, cudaMemcpyHostToDevice, sh.stream));
			} else {
				CUDA_TRY(cudaMemcpy2DAsync(d_y0, sizeof(double) * m, y0 + sh.off, sizeof(double) * M,
				                           sizeof(double) * m, n, cudaMemcpyHostToDevice, sh.stream));
			}
			if (p)
				CUDA_TRY(cudaMemcpy2DAsync(d_par, sizeof(double) * m, params + sh.off, sizeof(double) * M,
				                           sizeof(double) * m, p, cudaMemcpyHostToDevice, sh.stream));
			if (d_tstart) CUDA_TRY(cudaMemcpyAsync(d_tstart, t_start + sh.off, sizeof(double) * m, cudaMemcpyHostToDevice, sh.stream));
			if (d_state_in)
				CUDA_TRY(cudaMemcpy2DAsync(d_state_in, sizeof(double) * m, state_in + sh.off, sizeof(double) * M,
				                           sizeof(double) * m, REHUEL_NSTATE, cudaMemcpyHostToDevice, sh.stream));
			if (n_dense) {
				fill_nan_kernel<<<296, 256, 0, sh.stream>>>(sh.io.y_dense, n_dense * n * m);
				CUDA_TRY(cudaGetLastError());
				g_launches.fetch_add(1);
			}
			CUDA_TRY(cudaMemsetAsync(sh.d_stats, 0, sizeof(unsigned long long) * kStatWords, sh.stream));
			if (reorder) {
				unsigned *d_order = ds.take<unsigned>(m);
				void *ws = ds.take<char>(order_ws);
				int ro = locality_order(d_par, p, m, opts->handout_order == REHUEL_ORDER_LOCALITY_DESC, d_order, ws, order_ws, sh.stream);
				if (ro) return ro;
				sh.io.order = d_order;
			}
			sh.io.stats = sh.d_stats;
			KernelArgs k = ka;
			k.M = m;
			k.io = sh.io;
			if (n_samples_cap == 0) k.store_samples = 0; // final state only: skip the per-step sample bookkeeping
			CUDA_TRY(cudaEventRecord(sh.ev0, sh.stream));
			int rr = launch(k, tb, sh.stream, nullptr);
			if (rr) return rr;
			CUDA_TRY(cudaEventRecord(sh.ev1, sh.stream));
			CUDA_TRY(cudaMemcpyAsync(h_stats + si * kStatWords, sh.d_stats, sizeof(unsigned long long) * kStatWords,
			                         cudaMemcpyDeviceToHost, sh.stream));
			CUDA_TRY(cudaMemcpy2DAsync(hdbl + sh.off, sizeof(double) * M, ddbl, sizeof(double) * m, sizeof(double) * m,
			                           (size_t)n + 2, cudaMemcpyDeviceToHost, sh.stream));
			CUDA_TRY(cudaMemcpy2DAsync(hint + sh.off, sizeof(int) * M, dint, sizeof(int) * m, sizeof(int) * m,
			                           (size_t)1 + REHUEL_NCOUNTERS, cudaMemcpyDeviceToHost, sh.stream));
			if (keep_state)
				CUDA_TRY(cudaMemcpy2DAsync(out->state + sh.off, sizeof(double) * M, sh.io.state_out, sizeof(double) * m,
				                           sizeof(double) * m, REHUEL_NSTATE, cudaMemcpyDeviceToHost, sh.stream));
			if (n_dense)
				CUDA_TRY(cudaMemcpy2DAsync(out->y_dense + sh.off, sizeof(double) * M, sh.io.y_dense, sizeof(double) * m,
				                           sizeof(double) * m, n_dense * n, cudaMemcpyDeviceToHost, sh.stream));
			if (n_samples_cap)
				CUDA_TRY(cudaMemcpyAsync(out->n_samples + sh.off, sh.io.n_samples, sizeof(unsigned) * m, cudaMemcpyDeviceToHost, sh.stream));
			if (n_samples_cap) {
				CUDA_TRY(cudaMemcpy2DAsync(out->t_samples + sh.off, sizeof(double) * M, sh.io.t_samples, sizeof(double) * m,
				                           sizeof(double) * m, n_samples_cap, cudaMemcpyDeviceToHost, sh.stream));
				CUDA_TRY(cudaMemcpy2DAsync(out->err_samples + sh.off, sizeof(double) * M, sh.io.err_samples, sizeof(double) * m,
				                           sizeof(double) * m, n_samples_cap, cudaMemcpyDeviceToHost, sh.stream));
				CUDA_TRY(cudaMemcpy2DAsync(out->y_samples + sh.off, sizeof(double) * M, sh.io.y_samples, sizeof(double) * m,
				                           sizeof(double) * m, n_samples_cap * n, cudaMemcpyDeviceToHost, sh.stream));
			}
			t_mark[si + 1] = now_ms();
		}
		const bool timing = std::getenv("REHUEL_TIMING") != nullptr;
		const double t_issued = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tic).count();
		if (timing)
			for (DeviceShard &sh : shards)
				if (sh.m) {
					cudaSetDevice(sh.dev);
					cudaEventRecord(sh.ev2, sh.stream);
				}
		// wait; kernel_ms = per device, first kernel start -> last kernel end; max over devices
		std::vector<double> kms(n_gpus, 0.0);
		for (DeviceShard &sh : shards) {
			if (sh.m == 0) continue;
			CUDA_TRY(cudaSetDevice(sh.dev));
			CUDA_TRY(cudaStreamSynchronize(sh.stream));
			sh.synced = true;
		}
		for (int g = 0; g < n_gpus; ++g) {
			DeviceShard *first = nullptr;
			for (DeviceShard &sh : shards) {
				if (sh.dev != dev0 + g || sh.m == 0) continue;
				if (!first) first = &sh;
				float ms = 0.f;
				CUDA_TRY(cudaSetDevice(sh.dev));
				CUDA_TRY(cudaEventElapsedTime(&ms, first->ev0, sh.ev1));
				if (ms > kms[g]) kms[g] = ms;
			}
			if (kms[g] > out->kernel_ms) out->kernel_ms = kms[g];
		}
		if (timing) {
			const double t_synced = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tic).count();
			std::fprintf(stderr, "[rehuel timing] host: issue starts at %.3f ms; shards issued at", t_mark[0]);
			for (size_t si = 0; si < shards.size(); ++si) std::fprintf(stderr, " %.3f", t_mark[si + 1]);
			std::fprintf(stderr, " ms\n[rehuel timing] host: all work issued at %.3f ms, synced at %.3f ms\n", t_issued, t_synced);
			DeviceShard *first = nullptr;
			for (DeviceShard &sh : shards) {
				if (!sh.m) continue;
				if (!first) first = &sh;
				float a0 = 0, a1 = 0, a2 = 0;
				cudaEventElapsedTime(&a0, first->ev0, sh.ev0);
				cudaEventElapsedTime(&a1, first->ev0, sh.ev1);
				cudaEventElapsedTime(&a2, first->ev0, sh.ev2);
				std::fprintf(stderr, "[rehuel timing] shard off=%zu m=%zu: kernel %.3f..%.3f ms, copies done %.3f ms\n", sh.off, sh.m, a0, a1, a2);
			}
		}
		return REHUEL_OK;
	};
	rc = body();
	const size_t n_shards = shards.size();
	shards.clear(); // returns the device slabs to the cache
	cudaSetDevice(cur_dev);
	if (rc) {
		free_result(out);
		return rc;
	}

	// ---- ensemble statistics: sums of the device-side reductions
	rehuel_stats &s = out->sum;
	for (size_t si = 0; si < n_shards; ++si) {
		const unsigned long long *w = h_stats + si * kStatWords;
		s.attempt += w[REHUEL_CNT_ATTEMPT];
		s.reject_newton += w[REHUEL_CNT_REJECT_NEWTON];
		s.reject_err += w[REHUEL_CNT_REJECT_ERR];
		s.newton_success += w[REHUEL_CNT_NEWTON_SUCCESS];
		s.newton_maxit_exceed += w[REHUEL_CNT_NEWTON_MAXIT_EXCEED];
		s.fun_evals += w[REHUEL_CNT_FUN_EVALS];
		s.jac_evals += w[REHUEL_CNT_JAC_EVALS];
		s.newton_iters += w[REHUEL_CNT_NEWTON_ITERS];
		s.steps += w[REHUEL_CNT_STEPS];
		s.n_failed += w[REHUEL_NCOUNTERS];
		if (w[REHUEL_NCOUNTERS + 1] > s.max_attempt) s.max_attempt = w[REHUEL_NCOUNTERS + 1];
	}
	out->accept_frac = s.attempt ? (double)s.steps / (double)s.attempt : 0.0; // irk.hpp:864
	out->elapsed_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tic).count();
	return REHUEL_OK;
}

} // namespace rehuel

using namespace rehuel;

// =================================================================== C ABI
extern "C" {

const char *rehuel_last_error(void) { return g_err; }
const char *rehuel_version(void) { return "rehuel_b200 0.1 (sm_100a)"; }
unsigned long long rehuel_kernel_launches(void) { return g_launches.load(); }

int rehuel_fp64_peak(double *tflops, double *ms_out)
{
	if (!tflops) return fail(REHUEL_EINVAL, "tflops must not be NULL");
	int dev = 0, sms = 0;
	CUDA_TRY(cudaGetDevice(&dev));
	CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	double *sink = nullptr;
	CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
	cudaEvent_t e0, e1;
	CUDA_TRY(cudaEventCreate(&e0));
	CUDA_TRY(cudaEventCreate(&e1));
	const int iters = 4096, block = 256, grid = sms * 8;
	fp64_fma_peak_kernel<<<grid, block>>>(sink, iters / 8, 1.0000001, 1e-9); // warm-up
	float best = 1e30f;
	for (int rep = 0; rep < 5; ++rep) {
		CUDA_TRY(cudaEventRecord(e0));
		fp64_fma_peak_kernel<<<grid, block>>>(sink, iters, 1.0000001, 1e-9);
		CUDA_TRY(cudaEventRecord(e1));
		CUDA_TRY(cudaEventSynchronize(e1));
		float ms = 0.f;
		CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
		if (ms < best) best = ms;
	}
	CUDA_TRY(cudaGetLastError());
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(sink);
	const double flops = 2.0 * 64.0 * (double)iters * (double)block * (double)grid;
	*tflops = flops / (best * 1e-3) / 1e12;
	if (ms_out) *ms_out = best;
	return REHUEL_OK;
}

void rehuel_default_options(rehuel_options *o)
{
	if (!o) return;
	o->rel_tol = 1e-4;                 // options.hpp:50
	o->abs_tol = 10 * o->rel_tol;      // options.hpp:51
	o->max_dt = 0.0;                   // options.hpp:52
	o->max_steps = -1;                 // options.hpp:53
	o->adaptive_step_size = 1;         // irk.hpp:103
	o->out_interval = 0;               // options.hpp:55
	o->newton_tol = 1e-4;              // newton.hpp:58
	o->newton_dx_delta = 1e-4;         // newton.hpp:58
	o->newton_maxit = 5000;            // newton.hpp:58
	o->newton_refresh_jac = 10;        // newton.hpp:59
	o->output_mode = 1;                // options.hpp:103 STORE_IN_VECTORS
	o->output_interval = 1;            // options.hpp:104
	o->keep_state = 0;
	o->dense_samples = 0;
	o->dense_t0 = 0.0;
	o->dense_dt = 0.0;
	o->handout_order = REHUEL_ORDER_LOCALITY;
	o->shard_mode = REHUEL_SHARD_CONTIGUOUS;
}

int rehuel_name_to_method(const char *name) { return name_to_method(name); }
const char *rehuel_method_to_name(int method) { return method_to_name(method); }

int rehuel_name_to_problem(const char *name)
{
	if (!name) return 0;
	for (const ProblemInfo &p : kProblems)
		if (!std::strcmp(p.name, name)) return p.id;
	return 0;
}

int rehuel_problem_dims(int problem, int n_hint, int *n, int *p)
{
	const ProblemInfo *pi = find_problem(problem);
	if (!pi || !n || !p) return REHUEL_EINVAL;
	*n = pi->n ? pi->n : n_hint;
	*p = pi->p;
	if (*n <= 0) return REHUEL_EINVAL;
	return REHUEL_OK;
}

int rehuel_get_coefficients(int method, double *A, double *b, double *c, double *b2, double *gamma,
                            int *order, int *order2, double *d_weights, double *d2_weights)
{
	Tableau tb;
	if (!get_coefficients(method, tb)) return 0;
	for (int i = 0; i < tb.s; ++i) {
		for (int j = 0; j < tb.s; ++j)
			if (A) A[i * tb.s + j] = tb.A[i][j];
		if (b) b[i] = tb.b[i];
		if (c) c[i] = tb.c[i];
		if (b2) b2[i] = tb.b2[i];
		if (d_weights) d_weights[i] = tb.d[i];
		if (d2_weights) d2_weights[i] = tb.d2[i];
	}
	if (gamma) *gamma = tb.gamma;
	if (order) *order = tb.order;
	if (order2) *order2 = tb.order2;
	return tb.s;
}

int rehuel_project_b(int method, double theta, double *out)
{
	Tableau tb;
	if (!get_coefficients(method, tb) || !project_b(tb, theta, out)) return 0;
	return tb.s;
}

// Eigen-structure used by the kernels (for tests): T, Ti row-major s*s; lambda as
// (gamma_hat, alpha_0, beta_0, alpha_1, beta_1, ...).  Returns s.
int rehuel_get_eigen(int method, double *T, double *Ti, double *lambda, int *n_real, int *n_cplx, int *est_mode)
{
	Tableau tb;
	if (!get_coefficients(method, tb)) return 0;
	for (int i = 0; i < tb.s; ++i)
		for (int j = 0; j < tb.s; ++j) {
			T[i * tb.s + j] = tb.T[i][j];
			Ti[i * tb.s + j] = tb.Ti[i][j];
		}
	lambda[0] = tb.gamma_hat;
	for (int k = 0; k < tb.n_cplx; ++k) {
		lambda[1 + 2 * k] = tb.alpha[k];
		lambda[2 + 2 * k] = tb.beta[k];
	}
	*n_real = tb.n_real;
	*n_cplx = tb.n_cplx;
	*est_mode = tb.est_mode;
	return tb.s;
}

int rehuel_kernel_info(int problem, int method, int n_hint, int *regs, int *local_bytes, int *smem_bytes,
                       int *ctas_per_sm, int *block)
{
	Tableau tb;
	KernelArgs ka;
	rehuel_options o;
	rehuel_default_options(&o);
	int n = 0, p = 0;
	if (rehuel_problem_dims(problem, n_hint, &n, &p) != REHUEL_OK) return fail(REHUEL_EINVAL, "unknown problem id %d", problem);
	LaunchFn launch = find_launcher(problem, n);
	if (!launch) return REHUEL_EINVAL;
	int rc = prepare(method, 0.0, 1.0, 1e-6, 0, &o, tb, ka);
	if (rc) return rc;
	KernelInfo ki{};
	rc = launch(ka, tb, nullptr, &ki);
	if (rc) return rc;
	if (regs) *regs = ki.regs;
	if (local_bytes) *local_bytes = ki.local_bytes;
	if (smem_bytes) *smem_bytes = ki.smem_bytes;
	if (ctas_per_sm) *ctas_per_sm = ki.ctas_per_sm;
	if (block) *block = ki.block;
	return REHUEL_OK;
}

int rehuel_irk_ensemble_device(int problem, int method, double t0, double t1, double dt0, size_t M,
                               int n_hint, const rehuel_options *opts, const rehuel_device_io *io,
                               void *cuda_stream)
{
	int n = 0, p = 0;
	if (rehuel_problem_dims(problem, n_hint, &n, &p) != REHUEL_OK) return fail(REHUEL_EINVAL, "unknown problem id %d", problem);
	LaunchFn launch = find_launcher(problem, n);
	if (!launch) return REHUEL_EINVAL;
	return ensemble_device(launch, n, p, method, t0, t1, dt0, M, opts, io, static_cast<cudaStream_t>(cuda_stream));
}

int rehuel_irk_ensemble(int problem, int method, double t0, double t1, double dt0, size_t M, int n_hint,
                        const double *y0, int y0_broadcast, const double *params,
                        const rehuel_options *opts, size_t n_samples_cap, int n_gpus, rehuel_result *out)
{
	int n = 0, p = 0;
	if (rehuel_problem_dims(problem, n_hint, &n, &p) != REHUEL_OK) return fail(REHUEL_EINVAL, "unknown problem id %d", problem);
	LaunchFn launch = find_launcher(problem, n);
	if (!launch) return REHUEL_EINVAL;
	return ensemble_host(launch, n, p, method, t0, t1, dt0, M, y0, y0_broadcast, params, opts, n_samples_cap, n_gpus, out);
}

int rehuel_irk_ensemble_continue(int problem, int method, double t1, int n_hint, const double *params,
                                 const rehuel_options *opts, size_t n_samples_cap, int n_gpus,
                                 const rehuel_result *prev, rehuel_result *out)
{
	if (!prev || !prev->y || !prev->t_final || prev->M == 0)
		return fail(REHUEL_EINVAL, "continue: `prev` must be a (non-empty) result of this library");
	if (!prev->state) return fail(REHUEL_EINVAL, "continue: the previous call must run with options.keep_state != 0");
	int n = 0, p = 0;
	if (rehuel_problem_dims(problem, n_hint ? n_hint : prev->n, &n, &p) != REHUEL_OK) return fail(REHUEL_EINVAL, "unknown problem id %d", problem);
	if (n != prev->n) return fail(REHUEL_EINVAL, "continue: problem has n = %d but `prev` holds n = %d", n, prev->n);
	LaunchFn launch = find_launcher(problem, n);
	if (!launch) return REHUEL_EINVAL;
	// dt0 is only a placeholder here: every member's step size comes from its state
	return ensemble_host_from(launch, n, p, method, prev->t_final[0], t1, 1e-6, prev->M, prev->y, 0, params, opts, n_samples_cap,
	                          n_gpus, prev->t_final, prev->state, out);
}

int rehuel_eval_functor(int problem, int n_hint, double t, const double *y, const double *params, double *f, double *J)
{
	int n = 0, p = 0;
	if (rehuel_problem_dims(problem, n_hint, &n, &p) != REHUEL_OK) return fail(REHUEL_EINVAL, "unknown problem id %d", problem);
	if (!y || (p && !params) || !f || !J) return fail(REHUEL_EINVAL, "eval: y, params, f and J are required");
	EvalFn ev = find_evaluator(problem, n);
	if (!ev) return fail(REHUEL_EINVAL, "problem %d has no device functor for n = %d", problem, n);
	return ev(t, y, params, f, J);
}

void rehuel_release_cached_memory(void)
{
	g_cache.trim();
	g_streams.trim();
	int dev = 0;
	cudaMemPool_t pool;
	if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
}

void rehuel_debug_force_pivot(int on) { rehuel::g_force_pivot.store(on ? 1 : 0); }

void rehuel_result_free(rehuel_result *out) { free_result(out); }

int rehuel_irk_single(int problem, int method, double t0, double t1, double dt0, int n_hint,
                      const double *y0, const double *params, const rehuel_options *opts, size_t cap,
                      size_t *n_out, double *t_vals, double *y_vals, double *err_vals,
                      unsigned long long *counters, double *accept_frac, int *status)
{
	rehuel_result r;
	int rc = rehuel_irk_ensemble(problem, method, t0, t1, dt0, 1, n_hint, y0, 1, params, opts, cap, 1, &r);
	if (rc) return rc;
	const size_t produced = r.n_samples ? r.n_samples[0] : 0;
	const size_t valid = produced < r.n_samples_cap ? produced : r.n_samples_cap;
	if (n_out) *n_out = produced;
	for (size_t k = 0; k < valid; ++k) {
		if (t_vals) t_vals[k] = r.t_samples[k];
		if (err_vals) err_vals[k] = r.err_samples[k];
		if (y_vals)
			for (int c = 0; c < r.n; ++c) y_vals[k * r.n + c] = r.y_samples[k * r.n + c];
	}
	if (counters) { // order of rk_output::counters, irk.hpp:149-153
		counters[0] = r.attempt[0];
		counters[1] = r.reject_newton[0];
		counters[2] = r.reject_err[0];
		counters[3] = r.newton_success[0];
		counters[4] = 0; // newton_incr_diverge: never produced by the irk path
		counters[5] = 0; // newton_iter_error_too_large: idem
		counters[6] = r.newton_maxit_exceed[0];
		counters[7] = r.fun_evals[0];
		counters[8] = r.jac_evals[0];
	}
	if (accept_frac) *accept_frac = r.attempt[0] ? (double)r.steps[0] / (double)r.attempt[0] : 0.0;
	if (status) *status = r.status[0];
	rehuel_result_free(&r);
	return REHUEL_OK;
}

int rehuel_result_merge(rehuel_result *a, const rehuel_result *b)
{
	if (!a || !b || a->M != b->M || a->n != b->n) return fail(REHUEL_EINVAL, "merge: results must describe the same ensemble");
	const size_t M = a->M;
	const int n = a->n;
	ResultImpl *impl = static_cast<ResultImpl *>(a->impl);
	if (!impl) return fail(REHUEL_EINVAL, "merge: `a` was not produced by this library");
	// concatenate the sample rows: [cap_a + cap_b][...][M]
	const size_t cap = a->n_samples_cap + b->n_samples_cap;
	double *ts = nullptr, *ys = nullptr, *es = nullptr;
	if (cap) {
		int rc = pinned_alloc(impl, &ts, cap * M);
		if (!rc) rc = pinned_alloc(impl, &ys, cap * n * M);
		if (!rc) rc = pinned_alloc(impl, &es, cap * M);
		if (rc) return rc;
	}
	double w1 = 0.0, w2 = 0.0; // irk.cpp:767-771: weights are the numbers of stored samples
	unsigned *ns = a->n_samples;
	if (cap && !ns) { // leg 1 kept no samples but leg 2 did: leg 1 now needs a count array
		int rc = pinned_alloc(impl, &ns, M);
		if (rc) return rc;
		std::memset(ns, 0, sizeof(unsigned) * M);
	}
	for (size_t i = 0; i < M; ++i) {
		const size_t pa = a->n_samples ? a->n_samples[i] : 0, pb = b->n_samples ? b->n_samples[i] : 0;
		const size_t na = pa < a->n_samples_cap ? pa : a->n_samples_cap;
		const size_t nb = pb < b->n_samples_cap ? pb : b->n_samples_cap;
		for (size_t k = 0; k < na; ++k) {
			ts[k * M + i] = a->t_samples[k * M + i];
			es[k * M + i] = a->err_samples[k * M + i];
			for (int c = 0; c < n; ++c) ys[(k * n + c) * M + i] = a->y_samples[(k * n + c) * M + i];
		}
		for (size_t k = 0; k < nb; ++k) {
			ts[(na + k) * M + i] = b->t_samples[k * M + i];
			es[(na + k) * M + i] = b->err_samples[k * M + i];
			for (int c = 0; c < n; ++c) ys[((na + k) * n + c) * M + i] = b->y_samples[(k * n + c) * M + i];
		}
		w1 += (double)pa;
		w2 += (double)pb;
		if (ns) ns[i] = (unsigned)(na + nb);
		a->status[i] |= b->status[i];
		a->attempt[i] += b->attempt[i];
		a->reject_newton[i] += b->reject_newton[i];
		a->reject_err[i] += b->reject_err[i];
		a->newton_success[i] += b->newton_success[i];
		a->newton_maxit_exceed[i] += b->newton_maxit_exceed[i];
		a->newton_iters[i] += b->newton_iters[i];
		a->steps[i] += b->steps[i];
		// fun_evals / jac_evals are NOT merged by the reference (irk.cpp:773-780): keep leg 1's
		a->t_final[i] = b->t_final[i];
		a->err_last[i] = b->err_last[i];
		for (int c = 0; c < n; ++c) a->y[c * M + i] = b->y[c * M + i];
	}
	// controller state: the merged run stopped where leg 2 stopped
	if (a->state && b->state) std::memcpy(a->state, b->state, sizeof(double) * REHUEL_NSTATE * M);
	// dense output on a common grid: each leg filled the grid times it covered, the rest is NaN
	if (a->y_dense && b->y_dense && a->n_dense == b->n_dense) {
		const size_t cnt = a->n_dense * (size_t)n * M;
		for (size_t i = 0; i < cnt; ++i)
			if (a->y_dense[i] != a->y_dense[i]) a->y_dense[i] = b->y_dense[i];
	}
	a->n_samples = ns;
	a->t_samples = ts;
	a->y_samples = ys;
	a->err_samples = es;
	a->n_samples_cap = cap;
	a->elapsed_ms += b->elapsed_ms;
	a->kernel_ms += b->kernel_ms;
	a->accept_frac = (w1 + w2) > 0 ? (w1 * a->accept_frac + w2 * b->accept_frac) / (w1 + w2) : 0.0;
	a->sum.attempt += b->sum.attempt;
	a->sum.reject_newton += b->sum.reject_newton;
	a->sum.reject_err += b->sum.reject_err;
	a->sum.newton_success += b->sum.newton_success;
	a->sum.newton_maxit_exceed += b->sum.newton_maxit_exceed;
	a->sum.newton_iters += b->sum.newton_iters;
	a->sum.steps += b->sum.steps;
	a->sum.n_failed = 0;
	for (size_t i = 0; i < M; ++i) a->sum.n_failed += a->status[i] != 0;
	if (b->sum.max_attempt > a->sum.max_attempt) a->sum.max_attempt = b->sum.max_attempt;
	return REHUEL_OK;
}

} // extern "C"
