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
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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
	if (o->shard_mode != REHUEL_SHARD_CONTIGUOUS && o->shard_mode != REHUEL_SHARD_INTERLEAVED && o->shard_mode != REHUEL_SHARD_DYNAMIC)
		return fail(REHUEL_EINVAL, "shard_mode must be REHUEL_SHARD_CONTIGUOUS, _INTERLEAVED or _DYNAMIC");
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
	ka.trace_interval = o->out_interval > 0 ? o->out_interval : 1; // irk.hpp:805-806
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

// A stream with its timing events; recycled across calls like the memory blocks (creating and
// destroying 6-8 streams and 30-40 events per call costs more host time than issuing the work).
struct StreamSet {
	int dev = -1;
	cudaStream_t stream = nullptr;
	// ev[0] chunk start, ev[1] inputs on the device, ev[2] kernel start (hand-out order done), ev[3] kernel end,
	// ev[4] results on the host
	cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
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
		for (cudaEvent_t &e : out.ev) CUDA_TRY(cudaEventCreate(&e));
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
			for (cudaEvent_t e : s.ev)
				if (e) cudaEventDestroy(e);
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

// One unit of device work: a contiguous member range with its own stream, so that the H2D copy, the
// kernel and the D2H copies of different chunks overlap.  Which device runs it is decided when it is
// issued (statically dealt, or claimed from the shared cursor: REHUEL_SHARD_DYNAMIC).
struct DeviceShard {
	int dev = -1;
	size_t off = 0, m = 0;
	StreamSet ss;
	cudaStream_t stream = nullptr;
	void *slab = nullptr;
	rehuel_device_io io{};
	unsigned long long *d_stats = nullptr;
	size_t si = 0;             // index of the chunk (its statistics words in HostCall::h_stats)
	bool status_lazy = false;  // the status row was not copied back: see finish_chunk
	bool issued = false, synced = false;
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
constexpr size_t kTraceCap = size_t(1) << 18; // rows of the per-attempt log the host-buffer path keeps (10 MB)

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

// Everything the chunks of one host-buffer call share.
struct HostCall {
	LaunchFn launch;
	int n, p;
	size_t M;
	const double *y0, *params, *t_start, *state_in;
	int y0_broadcast;
	const rehuel_options *opts;
	size_t n_samples_cap, n_dense;
	bool keep_state, want_trace, timing;
	unsigned mask;
	Tableau tb;
	KernelArgs ka;
	rehuel_result *out;
	double *hdbl;  // [n + 2][M]: y rows, t_final, err_last
	int *hint;     // [1 + 9][M]: status, counters
	unsigned long long *h_stats; // [chunks][kStatWords]
	double *h_trace;
	unsigned *h_trace_len;
};

// Enqueue one chunk on its device (made current by the caller): inputs in, hand-out order, kernel, results out.
static int issue_chunk(const HostCall &hc, DeviceShard &sh, size_t si)
{
	const int n = hc.n, p = hc.p;
	const size_t M = hc.M, m = sh.m, n_samples_cap = hc.n_samples_cap, n_dense = hc.n_dense;
	const rehuel_options *opts = hc.opts;
	if (int rs = g_streams.get(sh.dev, sh.ss)) return rs;
	sh.stream = sh.ss.stream;
	sh.issued = true;
	const cudaEvent_t *ev = sh.ss.ev;
	const bool traced = hc.want_trace && opts->trace_member >= sh.off && opts->trace_member < sh.off + m;
	const size_t dbytes = Slab::need(hc.y0_broadcast ? (size_t)n : (size_t)n * m, 8) + Slab::need((size_t)p * m, 8) +
	                      Slab::need((size_t)(n + 2) * m, 8) + Slab::need((size_t)(1 + REHUEL_NCOUNTERS) * m, 4) +
	                      Slab::need(m, 4) + Slab::need(1, 8) + Slab::need(kStatWords, 8) +
	                      3 * Slab::need(n_samples_cap * m, 8) * (size_t)(n > 1 ? n : 1) +
	                      (hc.t_start ? Slab::need(m, 8) : 0) + (hc.state_in ? Slab::need((size_t)REHUEL_NSTATE * m, 8) : 0) +
	                      (hc.keep_state ? Slab::need((size_t)REHUEL_NSTATE * m, 8) : 0) + Slab::need(n_dense * n * m, 8) +
	                      (traced ? Slab::need(kTraceCap * 5, 8) + Slab::need(1, 4) : 0);
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
	double *d_y0 = ds.take<double>(hc.y0_broadcast ? (size_t)n : (size_t)n * m);
	double *d_par = ds.take<double>((size_t)p * m);
	double *ddbl = ds.take<double>((size_t)(n + 2) * m); // same row order as the host block
	int *dint = ds.take<int>((size_t)(1 + REHUEL_NCOUNTERS) * m);
	sh.io.y = ddbl;
	sh.io.t_final = ddbl + (size_t)n * m;
	sh.io.err_last = ddbl + (size_t)(n + 1) * m;
	sh.io.status = dint;
	// counters nobody asked for are not written either (their sums still come back through io.stats)
	sh.io.counters = (hc.mask & REHUEL_OUT_COUNTERS) ? reinterpret_cast<unsigned *>(dint + m) : nullptr;
	sh.io.n_samples = ds.take<unsigned>(m);
	sh.io.queue = ds.take<unsigned long long>(1);
	sh.d_stats = ds.take<unsigned long long>(kStatWords);
	double *d_tstart = hc.t_start ? ds.take<double>(m) : nullptr;
	double *d_state_in = hc.state_in ? ds.take<double>((size_t)REHUEL_NSTATE * m) : nullptr;
	sh.io.state_out = hc.keep_state ? ds.take<double>((size_t)REHUEL_NSTATE * m) : nullptr;
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
	if (traced) {
		sh.io.trace = ds.take<double>(kTraceCap * 5);
		sh.io.trace_len = ds.take<unsigned>(1);
		sh.io.trace_cap = kTraceCap;
		sh.io.trace_member = (size_t)opts->trace_member - sh.off;
	}
	sh.io.y0 = d_y0;
	sh.io.params = d_par;
	sh.io.y0_broadcast = hc.y0_broadcast;
	sh.io.n_samples_cap = n_samples_cap;
	if (hc.timing) CUDA_TRY(cudaEventRecord(ev[0], sh.stream));
	// H2D: SoA rows are strided by M on the host and by m on the device
	if (hc.y0_broadcast) {
		CUDA_TRY(cudaMemcpyAsync(d_y0, hc.y0, sizeof(double) * n, cudaMemcpyHostToDevice, sh.stream));
	} else {
		CUDA_TRY(cudaMemcpy2DAsync(d_y0, sizeof(double) * m, hc.y0 + sh.off, sizeof(double) * M, sizeof(double) * m, n,
		                           cudaMemcpyHostToDevice, sh.stream));
	}
	if (p)
		CUDA_TRY(cudaMemcpy2DAsync(d_par, sizeof(double) * m, hc.params + sh.off, sizeof(double) * M, sizeof(double) * m, p,
		                           cudaMemcpyHostToDevice, sh.stream));
	if (d_tstart) CUDA_TRY(cudaMemcpyAsync(d_tstart, hc.t_start + sh.off, sizeof(double) * m, cudaMemcpyHostToDevice, sh.stream));
	if (d_state_in)
		CUDA_TRY(cudaMemcpy2DAsync(d_state_in, sizeof(double) * m, hc.state_in + sh.off, sizeof(double) * M, sizeof(double) * m,
		                           REHUEL_NSTATE, cudaMemcpyHostToDevice, sh.stream));
	if (hc.timing) CUDA_TRY(cudaEventRecord(ev[1], sh.stream));
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
	KernelArgs k = hc.ka;
	k.M = m;
	k.io = sh.io;
	if (n_samples_cap == 0) k.store_samples = 0; // final state only: skip the per-step sample bookkeeping
	CUDA_TRY(cudaEventRecord(ev[2], sh.stream));
	int rr = hc.launch(k, hc.tb, sh.stream, nullptr);
	if (rr) return rr;
	CUDA_TRY(cudaEventRecord(ev[3], sh.stream));
	// D2H: the ensemble sums, y and status always; the other per-member rows as options.result_mask asks
	CUDA_TRY(cudaMemcpyAsync(hc.h_stats + si * kStatWords, sh.d_stats, sizeof(unsigned long long) * kStatWords,
	                         cudaMemcpyDeviceToHost, sh.stream));
	const size_t drows = (size_t)n + ((hc.mask & REHUEL_OUT_T_FINAL) ? ((hc.mask & REHUEL_OUT_ERR_LAST) ? 2 : 1) : 0);
	CUDA_TRY(cudaMemcpy2DAsync(hc.hdbl + sh.off, sizeof(double) * M, ddbl, sizeof(double) * m, sizeof(double) * m, drows,
	                           cudaMemcpyDeviceToHost, sh.stream));
	if ((hc.mask & REHUEL_OUT_ERR_LAST) && !(hc.mask & REHUEL_OUT_T_FINAL))
		CUDA_TRY(cudaMemcpyAsync(hc.hdbl + (size_t)(n + 1) * M + sh.off, sh.io.err_last, sizeof(double) * m, cudaMemcpyDeviceToHost, sh.stream));
	// The status row is constant on success, and the chunk's statistics say whether every member succeeded: when no
	// counter rows travel with it, finish_chunk fills it on the host and only fetches it for a chunk with failed members.
	sh.si = si;
	sh.status_lazy = !(hc.mask & REHUEL_OUT_COUNTERS);
	if (!sh.status_lazy)
		CUDA_TRY(cudaMemcpy2DAsync(hc.hint + sh.off, sizeof(int) * M, dint, sizeof(int) * m, sizeof(int) * m, (size_t)1 + REHUEL_NCOUNTERS,
		                           cudaMemcpyDeviceToHost, sh.stream));
	rehuel_result *out = hc.out;
	if (hc.keep_state)
		CUDA_TRY(cudaMemcpy2DAsync(out->state + sh.off, sizeof(double) * M, sh.io.state_out, sizeof(double) * m, sizeof(double) * m,
		                           REHUEL_NSTATE, cudaMemcpyDeviceToHost, sh.stream));
	if (n_dense)
		CUDA_TRY(cudaMemcpy2DAsync(out->y_dense + sh.off, sizeof(double) * M, sh.io.y_dense, sizeof(double) * m, sizeof(double) * m,
		                           n_dense * n, cudaMemcpyDeviceToHost, sh.stream));
	if (n_samples_cap) {
		CUDA_TRY(cudaMemcpyAsync(out->n_samples + sh.off, sh.io.n_samples, sizeof(unsigned) * m, cudaMemcpyDeviceToHost, sh.stream));
		CUDA_TRY(cudaMemcpy2DAsync(out->t_samples + sh.off, sizeof(double) * M, sh.io.t_samples, sizeof(double) * m,
		                           sizeof(double) * m, n_samples_cap, cudaMemcpyDeviceToHost, sh.stream));
		CUDA_TRY(cudaMemcpy2DAsync(out->err_samples + sh.off, sizeof(double) * M, sh.io.err_samples, sizeof(double) * m,
		                           sizeof(double) * m, n_samples_cap, cudaMemcpyDeviceToHost, sh.stream));
		CUDA_TRY(cudaMemcpy2DAsync(out->y_samples + sh.off, sizeof(double) * M, sh.io.y_samples, sizeof(double) * m,
		                           sizeof(double) * m, n_samples_cap * n, cudaMemcpyDeviceToHost, sh.stream));
	}
	if (traced) {
		CUDA_TRY(cudaMemcpyAsync(hc.h_trace_len, sh.io.trace_len, sizeof(unsigned), cudaMemcpyDeviceToHost, sh.stream));
		CUDA_TRY(cudaMemcpyAsync(hc.h_trace, sh.io.trace, sizeof(double) * 5 * kTraceCap, cudaMemcpyDeviceToHost, sh.stream));
	}
	if (hc.timing) CUDA_TRY(cudaEventRecord(ev[4], sh.stream));
	return REHUEL_OK;
}

// Wait for an issued chunk; with options.time_internals its phase times are added to `phase`.
static int finish_chunk(const HostCall &hc, DeviceShard &sh, double *phase)
{
	if (!sh.issued || sh.synced) return REHUEL_OK;
	CUDA_TRY(cudaStreamSynchronize(sh.stream));
	if (sh.status_lazy) {
		if (hc.h_stats[sh.si * kStatWords + REHUEL_NCOUNTERS] == 0) { // no failed member in this chunk
			std::memset(hc.hint + sh.off, 0, sizeof(int) * sh.m);
		} else {
			CUDA_TRY(cudaMemcpyAsync(hc.hint + sh.off, sh.io.status, sizeof(int) * sh.m, cudaMemcpyDeviceToHost, sh.stream));
			CUDA_TRY(cudaStreamSynchronize(sh.stream));
		}
	}
	sh.synced = true;
	if (hc.timing && phase)
		for (int ph = 0; ph < 4; ++ph) {
			float ms = 0.f;
			CUDA_TRY(cudaEventElapsedTime(&ms, sh.ss.ev[ph], sh.ss.ev[ph + 1]));
			phase[ph] += ms;
		}
	return REHUEL_OK;
}

int ensemble_host_from(LaunchFn launch, int n, int p, int method, double t0, double t1, double dt0, size_t M,
                       const double *y0, int y0_broadcast, const double *params, const rehuel_options *opts,
                       size_t n_samples_cap, int n_gpus, const double *t_start, const double *state_in, rehuel_result *out)
{
	auto tic = std::chrono::steady_clock::now();
	if (!out) return fail(REHUEL_EINVAL, "out must not be NULL");
	std::memset(out, 0, sizeof *out);
	HostCall hc{};
	int rc = prepare(method, t0, t1, dt0, M, opts, hc.tb, hc.ka);
	if (rc) return rc;
	KernelArgs &ka = hc.ka;
	if (!y0 || (p && !params)) return fail(REHUEL_EINVAL, "y0/params must not be NULL");
	const size_t n_dense = (size_t)opts->dense_samples;
	const bool keep_state = opts->keep_state != 0;
	if (n_dense && !hc.tb.has_interp) return fail(REHUEL_EINVAL, "Chosen method does not have dense output!"); // irk.cpp:733
	if (n_dense && !(opts->dense_dt > 0.0)) return fail(REHUEL_EINVAL, "dense_dt must be > 0");
	if ((opts->out_interval > 0 || (opts->output_mode & 2)) && opts->trace_member >= M)
		return fail(REHUEL_EINVAL, "trace_member %llu is not a member of this ensemble (M = %zu)", opts->trace_member, M);
	const bool to_file = (opts->output_mode & 2) != 0; // WRITE_TO_FILE, options.hpp:96-103
	if (to_file && !opts->output_file) return fail(REHUEL_EINVAL, "output_mode has WRITE_TO_FILE set but output_file is NULL (options.hpp:119-123)");
	if (to_file && n_samples_cap == 0) return fail(REHUEL_EINVAL, "WRITE_TO_FILE needs sample storage: n_samples_cap must be > 0");
	if (to_file) ka.store_samples = 1;
	int ndev = 0;
	CUDA_TRY(cudaGetDeviceCount(&ndev));
	if (n_gpus < 1) n_gpus = 1;
	if (n_gpus > ndev) return fail(REHUEL_EINVAL, "n_gpus=%d but only %d CUDA devices are visible", n_gpus, ndev);
	if (!ka.store_samples) n_samples_cap = 0;
	int cur_dev = 0;
	CUDA_TRY(cudaGetDevice(&cur_dev));
	const int dev0 = (n_gpus == 1) ? cur_dev : 0;
	hc.launch = launch;
	hc.n = n;
	hc.p = p;
	hc.M = M;
	hc.y0 = y0;
	hc.params = params;
	hc.t_start = t_start;
	hc.state_in = state_in;
	hc.y0_broadcast = y0_broadcast;
	hc.opts = opts;
	hc.n_samples_cap = n_samples_cap;
	hc.n_dense = n_dense;
	hc.keep_state = keep_state;
	hc.want_trace = opts->out_interval > 0;
	hc.timing = opts->time_internals != 0;
	hc.mask = opts->result_mask & REHUEL_OUT_ALL;
	hc.out = out;

	// ---- work units: cut [0, M) into chunks on separate streams so that the H2D copy of chunk i+1 and the D2H copies
	// of chunk i-1 overlap the kernel of chunk i.  One GPU: up to 8 equal chunks.  Several GPUs: contiguous shares of 8
	// chunks each (REHUEL_SHARD_CONTIGUOUS), chunks dealt in turn (_INTERLEAVED), or -- the default -- many small chunks
	// that the devices claim from a shared cursor as they finish (_DYNAMIC: SURVEY.md 8e).
	// tuning aids: REHUEL_CHUNKS=k (chunks per GPU), REHUEL_CHUNK_WEIGHTS="1,2,4,4,2,1" (relative chunk sizes, one GPU)
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
	const int mode = n_gpus > 1 ? opts->shard_mode : REHUEL_SHARD_CONTIGUOUS;
	std::deque<DeviceShard> shards; // deque: elements are never moved (DeviceShard owns CUDA handles)
	// members [lo, hi) cut into `chunks` pieces (relative sizes wts[] when `weighted`), owner = dev_first + (c mod owners),
	// or -1: claimed at run time
	auto add_chunks = [&](size_t lo, size_t hi, size_t chunks, bool weighted, int dev_first, int owners) {
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
			sh.dev = owners > 0 ? dev_first + (int)(c % (size_t)owners) : -1;
			sh.off = cut(wacc);
			wacc += weighted ? wts[c] : 1;
			sh.m = cut(wacc) - sh.off;
		}
	};
	auto chunks_for = [&](size_t members, size_t most) -> size_t { // >= 128 Ki members per chunk; 8 equal chunks measured best at 2^20
		size_t c = members / (128u << 10);
		c = c < 1 ? 1 : (c > most ? most : c);
		if (env_chunks >= 1 && env_chunks <= 64) c = (size_t)env_chunks;
		if (n_samples_cap) c = 1;
		return c;
	};
	if (mode == REHUEL_SHARD_DYNAMIC) {
		// small chunks (>= 16 Ki members, at most 64 per GPU): the tail a device can be left with is one chunk
		size_t c = M / (16u << 10);
		const size_t most = (size_t)n_gpus * 64;
		c = c < (size_t)n_gpus ? (size_t)n_gpus : (c > most ? most : c);
		if (env_chunks >= 1 && env_chunks <= 64) c = (size_t)env_chunks * (size_t)n_gpus;
		if (n_samples_cap) c = (size_t)n_gpus;
		add_chunks(0, M, c, false, dev0, 0);
	} else if (mode == REHUEL_SHARD_INTERLEAVED) {
		add_chunks(0, M, chunks_for(M / (size_t)n_gpus, 8) * (size_t)n_gpus, false, dev0, n_gpus); // chunk k -> GPU k mod G
	} else {
		// The thread-per-system kernels take their members in gangs of 32 per warp (irk_thread.cuh: kBatchLanes), so a
		// chunk costs whole gang rounds: cut a device's share in multiples of the persistent grid's lanes.
		size_t lanes = 0;
		if (n <= 4 && !n_samples_cap && env_chunks == 0 && n_wts == 0) {
			KernelInfo ki{};
			int sms = 0;
			if (launch(hc.ka, hc.tb, nullptr, &ki) == REHUEL_OK && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev0) == cudaSuccess)
				lanes = (size_t)sms * (size_t)ki.ctas_per_sm * (size_t)ki.block;
			if (lanes % 1024 != 0) lanes = 0; // chunk boundaries stay 4 KiB-aligned
		}
		for (int g = 0; g < n_gpus; ++g) {
			const size_t lo = M * (size_t)g / (size_t)n_gpus, hi = M * (size_t)(g + 1) / (size_t)n_gpus;
			const bool weighted = n_wts >= 1 && !n_samples_cap && n_gpus == 1;
			if (lanes && hi - lo >= 3 * lanes) {
				size_t per = ((hi - lo) / 8 + lanes / 2) / lanes; // rounds per chunk: about an eighth of the share
				per = (per < 1 ? 1 : per) * lanes;
				for (size_t at = lo; at < hi; at += per) {
					shards.emplace_back();
					DeviceShard &sh = shards.back();
					sh.dev = dev0 + g;
					sh.off = at;
					sh.m = (hi - at < per) ? hi - at : per;
				}
			} else {
				add_chunks(lo, hi, weighted ? n_wts : chunks_for(hi - lo, 8), weighted, dev0 + g, 1);
			}
		}
	}
	const size_t n_shards = shards.size();
	{ // the work units must tile [0, M) exactly: a gap would leave members silently unsolved
		size_t next = 0;
		for (const DeviceShard &sh : shards) {
			if (sh.off != next) return fail(REHUEL_EINVAL, "internal: chunk layout leaves a gap at member %zu", next);
			next += sh.m;
		}
		if (next != M) return fail(REHUEL_EINVAL, "internal: chunk layout covers %zu of %zu members", next, M);
	}

	// ---- result storage: ONE cached page-locked slab (so that the D2H copies are asynchronous)
	ResultImpl *impl = new (std::nothrow) ResultImpl();
	if (!impl) return fail(REHUEL_ENOMEM, "out of host memory");
	out->impl = impl;
	out->M = M;
	out->n = n;
	out->n_samples_cap = n_samples_cap;
	size_t hbytes = Slab::need(M, 8) + Slab::need((size_t)n * M, 8) + Slab::need(M, 8) + Slab::need(M, 4) +
	                Slab::need((size_t)REHUEL_NCOUNTERS * M, 4) + Slab::need(M, 4) +
	                Slab::need(n_shards * kStatWords, 8) + 3 * Slab::need(n_samples_cap * M, 8) * (size_t)(n > 1 ? n : 1) +
	                (keep_state ? Slab::need((size_t)REHUEL_NSTATE * M, 8) : 0) + Slab::need(n_dense * n * M, 8) +
	                (hc.want_trace ? Slab::need(kTraceCap * 5, 8) + Slab::need(1, 4) : 0);
	Slab hs;
	hs.base = static_cast<char *>(g_cache.get(-1, hbytes));
	hs.size = hbytes;
	if (!hs.base) {
		free_result(out);
		return REHUEL_ENOMEM;
	}
	impl->pinned.push_back(hs.base);
	// one block of doubles [n + 2][M] (y rows, t_final, err_last) and one of 32-bit words [1 + 9][M] (status,
	// counters): a chunk comes back with one pitched copy per block
	double *hdbl = hs.take<double>((size_t)(n + 2) * M);
	int *hint = hs.take<int>((size_t)(1 + REHUEL_NCOUNTERS) * M);
	if (!hdbl || !hint) {
		free_result(out);
		return fail(REHUEL_ENOMEM, "internal: result slab too small");
	}
	hc.hdbl = hdbl;
	hc.hint = hint;
	out->y = hdbl;
	out->t_final = (hc.mask & REHUEL_OUT_T_FINAL) ? hdbl + (size_t)n * M : nullptr;
	out->err_last = (hc.mask & REHUEL_OUT_ERR_LAST) ? hdbl + (size_t)(n + 1) * M : nullptr;
	out->status = hint;
	unsigned *counters = reinterpret_cast<unsigned *>(hint + M); // [NCOUNTERS][M]
	out->n_samples = n_samples_cap ? hs.take<unsigned>(M) : nullptr; // NULL: no samples were asked for
	unsigned long long *h_stats = hs.take<unsigned long long>(n_shards * kStatWords);
	std::memset(h_stats, 0, sizeof(unsigned long long) * n_shards * kStatWords);
	hc.h_stats = h_stats;
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
	if (hc.want_trace) {
		hc.h_trace = hs.take<double>(kTraceCap * 5);
		hc.h_trace_len = hs.take<unsigned>(1);
		*hc.h_trace_len = 0;
	}
	if (hc.mask & REHUEL_OUT_COUNTERS) {
		out->attempt = counters + (size_t)REHUEL_CNT_ATTEMPT * M;
		out->reject_newton = counters + (size_t)REHUEL_CNT_REJECT_NEWTON * M;
		out->reject_err = counters + (size_t)REHUEL_CNT_REJECT_ERR * M;
		out->newton_success = counters + (size_t)REHUEL_CNT_NEWTON_SUCCESS * M;
		out->newton_maxit_exceed = counters + (size_t)REHUEL_CNT_NEWTON_MAXIT_EXCEED * M;
		out->fun_evals = counters + (size_t)REHUEL_CNT_FUN_EVALS * M;
		out->jac_evals = counters + (size_t)REHUEL_CNT_JAC_EVALS * M;
		out->newton_iters = counters + (size_t)REHUEL_CNT_NEWTON_ITERS * M;
		out->steps = counters + (size_t)REHUEL_CNT_STEPS * M;
	}

	// ---- run.  One GPU: this thread issues every chunk up front and then waits.  Several GPUs: one issuing thread per
	// device, each keeping up to kDepth chunks in flight; in dynamic mode a thread claims its next chunk from the
	// shared cursor only when it has a free slot, so a device that finishes early takes more.
	constexpr size_t kDepth = 3;
	std::vector<double> phase((size_t)n_gpus * 4, 0.0);
	if (n_gpus == 1) {
		auto body = [&]() -> int {
			CUDA_TRY(cudaSetDevice(dev0));
			for (size_t si = 0; si < n_shards; ++si) {
				DeviceShard &sh = shards[si];
				if (sh.m == 0) continue;
				sh.dev = dev0;
				if (int r = issue_chunk(hc, sh, si)) return r;
			}
			for (DeviceShard &sh : shards)
				if (int r = finish_chunk(hc, sh, phase.data())) return r;
			return REHUEL_OK;
		};
		rc = body();
	} else {
		std::atomic<size_t> cursor{0};
		std::vector<int> wrc((size_t)n_gpus, REHUEL_OK);
		std::vector<std::string> werr((size_t)n_gpus);
		std::vector<std::thread> workers;
		for (int g = 0; g < n_gpus; ++g)
			workers.emplace_back([&, g]() {
				auto work = [&]() -> int {
					CUDA_TRY(cudaSetDevice(dev0 + g));
					std::deque<size_t> flying;
					size_t scan = 0; // static modes: next chunk of the list to look at
					for (;;) {
						size_t si = n_shards;
						if (mode == REHUEL_SHARD_DYNAMIC) {
							si = cursor.fetch_add(1);
						} else {
							while (scan < n_shards && shards[scan].dev != dev0 + g) ++scan;
							si = scan++;
						}
						if (si >= n_shards) break;
						DeviceShard &sh = shards[si];
						if (sh.m == 0) continue;
						sh.dev = dev0 + g;
						if (int r = issue_chunk(hc, sh, si)) return r;
						flying.push_back(si);
						if (flying.size() >= kDepth) {
							if (int r = finish_chunk(hc, shards[flying.front()], &phase[(size_t)g * 4])) return r;
							flying.pop_front();
						}
					}
					for (size_t si : flying)
						if (int r = finish_chunk(hc, shards[si], &phase[(size_t)g * 4])) return r;
					return REHUEL_OK;
				};
				wrc[(size_t)g] = work();
				if (wrc[(size_t)g]) werr[(size_t)g] = g_err; // the error text is per thread: hand it to the caller's
			});
		for (std::thread &w : workers) w.join();
		for (int g = 0; g < n_gpus && !rc; ++g)
			if (wrc[(size_t)g]) rc = fail(wrc[(size_t)g], "%s", werr[(size_t)g].c_str());
	}
	if (!rc) { // kernel_ms = per device, first kernel start -> last kernel end; max over devices
		for (int g = 0; g < n_gpus && !rc; ++g) {
			DeviceShard *first = nullptr;
			double kms = 0.0;
			for (DeviceShard &sh : shards) {
				if (sh.dev != dev0 + g || !sh.issued) continue;
				if (!first) first = &sh;
				float ms = 0.f;
				if (cudaSetDevice(sh.dev) != cudaSuccess || cudaEventElapsedTime(&ms, first->ss.ev[2], sh.ss.ev[3]) != cudaSuccess) {
					rc = fail(REHUEL_ECUDA, "cudaEventElapsedTime failed: %s", cudaGetErrorString(cudaGetLastError()));
					break;
				}
				if (ms > kms) kms = ms;
			}
			if (kms > out->kernel_ms) out->kernel_ms = kms;
		}
	}
	if (!rc && std::getenv("REHUEL_TIMING")) { // dev aid: where every chunk ran and when
		for (DeviceShard &sh : shards) {
			if (!sh.issued) continue;
			DeviceShard *first = nullptr;
			for (DeviceShard &o : shards)
				if (o.issued && o.dev == sh.dev) { first = &o; break; }
			float a0 = 0, a1 = 0;
			cudaSetDevice(sh.dev);
			cudaEventElapsedTime(&a0, first->ss.ev[2], sh.ss.ev[2]);
			cudaEventElapsedTime(&a1, first->ss.ev[2], sh.ss.ev[3]);
			std::fprintf(stderr, "[rehuel timing] chunk off=%zu m=%zu dev=%d: kernel %.3f..%.3f ms\n", sh.off, sh.m, sh.dev, a0, a1);
		}
	}
	shards.clear(); // returns the device slabs to the cache
	cudaSetDevice(cur_dev);
	if (rc) {
		free_result(out);
		return rc;
	}
	for (int g = 0; g < n_gpus; ++g)
		for (int ph = 0; ph < 4; ++ph) out->phase_ms[ph] += phase[(size_t)g * 4 + ph];
	if (hc.want_trace) {
		out->trace = hc.h_trace;
		out->trace_len = *hc.h_trace_len < kTraceCap ? *hc.h_trace_len : kTraceCap;
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

	// ---- WRITE_TO_FILE (irk.hpp:837-843): one "t y0 y1 ..." line per stored step of member trace_member, written at
	// the stream's default precision like the reference's operator<< (the initial state is stored but not written)
	if (to_file) {
		std::FILE *fh = std::fopen(opts->output_file, "w");
		if (!fh) {
			free_result(out);
			return fail(REHUEL_EINVAL, "cannot open output_file '%s' for writing", opts->output_file);
		}
		const size_t i = (size_t)opts->trace_member;
		const size_t produced = out->n_samples[i], rows = produced < n_samples_cap ? produced : n_samples_cap;
		for (size_t k = 1; k < rows; ++k) {
			std::fprintf(fh, "%g", out->t_samples[k * M + i]);
			for (int c = 0; c < n; ++c) std::fprintf(fh, " %g", out->y_samples[(k * n + c) * M + i]);
			std::fputc('\n', fh);
		}
		std::fclose(fh);
	}
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
	struct Scratch { // released on every exit path
		double *sink = nullptr;
		cudaEvent_t e0 = nullptr, e1 = nullptr;
		~Scratch()
		{
			if (e0) cudaEventDestroy(e0);
			if (e1) cudaEventDestroy(e1);
			cudaFree(sink);
		}
	} sc;
	CUDA_TRY(cudaMalloc(&sc.sink, sizeof(double)));
	CUDA_TRY(cudaEventCreate(&sc.e0));
	CUDA_TRY(cudaEventCreate(&sc.e1));
	const int iters = 4096, block = 256, grid = sms * 8;
	fp64_fma_peak_kernel<<<grid, block>>>(sc.sink, iters / 8, 1.0000001, 1e-9); // warm-up
	float best = 1e30f;
	for (int rep = 0; rep < 5; ++rep) {
		CUDA_TRY(cudaEventRecord(sc.e0));
		fp64_fma_peak_kernel<<<grid, block>>>(sc.sink, iters, 1.0000001, 1e-9);
		CUDA_TRY(cudaEventRecord(sc.e1));
		CUDA_TRY(cudaEventSynchronize(sc.e1));
		float ms = 0.f;
		CUDA_TRY(cudaEventElapsedTime(&ms, sc.e0, sc.e1));
		if (ms < best) best = ms;
	}
	CUDA_TRY(cudaGetLastError());
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
	o->shard_mode = REHUEL_SHARD_DYNAMIC;
	o->trace_member = 0;
	o->output_file = nullptr;           // options.hpp:107: output_stream is unset until set_output_stream()
	o->result_mask = REHUEL_OUT_ALL;
	o->time_internals = 0;              // irk.hpp:106
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
	if (!prev || !prev->y || prev->M == 0)
		return fail(REHUEL_EINVAL, "continue: `prev` must be a (non-empty) result of this library");
	if (!prev->state) return fail(REHUEL_EINVAL, "continue: the previous call must run with options.keep_state != 0");
	if (!prev->t_final) return fail(REHUEL_EINVAL, "continue: the previous call must return t_final (REHUEL_OUT_T_FINAL in options.result_mask)");
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

// ---- shared work cursor (one 64-bit word in POSIX shared memory)
struct rehuel_cursor {
	unsigned long long *word;
	std::string name;
};

rehuel_cursor *rehuel_cursor_open(const char *name, int create)
{
	if (!name || !*name || std::strchr(name, '/')) {
		fail(REHUEL_EINVAL, "cursor name must be a non-empty string without slashes");
		return nullptr;
	}
	const std::string path = std::string("/") + name;
	if (create) shm_unlink(path.c_str());
	const int fd = shm_open(path.c_str(), create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
	if (fd < 0) {
		fail(REHUEL_EINVAL, "shm_open(%s) failed: %s", path.c_str(), std::strerror(errno));
		return nullptr;
	}
	if (create && ftruncate(fd, sizeof(unsigned long long)) != 0) {
		fail(REHUEL_ENOMEM, "ftruncate(%s) failed: %s", path.c_str(), std::strerror(errno));
		close(fd);
		shm_unlink(path.c_str());
		return nullptr;
	}
	void *q = mmap(nullptr, sizeof(unsigned long long), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (q == MAP_FAILED) {
		fail(REHUEL_ENOMEM, "mmap(%s) failed: %s", path.c_str(), std::strerror(errno));
		if (create) shm_unlink(path.c_str());
		return nullptr;
	}
	rehuel_cursor *c = new (std::nothrow) rehuel_cursor{static_cast<unsigned long long *>(q), path};
	if (!c) {
		munmap(q, sizeof(unsigned long long));
		fail(REHUEL_ENOMEM, "out of host memory");
		return nullptr;
	}
	if (create) __atomic_store_n(c->word, 0ull, __ATOMIC_SEQ_CST);
	return c;
}

unsigned long long rehuel_cursor_next(rehuel_cursor *c, unsigned long long count)
{
	return c ? __atomic_fetch_add(c->word, count, __ATOMIC_SEQ_CST) : ~0ull;
}

void rehuel_cursor_reset(rehuel_cursor *c)
{
	if (c) __atomic_store_n(c->word, 0ull, __ATOMIC_SEQ_CST);
}

void rehuel_cursor_close(rehuel_cursor *c, int unlink)
{
	if (!c) return;
	munmap(c->word, sizeof(unsigned long long));
	if (unlink) shm_unlink(c->name.c_str());
	delete c;
}

void rehuel_debug_force_pivot(int on) { rehuel::g_force_pivot.store(on ? 1 : 0); }

void rehuel_result_free(rehuel_result *out) { free_result(out); }

int rehuel_irk_single(int problem, int method, double t0, double t1, double dt0, int n_hint,
                      const double *y0, const double *params, const rehuel_options *opts, size_t cap,
                      size_t *n_out, double *t_vals, double *y_vals, double *err_vals,
                      unsigned long long *counters, double *accept_frac, int *status)
{
	if (!opts) return fail(REHUEL_EINVAL, "options must not be NULL (the reference asserts newton_opts, irk.hpp:548)");
	rehuel_options o = *opts;
	o.result_mask = REHUEL_OUT_ALL; // this entry point reports every field of rk_output
	o.trace_member = 0;
	rehuel_result r;
	int rc = rehuel_irk_ensemble(problem, method, t0, t1, dt0, 1, n_hint, y0, 1, params, &o, cap, 1, &r);
	if (rc) return rc;
	if (o.out_interval > 0) { // irk.hpp:575-577, 805-810: the reference writes this log to output_opts.log_out (std::cout)
		const size_t need = rehuel_result_format_log(&r, 6, nullptr, 0);
		std::string text(need + 1, '\0');
		rehuel_result_format_log(&r, 6, &text[0], need + 1);
		std::fputs(text.c_str(), stdout);
	}
	if (o.time_internals) { // irk.hpp:866: print_timing_breakdown(timings, std::cerr, sc.name)
		const size_t need = rehuel_result_format_timings(&r, method_to_name(method), nullptr, 0);
		std::string text(need + 1, '\0');
		rehuel_result_format_timings(&r, method_to_name(method), &text[0], need + 1);
		std::fputs(text.c_str(), stderr);
	}
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

// snprintf-style appender for the two text formatters
struct TextOut {
	char *buf;
	size_t cap, len;
	void add(const char *fmt, ...)
	{
		char tmp[256];
		va_list ap;
		va_start(ap, fmt);
		const int k = vsnprintf(tmp, sizeof tmp, fmt, ap);
		va_end(ap);
		for (int i = 0; i < k && i < (int)sizeof tmp - 1; ++i, ++len)
			if (buf && len + 1 < cap) buf[len] = tmp[i];
	}
	size_t finish()
	{
		if (buf && cap) buf[len < cap ? len : cap - 1] = '\0';
		return len;
	}
};

size_t rehuel_result_format_log(const rehuel_result *res, int precision, char *buf, size_t cap)
{
	TextOut o{buf, cap, 0};
	if (!res) return o.finish();
	if (precision < 1) precision = 6;
	o.add("    Rehuel: step  t  dt   err   iters\n"); // irk.hpp:576
	for (size_t r = 0; r < res->trace_len; ++r) {
		const double *row = res->trace + 5 * r;
		if (row[3] < 0.0) continue; // failed Newton iteration: the reference's message is commented out (irk.hpp:649-656)
		o.add("    Rehuel: %lld %.*g %.*g %.*g %d\n", (long long)row[0], precision, row[1], precision, row[2], precision,
		      row[3], (int)row[4]); // irk.hpp:807-809
	}
	return o.finish();
}

size_t rehuel_result_format_timings(const rehuel_result *res, const char *method_name, char *buf, size_t cap)
{
	TextOut o{buf, cap, 0};
	if (!res) return o.finish();
	double total = 0.0;
	for (double v : res->phase_ms) total += v;
	// irk.hpp:346-359, with this path's four phases in place of the reference's six
	o.add("    Rehuel: %s done solving, timings (ms | %%):\n", method_name ? method_name : "IRK");
	const char *names[4] = {"Host-to-device copies:       ", "Hand-out ordering:           ", "Integration kernels:         ",
	                        "Device-to-host copies:       "};
	o.add("      %s%7.6g | %.3g\n", "Total time:                  ", total, 100.0);
	for (int ph = 0; ph < 4; ++ph) o.add("      %s%7.6g | %.3g\n", names[ph], res->phase_ms[ph], total > 0 ? 100.0 * res->phase_ms[ph] / total : 0.0);
	return o.finish();
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
		if (a->attempt && b->attempt) { // both legs carry the per-member counters (REHUEL_OUT_COUNTERS)
			a->attempt[i] += b->attempt[i];
			a->reject_newton[i] += b->reject_newton[i];
			a->reject_err[i] += b->reject_err[i];
			a->newton_success[i] += b->newton_success[i];
			a->newton_maxit_exceed[i] += b->newton_maxit_exceed[i];
			a->newton_iters[i] += b->newton_iters[i];
			a->steps[i] += b->steps[i];
			// fun_evals / jac_evals are NOT merged by the reference (irk.cpp:773-780): keep leg 1's
		}
		if (a->t_final && b->t_final) a->t_final[i] = b->t_final[i];
		if (a->err_last && b->err_last) a->err_last[i] = b->err_last[i];
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
