"""ctypes binding of include/rehuel_b200.h -- the reference-side stub a Python user would write.

Everything here goes through ``librehuel_b200.so`` (hand-written CUDA for sm_100a + the C++ host
layer).  There is no CPU fallback: if the library is missing the import fails, and if no GPU is
visible every solve returns REHUEL_ECUDA and raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# REHUEL_B200_LIB: development override used by tools/time_variants.py to time alternative builds
LIB_PATH = os.environ.get("REHUEL_B200_LIB") or os.path.join(_HERE, "librehuel_b200.so")

NCOUNTERS = 9
COUNTER_ROWS = ("attempt", "reject_newton", "reject_err", "newton_success", "newton_maxit_exceed",
                "fun_evals", "jac_evals", "newton_iters", "steps")

# method ids (enums.hpp:38-58)
IMPLICIT_EULER, RADAU_IIA_32, RADAU_IIA_53, RADAU_IIA_95, RADAU_IIA_137 = 200, 210, 211, 212, 213
LOBATTO_IIIC_43, LOBATTO_IIIC_64, LOBATTO_IIIC_85 = 230, 231, 232
# problem ids
ROBERTSON, VAN_DER_POL, BRUSSELATOR_1D, EXPONENTIAL, STIFF_EQ, BRUSSELATOR = 1, 2, 3, 5, 6, 7
LORENZ, HARMONIC, DIMER, KINETIC_4, THREE_BODY = 8, 9, 10, 11, 12
BRUSSELATOR_1D_DENSE = 4  # the same system with its Jacobian treated as dense (CTA-per-system kernel)
# status codes (enums.hpp:103-127)
SUCCESS, GENERAL_ERROR, ERROR_MAX_STEPS_EXCEEDED = 0, 4, 128


class Options(C.Structure):
    """``rehuel_options``: irk::solver_options + newton::options + output_options."""
    _fields_ = [
        ("rel_tol", C.c_double), ("abs_tol", C.c_double), ("max_dt", C.c_double),
        ("max_steps", C.c_longlong), ("adaptive_step_size", C.c_int), ("out_interval", C.c_int),
        ("newton_tol", C.c_double), ("newton_dx_delta", C.c_double),
        ("newton_maxit", C.c_int), ("newton_refresh_jac", C.c_int),
        ("output_mode", C.c_int), ("output_interval", C.c_ulonglong),
        ("keep_state", C.c_int), ("dense_samples", C.c_ulonglong), ("dense_t0", C.c_double), ("dense_dt", C.c_double),
        ("handout_order", C.c_int), ("shard_mode", C.c_int),
        ("trace_member", C.c_ulonglong), ("output_file", C.c_char_p), ("result_mask", C.c_uint),
        ("time_internals", C.c_int),
    ]


ORDER_INDEX, ORDER_LOCALITY, ORDER_LOCALITY_DESC = 0, 1, 2
SHARD_CONTIGUOUS, SHARD_INTERLEAVED, SHARD_DYNAMIC = 0, 1, 2
OUT_T_FINAL, OUT_ERR_LAST, OUT_COUNTERS, OUT_ALL = 1, 2, 4, 7  # rehuel_options.result_mask
PHASES = ("h2d", "order", "kernel", "d2h")                      # rehuel_result.phase_ms
NSTATE = 8  # rows of the controller state (REHUEL_STATE_*)


class Stats(C.Structure):
    _fields_ = [(k, C.c_ulonglong) for k in (
        "attempt", "reject_newton", "reject_err", "newton_success", "newton_incr_diverge",
        "newton_iter_error_too_large", "newton_maxit_exceed", "fun_evals", "jac_evals",
        "newton_iters", "steps", "n_failed", "max_attempt")]


_PD = C.POINTER(C.c_double)
_PU = C.POINTER(C.c_uint)
_PI = C.POINTER(C.c_int)


class Result(C.Structure):
    _fields_ = [
        ("M", C.c_size_t), ("n", C.c_int),
        ("t_final", _PD), ("y", _PD), ("err_last", _PD), ("status", _PI),
        ("attempt", _PU), ("reject_newton", _PU), ("reject_err", _PU), ("newton_success", _PU),
        ("newton_maxit_exceed", _PU), ("fun_evals", _PU), ("jac_evals", _PU), ("newton_iters", _PU),
        ("steps", _PU),
        ("n_samples_cap", C.c_size_t), ("n_samples", _PU),
        ("t_samples", _PD), ("y_samples", _PD), ("err_samples", _PD),
        ("state", _PD), ("n_dense", C.c_size_t), ("y_dense", _PD),
        ("elapsed_ms", C.c_double), ("kernel_ms", C.c_double), ("accept_frac", C.c_double),
        ("sum", Stats), ("trace_len", C.c_size_t), ("trace", _PD), ("phase_ms", C.c_double * 4), ("impl", C.c_void_p),
    ]


class DeviceIO(C.Structure):
    _fields_ = [
        ("y0", C.c_void_p), ("y0_broadcast", C.c_int), ("params", C.c_void_p),
        ("y", C.c_void_p), ("t_final", C.c_void_p), ("err_last", C.c_void_p), ("status", C.c_void_p),
        ("counters", C.c_void_p),
        ("n_samples_cap", C.c_size_t), ("n_samples", C.c_void_p),
        ("t_samples", C.c_void_p), ("y_samples", C.c_void_p), ("err_samples", C.c_void_p),
        ("queue", C.c_void_p),
        ("trace", C.c_void_p), ("trace_cap", C.c_size_t), ("trace_member", C.c_size_t),
        ("trace_len", C.c_void_p), ("rtol", C.c_void_p), ("atol", C.c_void_p), ("newton_tol", C.c_void_p),
        ("t_start", C.c_void_p), ("state_in", C.c_void_p), ("state_out", C.c_void_p),
        ("n_dense", C.c_size_t), ("dense_t0", C.c_double), ("dense_dt", C.c_double), ("y_dense", C.c_void_p),
        ("stats", C.c_void_p), ("order", C.c_void_p),
    ]


class RehuelError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rehuel_b200 error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` (nvcc, sm_100a). rehuel_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.rehuel_last_error.restype = C.c_char_p
    lib.rehuel_version.restype = C.c_char_p
    lib.rehuel_method_to_name.restype = C.c_char_p
    lib.rehuel_kernel_launches.restype = C.c_ulonglong
    lib.rehuel_irk_ensemble.argtypes = [
        C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_size_t, C.c_int, _PD, C.c_int, _PD,
        C.POINTER(Options), C.c_size_t, C.c_int, C.POINTER(Result)]
    lib.rehuel_irk_ensemble_device.argtypes = [
        C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_size_t, C.c_int,
        C.POINTER(Options), C.POINTER(DeviceIO), C.c_void_p]
    lib.rehuel_irk_single.argtypes = [
        C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, _PD, _PD, C.POINTER(Options),
        C.c_size_t, C.POINTER(C.c_size_t), _PD, _PD, _PD, C.POINTER(C.c_ulonglong), _PD, _PI]
    lib.rehuel_locality_order_workspace.restype = C.c_size_t
    lib.rehuel_locality_order_workspace.argtypes = [C.c_size_t]
    lib.rehuel_locality_order.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.c_void_p]
    lib.rehuel_irk_ensemble_continue.argtypes = [
        C.c_int, C.c_int, C.c_double, C.c_int, _PD, C.POINTER(Options), C.c_size_t, C.c_int, C.POINTER(Result),
        C.POINTER(Result)]
    lib.rehuel_result_free.argtypes = [C.POINTER(Result)]
    lib.rehuel_result_merge.argtypes = [C.POINTER(Result), C.POINTER(Result)]
    lib.rehuel_cursor_open.restype = C.c_void_p
    lib.rehuel_cursor_open.argtypes = [C.c_char_p, C.c_int]
    lib.rehuel_cursor_next.restype = C.c_ulonglong
    lib.rehuel_cursor_next.argtypes = [C.c_void_p, C.c_ulonglong]
    lib.rehuel_cursor_reset.argtypes = [C.c_void_p]
    lib.rehuel_cursor_close.argtypes = [C.c_void_p, C.c_int]
    lib.rehuel_result_format_log.restype = C.c_size_t
    lib.rehuel_result_format_log.argtypes = [C.POINTER(Result), C.c_int, C.c_char_p, C.c_size_t]
    lib.rehuel_result_format_timings.restype = C.c_size_t
    lib.rehuel_result_format_timings.argtypes = [C.POINTER(Result), C.c_char_p, C.c_char_p, C.c_size_t]
    return lib


lib = _load()

# every symbol include/rehuel_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = (
    "rehuel_default_options", "rehuel_name_to_method", "rehuel_method_to_name", "rehuel_name_to_problem",
    "rehuel_problem_dims", "rehuel_get_coefficients", "rehuel_irk_ensemble", "rehuel_result_free",
    "rehuel_irk_single", "rehuel_irk_ensemble_device", "rehuel_result_merge", "rehuel_kernel_launches",
    "rehuel_kernel_info", "rehuel_last_error", "rehuel_version", "rehuel_fp64_peak", "rehuel_release_cached_memory",
    "rehuel_locality_order_workspace", "rehuel_locality_order", "rehuel_irk_ensemble_continue", "rehuel_selftest_band",
    "rehuel_eval_functor", "rehuel_debug_force_pivot", "rehuel_result_format_log", "rehuel_result_format_timings",
    "rehuel_cursor_open", "rehuel_cursor_next", "rehuel_cursor_reset", "rehuel_cursor_close",
)


class Cursor:
    """rehuel_cursor: a work cursor shared by the ranks of one node (named POSIX shared memory)."""

    def __init__(self, name: str, create: bool):
        self.h = lib.rehuel_cursor_open(name.encode(), int(create))
        self.owner = create
        if not self.h:
            raise RehuelError(-1, last_error())

    def next(self, count: int = 1) -> int:
        return int(lib.rehuel_cursor_next(self.h, count))

    def reset(self) -> None:
        lib.rehuel_cursor_reset(self.h)

    def close(self) -> None:
        if self.h:
            lib.rehuel_cursor_close(self.h, int(self.owner))
            self.h = None


def last_error() -> str:
    return lib.rehuel_last_error().decode()


def check(rc: int) -> None:
    if rc != 0:
        raise RehuelError(rc, last_error())


def default_options() -> Options:
    o = Options()
    lib.rehuel_default_options(C.byref(o))
    return o


def make_options(rel_tol=1e-4, abs_tol=None, max_dt=0.0, max_steps=-1, adaptive=True, out_interval=0,
                 newton_tol=None, newton_dx_delta=1e-4, newton_maxit=5000, newton_refresh_jac=10,
                 output_mode=1, output_interval=1, handout_order=ORDER_LOCALITY, keep_state=False,
                 dense_samples=0, dense_t0=0.0, dense_dt=0.0, shard_mode=SHARD_DYNAMIC, trace_member=0, output_file=None,
                 result_mask=OUT_ALL, time_internals=False) -> Options:
    """Reference defaults, with newton_tol = 0.1*min(rtol, atol) unless given -- what the example
    driver and the 4-argument odeint do (examples/example_equations.cpp:220, irk.hpp:921)."""
    if abs_tol is None:
        abs_tol = 10 * rel_tol
    if newton_tol is None:
        newton_tol = 0.1 * min(rel_tol, abs_tol)
    return Options(rel_tol, abs_tol, max_dt, max_steps, int(adaptive), out_interval, newton_tol,
                   newton_dx_delta, newton_maxit, newton_refresh_jac, output_mode, output_interval, int(keep_state),
                   dense_samples, dense_t0, dense_dt, handout_order, shard_mode, trace_member,
                   output_file.encode() if isinstance(output_file, str) else output_file, result_mask, int(time_internals))


def name_to_method(name: str) -> int:
    return lib.rehuel_name_to_method(name.encode())


def method_to_name(method: int) -> str:
    return lib.rehuel_method_to_name(C.c_int(method)).decode()


def name_to_problem(name: str) -> int:
    return lib.rehuel_name_to_problem(name.encode())


def problem_dims(problem: int, n_hint: int = 0):
    n, p = C.c_int(), C.c_int()
    check(lib.rehuel_problem_dims(C.c_int(problem), C.c_int(n_hint), C.byref(n), C.byref(p)))
    return n.value, p.value


def get_coefficients(method: int):
    A = np.zeros(49); b = np.zeros(7); c = np.zeros(7); b2 = np.zeros(7); d = np.zeros(7); d2 = np.zeros(7)
    g = C.c_double(); o1 = C.c_int(); o2 = C.c_int()
    pd = lambda a: a.ctypes.data_as(_PD)
    s = lib.rehuel_get_coefficients(C.c_int(method), pd(A), pd(b), pd(c), pd(b2), C.byref(g), C.byref(o1),
                                    C.byref(o2), pd(d), pd(d2))
    if s <= 0:
        return None
    return dict(s=s, A=A[:s * s].reshape(s, s).copy(), b=b[:s].copy(), c=c[:s].copy(), b2=b2[:s].copy(),
                gamma=g.value, order=o1.value, order2=o2.value, d=d[:s].copy(), d2=d2[:s].copy())


def get_eigen(method: int):
    T = np.zeros(49); Ti = np.zeros(49); lam = np.zeros(7)
    nr, nc, em = C.c_int(), C.c_int(), C.c_int()
    pd = lambda a: a.ctypes.data_as(_PD)
    s = lib.rehuel_get_eigen(C.c_int(method), pd(T), pd(Ti), pd(lam), C.byref(nr), C.byref(nc), C.byref(em))
    if s <= 0:
        return None
    return dict(s=s, T=T[:s * s].reshape(s, s).copy(), Ti=Ti[:s * s].reshape(s, s).copy(), lam=lam.copy(),
                n_real=nr.value, n_cplx=nc.value, est_mode=em.value)


def project_b(method: int, theta: float):
    out = np.zeros(7)
    s = lib.rehuel_project_b(C.c_int(method), C.c_double(theta), out.ctypes.data_as(_PD))
    return out[:s].copy() if s > 0 else None


def eval_functor(problem: int, t: float, y, params, n_hint: int = 0):
    """f(t, y) and J(t, y) of a built-in device functor, evaluated on the GPU (diagnostic)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, p = problem_dims(problem, n_hint or y.shape[0])
    par = np.ascontiguousarray(params if p else [0.0], dtype=np.float64)
    f = np.zeros(n); J = np.zeros((n, n))
    pd = lambda a: a.ctypes.data_as(_PD)
    check(lib.rehuel_eval_functor(C.c_int(problem), C.c_int(n_hint or n), C.c_double(t), pd(y), pd(par), pd(f), pd(J)))
    return f, J


def kernel_info(problem: int, method: int, n_hint: int = 0) -> dict:
    v = [C.c_int() for _ in range(5)]
    check(lib.rehuel_kernel_info(C.c_int(problem), C.c_int(method), C.c_int(n_hint), *[C.byref(x) for x in v]))
    return dict(zip(("regs", "local_bytes", "smem_bytes", "ctas_per_sm", "block"), [x.value for x in v]))


def debug_force_pivot(on: bool) -> None:
    """Test hook: every attempt of the warp-per-system kernels takes the pivoted band fallback."""
    lib.rehuel_debug_force_pivot(C.c_int(1 if on else 0))


def fp64_peak_tflops() -> float:
    """Measured FP64 FMA peak of the current device (TFLOP/s)."""
    t, ms = C.c_double(), C.c_double()
    check(lib.rehuel_fp64_peak(C.byref(t), C.byref(ms)))
    return t.value


def kernel_launches() -> int:
    return int(lib.rehuel_kernel_launches())


@dataclass
class EnsembleOutput:
    """Host copy of ``rehuel_result`` (numpy arrays own their data; the C result is freed)."""
    y: np.ndarray
    t_final: np.ndarray
    err_last: np.ndarray
    status: np.ndarray
    counters: dict
    n_samples: np.ndarray
    t_samples: np.ndarray | None
    y_samples: np.ndarray | None
    err_samples: np.ndarray | None
    elapsed_ms: float
    kernel_ms: float
    accept_frac: float
    sum: dict
    state: np.ndarray | None = None    # [NSTATE, M] controller state at exit (options.keep_state)
    y_dense: np.ndarray | None = None  # [n_dense, n, M] dense-output samples (options.dense_samples)
    trace: np.ndarray | None = None    # [trace_len, 5] step, t, dt, err, iters of options.trace_member (options.out_interval)
    log: str = ""                      # the same as the reference prints it (irk.hpp:575-577, 805-810), 17 digits
    phase_ms: dict | None = None       # options.time_internals: device ms per phase


def _as_np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).astype(dtype, copy=True)


def irk_ensemble(problem: int, method: int, t0: float, t1: float, dt0: float, y0, params, opts: Options,
                 n_samples_cap: int = 0, n_gpus: int = 1, n_hint: int = 0, raw: bool = False):
    """Batched irk::odeint through the C ABI with HOST buffers (H2D/D2H inside the call)."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    n, p = problem_dims(problem, n_hint or y0.shape[0])
    broadcast = int(y0.ndim == 1)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(p, -1) if p else np.zeros((1, 1))
    M = params.shape[1] if p else (1 if broadcast else y0.shape[1])
    if y0.shape[0] != n or (not broadcast and y0.shape[1] != M):
        raise ValueError("y0 must be [n] or [n, M]")
    res = Result()
    check(lib.rehuel_irk_ensemble(problem, method, t0, t1, dt0, M, n_hint or n, y0.ctypes.data_as(_PD), broadcast,
                                  params.ctypes.data_as(_PD), C.byref(opts), n_samples_cap, n_gpus, C.byref(res)))
    if raw:
        return res
    try:
        out = to_output(res)
    finally:
        lib.rehuel_result_free(C.byref(res))
    return out


def to_output(res: Result) -> EnsembleOutput:
    """Copy a raw ``rehuel_result`` into numpy arrays (the C result stays owned by the caller)."""
    n, M, cap = res.n, res.M, res.n_samples_cap
    return EnsembleOutput(
        y=_as_np(res.y, (n, M), np.float64), t_final=_as_np(res.t_final, (M,), np.float64) if res.t_final else None,
        err_last=_as_np(res.err_last, (M,), np.float64) if res.err_last else None, status=_as_np(res.status, (M,), np.int32),
        counters={k: _as_np(getattr(res, k), (M,), np.uint32) for k in COUNTER_ROWS} if res.attempt else None,
        n_samples=_as_np(res.n_samples, (M,), np.uint32),
        t_samples=_as_np(res.t_samples, (cap, M), np.float64) if cap else None,
        y_samples=_as_np(res.y_samples, (cap, n, M), np.float64) if cap else None,
        err_samples=_as_np(res.err_samples, (cap, M), np.float64) if cap else None,
        elapsed_ms=res.elapsed_ms, kernel_ms=res.kernel_ms, accept_frac=res.accept_frac,
        sum={k: int(getattr(res.sum, k)) for k, _ in Stats._fields_},
        state=_as_np(res.state, (NSTATE, M), np.float64) if res.state else None,
        y_dense=_as_np(res.y_dense, (res.n_dense, n, M), np.float64) if res.n_dense else None,
        trace=_as_np(res.trace, (res.trace_len, 5), np.float64) if res.trace else None,
        log=format_log(res, 17) if res.trace else "",
        phase_ms=dict(zip(PHASES, list(res.phase_ms))))


def format_log(res: Result, precision: int = 6) -> str:
    """rehuel_result_format_log: the per-attempt log in the reference's words."""
    need = lib.rehuel_result_format_log(C.byref(res), precision, None, 0)
    buf = C.create_string_buffer(need + 1)
    lib.rehuel_result_format_log(C.byref(res), precision, buf, need + 1)
    return buf.value.decode()


def format_timings(res: Result, method_name: str) -> str:
    need = lib.rehuel_result_format_timings(C.byref(res), method_name.encode(), None, 0)
    buf = C.create_string_buffer(need + 1)
    lib.rehuel_result_format_timings(C.byref(res), method_name.encode(), buf, need + 1)
    return buf.value.decode()


def irk_ensemble_continue(problem: int, method: int, t1: float, params, opts: Options, prev: Result,
                          n_samples_cap: int = 0, n_gpus: int = 1, n_hint: int = 0) -> Result:
    """rehuel_irk_ensemble_continue: carry a stopped ensemble (raw result of a keep_state run) on to t1.
    Returns the RAW result of the new leg (free it with result_free; merge legs with result_merge)."""
    n, p = problem_dims(problem, n_hint or prev.n)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(p, -1) if p else np.zeros((1, 1))
    res = Result()
    check(lib.rehuel_irk_ensemble_continue(problem, method, t1, n_hint or n, params.ctypes.data_as(_PD), C.byref(opts),
                                           n_samples_cap, n_gpus, C.byref(prev), C.byref(res)))
    return res


def result_merge(a: Result, b: Result) -> None:
    check(lib.rehuel_result_merge(C.byref(a), C.byref(b)))


def result_free(res: Result) -> None:
    lib.rehuel_result_free(C.byref(res))


def irk_single(problem: int, method: int, t0: float, t1: float, dt0: float, y0, params, opts: Options,
               cap: int = 1 << 16, n_hint: int = 0):
    """irk::odeint for one system: returns (t_vals, y_vals[n_out, n], err_vals, counters, accept_frac, status)."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    n, p = problem_dims(problem, n_hint or y0.shape[0])
    params = np.ascontiguousarray(params if p else [0.0], dtype=np.float64)
    t = np.zeros(cap); y = np.zeros((cap, n)); e = np.zeros(cap)
    cnt = (C.c_ulonglong * 9)()
    n_out = C.c_size_t(); af = C.c_double(); st = C.c_int()
    pd = lambda a: a.ctypes.data_as(_PD)
    check(lib.rehuel_irk_single(problem, method, t0, t1, dt0, n_hint or n, pd(y0), pd(params), C.byref(opts), cap,
                                C.byref(n_out), pd(t), pd(y), pd(e), cnt, C.byref(af), C.byref(st)))
    L = min(n_out.value, cap)
    names = ("attempt", "reject_newton", "reject_err", "newton_success", "newton_incr_diverge",
             "newton_iter_error_too_large", "newton_maxit_exceed", "fun_evals", "jac_evals")
    return dict(t=t[:L].copy(), y=y[:L].copy(), err=e[:L].copy(), n_total=n_out.value,
                counters={k: int(cnt[i]) for i, k in enumerate(names)}, accept_frac=af.value, status=st.value)
