"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the two CPU checkers.

* ``RefSolver("reference")`` -> oracle/_ref/librehuel_ref.so: the UNMODIFIED reference solver
  (irk.hpp / irk.cpp / newton.cpp compiled from /root/reference against oracle/arma_shim).
* ``RefSolver("port")``      -> oracle/libirk_port.so: our plain-C restatement (oracle/irk_port.c).

Both export the same entry points, so tests can run the same case through either.  May be
imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs; the product package rehuel_b200 never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATHS = {
    "reference": os.path.join(HERE, "_ref", "librehuel_ref.so"),
    "port": os.path.join(HERE, "libirk_port.so"),
}

# problem ids shared by ref_driver.cpp and irk_port.c (NOT the product's ids)
ROBERTSON, VDPOL, BRUSS1D, ROBER_HARDCODED, EXPONENTIAL, STIFF_EQ, BRUSS2 = 1, 2, 3, 4, 5, 6, 7
LORENZ, HARMONIC, DIMER, KINETIC4, THREE_BODY = 8, 9, 10, 11, 12
PROBLEM_IDS = {
    "robertson": ROBERTSON, "van-der-pol": VDPOL, "brusselator-1d": BRUSS1D,
    "robertson-hardcoded": ROBER_HARDCODED, "exponential": EXPONENTIAL,
    "stiff-equation": STIFF_EQ, "brusselator": BRUSS2, "lorenz": LORENZ, "harmonic": HARMONIC, "dimer": DIMER,
    "kinetic-4": KINETIC4, "three-body": THREE_BODY,
}
COUNTER_NAMES = ("attempt", "reject_newton", "reject_err", "newton_success", "newton_incr_diverge",
                 "newton_iter_error_too_large", "newton_maxit_exceed", "fun_evals", "jac_evals")


class RefOptions(C.Structure):
    """Field-for-field twin of ``ref_options`` (ref_driver.cpp) == ``rehuel_options``."""
    _fields_ = [
        ("rel_tol", C.c_double), ("abs_tol", C.c_double), ("max_dt", C.c_double),
        ("max_steps", C.c_longlong), ("adaptive_step_size", C.c_int), ("out_interval", C.c_int),
        ("newton_tol", C.c_double), ("newton_dx_delta", C.c_double),
        ("newton_maxit", C.c_int), ("newton_refresh_jac", C.c_int),
        ("output_mode", C.c_int), ("output_interval", C.c_ulonglong),
    ]


def make_options(rel_tol=1e-4, abs_tol=None, max_dt=0.0, max_steps=-1, adaptive=True, out_interval=0,
                 newton_tol=None, newton_dx_delta=1e-4, newton_maxit=5000, newton_refresh_jac=10,
                 output_mode=1, output_interval=1) -> RefOptions:
    """Defaults = options.hpp:49-58, irk.hpp:103-107, newton.hpp:58-60; newton_tol defaults to
    0.1*min(rtol,atol) as the example driver sets it (examples/example_equations.cpp:220)."""
    if abs_tol is None:
        abs_tol = 10 * rel_tol
    if newton_tol is None:
        newton_tol = 0.1 * min(rel_tol, abs_tol)
    return RefOptions(rel_tol, abs_tol, max_dt, max_steps, int(adaptive), out_interval, newton_tol,
                      newton_dx_delta, newton_maxit, newton_refresh_jac, output_mode, output_interval)


@dataclass
class EnsembleResult:
    y: np.ndarray            # [n, M]
    t_final: np.ndarray      # [M]
    err_last: np.ndarray     # [M]
    status: np.ndarray       # [M] int32
    counters: dict           # name -> [M] uint64
    accept_frac: np.ndarray  # [M]
    elapsed_ms: float
    traj_t: np.ndarray | None = None
    traj_y: np.ndarray | None = None
    traj_err: np.ndarray | None = None
    log: str = ""
    extra: dict = field(default_factory=dict)


def _p(a, ty=C.c_double):
    return a.ctypes.data_as(C.POINTER(ty))


class RefSolver:
    def __init__(self, kind: str = "reference"):
        self.kind = kind
        path = LIB_PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (needs /root/reference for kind='reference')")
        self.lib = C.CDLL(path)
        self.lib.ref_last_error.restype = C.c_char_p
        self.lib.ref_method_to_name.restype = C.c_char_p

    # ------------------------------------------------------------------ tableaux
    def get_coefficients(self, method: int):
        A = np.zeros(49); b = np.zeros(7); c = np.zeros(7); b2 = np.zeros(7)
        d = np.zeros(7); d2 = np.zeros(7)
        gamma = C.c_double(); o1 = C.c_int(); o2 = C.c_int()
        s = self.lib.ref_get_coefficients(C.c_int(method), _p(A), _p(b), _p(c), _p(b2), C.byref(gamma),
                                          C.byref(o1), C.byref(o2), _p(d), _p(d2))
        if s <= 0:
            return None
        return dict(s=s, A=A[:s * s].reshape(s, s).copy(), b=b[:s].copy(), c=c[:s].copy(), b2=b2[:s].copy(),
                    gamma=gamma.value, order=o1.value, order2=o2.value, d=d[:s].copy(), d2=d2[:s].copy())

    def name_to_method(self, name: str) -> int:
        return self.lib.ref_name_to_method(name.encode())

    def method_to_name(self, m: int) -> str:
        return self.lib.ref_method_to_name(C.c_int(m)).decode()

    def project_b(self, method: int, theta: float):
        out = np.zeros(7)
        s = self.lib.ref_project_b(C.c_int(method), C.c_double(theta), _p(out))
        return out[:s].copy()

    # ------------------------------------------------------------------ ensemble
    def solve(self, problem, method, t0, t1, dt0, y0, params, opts: RefOptions, M=None,
              start=0, stride=1, traj_cap=0, want_log=False, n_hint=0) -> EnsembleResult:
        pid = PROBLEM_IDS[problem] if isinstance(problem, str) else problem
        n = C.c_int(); P = C.c_int()
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        if n_hint == 0 and pid == BRUSS1D:
            n_hint = y0.shape[0]
        if self.lib.ref_problem_dims(pid, n_hint, C.byref(n), C.byref(P)) != 0:
            raise ValueError("bad problem")
        n, P = n.value, P.value
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(P, -1) if P else np.zeros((0, 0))
        if M is None:
            M = params.shape[1] if P else (y0.shape[1] if y0.ndim == 2 else 1)
        broadcast = int(y0.ndim == 1)
        assert y0.shape[0] == n and (broadcast or y0.shape[1] == M)
        assert (not P) or params.shape == (P, M)
        y = np.full((n, M), np.nan); t_final = np.full(M, np.nan); err_last = np.full(M, np.nan)
        status = np.full(M, -1, dtype=np.int32)
        counters = np.zeros((9, M), dtype=np.uint64); acc = np.full(M, np.nan)
        el = C.c_double(0.0)
        tl = C.c_size_t(0)
        tt = np.zeros(max(traj_cap, 1)); ty = np.zeros((max(traj_cap, 1), n)); te = np.zeros(max(traj_cap, 1))
        log_cap = (1 << 22) if want_log else 0
        log_buf = C.create_string_buffer(max(log_cap, 1))
        dummy = np.zeros(1)
        rc = self.lib.ref_irk_ensemble(
            C.c_int(pid), C.c_int(method), C.c_double(t0), C.c_double(t1), C.c_double(dt0),
            C.c_size_t(M), C.c_int(n_hint), _p(y0), C.c_int(broadcast), _p(params if P else dummy), C.byref(opts),
            C.c_size_t(start), C.c_size_t(stride),
            _p(y), _p(t_final), _p(err_last), _p(status, C.c_int), _p(counters, C.c_ulonglong), _p(acc),
            C.byref(el), C.c_size_t(traj_cap), C.byref(tl), _p(tt), _p(ty), _p(te),
            log_buf, C.c_size_t(log_cap))
        if rc != 0:
            raise RuntimeError(f"ref_irk_ensemble rc={rc}: {self.lib.ref_last_error().decode()}")
        res = EnsembleResult(y, t_final, err_last, status,
                             {k: counters[i] for i, k in enumerate(COUNTER_NAMES)}, acc, el.value)
        if traj_cap:
            L = min(tl.value, traj_cap)
            res.traj_t, res.traj_y, res.traj_err = tt[:L].copy(), ty[:L].copy(), te[:L].copy()
            res.extra["traj_len"] = tl.value
        if want_log:
            res.log = log_buf.value.decode(errors="replace")
        return res

    # ------------------------------------------------------------------ the 4-argument odeint (reference library only)
    def odeint_defaults(self, problem, t0, t1, y0, params, n_hint=0):
        """irk::odeint(func, t0, t1, y0), irk.hpp:915-926: the reference picks every default itself.  One system."""
        assert self.kind == "reference", "only the unmodified reference knows its own defaults"
        pid = PROBLEM_IDS[problem] if isinstance(problem, str) else problem
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        params = np.ascontiguousarray(params if len(params) else [0.0], dtype=np.float64)
        y = np.full(y0.shape[0], np.nan); tf = np.zeros(1); el = np.zeros(1)
        st = np.full(1, -1, dtype=np.int32); cnt = np.zeros(9, dtype=np.uint64)
        rc = self.lib.ref_odeint_defaults(C.c_int(pid), C.c_int(n_hint), C.c_double(t0), C.c_double(t1), _p(y0), _p(params),
                                          _p(y), _p(tf), _p(el), _p(st, C.c_int), _p(cnt, C.c_ulonglong))
        if rc != 0:
            raise RuntimeError(f"ref_odeint_defaults rc={rc}: {self.lib.ref_last_error().decode()}")
        return EnsembleResult(y[:, None], tf, el, st, {k: cnt[i:i + 1] for i, k in enumerate(COUNTER_NAMES)}, np.full(1, np.nan), 0.0)

    # ------------------------------------------------------------------ one Newton solve (KAT)
    def newton_solve_stages(self, problem, method, variant, params, y, t, dt, maxit, refresh_jac, xtol, Rtol,
                            n_hint=0):
        pid = PROBLEM_IDS[problem] if isinstance(problem, str) else problem
        y = np.ascontiguousarray(y, dtype=np.float64)
        n = y.shape[0]
        params = np.ascontiguousarray(params if params is not None else [0.0], dtype=np.float64)
        Y = np.zeros(7 * n); y1 = np.zeros(n)
        iters = C.c_int(); res = C.c_double(); fe = C.c_ulonglong(); je = C.c_ulonglong()
        st = self.lib.ref_newton_solve_stages(
            C.c_int(pid), C.c_int(method), C.c_int(variant), C.c_int(n_hint or n), _p(params), _p(y),
            C.c_double(t), C.c_double(dt), C.c_int(maxit), C.c_int(refresh_jac), C.c_double(xtol), C.c_double(Rtol),
            _p(Y), _p(y1), C.byref(iters), C.byref(res), C.byref(fe), C.byref(je))
        if st < 0:
            raise RuntimeError(f"ref_newton_solve_stages rc={st}: {self.lib.ref_last_error().decode()}")
        s = self.get_coefficients(method)["s"]
        return dict(status=st, Y=Y[:s * n].copy(), y1=y1, iters=iters.value, res=res.value,
                    fun_evals=fe.value, jac_evals=je.value)

    # ------------------------------------------------------------------ fixed-step oracle
    def fixed_step(self, problem, method, params, y, t0, dt, nsteps, maxit=100, refresh_jac=25, xtol=1e-14,
                   Rtol=1e-14, n_hint=0):
        """Loop of newton_solve_stages<F,false,true> + y += YYs*d_weights (test/irk.cpp:271-286)."""
        pid = PROBLEM_IDS[problem] if isinstance(problem, str) else problem
        y = np.ascontiguousarray(y, dtype=np.float64)
        params = np.ascontiguousarray(params if params is not None else [0.0], dtype=np.float64)
        out = np.zeros_like(y)
        its = C.c_ulonglong()
        failed = self.lib.ref_fixed_step(
            C.c_int(pid), C.c_int(method), C.c_int(n_hint or y.shape[0]), _p(params), _p(y), C.c_double(t0),
            C.c_double(dt), C.c_longlong(nsteps), C.c_int(maxit), C.c_int(refresh_jac), C.c_double(xtol),
            C.c_double(Rtol), _p(out), C.byref(its))
        if failed < 0:
            raise RuntimeError(f"ref_fixed_step rc={failed}")
        return dict(y=out, failed=failed, newton_iters=its.value)

    def eval_functor(self, problem, t, y, params, n_hint=0):
        """``fun(t, y)`` and ``jac(t, y)`` (row-major n x n) of a test functor as the reference writes them."""
        pid = PROBLEM_IDS[problem] if isinstance(problem, str) else problem
        y = np.ascontiguousarray(y, dtype=np.float64)
        params = np.ascontiguousarray(params if params is not None else [0.0], dtype=np.float64)
        n = y.shape[0]
        f = np.zeros(n)
        J = np.zeros((n, n))
        rc = self.lib.ref_eval_functor(C.c_int(pid), C.c_int(n_hint or n), _p(params), C.c_double(t), _p(y), _p(f), _p(J))
        if rc:
            raise RuntimeError(f"ref_eval_functor rc={rc}")
        return f, J
