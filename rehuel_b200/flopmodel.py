"""Executed-flop models of the integration kernels: the arithmetic the kernels
actually perform, counted from their source.  Counting rule: FMA = 2, every other
FP64 operation (add, mul, reciprocal, sqrt, max, abs) = 1; a complex multiply-add is 4 FMA = 8.

Every model is ``fixed_per_attempt + per_newton_iteration * iterations (+ extra per use of formula
8.20) - skipped_residual per successful Newton solve`` so that the total comes from the counters the
kernels return (attempt, newton_iters, newton_success, fun_evals).  The kernels skip the final
residual evaluation of a solve whose step test already holds (irk_thread.cuh); how often that
happens is not among the counters, so the totals below assume EVERY successful solve skips it: a
lower bound on the executed flops (the share is 96-99.9 % on the Robertson sweeps).

Host-side arithmetic only; nothing here touches the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

EST_IDENTITY, EST_REUSE, EST_OWN_LU = 0, 1, 2

# method id -> (stages, real blocks, complex blocks, estimator mode); the kernels' template arguments (launch.cuh)
STRUCTURE = {200: (1, 1, 0, EST_IDENTITY), 210: (2, 0, 1, EST_OWN_LU), 211: (3, 1, 1, EST_REUSE), 212: (5, 1, 2, EST_REUSE),
             213: (7, 1, 3, EST_REUSE), 230: (3, 1, 1, EST_IDENTITY), 231: (4, 0, 2, EST_IDENTITY), 232: (5, 1, 2, EST_OWN_LU)}


@dataclass
class AttemptFlops:
    S: int                # stages (function evaluations per residual)
    fixed: float          # per attempt with a successful Newton solve (8.19 only)
    tail: float           # the part of `fixed` after the Newton solve (update, estimate, norm, controller)
    per_iteration: float  # per Newton iteration
    extra_820: float      # per use of formula 8.20
    residual: float       # one residual evaluation (what a skipped final residual saves)

    def total(self, members, attempt, newton_success, reject_err, fun_evals):
        """Lower bound on the executed flops of a run, from the SUMS of the counters the kernels return.
          uses of 8.20      = members + reject_err   (first attempt, and the first estimate after every error rejection:
                                                      irk.hpp:587, 763, 845)
          residuals booked  = (fun_evals - newton_success - uses_820) / S      (irk.hpp:443-446, 469-472, 702, 728)
          Newton iterations = residuals booked - attempt   (every attempt starts with one; failed solves included)
        Every successful solve is assumed to skip its final residual (see the module text)."""
        uses = members + reject_err
        iters = (fun_evals - newton_success - uses) / self.S - attempt
        failed = attempt - newton_success
        return (self.fixed * newton_success + (self.fixed - self.tail) * failed + self.per_iteration * iters
                + self.extra_820 * uses - self.residual * newton_success)


def _lu_dense(n, cplx):
    """Right-looking LU with reciprocal pivots (lu_factor, irk_thread.cuh): per step one reciprocal, (n-k-1)
    multipliers, (n-k-1)^2 FMAs; complex: pivot 1/(a+ib) = 6, multiplier 6, update 8."""
    f = 0.0
    for k in range(n):
        r = n - k - 1
        f += (6 + 6 * r + 8 * r * r) if cplx else (1 + r + 2 * r * r)
    return f


def _solve_dense(n, cplx):
    """lu_solve: forward n(n-1)/2 FMAs, backward the same plus n multiplications by the reciprocal pivots."""
    pairs = n * (n - 1)  # forward + backward
    return (8 * pairs + 6 * n) if cplx else (2 * pairs + n)


def _residual(n, S, f_fun, fun_calls):
    # y + Y_i (S n adds), the function evaluations, A (x) I product (S^2 n FMA), R = Y - dt*acc (S n FMA), |R|^2 (S n FMA)
    return S * n + fun_calls * f_fun + 2 * S * S * n + 2 * S * n + 2 * S * n


def _tail(n, S, est_solve, est, f_fun, autonomous):
    """New solution, estimate 8.19, norm, controller (irk.hpp:691-801) as the kernels evaluate them."""
    f = 4 * S * n            # dy, dalt: 2 S n FMA
    f += 1 + n               # gam = gamma*dt, yn = y + dy
    if est != EST_IDENTITY:
        f += 2 * n + (0 if autonomous else f_fun)  # gam*f(t,y) + dalt (FMA); f(t, y) is reused when autonomous
    f += n                   # delta_delta
    f += est_solve + n       # E^-1 (.) and its scaling
    f += 9 * n + 2           # scaled norm: abs, abs, max, FMA, reciprocal, mul, FMA per component; mean, sqrt
    f += 14                  # controller: one root (2), reciprocals and products (irk_thread.cuh: propose_dt)
    return f


def thread_kernel(n, method, f_fun, f_jac, autonomous=True) -> AttemptFlops:
    """ensemble_irk_thread (n <= 4): dense n x n LUs of the decoupled blocks in registers.
    Robertson / RADAU_IIA_53: fixed 259, per iteration 369, 8.20 +30, residual 132."""
    S, NR, NC, est = STRUCTURE[method]
    lu = NR * (_lu_dense(n, False) + n + 1) + NC * (_lu_dense(n, True) + n + 2) + 1  # + diagonal shifts, mu's, 1/dt
    if est == EST_OWN_LU:
        lu += _lu_dense(n, False) + n + 1
    est_solve = 0 if est == EST_IDENTITY else _solve_dense(n, False)
    first = (f_fun + S * (1 + 3 * n)) if autonomous else _residual(n, S, f_fun, S)
    tail = _tail(n, S, est_solve, est, f_fun, autonomous)
    fixed = f_jac + lu + first + tail
    transforms = 2 * S * S * n * 2                  # Ti (x) I and T (x) I products
    solves = NR * (_solve_dense(n, False) + n) + NC * (_solve_dense(n, True) + 6 * n)  # + the -mu scalings
    per_it = transforms + solves + 2 * S * n + S * n + _residual(n, S, f_fun, S)       # |dY|^2, Y += dY
    extra = (n + (f_fun if est != EST_IDENTITY else 0) + (2 * n if est != EST_IDENTITY else 0) + n + est_solve + n)
    return AttemptFlops(S, fixed, tail, per_it, extra, _residual(n, S, f_fun, S))


def _lu_band(n, kb, cplx):
    """Band LU without interchanges (band_twisted.cuh; the twisted sweeps do the same eliminations from both ends plus a
    kb x kb separator): per row one reciprocal pivot, kb multipliers, kb^2 updates."""
    return n * ((6 + 6 * kb + 8 * kb * kb) if cplx else (1 + kb + 2 * kb * kb))


def _solve_band(n, kb, cplx):
    return n * ((8 * 2 * kb + 6) if cplx else (2 * 2 * kb + 1))


def band_kernel(n, kb, method, f_fun_per_component, f_jac_per_row) -> AttemptFlops:
    """ensemble_irk_warp (banded Jacobian, half bandwidth kb): same frame, band factors/solves."""
    S, NR, NC, est = STRUCTURE[method]
    f_fun = n * f_fun_per_component
    mats_real = NR + (1 if est == EST_OWN_LU else 0)
    build = n * f_jac_per_row + (mats_real + NC) * n          # Jacobian band + diagonal shifts
    lu = mats_real * _lu_band(n, kb, False) + NC * _lu_band(n, kb, True)
    est_solve = 0 if est == EST_IDENTITY else _solve_band(n, kb, False)
    tail = _tail(n, S, est_solve, est, f_fun, False)
    fixed = build + lu + _residual(n, S, f_fun, S) + tail
    solves = NR * (_solve_band(n, kb, False) + n) + NC * (_solve_band(n, kb, True) + 6 * n)
    per_it = 4 * S * S * n + solves + 3 * S * n + _residual(n, S, f_fun, S)
    extra = n + (f_fun + 2 * n if est != EST_IDENTITY else 0) + n + est_solve + n
    return AttemptFlops(S, fixed, tail, per_it, extra, _residual(n, S, f_fun, S))


def dense_kernel(n, method, f_fun_per_component, f_jac) -> AttemptFlops:
    """ensemble_irk_cta / ensemble_irk_wdense (dense Jacobian, n > 4): dense n x n LUs of the decoupled blocks."""
    S, NR, NC, est = STRUCTURE[method]
    f_fun = n * f_fun_per_component
    mats_real = NR + (1 if est == EST_OWN_LU else 0)
    lu = mats_real * (_lu_dense(n, False) + n) + NC * (_lu_dense(n, True) + n)
    est_solve = 0 if est == EST_IDENTITY else _solve_dense(n, False)
    tail = _tail(n, S, est_solve, est, f_fun, False)
    fixed = f_jac + lu + _residual(n, S, f_fun, S) + tail
    solves = NR * (_solve_dense(n, False) + n) + NC * (_solve_dense(n, True) + 6 * n)
    per_it = 4 * S * S * n + solves + 3 * S * n + _residual(n, S, f_fun, S)
    extra = n + (f_fun + 2 * n if est != EST_IDENTITY else 0) + n + est_solve + n
    return AttemptFlops(S, fixed, tail, per_it, extra, _residual(n, S, f_fun, S))
