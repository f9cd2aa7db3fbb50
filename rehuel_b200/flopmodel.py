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

    def total(self, members, attempt, newton_success, reject_err, fun_evals, newton_iters=None):
        """Lower bound on the executed flops of a run, from the SUMS of the counters the kernels return.
          uses of 8.20      = members + reject_err   (first attempt, and the first estimate after every error rejection:
                                                      irk.hpp:587, 763, 845)
          residuals booked  = (fun_evals - newton_success - uses_820) / S      (irk.hpp:443-446, 469-472, 702, 728)
          Newton iterations = residuals booked - attempt   (every attempt starts with one; failed solves included)
        Every successful solve is assumed to skip its final residual (see the module text).

        A FAILED solve books the evaluations the reference would have spent (irk.hpp:458-488 spins to maxit), but a lane
        whose iterate has become non-finite stops early (irk_thread.cuh): booked is not executed.  For runs with Newton
        failures pass `newton_iters` (the counter: iterations of the SUCCESSFUL solves only); failed solves then count with
        their per-attempt part alone, which keeps the total a lower bound."""
        uses = members + reject_err
        iters = (fun_evals - newton_success - uses) / self.S - attempt if newton_iters is None else newton_iters
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


def _lu_real_zc(n, zc_rows):
    """lu_factor<false, ZCOL0>(LuReal): rows whose entry in column 0 is structurally zero skip the first step."""
    f = _lu_dense(n, False)
    return f - len(zc_rows) * (1 + 2 * (n - 1))


def _lu_cplx_diagim(n, zc_rows):
    """lu_factor<false, ZCOL0, true>(LuCplx): in the first elimination step the imaginary part is mi on the diagonal and
    zero elsewhere: multipliers are two real multiplications, updates one FMA + one FMA (diagonal) or one
    multiplication (off-diagonal); rows with a structurally zero entry in column 0 skip the step."""
    rows = n - 1 - len(zc_rows)
    first = 6 + rows * (2 + (n - 1) * 2 + 2 + (n - 2) * 1)  # pivot; per row: multiplier, re updates (FMA), im updates
    rest = 0.0
    for k in range(1, n):
        r = n - k - 1
        rest += 6 + 6 * r + 8 * r * r
    return first + rest


def thread_kernel(n, method, f_fun, f_jac, autonomous=True, zero_col0_rows=()) -> AttemptFlops:
    """ensemble_irk_thread (n <= 4) as built (irk_thread.cuh): dense n x n LUs of the decoupled blocks in registers,
    unpivoted path.  Counted from the source with its default switches: structural zeros (REHUEL_STRUCT_ZEROS: diagonal
    imaginary part in the first elimination step, real row 0 of U in the complex solves, the functor's kJacZeroMask --
    `zero_col0_rows`: Robertson (2,)), canonical tableau (d = e_s: no dy product; unit last row of T: additions),
    peeled first iteration (rank-one right-hand side, Y = dY), step-size proposal with one k-th root (counted as 2:
    one root, one reciprocal-free power, by the rule that div/sqrt/pow count 1)."""
    S, NR, NC, est = STRUCTURE[method]
    zc = tuple(zero_col0_rows)
    lu = 1  # 1/dt
    lu += NR * (_lu_real_zc(n, zc) + n + 1)          # + diagonal shifts, mu_r
    lu += NC * (_lu_cplx_diagim(n, zc) + n + 2)      # + diagonal shifts, mu_c (two products)
    if est == EST_OWN_LU:
        lu += _lu_real_zc(n, zc) + n + 1
    solve_r = _solve_dense(n, False) - 2 * len(zc)                     # forward term (i, 0) skipped
    solve_c = _solve_dense(n, True) - 4 * (n - 1) - 8 * len(zc)        # real row 0 of U; forward term (i, 0) skipped
    est_solve = 0 if est == EST_IDENTITY else solve_r
    # tail: dalt (S n FMA); gam; yn; gam*f + dalt (FMA) - dy; E^-1 and scaling; norm (abs, abs, max, FMA, rcp, mul, FMA per
    # component; mean); controller: root 2, ratio 5, fac 4, new_dt 2
    tail = 2 * S * n + 1 + n
    if est != EST_IDENTITY:
        tail += 2 * n + (0 if autonomous else f_fun)
    tail += n + est_solve + n + 9 * n + 1 + 13
    resid = _residual(n, S, f_fun, S) + (S - 1)                         # + the partial sums of |R|^2
    units = NR + NC                                                     # unit entries in the last row of T
    transforms_T = 2 * S * n * (S - 1) + (units - 1) * n                # T (x) I with the canonical last row
    solves = NR * (solve_r + n) + NC * (solve_c + 6 * n)                # + the -mu scalings
    per_it = 2 * S * S * n + solves + transforms_T + 2 * S * n + (S - 1) + S * n + resid
    if autonomous:   # first iteration peeled: Z = (-dt TiAsum_i) f(y): S + S n products; Y = dY
        first = f_fun - (2 * S * S * n - (S + S * n)) - S * n
    else:
        first = resid
    fixed = f_jac + lu + first + tail
    extra = n + (f_fun + 2 * n if est != EST_IDENTITY else 0) + n + est_solve + n
    return AttemptFlops(S, fixed, tail, per_it, extra, resid)


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
