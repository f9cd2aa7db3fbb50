"""Synthetic parameter-swept ensembles.  Host-side, numpy only.

Parameters are a pure function of the member index: draw k of the ensemble is
``splitmix64(seed + (k+1)*GOLDEN)`` mapped to [0,1) by ``(z >> 11) * 2**-53`` -- the counter
form of the splitmix64 stream seeded with ``seed``, so any rank can generate its own slice.
"""
from __future__ import annotations

import numpy as np

SEED = 0x5EED
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64_uniform(first_draw: int, count: int, seed: int = SEED) -> np.ndarray:
    """Draws first_draw .. first_draw+count-1 of the splitmix64 stream as doubles in [0,1)."""
    with np.errstate(over="ignore"):
        k = np.arange(first_draw + 1, first_draw + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def robertson_params(M: int, start: int = 0, seed: int = SEED) -> np.ndarray:
    """Config 2: k1 = 0.04*10^(u1-1/2), k2 = 1e4*10^(u2-1/2), k3 = 3e7*10^(u3-1/2); three draws
    per member in index order.  Returns [3, M] for members start .. start+M-1."""
    u = splitmix64_uniform(3 * start, 3 * M, seed).reshape(M, 3)
    base = np.array([0.04, 1e4, 3e7])
    return np.ascontiguousarray((base[None, :] * 10.0 ** (u - 0.5)).T)


ROBERTSON_Y0 = np.array([1.0, 0.0, 0.0])  # examples/example_equations.cpp:336
VDPOL_Y0 = np.array([2.0, 0.0])           # examples/example_equations.cpp:367


def vdpol_params(M: int, start: int = 0, total: int | None = None) -> np.ndarray:
    """Config 3: mu_ref = 10^(-6 i/(total-1)) (the reference's SMALL parameter; classical
    mu = 1 .. 1e3).  Returns [1, M]."""
    total = M if total is None else total
    i = np.arange(start, start + M, dtype=np.float64)
    return (10.0 ** (-6.0 * i / max(total - 1, 1)))[None, :].copy()


def bruss1d_params(M: int, start: int = 0, seed: int = SEED) -> np.ndarray:
    """Config 4: alpha in [0.01,0.05], A = 1, B in [2.5,3.5]; two draws per member."""
    u = splitmix64_uniform(2 * start, 2 * M, seed ^ 0xB055).reshape(M, 2)
    out = np.empty((3, M))
    out[0] = 0.01 + 0.04 * u[:, 0]
    out[1] = 1.0
    out[2] = 2.5 + u[:, 1]
    return out


def bruss1d_y0(Ng: int) -> np.ndarray:
    """u_i(0) = 1 + sin(2 pi x_i), v_i(0) = 3, x_i = i/(Ng+1), interleaved (u_i, v_i)."""
    x = np.arange(1, Ng + 1) / (Ng + 1.0)
    y = np.empty(2 * Ng)
    y[0::2] = 1.0 + np.sin(2.0 * np.pi * x)
    y[1::2] = 3.0
    return y


def flops_per_attempt_dense_model(n: int, s: int, f_fun: int, f_jac: int):
    """flops(attempt) = fixed + per_newton * m for the reference's dense
    algorithm.  Returns (fixed_with_8_19_only, extra_for_8_20, per_newton)."""
    N = n * s
    c_r = s * f_fun + s * n + 2 * s * s * n + 3 * N
    fixed = (f_jac + s * s + s * s * n * n + N + (2.0 / 3.0) * N ** 3) + c_r + (4 * N + f_fun + 5 * n) \
        + (n * n + n + (2.0 / 3.0) * n ** 3 + 2 * n * n + n) + (7 * n + 3) + 20
    extra_820 = f_fun + 4 * n + 2 * n * n
    per_newton = 2 * N * N + 4 * N + c_r
    return fixed, extra_820, per_newton


def mixed_tolerances(M: int, start: int = 0, seed: int = SEED) -> np.ndarray:
    """Config 5: rtol_i = atol_i = 10^(-3 - 6 v_i), v_i hashed from the member index (one draw per
    member, stream seed ^ 0x701)."""
    v = splitmix64_uniform(start, M, seed ^ 0x701)
    return 10.0 ** (-3.0 - 6.0 * v)
