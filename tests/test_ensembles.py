"""Host logic: synthetic ensemble generators and member sharding (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from rehuel_b200 import ensembles


def _splitmix64_scalar(seed, count):
    mask = (1 << 64) - 1
    x = seed
    out = []
    for _ in range(count):
        x = (x + 0x9E3779B97F4A7C15) & mask
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        z = z ^ (z >> 31)
        out.append((z >> 11) * 2.0 ** -53)
    return np.array(out)


def test_splitmix64_vectorised_equals_sequential_stream():
    ref = _splitmix64_scalar(0x5EED, 64)
    assert np.array_equal(ensembles.splitmix64_uniform(0, 64), ref)
    assert np.array_equal(ensembles.splitmix64_uniform(10, 20), ref[10:30])  # random access == stream


def test_robertson_sweep_is_index_addressable_and_in_range():
    P = ensembles.robertson_params(1000)
    assert P.shape == (3, 1000)
    base = np.array([0.04, 1e4, 3e7])[:, None]
    assert np.all(P >= base * 10 ** -0.5) and np.all(P < base * 10 ** 0.5)
    assert np.array_equal(ensembles.robertson_params(100, start=300), P[:, 300:400])
    u = ensembles.splitmix64_uniform(0, 3)
    np.testing.assert_array_equal(P[:, 0], np.array([0.04, 1e4, 3e7]) * 10.0 ** (u - 0.5))


def test_vdpol_and_brusselator_sweeps():
    P = ensembles.vdpol_params(11)
    assert P.shape == (1, 11) and P[0, 0] == 1.0 and abs(P[0, -1] - 1e-6) < 1e-20
    assert np.array_equal(ensembles.vdpol_params(4, start=3, total=11), P[:, 3:7])
    B = ensembles.bruss1d_params(50)
    assert np.all((B[0] >= 0.01) & (B[0] <= 0.05)) and np.all(B[1] == 1.0) and np.all((B[2] >= 2.5) & (B[2] <= 3.5))
    y0 = ensembles.bruss1d_y0(32)
    assert y0.shape == (64,) and np.all(y0[1::2] == 3.0)


def test_flop_model_matches_survey():
    fixed, extra, per = ensembles.flops_per_attempt_dense_model(3, 3, 11, 8)
    # Robertson/RADAU_IIA_53 = 873 + 321 m with formula 8.19 only
    assert round(fixed) == 873 and per == 321 and extra == 41


def test_sharding_world_size_2_gloo():
    """Two gloo ranks shard an ensemble, each 'solves' its slice (CPU stand-in for the kernel: the
    C port), and the all-reduced statistics equal the single-process ones."""
    script = os.path.join(ROOT, "tests", "_gloo_shard_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD-OK" in outs[0]


def test_executed_flop_model_of_the_thread_kernel():
    """rehuel_b200/flopmodel.py: the counts DESIGN.md section 6 quotes, and the counter identity behind `total`."""
    from rehuel_b200 import flopmodel as fm
    rob = fm.thread_kernel(3, 211, 11, 8, zero_col0_rows=(2,))
    assert (rob.S, rob.fixed, rob.tail, rob.per_iteration, rob.extra_820, rob.residual) == (3, 118, 88, 340, 39, 134)
    vdp = fm.thread_kernel(2, 230, 6, 7)
    assert vdp.per_iteration == 212 and vdp.extra_820 == 6          # LOBATTO_IIIC_43: E = I, no estimator solve
    # one member, 3 attempts, one of them rejected, 2 + 1 + 2 Newton iterations, every final residual booked:
    # fun_evals = S * (attempts + iterations) + newton_success + uses of 8.20 (first attempt + the one after the rejection)
    att, its, nok, rej = 3, 5, 3, 1
    fun = 3 * (att + its) + nok + (1 + rej)
    total = rob.total(1, att, nok, rej, fun)
    assert total == rob.fixed * 3 + rob.per_iteration * 5 + rob.extra_820 * 2 - rob.residual * 3
    # a failed Newton solve has no tail
    assert rob.total(1, att + 1, nok, rej, fun + 3 * (1 + 4)) == total + (rob.fixed - rob.tail) + 4 * rob.per_iteration
    # ... and what it books is not what a lane with a non-finite iterate executes: with the iterations of the SUCCESSFUL
    # solves given, the failed solve counts with its per-attempt part alone (a lower bound)
    assert rob.total(1, att + 1, nok, rej, fun + 3 * (1 + 4999), newton_iters=its) == total + (rob.fixed - rob.tail)
    band = fm.band_kernel(64, 2, 232, 10, 8)
    dense = fm.dense_kernel(64, 232, 10, 64 * 8)
    assert dense.fixed > 50 * band.fixed                            # three dense 64 x 64 LUs against band LUs
