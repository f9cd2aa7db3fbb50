#!/usr/bin/env python
"""Generate tests/golden/config_cases.json: members of the BASELINE configs whose oracle run is too
slow to repeat inside a test, integrated once by the UNMODIFIED reference (oracle/_ref), plus the
reference's own per-attempt log (irk.hpp:575-577, 805-810) for two members.

  * config 4: 16 members spread over the 65 536-member Brusselator-MOL sweep
    (n = 64, LOBATTO_IIIC_85, t in [0,10], rtol = atol = 1e-6): ~50 s each through the reference's
    dense 320 x 320 LU.
  * config 5: 16 members of the 10 485 760-member mixed Robertson / Van der Pol ensemble at the
    extreme ends of the tolerance sweep, two per octant of the index range, each run with ITS
    rtol = atol = 10^(-3-6v), newton tol = 0.1 rtol.
  * trace: the "    Rehuel: step t dt err iters" log of one Robertson and one Van der Pol member.

Floats are C99 hex strings (exact round trip).  Run (only where /root/reference exists):
    make -C oracle && python tools/make_golden_configs.py
"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refsolver import RefSolver, make_options  # noqa: E402


def _ensembles():
    """rehuel_b200/ensembles.py loaded by path: the host-side parameter generators, without importing the product."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ensembles", os.path.join(ROOT, "rehuel_b200", "ensembles.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


ensembles = _ensembles()
CONFIG4_M = 1 << 16
CONFIG5_M = 10485760


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def config4_indices():
    """every 6th of the 96 members the full-size GPU test re-runs alone"""
    return np.arange(0, CONFIG4_M, CONFIG4_M // 96)[::6][:16]


def config5_indices():
    """per octant of the index range: the member with the tightest and the one with the loosest tolerance, problem
    classes alternating (even index = Robertson, odd = Van der Pol)"""
    tol = ensembles.mixed_tolerances(CONFIG5_M)
    idx = []
    for k in range(8):
        lo, hi = CONFIG5_M * k // 8, CONFIG5_M * (k + 1) // 8
        seg = np.arange(lo, hi)
        for want_min, parity in ((True, k % 2), (False, 1 - k % 2)):
            cand = seg[seg % 2 == parity]
            t = tol[cand]
            idx.append(int(cand[np.argmin(t) if want_min else np.argmax(t)]))
    return idx


def run_config4(i):
    ref = RefSolver("reference")
    P = ensembles.bruss1d_params(1, start=int(i))
    r = ref.solve("brusselator-1d", 232, 0.0, 10.0, 1e-6, ensembles.bruss1d_y0(32), P, make_options(1e-6, 1e-6), M=1)
    return dict(kind="config4", index=int(i), params=hx(P[:, 0]), y=hx(r.y[:, 0]), status=int(r.status[0]),
                t_final=hx(r.t_final), counters={k: int(v[0]) for k, v in r.counters.items()}, seconds=r.elapsed_ms * 1e-3)


def run_config5(i):
    ref = RefSolver("reference")
    tol = float(ensembles.mixed_tolerances(1, start=int(i))[0])
    half = CONFIG5_M // 2
    j = i // 2
    if i % 2 == 0:
        prob, method, t1, y0 = "robertson", 211, 40.0, [1.0, 0.0, 0.0]
        P = ensembles.robertson_params(1, start=j)
    else:
        prob, method, t1, y0 = "van-der-pol", 230, 2.0, [2.0, 0.0]
        P = ensembles.vdpol_params(1, start=j, total=half)
    r = ref.solve(prob, method, 0.0, t1, 1e-6, np.array(y0), P, make_options(tol, tol, newton_tol=0.1 * tol), M=1)
    return dict(kind="config5", index=int(i), problem=prob, method=method, t1=t1, y0=y0, tol=float(tol).hex(),
                params=hx(P[:, 0]), y=hx(r.y[:, 0]), status=int(r.status[0]), t_final=hx(r.t_final),
                counters={k: int(v[0]) for k, v in r.counters.items()})


def run_trace(case):
    prob, par, y0, method, t1, out_interval = case
    ref = RefSolver("reference")
    o = make_options(1e-6, 1e-6, out_interval=out_interval)
    r = ref.solve(prob, method, 0.0, t1, 1e-6, np.array(y0), np.array(par)[:, None], o, M=1, want_log=True)
    return dict(kind="trace", problem=prob, params=par, y0=y0, method=method, t1=t1, out_interval=out_interval,
                log=r.log, counters={k: int(v[0]) for k, v in r.counters.items()})


def _job(j):
    return {"config4": run_config4, "config5": run_config5, "trace": run_trace}[j[0]](j[1])


def main():
    jobs = [("config4", int(i)) for i in config4_indices()] + [("config5", int(i)) for i in config5_indices()]
    jobs += [("trace", ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 211, 40.0, 1)),
             ("trace", ("van-der-pol", [1e-5], [2.0, 0.0], 230, 2.0, 1)),
             ("trace", ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 211, 1e11, 7))]
    with ProcessPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        cases = list(ex.map(_job, jobs))
    out = {"generator": "tools/make_golden_configs.py",
           "source": "oracle/_ref/librehuel_ref.so (unmodified reference over oracle/arma_shim)", "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "config_cases.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print(f"wrote {path}: {len(cases)} cases, {os.path.getsize(path)} bytes")
    for c in cases:
        if c["kind"] == "config4":
            print("config4", c["index"], c["counters"]["attempt"], f"{c['seconds']:.1f} s")
        elif c["kind"] == "config5":
            print("config5", c["index"], c["problem"], float.fromhex(c["tol"]), c["counters"]["attempt"], c["counters"]["reject_newton"])


if __name__ == "__main__":
    main()
