#!/usr/bin/env python
"""Generate tests/golden/oracle_cases.json by running the UNMODIFIED reference (oracle/_ref,
built from /root/reference by oracle/Makefile) in this container.  The fixture pins

  * the CPU port (oracle/irk_port.c) bit for bit   -> tests/test_oracle.py
  * the CUDA path to the stated tolerance           -> tests/test_gpu_parity.py

/root/reference does not exist on the GPU box, hence the committed vectors.  Floats are stored
as C99 hex strings so that they round-trip exactly.

Run (only where /root/reference exists):  make -C oracle && python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refsolver import RefSolver, make_options  # noqa: E402
from rehuel_b200 import ensembles  # noqa: E402  (host-side parameter generators only)


THREE_BODY_Y0 = [1.0, 0.0, 3.0, 2.0, 0.0, 3.0, 0.0, 2.0, -1.0, 0.0, -1.0, 0.0]


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def counters(res, i=0):
    return {k: int(v[i]) for k, v in res.counters.items()}


def main():
    ref = RefSolver("reference")
    out = {"generator": "tools/make_golden.py", "source": "oracle/_ref/librehuel_ref.so (unmodified reference over oracle/arma_shim)",
           "cases": []}
    add = out["cases"].append

    # 1. the reference's own known-answer test, test/irk.cpp:250-296
    for variant in (0, 1):
        k = ref.newton_solve_stages("robertson-hardcoded", 211, variant, None, [1.0, 0.0, 0.0], 0.0, 1.5e-3, 100, 25,
                                    1e-8, 1e-10)
        add(dict(kind="newton_solve_stages", name=f"irk_calc_stages_variant{variant}", problem="robertson-hardcoded",
                 method=211, variant=variant, y=[1.0, 0.0, 0.0], t=0.0, dt=1.5e-3, maxit=100, refresh_jac=25,
                 xtol=1e-8, Rtol=1e-10, status=k["status"], iters=k["iters"], fun_evals=k["fun_evals"],
                 jac_evals=k["jac_evals"], Y=hx(k["Y"]), y1=hx(k["y1"]), res=float(k["res"]).hex()))

    # 2. single systems, headline settings
    o6 = dict(rel_tol=1e-6, abs_tol=1e-6)
    singles = [
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 211, 0.0, 40.0, 1e-6, o6),
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 211, 0.0, 1e11, 1e-6, o6),
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 230, 0.0, 40.0, 1e-6, o6),
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 232, 0.0, 40.0, 1e-6, o6),
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 210, 0.0, 40.0, 1e-6, o6),
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 231, 0.0, 40.0, 1e-6, o6),
        # example-driver defaults (examples/example_equations.cpp:40-43): rtol 1e-5, atol 1e-4, dt 1e-2
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 232, 0.0, 10.0, 1e-2, dict(rel_tol=1e-5, abs_tol=1e-4)),
        ("van-der-pol", [1.0], [2.0, 0.0], 230, 0.0, 2.0, 1e-6, o6),
        ("van-der-pol", [1e-2], [2.0, 0.0], 230, 0.0, 2.0, 1e-6, o6),
        ("van-der-pol", [1e-4], [2.0, 0.0], 230, 0.0, 2.0, 1e-6, o6),
        ("van-der-pol", [1e-6], [2.0, 0.0], 230, 0.0, 2.0, 1e-6, o6),
        ("van-der-pol", [1e-3], [2.0, 0.0], 211, 0.0, 2.0, 1e-6, o6),
        ("exponential", [-0.4], [1.0], 211, 0.0, 2.0, 1e-6, dict(rel_tol=1e-4)),  # test/irk.cpp:218-228
        ("stiff-equation", None, [1.0, 0.0], 211, 0.0, 10.0, 1e-6, o6),
        ("stiff-equation", None, [1.0, 0.0], 232, 0.0, 10.0, 1e-2, dict(rel_tol=1e-5, abs_tol=1e-4)),
        ("brusselator", [1.0, 2.5], [1.0, 1.0], 211, 0.0, 20.0, 1e-6, o6),
        ("brusselator", [1.0, 3.5], [1.0, 1.0], 230, 0.0, 20.0, 1e-6, o6),
        # the decimal-defined high-order Radau tableaux (irk.cpp:393-533)
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 212, 0.0, 40.0, 1e-6, o6),
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 213, 0.0, 40.0, 1e-6, o6),
        ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 212, 0.0, 1e11, 1e-6, o6),
        ("van-der-pol", [1e-3], [2.0, 0.0], 212, 0.0, 2.0, 1e-6, o6),
        ("van-der-pol", [1e-2], [2.0, 0.0], 213, 0.0, 2.0, 1e-6, o6),
        ("brusselator", [1.0, 3.0], [1.0, 1.0], 212, 0.0, 20.0, 1e-6, o6),
        ("exponential", [-0.4], [1.0], 213, 0.0, 2.0, 1e-6, dict(rel_tol=1e-4)),
        # the remaining small test functors of test_equations.hpp (kinetic-4 is the n = 4 case)
        ("lorenz", [10.0, 28.0, 8.0 / 3.0], [1.0, 1.0, 1.0], 211, 0.0, 2.0, 1e-6, o6),
        ("lorenz", [10.0, 28.0, 8.0 / 3.0], [1.0, 1.0, 1.0], 230, 0.0, 1.0, 1e-6, o6),
        ("harmonic", [3.0], [0.0, 1.0], 211, 0.0, 5.0, 1e-6, o6),
        ("dimer", [5.0], [1.0, 0.0], 230, 0.0, 3.0, 1e-6, o6),
        ("kinetic-4", [0.5, 0.3, 0.2], [1.0, 0.0, 0.0, 0.0], 211, 0.0, 10.0, 1e-6, o6),
        ("kinetic-4", [0.5, 0.3, 0.2], [1.0, 0.0, 0.0, 0.0], 230, 0.0, 10.0, 1e-6, o6),
        ("kinetic-4", [0.5, 0.3, 0.2], [1.0, 0.0, 0.0, 0.0], 232, 0.0, 2.0, 1e-6, o6),
        ("kinetic-4", [0.5, 0.3, 0.2], [1.0, 0.0, 0.0, 0.0], 212, 0.0, 10.0, 1e-6, o6),
        # three_body (n = 12, test_equations.hpp:254-421) from the example driver's initial state
        # (examples/example_equations.cpp:341-343); its jac() is not the derivative of its fun() -- reproduced as written
        ("three-body", [1.0, 1.0, 1.0], THREE_BODY_Y0, 211, 0.0, 5.0, 1e-6, o6),  # includes a Newton failure
        ("three-body", [1.0, 1.0, 1.0], THREE_BODY_Y0, 230, 0.0, 2.0, 1e-6, o6),
        ("three-body", [1.0, 1.0, 1.0], THREE_BODY_Y0, 232, 0.0, 2.0, 1e-6, o6),
        ("three-body", [0.5, 2.0, 3.0], THREE_BODY_Y0, 211, 0.0, 2.0, 1e-6, o6),
        ("brusselator-1d", [0.02, 1.0, 3.0], list(ensembles.bruss1d_y0(4)), 211, 0.0, 1.0, 1e-6, o6),
        ("brusselator-1d", [0.02, 1.0, 3.0], list(ensembles.bruss1d_y0(4)), 232, 0.0, 1.0, 1e-6, o6),
    ]
    for prob, par, y0, method, t0, t1, dt0, okw in singles:
        opts = make_options(**okw)
        P = np.array(par, dtype=np.float64)[:, None] if par is not None else None
        traj_cap = 4096
        r = ref.solve(prob, method, t0, t1, dt0, np.array(y0), P, opts, M=1, traj_cap=traj_cap)
        L = r.extra["traj_len"]
        add(dict(kind="single", name=f"{prob}_{method}_t{t1:g}", problem=prob, params=par, y0=y0, method=method,
                 t0=t0, t1=t1, dt0=dt0, options=okw, status=int(r.status[0]), y=hx(r.y[:, 0]), t_final=hx(r.t_final),
                 err_last=hx(r.err_last), counters=counters(r), accept_frac=float(r.accept_frac[0]).hex(),
                 traj_len=int(L), traj_t=hx(r.traj_t[:64]), traj_y=hx(r.traj_y[:64]), traj_err=hx(r.traj_err[:64])))

    # 3. swept Robertson ensemble, first 64 members of config 2
    M = 64
    P = ensembles.robertson_params(M)
    opts = make_options(**o6)
    r = ref.solve("robertson", 211, 0.0, 40.0, 1e-6, ensembles.ROBERTSON_Y0, P, opts)
    add(dict(kind="ensemble", name="robertson_sweep64_211_t40", problem="robertson", method=211, t0=0.0, t1=40.0,
             dt0=1e-6, options=o6, M=M, params=hx(P), y0=[1.0, 0.0, 0.0], y=hx(r.y), status=[int(s) for s in r.status],
             attempt=[int(v) for v in r.counters["attempt"]], reject_err=[int(v) for v in r.counters["reject_err"]],
             fun_evals=[int(v) for v in r.counters["fun_evals"]], jac_evals=[int(v) for v in r.counters["jac_evals"]]))
    M = 8
    P = ensembles.robertson_params(M)
    r = ref.solve("robertson", 211, 0.0, 1e11, 1e-6, ensembles.ROBERTSON_Y0, P, opts)
    add(dict(kind="ensemble", name="robertson_sweep8_211_t1e11", problem="robertson", method=211, t0=0.0, t1=1e11,
             dt0=1e-6, options=o6, M=M, params=hx(P), y0=[1.0, 0.0, 0.0], y=hx(r.y), status=[int(s) for s in r.status],
             attempt=[int(v) for v in r.counters["attempt"]], reject_err=[int(v) for v in r.counters["reject_err"]],
             fun_evals=[int(v) for v in r.counters["fun_evals"]], jac_evals=[int(v) for v in r.counters["jac_evals"]]))
    # Van der Pol sweep, config 3 shape (16 members)
    M = 16
    P = ensembles.vdpol_params(M)
    r = ref.solve("van-der-pol", 230, 0.0, 2.0, 1e-6, ensembles.VDPOL_Y0, P, opts)
    add(dict(kind="ensemble", name="vdpol_sweep16_230_t2", problem="van-der-pol", method=230, t0=0.0, t1=2.0,
             dt0=1e-6, options=o6, M=M, params=hx(P), y0=[2.0, 0.0], y=hx(r.y), status=[int(s) for s in r.status],
             attempt=[int(v) for v in r.counters["attempt"]], reject_err=[int(v) for v in r.counters["reject_err"]],
             reject_newton=[int(v) for v in r.counters["reject_newton"]],
             fun_evals=[int(v) for v in r.counters["fun_evals"]], jac_evals=[int(v) for v in r.counters["jac_evals"]]))

    # 4. fixed-step oracle
    f = ref.fixed_step("robertson-hardcoded", 211, None, [1.0, 0.0, 0.0], 0.0, 1e-3, 100, maxit=100, refresh_jac=25,
                       xtol=1e-14, Rtol=1e-14)
    add(dict(kind="fixed_step", name="robertson_fixed_211_100x1e-3", problem="robertson", params=[0.04, 1e4, 3e7],
             method=211, y0=[1.0, 0.0, 0.0], t0=0.0, dt=1e-3, nsteps=100, xtol=1e-14, Rtol=1e-14, y=hx(f["y"]),
             failed=f["failed"], newton_iters=f["newton_iters"]))
    f = ref.fixed_step("van-der-pol", 230, [1e-2], [2.0, 0.0], 0.0, 1e-3, 200, maxit=100, refresh_jac=25, xtol=1e-14,
                       Rtol=1e-14)
    add(dict(kind="fixed_step", name="vdpol_fixed_230_200x1e-3", problem="van-der-pol", params=[1e-2], method=230,
             y0=[2.0, 0.0], t0=0.0, dt=1e-3, nsteps=200, xtol=1e-14, Rtol=1e-14, y=hx(f["y"]), failed=f["failed"],
             newton_iters=f["newton_iters"]))

    # implicit Euler and the high-order Radau methods through the sanctioned fixed-step loop
    for prob, par, y0, method, dt, nsteps in (("robertson-hardcoded", None, [1.0, 0.0, 0.0], 200, 1e-4, 100),
                                              ("van-der-pol", [1e-2], [2.0, 0.0], 200, 1e-3, 200),
                                              ("robertson-hardcoded", None, [1.0, 0.0, 0.0], 212, 1e-3, 50),
                                              ("van-der-pol", [1e-2], [2.0, 0.0], 213, 1e-2, 50)):
        f = ref.fixed_step(prob, method, par, y0, 0.0, dt, nsteps, maxit=100, refresh_jac=25, xtol=1e-14, Rtol=1e-14)
        gp = "robertson" if prob == "robertson-hardcoded" else prob
        add(dict(kind="fixed_step", name=f"{gp}_fixed_{method}_{nsteps}x{dt:g}", problem=gp,
                 params=[0.04, 1e4, 3e7] if par is None else par, method=method, y0=y0, t0=0.0, dt=dt, nsteps=nsteps,
                 xtol=1e-14, Rtol=1e-14, y=hx(f["y"]), failed=f["failed"], newton_iters=f["newton_iters"]))

    # 5. tableaux + derived weights
    for m in (200, 210, 211, 212, 213, 230, 231, 232):
        c = ref.get_coefficients(m)
        add(dict(kind="tableau", name=ref.method_to_name(m), method=m, s=c["s"], A=hx(c["A"]), b=hx(c["b"]), c=hx(c["c"]),
                 b2=hx(c["b2"]), gamma=float(c["gamma"]).hex(), order=c["order"], order2=c["order2"], d=hx(c["d"]),
                 d2=hx(c["d2"])))
    # 6. dense output identities (test/test_interpolate.cpp)
    for m, thetas in ((210, [1.0, 1.0 / 3.0, 0.5]), (211, [1.0, 0.3, 0.75]), (232, [1.0, 0.5]), (212, [1.0, 0.25]),
                      (213, [1.0, 0.6])):
        for th in thetas:
            add(dict(kind="project_b", method=m, theta=float(th).hex(), b=hx(ref.project_b(m, th))))

    path = os.path.join(ROOT, "tests", "golden", "oracle_cases.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print(f"wrote {path}: {len(out['cases'])} cases, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
