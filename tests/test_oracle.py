"""CPU tests of the oracle itself: the C port must reproduce, BIT FOR BIT, golden vectors that were
produced by the unmodified reference (tools/make_golden.py), and the reference's own known-answer
test (test/irk.cpp:250-296).  When oracle/_ref is present the two are also compared live."""
import numpy as np
import pytest

from conftest import cases_of, unhex
from oracle.refsolver import make_options


def _solve_single(solver, c, traj_cap=4096):
    P = np.array(c["params"], dtype=np.float64)[:, None] if c["params"] is not None else None
    return solver.solve(c["problem"], c["method"], c["t0"], c["t1"], c["dt0"], np.array(c["y0"]), P,
                        make_options(**c["options"]), M=1, traj_cap=traj_cap)


def test_reference_known_answer_irk_calc_stages(port, golden):
    """test/irk.cpp:250-296: y1 ~ {0.99994, 3.39216e-05, 2.60729e-05} to 1e-2 (Octave), variant 0 =
    the template defaults the reference test uses."""
    k = port.newton_solve_stages("robertson-hardcoded", 211, 0, None, [1.0, 0.0, 0.0], 0.0, 1.5e-3, 100, 25, 1e-8, 1e-10)
    true_y1 = np.array([0.99994, 3.39216e-05, 2.60729e-05])
    assert k["status"] == 0
    np.testing.assert_allclose(k["y1"], true_y1, rtol=1e-2)
    # and far tighter than the reference asks: 5 digits of the Octave values
    np.testing.assert_allclose(k["y1"], true_y1, rtol=2e-5)


def test_port_matches_golden_newton(port, golden):
    for c in cases_of(golden, "newton_solve_stages"):
        k = port.newton_solve_stages(c["problem"], c["method"], c["variant"], None, c["y"], c["t"], c["dt"], c["maxit"],
                                     c["refresh_jac"], c["xtol"], c["Rtol"])
        assert (k["status"], k["iters"], k["fun_evals"], k["jac_evals"]) == (c["status"], c["iters"], c["fun_evals"], c["jac_evals"])
        assert np.array_equal(k["Y"], unhex(c["Y"]))
        assert np.array_equal(k["y1"], unhex(c["y1"]))


def test_port_matches_golden_singles_bitwise(port, golden):
    singles = cases_of(golden, "single")
    assert len(singles) >= 15
    for c in singles:
        r = _solve_single(port, c)
        assert int(r.status[0]) == c["status"], c["name"]
        assert np.array_equal(r.y[:, 0], unhex(c["y"])), c["name"]
        assert np.array_equal(r.t_final, unhex(c["t_final"])), c["name"]
        assert np.array_equal(r.err_last, unhex(c["err_last"])), c["name"]
        assert {k: int(v[0]) for k, v in r.counters.items()} == c["counters"], c["name"]
        assert r.extra["traj_len"] == c["traj_len"]
        assert np.array_equal(r.traj_t[:64], unhex(c["traj_t"])), c["name"]
        assert np.array_equal(r.traj_y[:64].ravel(), unhex(c["traj_y"])), c["name"]
        assert np.array_equal(r.traj_err[:64], unhex(c["traj_err"])), c["name"]


def test_port_matches_golden_ensembles_bitwise(port, golden):
    for c in cases_of(golden, "ensemble"):
        M = c["M"]
        P = unhex(c["params"]).reshape(-1, M)
        r = port.solve(c["problem"], c["method"], c["t0"], c["t1"], c["dt0"], np.array(c["y0"]), P, make_options(**c["options"]))
        assert np.array_equal(r.y.ravel(), unhex(c["y"])), c["name"]
        assert [int(v) for v in r.counters["attempt"]] == c["attempt"]
        assert [int(v) for v in r.counters["fun_evals"]] == c["fun_evals"]
        assert [int(v) for v in r.counters["jac_evals"]] == c["jac_evals"]
        assert [int(v) for v in r.status] == c["status"]


def test_port_matches_golden_fixed_step(port, golden):
    for c in cases_of(golden, "fixed_step"):
        f = port.fixed_step(c["problem"], c["method"], c["params"], c["y0"], c["t0"], c["dt"], c["nsteps"], maxit=100,
                            refresh_jac=25, xtol=c["xtol"], Rtol=c["Rtol"])
        assert np.array_equal(f["y"], unhex(c["y"]))
        assert (f["failed"], f["newton_iters"]) == (c["failed"], c["newton_iters"])


def test_port_tableaux_match_golden(port, golden):
    for c in cases_of(golden, "tableau"):
        t = port.get_coefficients(c["method"])
        assert t["s"] == c["s"] and t["order"] == c["order"] and t["order2"] == c["order2"]
        for k in ("A", "b", "c", "b2", "d", "d2"):
            assert np.array_equal(np.asarray(t[k]).ravel(), unhex(c[k])), (c["name"], k)
        assert t["gamma"] == float.fromhex(c["gamma"])


def test_reference_live_equals_port(reference, port):
    """Where oracle/_ref exists: unmodified reference == C port, bit for bit, on fresh inputs
    (members 100..131 of the config-2 sweep, not in the fixture)."""
    from rehuel_b200 import ensembles
    P = ensembles.robertson_params(32, start=100)
    o = make_options(rel_tol=1e-6, abs_tol=1e-6)
    a = reference.solve("robertson", 211, 0.0, 40.0, 1e-6, ensembles.ROBERTSON_Y0, P, o)
    b = port.solve("robertson", 211, 0.0, 40.0, 1e-6, ensembles.ROBERTSON_Y0, P, o)
    assert np.array_equal(a.y, b.y)
    for k in a.counters:
        assert np.array_equal(a.counters[k], b.counters[k]), k
    P = ensembles.vdpol_params(8)
    a = reference.solve("van-der-pol", 230, 0.0, 2.0, 1e-6, ensembles.VDPOL_Y0, P, o)
    b = port.solve("van-der-pol", 230, 0.0, 2.0, 1e-6, ensembles.VDPOL_Y0, P, o)
    assert np.array_equal(a.y, b.y)
    assert np.array_equal(a.counters["fun_evals"], b.counters["fun_evals"])


def test_functor_evaluations_port_equals_reference(reference, port):
    """fun(t, y) and jac(t, y) of every restated test functor, bit for bit against the reference's own
    (test_equations.hpp) at random states -- including three_body, whose jac() is NOT the derivative of its fun()
    (test_equations.hpp:304-421) and is reproduced as written."""
    rng = np.random.default_rng(11)
    cases = [("robertson", [0.04, 1e4, 3e7], 3), ("van-der-pol", [1e-2], 2), ("exponential", [-0.4], 1),
             ("stiff-equation", None, 2), ("brusselator", [1.0, 3.5], 2), ("lorenz", [10.0, 28.0, 8.0 / 3.0], 3),
             ("harmonic", [3.0], 2), ("dimer", [5.0], 2), ("kinetic-4", [0.5, 0.3, 0.2], 4),
             ("three-body", [0.5, 2.0, 3.0], 12), ("three-body", [1.0, 1.0, 1000.0], 12), ("brusselator-1d", [0.02, 1.0, 3.0], 16)]
    for prob, par, n in cases:
        for _ in range(4):
            y = rng.uniform(-1.5, 1.5, n)
            fa, Ja = reference.eval_functor(prob, 0.3, y, par)
            fb, Jb = port.eval_functor(prob, 0.3, y, par)
            assert np.array_equal(fa, fb) and np.array_equal(Ja, Jb), prob
    # the three-body "Jacobian" really is not one: the momentum rows of fun() depend on the positions only
    f, J = port.eval_functor("three-body", 0.0, rng.uniform(-1.5, 1.5, 12), [1.0, 1.0, 1.0])
    assert np.all(J[6:, :6] == 0.0) and np.any(J[6:, 6:] != 0.0) and np.all(np.diag(J)[:6] == -1.0)


def test_reference_rober_p_equals_hardcoded_rober(reference):
    """Our parameterised Robertson functor with (0.04,1e4,3e7) is the reference's `rober` bit for bit."""
    o = make_options(rel_tol=1e-6, abs_tol=1e-6)
    a = reference.solve("robertson-hardcoded", 211, 0.0, 40.0, 1e-6, [1.0, 0.0, 0.0], None, o)
    b = reference.solve("robertson", 211, 0.0, 40.0, 1e-6, [1.0, 0.0, 0.0], np.array([[0.04], [1e4], [3e7]]), o)
    assert np.array_equal(a.y, b.y) and np.array_equal(a.counters["fun_evals"], b.counters["fun_evals"])


def test_fixed_step_irk_guts_throws_like_reference(port, reference):
    """adaptive_step_size=false makes the reference throw (size mismatch)."""
    o = make_options(rel_tol=1e-6, abs_tol=1e-6, adaptive=False)
    for s in (port, reference):
        with pytest.raises(RuntimeError, match="incompatible"):
            s.solve("exponential", 211, 0.0, 1.0, 1e-2, [1.0], np.array([[-1.0]]), o)


def test_analytic_sanity(port):
    """Independent of the reference: exponential and the linear stiff problem against closed forms
    (test_equations.hpp:43-46, 227-235); Robertson conserves mass."""
    o = make_options(rel_tol=1e-8, abs_tol=1e-10)
    r = port.solve("exponential", 211, 0.0, 2.0, 1e-6, [1.0], np.array([[-0.4]]), o)
    assert abs(r.y[0, 0] - np.exp(-0.8)) < 1e-7
    r = port.solve("stiff-equation", 211, 0.0, 1.0, 1e-6, [1.0, 0.0], None, o)
    ysol = -np.exp(-1000.0) / 999.0 + 1000 * np.exp(-1.0) / 999.0
    assert abs(r.y[0, 0] - ysol) < 1e-7
    r = port.solve("robertson-hardcoded", 211, 0.0, 1e5, 1e-6, [1.0, 0.0, 0.0], None, make_options(rel_tol=1e-6, abs_tol=1e-6))
    assert abs(r.y[:, 0].sum() - 1.0) < 1e-12


def test_method_names(port):
    assert port.name_to_method("RADAU_IIA_53") == 211
    assert port.name_to_method("LOBATTO_IIIC_43") == 230
    assert port.name_to_method("LOBATTO_IIIC_85") == 232
    assert port.name_to_method("no-such-method") == 0
    assert port.method_to_name(211) == "RADAU_IIA_53"


def test_high_order_step_counts_are_noise(port):
    """Why step counts of RADAU_IIA_95/137 are compared with a wide band (tests/test_gpu_parity.py): at rtol 1e-6 the
    embedded estimate of a 9th/13th order pair is rounding noise, and the REFERENCE algorithm's own attempt count on
    Robertson to t = 1e11 changes by >10 % when k1 moves by parts in 1e16 -- while RADAU_IIA_53 does not move at all."""
    import numpy as np
    from oracle.refsolver import make_options
    o = make_options(rel_tol=1e-6, abs_tol=1e-6)
    counts = {}
    for method in (211, 212):
        counts[method] = []
        for eps in (0.0, -1e-16, 2e-16, 1e-15, -1e-15):
            P = np.array([0.04 * (1 + eps), 1e4, 3e7])[:, None]
            r = port.solve("robertson", method, 0.0, 1e11, 1e-6, np.array([1.0, 0.0, 0.0]), P, o, M=1)
            counts[method].append(int(r.counters["attempt"][0]))
    assert len(set(counts[211])) == 1, counts
    assert max(counts[212]) - min(counts[212]) > 0.1 * max(counts[212]), counts


def test_port_matches_reference_on_config_members(port, golden_configs):
    """tests/golden/config_cases.json (the unmodified reference on members of BASELINE configs 4 and 5): the C port
    reproduces every config-5 member and the cheapest config-4 member (n = 64, LOBATTO_IIIC_85, t in [0,10]: the
    reference's dense 320 x 320 LU, ~15 s through the port) bit for bit."""
    import importlib.util, os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("_ens", os.path.join(ROOT, "rehuel_b200", "ensembles.py"))
    ens = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ens)
    c5 = cases_of(golden_configs, "config5")
    assert len(c5) == 16
    for c in c5:
        tol = float.fromhex(c["tol"])
        P = unhex(c["params"])[:, None]
        r = port.solve(c["problem"], c["method"], 0.0, c["t1"], 1e-6, np.array(c["y0"]), P,
                       make_options(tol, tol, newton_tol=0.1 * tol), M=1)
        assert np.array_equal(r.y[:, 0], unhex(c["y"])), c["index"]
        assert {k: int(v[0]) for k, v in r.counters.items()} == c["counters"], c["index"]
    c4 = cases_of(golden_configs, "config4")
    assert len(c4) == 16
    c = min(c4, key=lambda c: c["counters"]["attempt"])
    P = unhex(c["params"])[:, None]
    assert np.array_equal(P, ens.bruss1d_params(1, start=c["index"]))
    r = port.solve("brusselator-1d", 232, 0.0, 10.0, 1e-6, ens.bruss1d_y0(32), P, make_options(1e-6, 1e-6), M=1)
    assert np.array_equal(r.y[:, 0], unhex(c["y"])) and int(r.status[0]) == c["status"] == 0
    assert {k: int(v[0]) for k, v in r.counters.items()} == c["counters"]


def test_reference_log_golden_is_consistent(golden_configs):
    """The per-attempt logs in the fixture: one line per attempt that reached the error estimate (Newton failures print
    nothing, irk.hpp:649-656) whose step count is a multiple of out_interval."""
    from conftest import parse_reference_log
    tr = cases_of(golden_configs, "trace")
    assert len(tr) == 3
    for c in tr:
        rows = parse_reference_log(c["log"])
        k = c["counters"]
        if c["out_interval"] == 1:
            assert len(rows) == k["attempt"] - k["reject_newton"]
            assert int((rows[:, 3] > 1.0).sum()) == k["reject_err"]
        assert np.all(rows[:, 0] % c["out_interval"] == 0) and rows[0, 1] == 0.0


def test_four_argument_odeint_defaults_are_pinned(port, golden_defaults):
    """irk.hpp:915-926: RADAU_IIA_53, dt = 1e-6, rel_tol = 1e-4, abs_tol = 1e-3 (options.hpp), refresh_jac = 25,
    Newton tol = 0.1*min(abs_tol, rel_tol).  The fixture comes from the reference's own 4-argument call, which is not
    told any of these: stating them to the port must reproduce it bit for bit."""
    from oracle import refsolver
    o = refsolver.make_options(rel_tol=1e-4, abs_tol=1e-3, newton_tol=1e-5, newton_refresh_jac=25)
    for c in golden_defaults:
        r = port.solve(c["problem"], 211, c["t0"], c["t1"], 1e-6, np.array(c["y0"]), np.array(c["params"])[:, None], o)
        assert int(r.status[0]) == c["status"] == 0
        assert np.array_equal(r.y[:, 0], unhex(c["y"])), c["name"]
        assert r.t_final[0] == unhex(c["t_final"])[0] and r.err_last[0] == unhex(c["err_last"])[0]
        assert {k: int(v[0]) for k, v in r.counters.items()} == c["counters"], c["name"]


def test_four_argument_odeint_of_the_reference_library(reference, golden_defaults):
    """The fixture regenerates from the unmodified reference (skipped where oracle/_ref was built before this entry point existed)."""
    if not hasattr(reference.lib, "ref_odeint_defaults"):
        pytest.skip("oracle/_ref predates ref_odeint_defaults")
    for c in golden_defaults[:3]:
        r = reference.odeint_defaults(c["problem"], c["t0"], c["t1"], c["y0"], c["params"])
        assert np.array_equal(r.y[:, 0], unhex(c["y"])) and {k: int(v[0]) for k, v in r.counters.items()} == c["counters"]
