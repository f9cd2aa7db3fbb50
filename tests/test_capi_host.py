"""CPU tests of the C-ABI library: it loads, exports every symbol include/rehuel_b200.h declares,
its host-side pieces (defaults, tableaux, name maps, eigen-structure, dense-output weights) agree
with the oracle, and -- no GPU here -- every compute entry point fails LOUDLY (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, cases_of, unhex


@pytest.fixture(scope="module")
def capi():
    from rehuel_b200 import capi as m
    return m


def test_library_exports_every_declared_symbol(capi):
    header = open(os.path.join(ROOT, "include", "rehuel_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rehuel_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(capi.lib, name), f"{name} declared in the header but not exported"
    assert set(capi.EXPORTS) <= declared


def test_struct_layouts_match_header(capi, tmp_path):
    # the sizes and a few offsets gcc gives the header's structs must be the ctypes mirror's
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rehuel_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rehuel_options), sizeof(rehuel_stats),'
                   ' sizeof(rehuel_result), sizeof(rehuel_device_io), offsetof(rehuel_options, handout_order),'
                   ' offsetof(rehuel_result, sum), offsetof(rehuel_device_io, order)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [C.sizeof(capi.Options), C.sizeof(capi.Stats), C.sizeof(capi.Result), C.sizeof(capi.DeviceIO),
                   capi.Options.handout_order.offset, capi.Result.sum.offset, capi.DeviceIO.order.offset]
    assert C.sizeof(capi.Stats) == 13 * 8


def test_default_options_are_the_reference_defaults(capi):
    o = capi.default_options()
    # options.hpp:49-58, irk.hpp:103-107, newton.hpp:58-60, options.hpp:103-104
    assert (o.rel_tol, o.abs_tol, o.max_dt, o.max_steps) == (1e-4, 1e-3, 0.0, -1)
    assert (o.adaptive_step_size, o.out_interval) == (1, 0)
    assert (o.newton_tol, o.newton_dx_delta, o.newton_maxit, o.newton_refresh_jac) == (1e-4, 1e-4, 5000, 10)
    assert (o.output_mode, o.output_interval) == (1, 1)


def test_method_and_problem_names(capi, port):
    for name in ("IMPLICIT_EULER", "RADAU_IIA_32", "RADAU_IIA_53", "RADAU_IIA_95", "RADAU_IIA_137", "LOBATTO_IIIA_43",
                 "LOBATTO_IIIC_43", "LOBATTO_IIIC_64", "LOBATTO_IIIC_85", "GAUSS_LEGENDRE_147", "bogus"):
        assert capi.name_to_method(name) == port.name_to_method(name)
    for m in (200, 210, 211, 212, 213, 230, 231, 232, 243, 999):
        assert capi.method_to_name(m) == port.method_to_name(m)
    assert capi.name_to_problem("robertson") == capi.ROBERTSON
    assert capi.name_to_problem("van-der-pol") == capi.VAN_DER_POL
    assert capi.name_to_problem("brusselator-1d") == capi.BRUSSELATOR_1D
    assert [capi.name_to_problem(k) for k in ("lorenz", "harmonic", "dimer", "kinetic-4", "brusselator-1d-dense")] == [8, 9, 10, 11, 4]
    assert capi.problem_dims(capi.KINETIC_4) == (4, 3) and capi.problem_dims(capi.LORENZ) == (3, 3)
    assert capi.name_to_problem("three-body") == capi.THREE_BODY == 12 and capi.problem_dims(capi.THREE_BODY) == (12, 3)
    assert capi.name_to_problem("nope") == 0
    assert capi.problem_dims(capi.ROBERTSON) == (3, 3)
    assert capi.problem_dims(capi.VAN_DER_POL) == (2, 1)
    assert capi.problem_dims(capi.BRUSSELATOR_1D, 64) == (64, 3)


def test_tableaux_match_reference_golden(capi, golden):
    for c in cases_of(golden, "tableau"):
        t = capi.get_coefficients(c["method"])
        assert t is not None, c["name"]
        assert (t["s"], t["order"], t["order2"]) == (c["s"], c["order"], c["order2"])
        for k in ("A", "b", "c", "b2"):
            assert np.array_equal(np.asarray(t[k]).ravel(), unhex(c[k])), (c["name"], k)
        assert t["gamma"] == float.fromhex(c["gamma"])
        # d = A^-T b goes through a different elimination order: equal to rounding
        np.testing.assert_allclose(t["d"], unhex(c["d"]), rtol=0, atol=4e-15)
        np.testing.assert_allclose(t["d2"], unhex(c["d2"]), rtol=0, atol=4e-14)
    assert capi.get_coefficients(12345) is None


def test_eigen_structure_reconstructs_inverse_of_A(capi):
    """inv(A) = T Lambda Ti with the block structure the kernels assume."""
    for m in (200, 210, 211, 212, 213, 230, 231, 232):
        t = capi.get_coefficients(m)
        e = capi.get_eigen(m)
        s = t["s"]
        assert e["n_real"] + 2 * e["n_cplx"] == s
        L = np.zeros((s, s))
        j = 0
        if e["n_real"]:
            L[0, 0] = e["lam"][0]
            j = 1
        for k in range(e["n_cplx"]):
            a, b = e["lam"][1 + 2 * k], e["lam"][2 + 2 * k]
            assert b > 0
            L[j:j + 2, j:j + 2] = [[a, -b], [b, a]]
            j += 2
        Ai = np.linalg.inv(t["A"])
        # T and Ti are 60-digit results rounded to double: the products carry cond(T)*eps of rounding
        cond = np.linalg.norm(e["T"], np.inf) * np.linalg.norm(e["Ti"], np.inf)  # 1.9 (LOBATTO_IIIC_43) .. 299 (RADAU_IIA_137)
        np.testing.assert_allclose(e["T"] @ L @ e["Ti"], Ai, rtol=0, atol=max(2e-13, 2e-15 * cond) * np.abs(Ai).max())
        np.testing.assert_allclose(e["T"] @ e["Ti"], np.eye(s), atol=max(1e-14, 4e-16 * cond))
    # estimator mode: LOBATTO_IIIC_43/64 keep gamma = 0 (irk.cpp:231) -> identity; RADAU_IIA_53's gamma is
    # 1/gamma_hat -> reuse; LOBATTO_IIIC_85's printed gamma is NOT (4e-6 off) -> own LU
    # RADAU_IIA_95/137 print gamma to >= 16 digits -> reuse; IMPLICIT_EULER has no estimator at all (gamma = 0)
    assert [capi.get_eigen(m)["est_mode"] for m in (230, 231, 211, 232, 210, 212, 213, 200)] == [0, 0, 1, 2, 2, 1, 1, 0]
    e = capi.get_eigen(211)
    assert abs(capi.get_coefficients(211)["gamma"] * e["lam"][0] - 1.0) < 1e-14


def test_project_b_matches_reference_and_identities(capi, golden):
    """test/test_interpolate.cpp:5-44: b(1) = b, b(c_i) = A(i,:)."""
    for c in cases_of(golden, "project_b"):
        got = capi.project_b(c["method"], float.fromhex(c["theta"]))
        # the monomial coefficients of the 7-stage interpolant reach ~1e3 with alternating signs: both sides carry
        # ~1e3 ulp of cancellation noise, in different summation orders (irk.cpp:178-206 vs tableau.cpp)
        np.testing.assert_allclose(got, unhex(c["b"]), rtol=0, atol=5e-14 if c["method"] != 213 else 1e-12)
    for m, tol in ((210, 1e-14), (211, 1e-14), (212, 1e-13), (213, 1e-11)):
        t = capi.get_coefficients(m)
        np.testing.assert_allclose(capi.project_b(m, 1.0), t["b"], atol=tol)
        for i in range(t["s"]):
            np.testing.assert_allclose(capi.project_b(m, t["c"][i]), t["A"][i], atol=tol)
    assert capi.project_b(230, 0.5) is None  # no b_interp for LOBATTO_IIIC_43 (irk.cpp:261-272)


def test_argument_errors_are_reported_before_any_gpu_work(capi):
    """Bad arguments come back as REHUEL_EINVAL with a message (no GPU needed to find out)."""
    from rehuel_b200 import ensembles
    P = ensembles.robertson_params(4)
    for kw, needle in ((dict(dense_samples=4, dense_t0=0.0, dense_dt=1.0, method=230), "dense output"),  # no b_interp (irk.cpp:733)
                       (dict(dense_samples=4, dense_t0=0.0, dense_dt=0.0, method=211), "dense_dt"),
                       (dict(handout_order=7, method=211), "handout_order"),
                       (dict(newton_refresh_jac=0, method=211), "refresh_jac"),
                       (dict(method=220), "not supported")):  # LOBATTO_IIIA_43: singular A
        method = kw.pop("method")
        with pytest.raises(capi.RehuelError) as ei:
            capi.irk_ensemble(capi.ROBERTSON, method, 0.0, 1.0, 1e-6, ensembles.ROBERTSON_Y0, P, capi.make_options(1e-6, 1e-6, **kw))
        assert ei.value.code == -1 and needle in str(ei.value), (needle, str(ei.value))
    with pytest.raises(capi.RehuelError) as ei:  # nothing to continue from
        capi.irk_ensemble_continue(capi.ROBERTSON, 211, 2.0, P, capi.make_options(1e-6, 1e-6), capi.Result())
    assert ei.value.code == -1 and "continue" in str(ei.value)
    assert capi.lib.rehuel_locality_order(None, 3, 16, 0, None, None, 0, None) == -1


def test_no_gpu_means_loud_failure_not_a_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rehuel_b200 import ensembles
    with pytest.raises(capi.RehuelError) as ei:
        capi.irk_ensemble(capi.ROBERTSON, 211, 0.0, 40.0, 1e-6, ensembles.ROBERTSON_Y0, ensembles.robertson_params(8),
                          capi.make_options(1e-6, 1e-6))
    assert ei.value.code == -2 and "cuda" in str(ei.value).lower()
    assert capi.kernel_launches() == 0


def test_argument_validation(capi):
    from rehuel_b200 import ensembles
    P = ensembles.robertson_params(4)
    o = capi.make_options(1e-6, 1e-6)
    for kwargs, frag in ((dict(method=999), "not supported"), (dict(dt0=0.0), "step size"), (dict(problem=77), "problem")):
        args = dict(problem=capi.ROBERTSON, method=211, dt0=1e-6)
        args.update(kwargs)
        res = capi.Result()
        rc = capi.lib.rehuel_irk_ensemble(args["problem"], args["method"], 0.0, 1.0, args["dt0"], 4, 3,
                                          ensembles.ROBERTSON_Y0.ctypes.data_as(C.POINTER(C.c_double)), 1,
                                          P.ctypes.data_as(C.POINTER(C.c_double)), C.byref(o), 0, 1, C.byref(res))
        assert rc == -1 and frag in capi.last_error()
    bad = capi.make_options(1e-6, 1e-6, newton_refresh_jac=0)  # the reference would divide by zero (irk.hpp:485)
    res = capi.Result()
    rc = capi.lib.rehuel_irk_ensemble(capi.ROBERTSON, 211, 0.0, 1.0, 1e-6, 4, 3,
                                      ensembles.ROBERTSON_Y0.ctypes.data_as(C.POINTER(C.c_double)), 1,
                                      P.ctypes.data_as(C.POINTER(C.c_double)), C.byref(bad), 0, 1, C.byref(res))
    assert rc == -1


def test_c_program_fails_loudly_without_gpu():
    """examples/c_interface_test.c (plain C against the C ABI): on a box without a GPU it must exit
    non-zero with the CUDA error text, not produce numbers from somewhere else."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = os.path.join(ROOT, "build", "c_interface_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ROOT, "build/c_interface_test"], stdout=subprocess.DEVNULL)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "failed" in r.stderr


def test_twisted_band_factorisation_on_the_host(tmp_path):
    """rehuel_b200/csrc/band_twisted.cuh (the lane-pair band LU / solve of the warp-per-system kernel) is
    __host__ __device__: tests/band_twisted_host.cu runs those sweeps on the CPU, the two lanes one after the other,
    against a dense complex solve with partial pivoting -- real and complex, n = 8 ... 64, half bandwidths 1 ... 4,
    odd and even splits; matrices that need interchanges must raise the would-pivot flag."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    exe = tmp_path / "band_twisted_host"
    cmd = [nvcc, "-O1", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", str(exe), os.path.join(ROOT, "tests", "band_twisted_host.cu")]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    print(out.stdout)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK")


def test_new_option_defaults_and_argument_errors(capi):
    """Round-2 additions to rehuel_options: defaults, and the combinations that are refused before any GPU work."""
    from rehuel_b200 import ensembles
    o = capi.default_options()
    assert (o.trace_member, o.output_file, o.result_mask, o.time_internals, o.shard_mode) == (0, None, capi.OUT_ALL, 0, capi.SHARD_DYNAMIC)
    P = ensembles.robertson_params(4)
    for kw, needle in ((dict(out_interval=1, trace_member=4), "trace_member"),               # not a member
                       (dict(output_mode=3), "output_file"),                                # WRITE_TO_FILE without a path
                       (dict(output_mode=3, output_file="/tmp/x.txt"), "n_samples_cap"),    # ... without sample storage
                       (dict(shard_mode=5), "shard_mode")):
        with pytest.raises(capi.RehuelError) as ei:
            capi.irk_ensemble(capi.ROBERTSON, 211, 0.0, 1.0, 1e-6, ensembles.ROBERTSON_Y0, P, capi.make_options(1e-6, 1e-6, **kw))
        assert ei.value.code == -1 and needle in str(ei.value), (needle, str(ei.value))


def test_log_and_timing_text_of_a_result(capi):
    """rehuel_result_format_log / _format_timings are host-side: the reference's words around the numbers of a
    result (irk.hpp:576, 807-809; 346-359)."""
    res = capi.Result()
    rows = np.array([[0, 0.0, 1e-6, 1e-17, 1], [1, 1e-6, 7.2e-6, -1.0, 4999], [1, 1e-6, 5.04e-6, 0.25, 2]])  # the middle one: a failed Newton solve
    res.trace = rows.ctypes.data_as(C.POINTER(C.c_double))
    res.trace_len = 3
    text = capi.format_log(res, 6).splitlines()
    assert text == ["    Rehuel: step  t  dt   err   iters", "    Rehuel: 0 0 1e-06 1e-17 1", "    Rehuel: 1 1e-06 5.04e-06 0.25 2"]
    assert capi.format_log(res, 17).splitlines()[2] == "    Rehuel: 1 9.9999999999999995e-07 5.04e-06 0.25 2"
    assert capi.lib.rehuel_result_format_log(C.byref(res), 6, None, 0) == len("\n".join(text)) + 1   # snprintf-like length
    res.phase_ms[0], res.phase_ms[1], res.phase_ms[2], res.phase_ms[3] = 1.0, 0.5, 8.0, 0.5
    t = capi.format_timings(res, "RADAU_IIA_53").splitlines()
    assert t[0] == "    Rehuel: RADAU_IIA_53 done solving, timings (ms | %):" and len(t) == 6
    assert t[1].split("|")[0].split()[-1] == "10" and t[4].strip().startswith("Integration kernels:") and t[4].strip().endswith("80")


def test_shared_work_cursor_across_processes(capi):
    """rehuel_cursor: four processes claim indices from one counter in POSIX shared memory; every index is handed out
    exactly once (the cross-process form of REHUEL_SHARD_DYNAMIC)."""
    import multiprocessing as mp
    name = f"rehuel_test_{os.getpid()}"
    cur = capi.Cursor(name, True)
    ctx = mp.get_context("fork")
    q = ctx.Queue()

    def worker():
        c = capi.Cursor(name, False)
        got = []
        while True:
            i = c.next(2)          # two indices per claim
            if i >= 1000:
                break
            got += [i, i + 1]
        c.close()
        q.put(got)

    procs = [ctx.Process(target=worker) for _ in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join()
    assert sorted(sum(res, [])) == list(range(1000))
    cur.reset()
    assert cur.next(1) == 0
    cur.close()
    with pytest.raises(capi.RehuelError):
        capi.Cursor(name, False)   # unlinked by its creator
    with pytest.raises(capi.RehuelError):
        capi.Cursor("a/b", True)
