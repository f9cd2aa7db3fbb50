#!/usr/bin/env python
"""Developer timing aids for the GPU box, one CLI (run under gpurun; device-timed with CUDA events).

  python tools/gpu_dev.py variants [LIB.so ...]      headline kernel, short + long span, Van der Pol: per library variant
                                                     (tools/build_variant.sh builds them; "default" = the in-tree library)
  python tools/gpu_dev.py e2e                        rehuel_irk_ensemble with host buffers: chunks x result mask sweep + phases
  python tools/gpu_dev.py methods                    every tableau on the Robertson sweep
  python tools/gpu_dev.py solve PROBLEM METHOD M T1 [N]   one ensemble (for ncu captures): e.g. solve 3 232 1184 10 64
  python tools/gpu_dev.py config4 [M]                config 4 on the band kernel and on the dense CTA kernel
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _timed(torch, fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def _ensemble(problem, method, M, t1, n=0, handout=1):
    import torch
    from rehuel_b200 import capi, ensembles
    from rehuel_b200.device import DeviceEnsemble
    cu = lambda a: torch.from_numpy(a.copy()).cuda()
    if problem == capi.ROBERTSON:
        y0, P = ensembles.ROBERTSON_Y0, ensembles.robertson_params(M)
    elif problem == capi.VAN_DER_POL:
        y0, P = ensembles.VDPOL_Y0, ensembles.vdpol_params(M)
    else:
        y0, P = ensembles.bruss1d_y0((n or 64) // 2), ensembles.bruss1d_params(M)
    return DeviceEnsemble(problem, method, M, cu(y0), cu(P), n_hint=n, handout_order=handout if problem in (1, 2) else 0)


def cmd_variants_child():
    import torch
    from rehuel_b200 import capi
    o = capi.make_options(1e-6, 1e-6)
    out = []
    for span, t1, M in (("short", 40.0, 1 << 20), ("long", 1e11, 1 << 18)):
        de = _ensemble(capi.ROBERTSON, 211, M, t1, handout=int(os.environ.get("TV_HANDOUT", "2")))
        ms = _timed(torch, lambda: de.solve(0.0, t1, 1e-6, o), warm=2)
        out.append(f"{span} {ms:.3f} ms (att {float(de.counter('attempt').double().mean()):.2f})")
        if os.environ.get("TV_DIV"):  # libraries built with -DREHUEL_COUNT_SKIPS=3: iterations the warp ran / iterations the lane needed
            out.append(f"warp-its/lane-its {float(de.counter('newton_maxit_exceed').double().sum() / de.counter('newton_iters').double().sum()):.3f}"
                       f" its/att {float(de.counter('newton_iters').double().sum() / de.counter('newton_success').double().sum()):.3f}")
    vm = int(os.environ.get("TV_VDP_M", str(1 << 18)))
    de = _ensemble(capi.VAN_DER_POL, 230, vm, 2.0)
    ms = _timed(torch, lambda: de.solve(0.0, 2.0, 1e-6, o), reps=2)
    out.append(f"vdpol-{vm >> 10}k {ms:.2f} ms (rejN {int(de.counter('reject_newton').sum())}, fun {float(de.counter('fun_evals').double().sum()):.4g})")
    print(os.path.basename(os.environ.get("REHUEL_B200_LIB", "default")), capi.kernel_info(capi.ROBERTSON, 211), " | ".join(out), flush=True)


def cmd_variants(libs):
    for lib in libs or ["default"]:
        env = dict(os.environ)
        if lib != "default":
            env["REHUEL_B200_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "_variants_child"], env=env)


def cmd_e2e_child():
    import ctypes as C
    import time
    import torch
    from rehuel_b200 import capi, ensembles
    M = 1 << 20
    P = torch.from_numpy(ensembles.robertson_params(M)).pin_memory()
    y0 = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).pin_memory()
    for mask in (0, 7):
        o = capi.make_options(1e-6, 1e-6, handout_order=2, result_mask=mask, time_internals=True)

        def call():
            res = capi.Result()
            capi.check(capi.lib.rehuel_irk_ensemble(1, 211, 0.0, 40.0, 1e-6, M, 3, C.cast(y0.data_ptr(), capi._PD), 1,
                                                    C.cast(P.data_ptr(), capi._PD), C.byref(o), 0, 1, C.byref(res)))
            k, ph = res.kernel_ms, list(res.phase_ms)
            capi.lib.rehuel_result_free(C.byref(res))
            return k, ph
        for _ in range(3):
            call()
        t = time.perf_counter()
        rs = [call() for _ in range(10)]
        el = (time.perf_counter() - t) / 10
        print(f"chunks {os.environ.get('REHUEL_CHUNKS', 'auto'):>4} mask {mask}: e2e {el * 1e3:.3f} ms, kernels first-start..last-end {sum(r[0] for r in rs) / 10:.3f} ms, "
              f"phase sums h2d/order/kernel/d2h {[round(sum(r[1][i] for r in rs) / 10, 3) for i in range(4)]}", flush=True)


def cmd_e2e():
    for c in ("2", "4", "6", "8", "12", "16"):
        subprocess.run([sys.executable, os.path.abspath(__file__), "_e2e_child"], env=dict(os.environ, REHUEL_CHUNKS=c))


def cmd_methods():
    import torch
    from rehuel_b200 import capi
    M, o = 1 << 18, None
    from rehuel_b200 import capi
    o = capi.make_options(1e-6, 1e-6)
    for method in (210, 211, 212, 213, 230, 231, 232):
        de = _ensemble(capi.ROBERTSON, method, M, 40.0, handout=2)
        ms = _timed(torch, lambda: de.solve(0.0, 40.0, 1e-6, o), reps=2)
        att = float(de.counter("attempt").double().mean())
        print(f"{capi.method_to_name(method):16s} {ms:9.3f} ms  {M / ms * 1e3:11.4g} systems/s  attempts/member {att:8.1f}  {capi.kernel_info(capi.ROBERTSON, method)}", flush=True)


def cmd_solve(problem, method, M, t1, n=0):
    import torch
    from rehuel_b200 import capi
    o = capi.make_options(1e-6, 1e-6)
    de = _ensemble(problem, method, M, t1, n)
    ms = _timed(torch, lambda: de.solve(0.0, t1, 1e-6, o), reps=1)
    att = float(de.counter("attempt").double().sum())
    print(f"problem {problem} method {method} n={n} M={M} t1={t1}: {ms:.3f} ms, {M / ms * 1e3:.4g} systems/s, attempts/member {att / M:.1f}, "
          f"iters/attempt {float(de.counter('newton_iters').double().sum()) / att:.2f}, failed {int((de.status != 0).sum())}, {capi.kernel_info(problem, method, n)}")


def cmd_config4(M):
    from rehuel_b200 import capi
    cmd_solve(capi.BRUSSELATOR_1D, 232, M, 10.0, 64)
    cmd_solve(capi.BRUSSELATOR_1D_DENSE, 232, min(M, 296), 0.5, 64)


if __name__ == "__main__":
    a = sys.argv[1:]
    if not a:
        sys.exit(__doc__)
    if a[0] == "_variants_child":
        cmd_variants_child()
    elif a[0] == "_e2e_child":
        cmd_e2e_child()
    elif a[0] == "variants":
        cmd_variants(a[1:])
    elif a[0] == "e2e":
        cmd_e2e()
    elif a[0] == "methods":
        cmd_methods()
    elif a[0] == "solve":
        cmd_solve(int(a[1]), int(a[2]), int(a[3]), float(a[4]), int(a[5]) if len(a) > 5 else 0)
    elif a[0] == "config4":
        cmd_config4(int(a[1]) if len(a) > 1 else 4096)
    else:
        sys.exit(__doc__)
