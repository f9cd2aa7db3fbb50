"""Quick on-GPU sanity run (development aid; the real checks live in tests/)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble
from oracle.refsolver import RefSolver, make_options as ref_options

def scaled_err(y, yref, rtol, atol):
    return np.max(np.abs(y - yref) / (atol + rtol * np.abs(yref)), axis=0)

def main():
    print(torch.cuda.get_device_name(0))
    print("kernel info rober/radau53:", capi.kernel_info(capi.ROBERTSON, capi.RADAU_IIA_53))
    ref = RefSolver("reference"); port = RefSolver("port")
    o = capi.make_options(1e-6, 1e-6); ro = ref_options(1e-6, 1e-6)
    p1 = np.array([[0.04], [1e4], [3e7]])
    for t1 in (40.0, 1e11):
        for m in (211, 230, 232):
            if m != 211 and t1 > 40: continue
            g = capi.irk_ensemble(capi.ROBERTSON, m, 0.0, t1, 1e-6, ensembles.ROBERTSON_Y0, p1, o)
            r = ref.solve("robertson", m, 0.0, t1, 1e-6, ensembles.ROBERTSON_Y0, p1, ro)
            print(f"m={m} t1={t1:g} gpu y={g.y[:,0]} ref y={r.y[:,0]} serr={scaled_err(g.y, r.y, 1e-6, 1e-6)}")
            print("   gpu:", {k: int(v[0]) for k, v in g.counters.items()}, "status", g.status)
            print("   ref:", {k: int(v[0]) for k, v in r.counters.items()})
    # ensemble parity on 4096 members, short span
    M = 4096
    P = ensembles.robertson_params(M)
    g = capi.irk_ensemble(capi.ROBERTSON, 211, 0.0, 40.0, 1e-6, ensembles.ROBERTSON_Y0, P, o)
    r = port.solve("robertson", 211, 0.0, 40.0, 1e-6, ensembles.ROBERTSON_Y0, P, ro)
    se = scaled_err(g.y, r.y, 1e-6, 1e-6)
    same_steps = np.mean(g.counters["attempt"] == r.counters["attempt"])
    print(f"4096 members: max scaled err {se.max():.3e}, median {np.median(se):.3e}, same attempt count {same_steps:.4f}, "
          f"gpu attempts mean {g.counters['attempt'].mean():.2f} ref {r.counters['attempt'].mean():.2f} kernel_ms {g.kernel_ms:.3f}")
    # timing, device-resident
    for span, t1 in (("short", 40.0), ("long", 1e11)):
        for M in (1 << 16, 1 << 20):
            if span == "long" and M > (1 << 18): M = 1 << 18
            P = torch.from_numpy(ensembles.robertson_params(M)).cuda()
            y0 = torch.from_numpy(ensembles.ROBERTSON_Y0).cuda()
            de = DeviceEnsemble(capi.ROBERTSON, 211, M, y0, P)
            for _ in range(2):
                de.solve(0.0, t1, 1e-6, o)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); de.solve(0.0, t1, 1e-6, o); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            att = de.counter("attempt").double().sum().item(); its = de.counter("newton_iters").double().sum().item()
            flops = 873 * att + 321 * its
            ok = int((de.status == 0).sum().item())
            print(f"{span} M={M}: {ms:.3f} ms, {M/ms*1e3:.4g} systems/s, attempts/member {att/M:.1f}, iters/attempt {its/att:.2f}, "
                  f"dense-model {flops/ms/1e9:.2f} TFLOP/s, ok {ok}/{M}, masscons {float((de.y.sum(0)-1).abs().max()):.2e}")

if __name__ == "__main__":
    main()
