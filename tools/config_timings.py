"""Device-timed runs of BASELINE.json configs 3, 4, 5 (the non-headline configs) -> one JSON line.
These are parity-test configurations, not bench lines; the numbers document where the kernels stand."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble

def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
o = capi.make_options(1e-6, 1e-6)
# config 3: Van der Pol sweep, LOBATTO_IIIC_43, 2^20 members (one GPU's worth; 8-GPU sharding is i mod 8)
M = 1 << 20
P = torch.from_numpy(ensembles.vdpol_params(M)).cuda(); y0 = torch.from_numpy(ensembles.VDPOL_Y0.copy()).cuda()
for name, order in (("index_order", None), ("stiff_first", torch.arange(M - 1, -1, -1, dtype=torch.int32, device="cuda"))):
    de = DeviceEnsemble(capi.VAN_DER_POL, 230, M, y0, P, order=order)
    ms = timed(lambda: de.solve(0.0, 2.0, 1e-6, o))
    out[f"config3_vdpol_2^20_{name}"] = dict(ms=ms, systems_per_s=M / ms * 1e3, newton_rejections=int(de.counter("reject_newton").sum()),
                                             fun_evals=float(de.counter("fun_evals").double().sum()), failed=int((de.status != 0).sum()))
# config 4: Brusselator MOL n=64, LOBATTO_IIIC_85, t in [0,10]; 1184 members = 8 per SM (65536 at full size)
M = 148 * 8
P = torch.from_numpy(ensembles.bruss1d_params(M)).cuda(); y0 = torch.from_numpy(ensembles.bruss1d_y0(32)).cuda()
de = DeviceEnsemble(capi.BRUSSELATOR_1D, 232, M, y0, P, n_hint=64)
ms = timed(lambda: de.solve(0.0, 10.0, 1e-6, o), reps=1)
att = float(de.counter("attempt").double().sum())
out["config4_bruss1d_n64_lobatto85"] = dict(members=M, ms=ms, systems_per_s=M / ms * 1e3, attempts_per_member=att / M,
                                            us_per_attempt_per_sm=ms * 1e3 / (att / 148), failed=int((de.status != 0).sum()),
                                            extrapolated_s_for_65536=65536 / (M / ms * 1e3),
                                            kernel=capi.kernel_info(capi.BRUSSELATOR_1D, 232, 64))
de = DeviceEnsemble(capi.BRUSSELATOR_1D, 211, M, y0, P, n_hint=64)
ms = timed(lambda: de.solve(0.0, 10.0, 1e-6, o), reps=1)
out["config4_bruss1d_n64_radau53"] = dict(members=M, ms=ms, systems_per_s=M / ms * 1e3, attempts_per_member=float(de.counter("attempt").double().mean()))
# config 5: mixed Robertson + Van der Pol, per-member rtol 1e-3..1e-9, two kernels on two streams
M = 1 << 21
tol = ensembles.mixed_tolerances(M)
halves = []
for which, (prob, method, t1, y0h, P) in enumerate(((capi.ROBERTSON, 211, 40.0, ensembles.ROBERTSON_Y0, ensembles.robertson_params(M // 2)),
                                                    (capi.VAN_DER_POL, 230, 2.0, ensembles.VDPOL_Y0, ensembles.vdpol_params(M // 2)))):
    tl = torch.from_numpy(tol[which::2].copy()).cuda()
    halves.append((DeviceEnsemble(prob, method, M // 2, torch.from_numpy(y0h.copy()).cuda(), torch.from_numpy(P).cuda(),
                                  rtol=tl, atol=tl.clone(), newton_tol=(0.1 * tl).contiguous()), t1))
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def run5():
    cur = torch.cuda.current_stream()
    for (de, t1), st in zip(halves, streams):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            de.solve(0.0, t1, 1e-6, o)
    for st in streams:
        cur.wait_stream(st)
ms = timed(run5, reps=1)
out["config5_mixed_2^21_rtol_sweep"] = dict(members=M, ms=ms, systems_per_s=M / ms * 1e3,
                                            failed=int(sum(int((de.status != 0).sum()) for de, _ in halves)),
                                            attempts_per_member=[float(de.counter("attempt").double().mean()) for de, _ in halves])
print(json.dumps(out))
