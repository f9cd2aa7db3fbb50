"""Device-timed run of BASELINE.json config 4 (1-D Brusselator MOL, n = 64, LOBATTO_IIIC_85, t in [0,10]) on a
sample of the sweep: `waves` full waves of the persistent grid (default 2) -> one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble

waves = int(sys.argv[1]) if len(sys.argv) > 1 else 2
info = capi.kernel_info(capi.BRUSSELATOR_1D, 232, 64)
M = 148 * info["ctas_per_sm"] * waves
o = capi.make_options(1e-6, 1e-6)
P = torch.from_numpy(ensembles.bruss1d_params(M)).cuda(); y0 = torch.from_numpy(ensembles.bruss1d_y0(32)).cuda()
out = {"kernel": info}
for method, name in ((232, "lobatto85"), (211, "radau53")):
    de = DeviceEnsemble(capi.BRUSSELATOR_1D, method, M, y0, P, n_hint=64)
    de.solve(0.0, 10.0, 1e-6, o); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); de.solve(0.0, 10.0, 1e-6, o); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    att = float(de.counter("attempt").double().sum())
    out[f"config4_bruss1d_n64_{name}"] = dict(members=M, ms=ms, systems_per_s=M / ms * 1e3, attempts_per_member=att / M,
                                              newton_iters_per_attempt=float(de.counter("newton_iters").double().sum()) / att,
                                              us_per_attempt_per_sm=ms * 1e3 / (att / 148), failed=int((de.status != 0).sum()),
                                              extrapolated_s_for_65536=65536 / (M / ms * 1e3))
print(json.dumps(out))
