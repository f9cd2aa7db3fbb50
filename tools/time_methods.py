"""Dev aid: device-timed Robertson sweep (2^18 members, t in [0,40]) for every tableau -> systems/s, attempts, kernel info."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble
M = 1 << 18
o = capi.make_options(1e-6, 1e-6)
P = torch.from_numpy(ensembles.robertson_params(M)).cuda(); y0 = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).cuda()
for method in (210, 211, 212, 213, 230, 231, 232):
    de = DeviceEnsemble(capi.ROBERTSON, method, M, y0, P, handout_order=capi.ORDER_LOCALITY_DESC)
    de.solve(0.0, 40.0, 1e-6, o); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); de.solve(0.0, 40.0, 1e-6, o); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    att = float(de.counter("attempt").double().mean()); its = float(de.counter("newton_iters").double().sum()) / float(de.counter("attempt").double().sum())
    print(f"{capi.method_to_name(method):16s} {best:9.3f} ms  {M/best*1e3:11.4g} systems/s  attempts/member {att:8.1f}  iters/attempt {its:.2f}  "
          f"us per 1000 attempts {best*1e3/(att*M)*1e3:.3f}  {capi.kernel_info(capi.ROBERTSON, method)}", flush=True)
