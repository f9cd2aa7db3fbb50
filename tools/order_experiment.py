import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble
o = capi.make_options(1e-6, 1e-6)
for span, t1, M in (("short", 40.0, 1 << 20), ("long", 1e11, 1 << 18)):
    P = torch.from_numpy(ensembles.robertson_params(M)).cuda(); y0 = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).cuda()
    L = torch.log10(P)
    keys = {"none": None, "k1": L[0], "k2": L[1], "k3": L[2], "lin": 10.185 * L[0] + 1.079 * L[1] - 1.544 * L[2]}
    for name, key in keys.items():
        order = None if key is None else torch.argsort(key).to(torch.int32)
        de = DeviceEnsemble(capi.ROBERTSON, 211, M, y0, P, order=order)
        for _ in range(2): de.solve(0.0, t1, 1e-6, o)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); de.solve(0.0, t1, 1e-6, o); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{span} M={M} order={name}: {best:.3f} ms  ({M/best*1e3:.4g} systems/s)", flush=True)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
torch.argsort(L[0]); torch.cuda.synchronize()
e0.record(); torch.argsort(L[0]).to(torch.int32); e1.record(); torch.cuda.synchronize(); print("argsort ms", e0.elapsed_time(e1))
