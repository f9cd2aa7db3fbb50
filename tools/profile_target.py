"""Minimal launch sequence for ncu: N device-resident solves of the headline workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
span = sys.argv[2] if len(sys.argv) > 2 else "short"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t1 = 40.0 if span == "short" else 1e11
P = torch.from_numpy(ensembles.robertson_params(M)).cuda()
y0 = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).cuda()
ens = DeviceEnsemble(capi.ROBERTSON, capi.RADAU_IIA_53, M, y0, P)
o = capi.make_options(1e-6, 1e-6)
for _ in range(reps):
    ens.solve(0.0, t1, 1e-6, o)
torch.cuda.synchronize()
print("ok", int((ens.status == 0).sum()), float(ens.counter("attempt").double().mean()))
