"""Config-3 shaped timing: Van der Pol sweep, LOBATTO_IIIC_43, with/without cost-ordered hand-out (dev aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble
o = capi.make_options(1e-6, 1e-6)
for M in (1 << 20,):
    P = torch.from_numpy(ensembles.vdpol_params(M)).cuda(); y0 = torch.from_numpy(ensembles.VDPOL_Y0.copy()).cuda()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for name, order in (("index order", None), ("stiff first", torch.arange(M - 1, -1, -1, dtype=torch.int32, device="cuda")),
                        ("shuffled", torch.randperm(M, device="cuda", generator=g).to(torch.int32))):
        de = DeviceEnsemble(capi.VAN_DER_POL, 230, M, y0, P, order=order)
        de.solve(0.0, 2.0, 1e-6, o); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); de.solve(0.0, 2.0, 1e-6, o); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{os.environ.get('REHUEL_B200_LIB','default')[-16:]} M={M} {name}: {ms:.1f} ms, {M/ms*1e3:.3g} systems/s, rejN {int(de.counter('reject_newton').sum())}, "
              f"fun {float(de.counter('fun_evals').double().sum()):.3g}, ok {int((de.status==0).sum())}", flush=True)
