"""Dev aid: one device-timed solve of the config-4 shape (1-D Brusselator MOL, n = 64) -- for ncu captures and timing.
usage: run_config4.py [M] [t1] [method] [problem-id] [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble
M = int(sys.argv[1]) if len(sys.argv) > 1 else 888
t1 = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
method = int(sys.argv[3]) if len(sys.argv) > 3 else 232
prob = int(sys.argv[4]) if len(sys.argv) > 4 else capi.BRUSSELATOR_1D
n = int(sys.argv[5]) if len(sys.argv) > 5 else 64
o = capi.make_options(1e-6, 1e-6)
P = torch.from_numpy(ensembles.bruss1d_params(M)).cuda(); y0 = torch.from_numpy(ensembles.bruss1d_y0(n // 2)).cuda()
de = DeviceEnsemble(prob, method, M, y0, P, n_hint=n)
for rep in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); de.solve(0.0, t1, 1e-6, o); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
att = float(de.counter("attempt").double().sum()); its = float(de.counter("newton_iters").double().sum())
info = capi.kernel_info(prob, method, n)
waves = -(-M // (148 * info["ctas_per_sm"]))
print(f"n={n} M={M} t1={t1} method={method}: {ms:.2f} ms, attempts/member {att/M:.1f}, newton iters/attempt {its/att:.2f}, "
      f"{ms*1e3/(att/M)/waves:.1f} us per attempt per {'warp' if info['block']==32 else 'CTA'} ({waves} wave(s)), failed {int((de.status!=0).sum())}, {info}")
