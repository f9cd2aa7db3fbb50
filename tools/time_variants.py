"""Time the headline kernel for every library variant given on the command line (dev aid)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import torch
from rehuel_b200 import capi, ensembles
from rehuel_b200.device import DeviceEnsemble
o = capi.make_options(1e-6, 1e-6)
out = []
for span, t1, M in (("short", 40.0, 1 << 20), ("long", 1e11, 1 << 18)):
    P = torch.from_numpy(ensembles.robertson_params(M)).cuda(); y0 = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).cuda()
    de = DeviceEnsemble(capi.ROBERTSON, 211, M, y0, P, handout_order=int(os.environ.get('TV_HANDOUT', '1')))
    for _ in range(2): de.solve(0.0, t1, 1e-6, o)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); de.solve(0.0, t1, 1e-6, o); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out.append("%%s %%.3f ms" %% (span, best))
M = 1 << 16
P = torch.from_numpy(ensembles.vdpol_params(M)).cuda(); y0 = torch.from_numpy(ensembles.VDPOL_Y0.copy()).cuda()
de = DeviceEnsemble(capi.VAN_DER_POL, 230, M, y0, P, handout_order=int(os.environ.get('TV_HANDOUT', '1')))
de.solve(0.0, 2.0, 1e-6, o); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); de.solve(0.0, 2.0, 1e-6, o); e1.record(); torch.cuda.synchronize()
out.append("vdpol-64k %%.1f ms (rejN %%d, fun %%.3g)" %% (e0.elapsed_time(e1), int(de.counter("reject_newton").sum()), float(de.counter("fun_evals").double().sum())))
info = capi.kernel_info(capi.ROBERTSON, 211)
print(os.environ.get("REHUEL_B200_LIB", "default"), info, " | ".join(out), flush=True)
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default":
        env["REHUEL_B200_LIB"] = os.path.abspath(lib)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
