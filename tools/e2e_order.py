"""Dev aid: end-to-end time of rehuel_irk_ensemble (host buffers, 2^20 Robertson members) for every hand-out order and a few
chunk counts -- does the per-chunk ordering pay for itself on the pipelined host path?"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch, ctypes as C
from rehuel_b200 import capi, ensembles
M = 1 << 20
P = torch.from_numpy(ensembles.robertson_params(M)).pin_memory(); y0 = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).pin_memory()
for order in (capi.ORDER_INDEX, capi.ORDER_LOCALITY, capi.ORDER_LOCALITY_DESC):
    o = capi.make_options(1e-6, 1e-6, handout_order=order)
    def call():
        res = capi.Result()
        capi.check(capi.lib.rehuel_irk_ensemble(1, 211, 0.0, 40.0, 1e-6, M, 3, C.cast(y0.data_ptr(), capi._PD), 1, C.cast(P.data_ptr(), capi._PD), C.byref(o), 0, 1, C.byref(res)))
        k = res.kernel_ms
        capi.lib.rehuel_result_free(C.byref(res)); return k
    for _ in range(3): call()
    t = time.perf_counter(); ks = [call() for _ in range(20)]; el = (time.perf_counter() - t) / 20
    print("chunks", os.environ.get("REHUEL_CHUNKS"), "order", order, "e2e ms %%.3f kernel_ms %%.3f" %% (el * 1e3, sum(ks) / len(ks)), flush=True)
''' % ROOT
for c in ("4", "6", "8", "12"):
    subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, REHUEL_CHUNKS=c))
