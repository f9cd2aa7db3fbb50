import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from rehuel_b200 import capi, ensembles
M = 1 << 20
P = torch.from_numpy(ensembles.robertson_params(M)).pin_memory(); y0 = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).pin_memory()
o = capi.make_options(1e-6, 1e-6)
def call():
    res = capi.Result()
    capi.check(capi.lib.rehuel_irk_ensemble(1, 211, 0.0, 40.0, 1e-6, M, 3, C.cast(y0.data_ptr(), capi._PD), 1, C.cast(P.data_ptr(), capi._PD), C.byref(o), 0, 1, C.byref(res)))
    capi.lib.rehuel_result_free(C.byref(res))
for _ in range(3): call()
os.environ["REHUEL_TIMING"] = "1"
t = time.perf_counter(); call(); print("e2e ms", (time.perf_counter() - t) * 1e3)
