"""BASELINE config 4 (65 536 members of the 64-equation Brusselator, LOBATTO_IIIC_85, t in [0,10]) through the C ABI
with host buffers on 1, 2, 4, 8 GPUs of one box (one process, members sharded by index, no data-path collective):
results must be bit-identical to the 1-GPU run -> one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rehuel_b200 import capi, ensembles

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
P, y0 = ensembles.bruss1d_params(M), ensembles.bruss1d_y0(32)
o = capi.make_options(1e-6, 1e-6)
out, base = {"members": M}, None
for g in (1, 2, 4, 8):
    if g > torch.cuda.device_count():
        break
    capi.irk_ensemble(capi.BRUSSELATOR_1D, 232, 0.0, 0.02, 1e-6, y0, P, o, n_gpus=g, n_hint=64)  # warm-up at full size: contexts, pools, cached slabs
    t = time.perf_counter()
    r = capi.irk_ensemble(capi.BRUSSELATOR_1D, 232, 0.0, 10.0, 1e-6, y0, P, o, n_gpus=g, n_hint=64)
    wall = (time.perf_counter() - t) * 1e3
    if base is None:
        base = r
    same = bool(np.array_equal(r.y, base.y) and all(np.array_equal(r.counters[k], base.counters[k]) for k in r.counters))
    out[f"gpus_{g}"] = dict(kernel_ms=r.kernel_ms, call_ms=wall, systems_per_s=M / wall * 1e3, failed=int((r.status != 0).sum()),
                            bit_identical_to_1gpu=same, speedup=out["gpus_1"]["call_ms"] / wall if g > 1 else 1.0)
print(json.dumps(out))
