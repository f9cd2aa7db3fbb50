"""Dev aid: device-timed throughput of a mass-swept three_body ensemble (n = 12, dense Jacobian: the warp-per-system
register-LU kernel) for the three tableaux of the BASELINE configs.  python tools/three_body_timing.py [M]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rehuel_b200 import capi  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
rng = np.random.default_rng(3)
P = np.ascontiguousarray(rng.uniform(0.5, 2.0, (3, M)))
y0 = np.array([1.0, 0.0, 3.0, 2.0, 0.0, 3.0, 0.0, 2.0, -1.0, 0.0, -1.0, 0.0])
out = {}
for method in (211, 230, 232):
    o = capi.make_options(1e-6, 1e-6)
    ms = []
    for _ in range(3):
        r = capi.irk_ensemble(capi.THREE_BODY, method, 0.0, 1.0, 1e-6, y0, P, o)
        ms.append(r.kernel_ms)
    att = float(r.counters["attempt"].sum())
    out[capi.method_to_name(method)] = {"members": M, "kernel_ms": min(ms), "systems_per_s": M / (min(ms) * 1e-3),
                                        "attempts_per_member": att / M, "attempts_per_s": att / (min(ms) * 1e-3),
                                        "failed": int(r.sum["n_failed"]), "kernel": capi.kernel_info(capi.THREE_BODY, method)}
print(json.dumps(out))
