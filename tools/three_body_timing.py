"""Dev aid: device-timed throughput of a mass-swept three_body ensemble (n = 12, dense Jacobian: the warp-per-system
register-LU kernel) for the three tableaux of the BASELINE configs.  python tools/three_body_timing.py [M]
Masses are swept over [0.8, 2]: with m2 < 0.65 two bodies collide before t = 1 and the integration degenerates exactly as
the reference's does (oracle/irk_port.c on such a member: ~960 Newton failures within 200 accepted steps, and since
`maxit` only ever grows by maxit0 per failure, irk.hpp:634-665, the cost is quadratic in the failures) -- 65 536 members
over [0.5, 2] did not finish in 5 minutes on the GPU, nor does a single such member on the CPU.  newton_maxit = 50 and
max_steps = 5000 bound the run in case a member still gets there (timing aid, not a parity setting)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rehuel_b200 import capi  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(3)
P = np.ascontiguousarray(rng.uniform(0.8, 2.0, (3, M)))  # see the note above: m2 < 0.65 collides before t = 1
y0 = np.array([1.0, 0.0, 3.0, 2.0, 0.0, 3.0, 0.0, 2.0, -1.0, 0.0, -1.0, 0.0])
out = {}
for method in (211, 230, 232):
    o = capi.make_options(1e-6, 1e-6, newton_maxit=50, max_steps=5000)
    ms = []
    for _ in range(3):
        r = capi.irk_ensemble(capi.THREE_BODY, method, 0.0, 1.0, 1e-6, y0, P, o)
        ms.append(r.kernel_ms)
    att = float(r.counters["attempt"].sum())
    out[capi.method_to_name(method)] = {"members": M, "kernel_ms": min(ms), "systems_per_s": M / (min(ms) * 1e-3),
                                        "attempts_per_member": att / M, "attempts_per_s": att / (min(ms) * 1e-3),
                                        "failed": int(r.sum["n_failed"]), "kernel": capi.kernel_info(capi.THREE_BODY, method)}
print(json.dumps(out))
