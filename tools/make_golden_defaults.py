#!/usr/bin/env python
"""Generate tests/golden/odeint_defaults.json: the reference's 4-argument irk::odeint(func, t0, t1, y0)
(irk.hpp:915-926), which picks method, first step, tolerances and Newton options itself, run UNMODIFIED
(oracle/_ref) on a few systems.  The fixture pins what those defaults ARE: tests/test_oracle.py reproduces it
bit for bit with the port under explicitly stated options, tests/test_gpu_parity.py with the CUDA path.

Run (only where /root/reference exists):  make -C oracle && python tools/make_golden_defaults.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refsolver import RefSolver  # noqa: E402

CASES = [("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 0.0, 40.0),
         ("robertson", [0.04, 1e4, 3e7], [1.0, 0.0, 0.0], 0.0, 1e6),
         ("robertson", [0.02, 3e4, 1e7], [1.0, 0.0, 0.0], 0.0, 1e3),
         ("van-der-pol", [1e-3], [2.0, -0.66], 0.0, 2.0),
         ("van-der-pol", [1.0], [2.0, -0.66], 0.0, 10.0),
         ("exponential", [-2.0], [1.0], 0.0, 3.0),
         ("brusselator", [1.0, 2.5], [1.5, 3.0], 0.0, 20.0)]


def main():
    ref = RefSolver("reference")
    hx = lambda a: [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]
    out = []
    for prob, par, y0, t0, t1 in CASES:
        r = ref.odeint_defaults(prob, t0, t1, y0, par)
        out.append(dict(name=f"{prob}_defaults_t{t1:g}", problem=prob, params=par, y0=y0, t0=t0, t1=t1, status=int(r.status[0]),
                        y=hx(r.y[:, 0]), t_final=hx(r.t_final), err_last=hx(r.err_last),
                        counters={k: int(v[0]) for k, v in r.counters.items()}))
        print(out[-1]["name"], out[-1]["status"], r.y[:, 0], out[-1]["counters"])
    with open(os.path.join(ROOT, "tests", "golden", "odeint_defaults.json"), "w") as fh:
        json.dump(dict(source="unmodified reference, irk::odeint(func, t0, t1, y0)", cases=out), fh, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
