cd $GRAFT_REPO_ROOT
python - <<'PY' 2>&1 | tail -12
import os, time
import numpy as np
from rehuel_b200 import capi, ensembles
M = 1 << 14
P, y0 = ensembles.bruss1d_params(M), ensembles.bruss1d_y0(32)
o = capi.make_options(1e-6, 1e-6)
for g in (1, 2):
    capi.irk_ensemble(capi.BRUSSELATOR_1D, 232, 0.0, 0.05, 1e-6, y0, P, o, n_gpus=g, n_hint=64)
os.environ["REHUEL_TIMING"] = "1"
for g in (1, 2, 2):
    t = time.perf_counter()
    r = capi.irk_ensemble(capi.BRUSSELATOR_1D, 232, 0.0, 2.0, 1e-6, y0, P, o, n_gpus=g, n_hint=64)
    print("gpus", g, "call ms", (time.perf_counter() - t) * 1e3, "kernel ms", r.kernel_ms, flush=True)
PY
