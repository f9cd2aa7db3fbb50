set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/w1_small.log 2>&1
import numpy as np
from rehuel_b200 import capi, ensembles
print(capi.kernel_info(capi.BRUSSELATOR_1D, 232, 64))
P, y0 = ensembles.bruss1d_params(4), ensembles.bruss1d_y0(32)
g = capi.irk_ensemble(capi.BRUSSELATOR_1D, 211, 0.0, 1.0, 1e-6, y0, P, capi.make_options(1e-6, 1e-6))
print(g.status, g.t_final, g.counters["attempt"], g.y[:4, 0])
PY
tail -5 gpurun_out/w1_small.log
cat > /tmp/small.py <<'PY'
from rehuel_b200 import capi, ensembles
P, y0 = ensembles.bruss1d_params(2), ensembles.bruss1d_y0(32)
o = capi.make_options(1e-6, 1e-6)
g = capi.irk_ensemble(capi.BRUSSELATOR_1D, 232, 0.0, 0.02, 1e-6, y0, P, o)
print(g.status, g.counters["attempt"])
capi.debug_force_pivot(True)
g = capi.irk_ensemble(capi.BRUSSELATOR_1D, 232, 0.0, 0.02, 1e-6, y0, P, o)
print(g.status, g.counters["attempt"])
PY
timeout 300 compute-sanitizer --tool memcheck python /tmp/small.py > gpurun_out/w1_memcheck.log 2>&1; tail -8 gpurun_out/w1_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python /tmp/small.py > gpurun_out/w1_racecheck.log 2>&1; tail -8 gpurun_out/w1_racecheck.log
timeout 900 python -m pytest tests -m gpu -x -q -s -k "bruss or band or continue or dense or concurrent or pivoted" > gpurun_out/w1_pytest.log 2>&1; tail -30 gpurun_out/w1_pytest.log
timeout 300 python tools/config4_timing.py 2 > gpurun_out/w1_config4.json 2> gpurun_out/w1_config4.err; cat gpurun_out/w1_config4.json; tail -3 gpurun_out/w1_config4.err
