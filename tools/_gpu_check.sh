cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ID=$(python -c "from rehuel_b200 import capi; print(capi.BRUSSELATOR_1D_DENSE)")
python tools/run_config4.py 148 1.0 232 $ID 2>&1 | tail -1
python tools/run_config4.py 148 1.0 211 $ID 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -s -k "bruss or dense or user_functor or jacobians" 2>&1 | grep "cta \|passed\|failed\|Error\|error" | tail -25
