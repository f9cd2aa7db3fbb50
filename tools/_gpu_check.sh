cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/config4_timing.py 2 2>&1 | tail -1 | tee gpurun_out/w4_config4.json
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
