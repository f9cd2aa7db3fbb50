cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ID=$(python -c "from rehuel_b200 import capi; print(capi.BRUSSELATOR_1D_DENSE)")
python tools/run_config4.py 16384 10.0 211 $ID 8 2>&1 | tail -1 | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ensemble_irk_wdense -c 1 -f -o gpurun_out/wdense_v1 python tools/run_config4.py 16384 10.0 211 $ID 8 > gpurun_out/wd_ncu.log 2>&1; tail -2 gpurun_out/wd_ncu.log
