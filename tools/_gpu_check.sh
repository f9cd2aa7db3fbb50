cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/config_timings.py > gpurun_out/configs_v4.json 2> gpurun_out/configs_v4.err; tail -c 1500 gpurun_out/configs_v4.json; tail -3 gpurun_out/configs_v4.err
python tools/run_config4.py 65536 10.0 232 2>&1 | tail -1 | tee gpurun_out/config4_full.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ensemble_irk_warp -c 1 -f -o gpurun_out/warp_v5 python tools/run_config4.py 1628 1.0 232 > gpurun_out/w5_ncu.log 2>&1; tail -2 gpurun_out/w5_ncu.log
