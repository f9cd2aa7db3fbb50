python tools/run_config4.py 888 10.0 232
python tools/run_config4.py 888 10.0 211
ncu --set full --clock-control none --import-source on -k regex:ensemble_irk_warp -c 1 -o gpurun_out/prof_r1_warp_v1 python tools/run_config4.py 888 0.5 232 > gpurun_out/ncu_warp1.log 2>&1; tail -2 gpurun_out/ncu_warp1.log
