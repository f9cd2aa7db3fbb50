python -m pytest tests -m gpu -q > gpurun_out/gputests_final.log 2>&1; tail -3 gpurun_out/gputests_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>> gpurun_out/bench_final.err
python bench.py --span long --members 262144 --steps 5 --no-cpu-baseline > gpurun_out/bench_final_long.json 2>> gpurun_out/bench_final.err
python tools/config_timings.py > gpurun_out/configs_final.json 2> gpurun_out/configs_final.err; tail -c 300 gpurun_out/configs_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ensemble_irk_thread -s 3 -c 1 -o gpurun_out/prof_r1_v7 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu7.log 2>&1; tail -2 gpurun_out/ncu7.log
