python -m pytest tests -m gpu -x -q > gpurun_out/gputests4.log 2>&1; tail -5 gpurun_out/gputests4.log
REHUEL_TIMING=1 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench5_timing.json 2> gpurun_out/bench5_timing.err; tail -40 gpurun_out/bench5_timing.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
