python -m pytest tests -m gpu -x -q > gpurun_out/gputests3.log 2>&1; tail -5 gpurun_out/gputests3.log
python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -c 600 gpurun_out/bench4.err
python bench.py --handout index --no-cpu-baseline > gpurun_out/bench4_index.json 2>> gpurun_out/bench4.err
python bench.py --span long --members 262144 --steps 5 --no-cpu-baseline > gpurun_out/bench4_long.json 2>> gpurun_out/bench4.err
