python -m pytest tests -m gpu -q > gpurun_out/gputests13.log 2>&1; tail -3 gpurun_out/gputests13.log
python bench.py --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -c 300 gpurun_out/bench8.err
