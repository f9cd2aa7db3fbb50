python -m pytest tests -m gpu -q > gpurun_out/gputests7.log 2>&1; tail -5 gpurun_out/gputests7.log
python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -c 300 gpurun_out/bench7.err
python tools/config_timings.py > gpurun_out/configs_r1c.json 2> gpurun_out/configs_r1c.err; tail -c 300 gpurun_out/configs_r1c.err; cat gpurun_out/configs_r1c.json
