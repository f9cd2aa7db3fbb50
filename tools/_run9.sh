timeout 600 python -m pytest tests -m gpu -x -q -k "brusselator or golden_single" -s > gpurun_out/gputests8.log 2>&1; grep "bruss1d\|passed\|failed\|Error" gpurun_out/gputests8.log | head -40
timeout 600 python tools/config_timings.py > gpurun_out/configs_r1c.json 2> gpurun_out/configs_r1c.err; tail -c 300 gpurun_out/configs_r1c.err; cat gpurun_out/configs_r1c.json
