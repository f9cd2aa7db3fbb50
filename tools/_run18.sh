timeout 900 python -m pytest tests -m gpu -x -q -s -k "brusselator or continue or dense" > gpurun_out/gputests12.log 2>&1; grep "bruss1d\|passed\|failed\|Error" gpurun_out/gputests12.log | head -30
python tools/run_config4.py 888 10.0 232
python tools/run_config4.py 888 10.0 211
python tools/run_config4.py 7104 10.0 211
