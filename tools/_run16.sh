timeout 900 python -m pytest tests -m gpu -x -q -s -k "continue or dense" > gpurun_out/gputests10.log 2>&1; tail -30 gpurun_out/gputests10.log
