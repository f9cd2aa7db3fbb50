timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputests11.log 2>&1; tail -30 gpurun_out/gputests11.log
