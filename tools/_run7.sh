python tools/time_variants.py build/var/spec1.so build/var/spec2.so > gpurun_out/variants5.log 2>&1; cat gpurun_out/variants5.log
REHUEL_B200_LIB=$PWD/build/var/spec2.so python -m pytest tests -m gpu -x -q > gpurun_out/gputests6.log 2>&1; tail -5 gpurun_out/gputests6.log
