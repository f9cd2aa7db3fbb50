python tools/time_variants.py build/var/piv1.so build/var/piv2.so build/var/piv0.so > gpurun_out/variants4.log 2>&1; cat gpurun_out/variants4.log
TV_HANDOUT=0 python tools/time_variants.py build/var/piv1.so build/var/piv2.so build/var/piv0.so > gpurun_out/variants4i.log 2>&1; cat gpurun_out/variants4i.log
