python tools/explore_order.py > gpurun_out/explore_order.log 2>&1
python tools/time_variants.py default build/var/minb5.so build/var/minb6.so build/var/b64m8.so build/var/b256m2.so > gpurun_out/variants1.log 2>&1
