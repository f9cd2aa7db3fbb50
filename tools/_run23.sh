python tools/time_variants.py build/var/base.so build/var/loop2.so build/var/base.so build/var/loop2.so 2>&1 | tee gpurun_out/variants6.log
