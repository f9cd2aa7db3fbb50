REHUEL_CHUNK_TAPER=0 REHUEL_TIMING=1 python bench.py --steps 2 --no-cpu-baseline 2>&1 >/dev/null | tail -11
