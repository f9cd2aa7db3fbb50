python -m pytest tests -m gpu -q -x -k "locality or handout or placement or full_size" 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/bench_ord.json 2> gpurun_out/bench_ord.err; tail -c 300 gpurun_out/bench_ord.err
for b in 12 15 18 21; do REHUEL_ORDER_BITS=$b python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bits $b step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms_per_launch'],4), 'e2e', round(d['e2e']['ms_per_step'],3))"; done
