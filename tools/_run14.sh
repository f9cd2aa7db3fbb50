run() { python bench.py --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 e2e', round(d['e2e']['ms_per_step'],3))"; }
for c in 8 9 10 12 14 16; do REHUEL_CHUNKS=$c run eq$c; done
REHUEL_CHUNKS=12 REHUEL_TIMING=1 python bench.py --steps 2 --no-cpu-baseline 2>&1 >/dev/null | tail -15
