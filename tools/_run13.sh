run() { python bench.py --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 e2e', round(d['e2e']['ms_per_step'],3))"; }
for rep in 1 2; do
REHUEL_CHUNK_WEIGHTS=1,1,1,1,1,1,1,1 run eq8
REHUEL_CHUNK_WEIGHTS=4,4,4,4,4,3,2,1 run tail8
REHUEL_CHUNK_WEIGHTS=2,4,4,4,4,3,2,1 run both8
REHUEL_CHUNK_WEIGHTS=1,2,4,4,4,4,2,1 run sym8
REHUEL_CHUNK_WEIGHTS=1,1,1,1,1,1 run eq6
REHUEL_CHUNK_WEIGHTS=2,3,3,3,2,1 run t6
REHUEL_CHUNK_WEIGHTS=1,3,4,4,3,2,1 run t7
done
