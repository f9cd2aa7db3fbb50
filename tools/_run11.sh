python -m pytest tests -m gpu -q > gpurun_out/gputests9.log 2>&1; tail -4 gpurun_out/gputests9.log
REHUEL_TIMING=1 python bench.py --steps 2 --no-cpu-baseline 2>&1 >/dev/null | tail -9
for c in 5 6 7 8; do REHUEL_CHUNKS=$c python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c e2e', d['e2e']['ms_per_step'], 'step', d['ms_per_step'])"; done
REHUEL_CHUNK_TAPER=0 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no taper e2e', d['e2e']['ms_per_step'])"
