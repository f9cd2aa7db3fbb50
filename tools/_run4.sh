python -m pytest tests -m gpu -q > gpurun_out/gputests5.log 2>&1; tail -15 gpurun_out/gputests5.log
python tools/time_variants.py build/var/base.so default > gpurun_out/variants2.log 2>&1; cat gpurun_out/variants2.log
python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -c 300 gpurun_out/bench6.err
REHUEL_TIMING=1 python bench.py --steps 2 --no-cpu-baseline 2>&1 >/dev/null | tail -9
for c in 4 6 8; do REHUEL_CHUNKS=$c python bench.py --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks $c', d['e2e']['ms_per_step'])"; done
REHUEL_CHUNK_TAPER=0 python bench.py --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no taper', d['e2e']['ms_per_step'])"
