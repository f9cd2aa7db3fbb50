for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/scale2_$n.json 2> gpurun_out/scale2_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale2_$n.json 2> gpurun_out/scale2_$n.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale2_$n.json").read().strip().splitlines()[-1])
    print($n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
except Exception as e: print($n, "failed", e)
PY
done
