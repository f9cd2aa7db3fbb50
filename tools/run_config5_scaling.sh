python tools/config5.py > gpurun_out/config5_n1.json 2> gpurun_out/config5_n1.err; cat gpurun_out/config5_n1.json
for n in 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 tools/config5.py > gpurun_out/config5_n$n.json 2> gpurun_out/config5_n$n.err; tail -1 gpurun_out/config5_n$n.json; done
