cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; tail -c 600 gpurun_out/bench_s3.json; tail -2 gpurun_out/bench_s3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_s3_ref.json 2> gpurun_out/bench_s3_ref.err; tail -c 400 gpurun_out/bench_s3_ref.json
