cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python tools/config4_multi_gpu.py > gpurun_out/config4_multi_gpu.json 2> gpurun_out/config4_multi_gpu.err; cat gpurun_out/config4_multi_gpu.json; tail -3 gpurun_out/config4_multi_gpu.err
