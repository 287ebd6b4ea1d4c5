set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_final_n8.json 2> gpurun_out/r2_bench_final_n8.err; echo rc=$?
tail -c 800 gpurun_out/r2_bench_final_n8.json; tail -3 gpurun_out/r2_bench_final_n8.err
