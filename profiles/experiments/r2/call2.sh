set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "synthetic" 2>&1 | tail -15) > gpurun_out/r2_tests2a.log 2>&1
(timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -rs 2>&1 | tail -25) > gpurun_out/r2_multi_gpu_test.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_tests2a.log; tail -8 gpurun_out/r2_multi_gpu_test.log; tail -c 3000 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
