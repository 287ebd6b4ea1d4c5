set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
df -h /tmp | tail -2; free -g | head -2
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
echo rc=$?
tail -c 5000 gpurun_out/r2_bench3.json; tail -20 gpurun_out/r2_bench3.err
