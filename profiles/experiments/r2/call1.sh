set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2_tests1.log 2>&1
cd profiles/microbench
for m in "0 1" "1 1" "2 1" "2 4" "3 1"; do timeout 120 ./gather_tma4 $m; echo "rc=$?"; done > ../../gpurun_out/r2_gather_tma4.log 2>&1
cd ../..
timeout 600 python tests/perf_other_configs.py mlst > gpurun_out/r2_mlst1.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -3 gpurun_out/r2_tests1.log; cat gpurun_out/r2_gather_tma4.log; tail -3 gpurun_out/r2_mlst1.log; tail -c 600 gpurun_out/r2_bench1.json
