set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2_tests5.log 2>&1
timeout 600 python tests/perf_other_configs.py mlst > gpurun_out/r2_mlst5.log 2>&1
XS_NO_PAGES_KERNEL=1 timeout 600 python tests/perf_other_configs.py mlst > gpurun_out/r2_mlst5_nopages.log 2>&1
XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
tail -4 gpurun_out/r2_tests5.log; tail -2 gpurun_out/r2_mlst5.log; tail -2 gpurun_out/r2_mlst5_nopages.log; tail -c 1500 gpurun_out/r2_bench5.json; tail -5 gpurun_out/r2_bench5.err
