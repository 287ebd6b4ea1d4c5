set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2_tests4.log 2>&1
timeout 600 python tests/perf_other_configs.py mlst > gpurun_out/r2_mlst4.log 2>&1
XS_NO_PAGES_KERNEL=1 timeout 600 python tests/perf_other_configs.py mlst > gpurun_out/r2_mlst4_nopages.log 2>&1
tail -4 gpurun_out/r2_tests4.log; tail -2 gpurun_out/r2_mlst4.log; tail -2 gpurun_out/r2_mlst4_nopages.log
