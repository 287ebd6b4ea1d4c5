set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 XS_BENCH_FILE=0 XS_BENCH_CPU_SAMPLE=20000 XS_BENCH_READS=2000000
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_bucket_ -s 48 -c 3 -o gpurun_out/r2_prof_bucket python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_bucket.log 2>&1
timeout 600 python tests/perf_other_configs.py mlst > gpurun_out/r2_prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_cobs_pages -s 7 -c 1 -o gpurun_out/r2_prof_pages python tests/perf_other_configs.py mlst > gpurun_out/r2_ncu_pages.log 2>&1
timeout 600 python profiles/experiments/r2/cfg5_tile.py 8000000 > gpurun_out/r2_prof_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_cobs_wide|k_sharded_reduce" -s 2 -c 2 -o gpurun_out/r2_prof_cfg5 python profiles/experiments/r2/cfg5_tile.py 8000000 > gpurun_out/r2_ncu_cfg5.log 2>&1
timeout 600 python profiles/experiments/r2/cfg5_tile.py 8000000 0 1280 > gpurun_out/r2_prof_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_cobs_wide" -s 1 -c 1 -o gpurun_out/r2_prof_cfg5_shard8 python profiles/experiments/r2/cfg5_tile.py 8000000 0 1280 > gpurun_out/r2_ncu_cfg5_shard8.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2_ncu_*.log
