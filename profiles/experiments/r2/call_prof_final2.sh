set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 XS_BENCH_FILE=0 XS_BENCH_CPU_SAMPLE=20000 XS_BENCH_READS=2000000
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_(emit|fetch|reduce)" -s 48 -c 3 -f -o gpurun_out/r2b_prof_bucket_final python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_ncu_bucket.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 3000 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_ncu_launches.log 2>&1
XS_BENCH_BUCKETED=0 timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_prof_plain6.log 2>&1 &&
XS_BENCH_BUCKETED=0 ncu --set full --clock-control none --import-source on -k regex:"k_cobs_narrow" -s 4 -c 1 -f -o gpurun_out/r2b_prof_narrow_final python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_ncu_narrow.log 2>&1
ls -la gpurun_out/r2b_*final.ncu-rep gpurun_out/r2b_launches.csv
