set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; echo rc=$?
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err; echo rc=$?
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
tail -c 1500 gpurun_out/r2_bench_final_n1.json; tail -c 600 gpurun_out/r2_bench_final_ref.json
