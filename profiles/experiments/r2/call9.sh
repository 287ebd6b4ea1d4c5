set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/r2_tests9.log 2>&1
tail -3 gpurun_out/r2_tests9.log
export XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 XS_BENCH_FILE=0 XS_BENCH_CPU_SAMPLE=50000
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench9_$name.json 2> gpurun_out/r2_bench9_$name.err; }
run tma_v1 XS_BK_TMA=1 XS_BK_TMA_VAR=1
run tma_v2 XS_BK_TMA=1 XS_BK_TMA_VAR=2
run tma_v3 XS_BK_TMA=1 XS_BK_TMA_VAR=3
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench9_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench9_")[1], round(d["value"]/1e9,3), round(d["ms_per_step"],1), [round(p["ms_per_step"],1) for p in d["roofline"]["phases"]], d["parity"]["mismatches"])
    except Exception as e:
        print(f, "failed", e)
PY
