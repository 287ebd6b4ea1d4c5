set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 XS_BENCH_CPU_SAMPLE=50000
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench10_$name.json 2> gpurun_out/r2_bench10_$name.err; }
XS_BENCH_FILE=1 run m0 XS_BK_TMA=0
export XS_BENCH_FILE=0
run tma1 XS_BK_TMA=1
run tma2 XS_BK_TMA=2
run tma1_f6 XS_BK_TMA=1 XS_BK_FETCH_CTAS=6
(XS_BK_TMA=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz_bucketed.py -q -m gpu -k "cobs and bucket or fuzz" 2>&1 | tail -3) > gpurun_out/r2_tests10_tma1.log 2>&1
cat gpurun_out/r2_tests10_tma1.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench10_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench10_")[1], round(d["value"]/1e9,3), round(d["e2e"]["value"]/1e9,3), round(d["ms_per_step"],1), [round(p["ms_per_step"],1) for p in d["roofline"]["phases"]], round(d["roofline"]["frac"],4), d["parity"]["mismatches"], (d.get("file_e2e") or {}).get("reads_per_sec"), (d.get("file_e2e") or {}).get("parse_s_inside"), (d.get("file_e2e") or {}).get("runs_s"))
    except Exception as e:
        print(f, "failed", e)
PY
