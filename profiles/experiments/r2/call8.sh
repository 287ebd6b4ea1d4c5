set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for i in 1 2 3; do
(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_bloom_bucketed_long_low_complexity_and_sub_batches" 2>&1 | grep -E "AssertionError|passed|failed" | head -3) >> gpurun_out/r2_tests8_bloom_mode0.log 2>&1
(XS_BK_HASH_AHEAD=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_bloom_bucketed_long_low_complexity_and_sub_batches" 2>&1 | grep -E "AssertionError|passed|failed" | head -3) >> gpurun_out/r2_tests8_bloom_mode1.log 2>&1
done
cat gpurun_out/r2_tests8_bloom_mode0.log gpurun_out/r2_tests8_bloom_mode1.log
for m in 1 2; do
(XS_BK_TMA=$m timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz_bucketed.py tests/test_gpu_properties.py -q -m gpu -k "cobs and bucket or fuzz or propert" 2>&1 | tail -5) > gpurun_out/r2_tests8_tma$m.log 2>&1
tail -3 gpurun_out/r2_tests8_tma$m.log
done
export XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 XS_BENCH_CPU_SAMPLE=100000
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench8_$name.json 2> gpurun_out/r2_bench8_$name.err; }
XS_BENCH_FILE=1 run m0 XS_BK_TMA=0
export XS_BENCH_FILE=0
run tma1 XS_BK_TMA=1
run tma2 XS_BK_TMA=2
run tma1_f6 XS_BK_TMA=1 XS_BK_FETCH_CTAS=6
run m0_b XS_BK_TMA=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench8_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench8_")[1], round(d["value"]/1e9,3), round(d["e2e"]["value"]/1e9,3), round(d["ms_per_step"],1), [round(p["ms_per_step"],1) for p in d["roofline"]["phases"]], round(d["roofline"]["frac"],4), d["parity"]["mismatches"], (d.get("file_e2e") or {}).get("reads_per_sec"), (d.get("file_e2e") or {}).get("parse_s_inside"), (d.get("file_e2e") or {}).get("runs_s"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2_bench8_tma1.err
