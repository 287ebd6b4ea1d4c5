set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for m in 1 2; do
(XS_BK_HASH_AHEAD=$m timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz_bucketed.py tests/test_gpu_properties.py -x -q -m gpu -k "bucket or fuzz or propert" 2>&1 | tail -5) > gpurun_out/r2_tests7_mode$m.log 2>&1
tail -2 gpurun_out/r2_tests7_mode$m.log
done
(timeout 600 python -m pytest tests/test_models_gpu.py -x -q -m gpu -k "streamed or summary" 2>&1 | tail -5) > gpurun_out/r2_tests7_stream.log 2>&1
tail -2 gpurun_out/r2_tests7_stream.log
export XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 XS_BENCH_CPU_SAMPLE=100000
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench7_$name.json 2> gpurun_out/r2_bench7_$name.err; }
XS_BENCH_FILE=1 run m0 XS_BK_HASH_AHEAD=0
export XS_BENCH_FILE=0
run m1 XS_BK_HASH_AHEAD=1
run m2_f8_h1 XS_BK_HASH_AHEAD=2 XS_BK_FETCH_CTAS=8 XS_BK_HASH_CTAS=1
run m2_f5_h1 XS_BK_HASH_AHEAD=2 XS_BK_FETCH_CTAS=5 XS_BK_HASH_CTAS=1
run m2_f4_h2 XS_BK_HASH_AHEAD=2 XS_BK_FETCH_CTAS=4 XS_BK_HASH_CTAS=2
run m2_f4_h1 XS_BK_HASH_AHEAD=2 XS_BK_FETCH_CTAS=4 XS_BK_HASH_CTAS=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench7_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench7_")[1], round(d["value"]/1e9,3), round(d["e2e"]["value"]/1e9,3), round(d["ms_per_step"],1), [round(p["ms_per_step"],1) for p in d["roofline"]["phases"]], round(d["roofline"]["frac"],4), d["parity"]["mismatches"], (d.get("file_e2e") or {}).get("reads_per_sec"), (d.get("file_e2e") or {}).get("parse_s_inside"))
    except Exception as e:
        print(f, "failed", e)
PY
