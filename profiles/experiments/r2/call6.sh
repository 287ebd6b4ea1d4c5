set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_fuzz_bucketed.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2_tests6.log 2>&1
tail -4 gpurun_out/r2_tests6.log
export XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 XS_BENCH_FILE=0 XS_BENCH_CPU_SAMPLE=200000
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench6_ahead.json 2> gpurun_out/r2_bench6_ahead.err
XS_BK_HASH_AHEAD=0 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench6_base.json 2> gpurun_out/r2_bench6_base.err
python - <<'PY'
import json
for f in ("ahead","base"):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench6_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,3), round(d["e2e"]["value"]/1e9,3), round(d["ms_per_step"],1), [round(p["ms_per_step"],1) for p in d["roofline"]["phases"]], d["roofline"]["frac"], d["parity"]["mismatches"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2_bench6_ahead.err
