set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export XS_BENCH_CFG5=0 XS_BENCH_CFG3=0 XS_BENCH_FILE=0 XS_BENCH_CPU_SAMPLE=50000
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench13_$name.json 2> gpurun_out/r2_bench13_$name.err; }
run m0 XS_BK_TMA=0
run tma1 XS_BK_TMA=1
run tma2 XS_BK_TMA=2
run ahead1 XS_BK_HASH_AHEAD=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench13_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench13_")[1], round(d["value"]/1e9,3), round(d["e2e"]["value"]/1e9,3), round(d["ms_per_step"],1), [round(p["ms_per_step"],1) for p in d["roofline"]["phases"]], round(d["roofline"]["frac"],4), d["parity"]["mismatches"])
    except Exception as e:
        print(f, "failed", e)
PY
export XS_BENCH_CPU_SAMPLE=20000 XS_BENCH_READS=2000000
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bucket_(emit|fetch|reduce)" -s 48 -c 6 -o gpurun_out/r2_prof_bucket python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_bucket.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
ls -la gpurun_out/r2_prof_bucket.ncu-rep
