set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" 
cd profiles/experiments/r2 && g++ -O3 -std=c++17 -pthread -I../../../include -o /tmp/parse_harness parse_harness.cpp && cd ../../..
python - <<'PY'
import numpy as np
n=4_000_000; L=150
width = 1 + 8 + 1 + L + 3 + L + 1
rec = np.empty((n, width), dtype=np.uint8)
rec[:, 0] = ord("@")
ids = np.arange(n, dtype=np.int64)
for j in range(8):
    rec[:, 8 - j] = (ids // 10 ** j % 10 + ord("0")).astype(np.uint8)
rec[:, 9] = 10
rng=np.random.default_rng(0)
rec[:, 10:10+L] = np.frombuffer(b"ACGT",np.uint8)[rng.integers(0,4,(n,L))]
rec[:, 10+L:13+L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
rec[:, 13+L:13+2*L] = ord("I")
rec[:, -1] = 10
rec.tofile('/tmp/t.fastq')
PY
for t in 1 2 4 8 16 32; do /tmp/parse_harness /tmp/t.fastq $t | tail -1; done > gpurun_out/r2_parse_scaling.log 2>&1
cat gpurun_out/r2_parse_scaling.log
