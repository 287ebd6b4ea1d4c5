"""profiles/r2_scaling.md from profiles/r2_bench_n{1,2,4,8}.json (bench.py lines of the final round-2 runs)."""
import json
from pathlib import Path

P = Path(__file__).resolve().parents[2]
rows = {}
for n in (1, 2, 4, 8):
    f = P / f"r2_bench_n{n}.json"
    if f.exists():
        rows[n] = json.loads(f.read_text().strip().splitlines()[-1])
out = ["# Round 2 — `bench.py` at 1 / 2 / 4 / 8 B200s of one box (final code of the round; `profiles/r2_bench_n*.json`)\n",
       "Config 2 is read-sharded (weak scaling: 10 M reads per GPU, no data-path collective); config 5 is document-column sharded",
       "with the NCCL score all-gather (strong scaling: the same 2 M reads per step at every N, S = 96 M rows); config 3 is read-sharded",
       "over a fixed 50 M-read set (strong scaling, one 90-element all-reduce).  Every line carries its parity gate (0 mismatches in all of them).\n",
       "| N | cfg2 G lookups/s (device) | cfg2 e2e | efficiency | cfg5 G lookups/s | ms/step | scoring kernel ms | all-gather ms/step | GB/s in per GPU | speed-up | efficiency | bound | cfg3 M reads/s | speed-up |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
b = rows.get(1)
for n, d in rows.items():
    c5, c3 = d.get("cfg5", {}), d.get("cfg3", {})
    s5 = c5.get("lookups_per_s", 0) / b["cfg5"]["lookups_per_s"] if b and "cfg5" in b else float("nan")
    s3 = c3.get("reads_per_s", 0) / b["cfg3"]["reads_per_s"] if b and "cfg3" in b else float("nan")
    out.append(f"| {n} | {d['value'] / 1e9:.2f} | {d['e2e']['value'] / 1e9:.2f} | {d['value'] / (n * b['value']):.3f} | "
               f"{c5.get('lookups_per_s', 0) / 1e9:.3f} | {c5.get('ms_per_step', 0):.1f} | {c5.get('scoring_kernel_ms_per_step', 0):.1f} | "
               f"{c5.get('allgather_ms_per_step', 0):.1f} | {c5.get('allgather_GBps_in_per_gpu') or 0:.0f} | {s5:.2f} | {s5 / n:.2f} | "
               f"{c5.get('strong_scaling_bound_from_128B_fetches', 1)} | {c3.get('reads_per_s', 0) / 1e6:.1f} | {s3:.2f} |")
out += ["", "Parity gates: " + "; ".join(f"N={n}: cfg2 {d['parity']['mismatches']}/{d['parity']['reads']} reads, cfg5 {d.get('cfg5', {}).get('parity', {}).get('mismatches')}, "
                                         f"cfg3 {d.get('cfg3', {}).get('parity', {}).get('mismatches')}" for n, d in rows.items()) + " mismatches.",
        "", "Clocks: " + "; ".join(f"N={n}: {d['clocks'].get('sm_mhz')} MHz, reasons {d['clocks'].get('reasons')}" for n, d in rows.items())]
(P / "r2_scaling.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
