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
       "Config 2 is read-sharded (weak scaling: 10 M reads per GPU, no data-path collective).  Config 5 (strong scaling: the same 2 M reads",
       "per step at every N, S = 96 M rows, 123 GB) is document-column sharded with the NCCL score all-gather, in two layouts: `cfg5` =",
       "C column groups x N / C read groups with the fewest column groups that keep a shard within half of a GPU's HBM (C = 2), and",
       "`cfg5_columns_only` = every GPU its own column range (C = N; identical to `cfg5` at N = 2).  Config 3 is read-sharded over a fixed",
       "50 M-read set (strong scaling, one 90-element all-reduce), the records streaming through each GPU in blocks.  Every line carries",
       "its parity gates (0 mismatches in all of them).\n",
       "| N | cfg2 G lookups/s (device) | cfg2 e2e | efficiency | cfg5 layout | cfg5 G lookups/s | ms/step | scoring kernel ms | all-gather ms/step | speed-up | efficiency | bound | columns only: G lookups/s | efficiency | bound | cfg3 M reads/s | speed-up |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
b = rows.get(1)
for n, d in rows.items():
    c5, c3 = d.get("cfg5", {}), d.get("cfg3", {})
    cc = d.get("cfg5_columns_only") or c5
    base5 = b["cfg5"]["lookups_per_s"]
    s5 = c5.get("lookups_per_s", 0) / base5
    sc = cc.get("lookups_per_s", 0) / base5
    s3 = c3.get("reads_per_s", 0) / b["cfg3"]["reads_per_s"]
    out.append(f"| {n} | {d['value'] / 1e9:.2f} | {d['e2e']['value'] / 1e9:.2f} | {d['value'] / (n * b['value']):.3f} | "
               f"{c5.get('column_groups', 1)} x {c5.get('read_groups', 1)} | {c5.get('lookups_per_s', 0) / 1e9:.3f} | {c5.get('ms_per_step', 0):.1f} | "
               f"{c5.get('scoring_kernel_ms_per_step', 0):.1f} | {c5.get('allgather_ms_per_step', 0):.1f} | {s5:.2f} | {s5 / n:.2f} | "
               f"{c5.get('strong_scaling_bound_from_128B_fetches', 1)} | {cc.get('lookups_per_s', 0) / 1e9:.3f} | {sc / n:.2f} | "
               f"{cc.get('strong_scaling_bound_from_128B_fetches', 1)} | {c3.get('reads_per_s', 0) / 1e6:.1f} | {s3:.2f} |")
out += ["", "Parity gates: " + "; ".join(f"N={n}: cfg2 {d['parity']['mismatches']}/{d['parity']['reads']} reads, cfg5 {d.get('cfg5', {}).get('parity', {}).get('mismatches')}, "
                                         f"columns only {(d.get('cfg5_columns_only') or {}).get('parity', {}).get('mismatches', '-')}, "
                                         f"cfg3 {d.get('cfg3', {}).get('parity', {}).get('mismatches')}" for n, d in rows.items()) + " mismatches.",
        "", "Clocks: " + "; ".join(f"N={n}: {d['clocks'].get('sm_mhz')} MHz, reasons {d['clocks'].get('reasons')}" for n, d in rows.items()),
        "", "(The N = 8 line was taken before `genus_then_species` chose about eight blocks per pass for small per-rank inputs; at N = 8 a rank's",
        "6.25 M reads then went through 4 blocks of 2 M.  Earlier in the round, with one monolithic pass per rank and pure column sharding:",
        "cfg5 0.600 / 1.280 / 1.773 / 2.547 G lookups/s and cfg3 23.1 / 45.2 / 92.6 / 172.4 M reads/s at N = 1 / 2 / 4 / 8.)"]
(P / "r2_scaling.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
