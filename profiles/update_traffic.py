#!/usr/bin/env python
"""Rebuild profiles/traffic.json from an `ncu --set full` report of the bench's scoring kernels.

    python profiles/update_traffic.py gpurun_out/prof.ncu-rep --reads-per-launch N [--source-note "..."]

Each kernel's entry holds dram__bytes_read.sum / dram__bytes_write.sum (averaged over the captured launches) and the
read count one captured launch covered, so bench.py can scale the figures to a step.  The file also records the
SHA-256 of the kernel sources the capture was taken from; bench.py reports `roofline.traffic` only while that hash
matches the sources it runs (a changed kernel with an unchanged capture must not report stale traffic).
"""
from __future__ import annotations

import argparse
import csv
import hashlib
import io
import json
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KERNEL_SOURCES = ["xspect2_b200/csrc/xs_kernels.cuh", "xspect2_b200/csrc/xs_device.cuh"]


def source_hash() -> str:
    h = hashlib.sha256()
    for rel in KERNEL_SOURCES:
        h.update((ROOT / rel).read_bytes())
    return h.hexdigest()


def short_name(full: str) -> str:
    """'void xs::k_bucket_emit<(int)21, (int)7>(xs::BucketParams)' -> 'k_bucket_emit<21,7>'"""
    m = re.search(r"(k_\w+)(<[^>]*>)?", full)
    name, targs = m.group(1), m.group(2) or ""
    targs = targs.replace("(int)", "").replace("(bool)", "").replace("unsigned char", "u8").replace(
        "unsigned short", "u16").replace("unsigned int", "u32").replace(" ", "")
    return name + targs


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--reads-per-launch", type=float, required=True)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--sig-size", type=int, default=150_000_001)
    ap.add_argument("--source-note", default="")
    ap.add_argument("--first-launch-only", action="store_true", help="use only the first captured launch of every kernel")
    ap.add_argument("--merge", action="store_true", help="keep the entries of kernels this report does not contain")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    units = rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    acc: dict[str, dict] = {}
    for r in rows[2:]:
        name = short_name(r[col["Kernel Name"]])
        e = acc.setdefault(name, {"n": 0, "rd": 0.0, "wr": 0.0, "sect": 0.0, "ms": 0.0})
        if args.first_launch_only and e["n"]:
            continue
        e["n"] += 1
        for key, metric in (("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum")):
            e[key] += float(r[col[metric]].replace(",", "")) * scale.get(units[col[metric]], 1)
        if "lts__t_sectors_srcunit_tex_op_read.sum" in col:
            e["sect"] += float(r[col["lts__t_sectors_srcunit_tex_op_read.sum"]].replace(",", "") or 0)
        e["ms"] += float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    out = {}
    if args.merge and (ROOT / "profiles" / "traffic.json").exists():
        out = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    out["_source_sha256"], out["_sources"] = source_hash(), KERNEL_SOURCES
    for name, e in acc.items():
        out[name] = {"n_reads": args.reads_per_launch, "read_len": args.read_len, "sig_size": args.sig_size,
                     "dram_bytes_read": e["rd"] / e["n"], "dram_bytes_write": e["wr"] / e["n"],
                     "global_load_sectors": e["sect"] / e["n"] or None, "ncu_ms_per_launch": e["ms"] / e["n"],
                     "launches_captured": e["n"], "scales_with_reads": True,
                     "source": args.source_note or f"{args.report} (ncu --set full --clock-control none)"}
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
