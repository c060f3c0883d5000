"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list into a per-kernel table (launches, total us, share, DRAM bytes per launch) — the tables
committed under profiles/.
    python scripts/summarize_launches.py gpurun_out/launches.csv [traffic.json] > profiles/rNN_launches.md
With a second argument the per-family DRAM traffic (bytes per launch, mean) is also written as JSON
(bench.py reads profiles/traffic.json for the `traffic` key of its roofline objects)."""
import csv
import json
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"vitssl::\(anonymous namespace\)::|vitssl::<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def to_us(v, unit):
    return v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)


def to_bytes(v, unit):
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * mult.get(unit, 1.0)


def main(path, traffic_out=None):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ii, ki, mi, vi, ui = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launches = {}  # id -> [name, us, bytes]
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        rec = launches.setdefault(r[ii], [short(r[ki]), 0.0, 0.0])
        if r[mi].startswith("gpu__time_duration"):
            rec[1] += to_us(v, r[ui])
        elif r[mi].startswith("dram__bytes"):
            rec[2] += to_bytes(v, r[ui])
    agg = {}
    for name, us, by in launches.values():
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += by
    tot = sum(v[1] for v in agg.values())
    have_bytes = any(v[2] for v in agg.values())
    print("| kernel | launches | total us | avg us | share |" + (" DRAM MB / launch | DRAM GB/s |" if have_bytes else ""))
    print("|---|---:|---:|---:|---:|" + ("---:|---:|" if have_bytes else ""))
    for k, (n, t, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        extra = f" {by / n / 1e6:.1f} | {by / t / 1e3:.0f} |" if have_bytes else ""
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |" + extra)
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | | 100% |" + (" | |" if have_bytes else ""))
    if traffic_out and have_bytes:
        fam = {"gemm_tcgen05_kernel": [0, 0.0], "ln_kernel": [0, 0.0], "attn_bwd_kernel": [0, 0.0], "attn_fwd_kernel": [0, 0.0]}
        for k, (n, t, by) in agg.items():
            key = ("gemm_tcgen05_kernel" if k.startswith("gemm_tcgen05") else "ln_kernel" if k.startswith("ln_") else
                   "attn_bwd_kernel" if k.startswith("attn_bwd") else "attn_fwd_kernel" if k.startswith("attn_fwd") else None)
            if key:
                fam[key][0] += n; fam[key][1] += by
        out = {k: {"launches": n, "dram_bytes_per_launch": round(by / n) if n else None} for k, (n, by) in fam.items()}
        out["source"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                         "over the kernels of python bench.py --steps 3 --warmup 3 --no-cpu-baseline; mean over each family's launches")
        json.dump(out, open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
