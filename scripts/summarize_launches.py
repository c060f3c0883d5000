"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table
(launches, total us, share) — the tables committed under profiles/.
    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"vitssl::\(anonymous namespace\)::|vitssl::<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, order = {}, []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1e3 if r[ui] in ("ns", "nsecond") else v * (1e3 if r[ui] in ("ms", "msecond") else 1.0)
        k = short(r[ki])
        if k not in agg:
            agg[k] = [0, 0.0]
            order.append(k)
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | | 100% |")


if __name__ == "__main__":
    main(sys.argv[1])
