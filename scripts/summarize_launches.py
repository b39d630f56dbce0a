"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, for the
DMMA GEMM, a breakdown by grid size class.  Usage: summarize_launches.py launches.csv [out.md]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
gemm = collections.OrderedDict()
total = 0.0
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1e3 if unit == "ns" else v * (1e3 if unit == "ms" else 1e6 if unit == "s" else 1.0)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"void |gpb::\(anonymous namespace\)::|gpb::|<unnamed>::", "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
    n += 1
    if "gemm_f64_kernel" in name:
        g = [int(x) for x in row["Grid Size"].strip("()").split(",")]
        ctas = g[0] * g[1] * g[2]
        cls = "<=16 CTAs" if ctas <= 16 else "<=296 CTAs (one wave)" if ctas <= 296 else "<=4096 CTAs" if ctas <= 4096 else ">4096 CTAs"
        b = gemm.setdefault(cls, [0, 0.0])
        b[0] += 1
        b[1] += us
out = [f"launches: {n}, total kernel time {total/1e3:.2f} ms (ncu: cold-cache, serialised; compare SHARES)", "",
       "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k[:70]}` | {c} | {t/1e3:.2f} | {100*t/total:.1f} % | {t/c:.1f} |")
out += ["", "DMMA GEMM launches by grid size:", "", "| class | launches | total ms | share of all |", "|---|---:|---:|---:|"]
for k, (c, t) in gemm.items():
    out.append(f"| {k} | {c} | {t/1e3:.2f} | {100*t/total:.1f} % |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
