#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small JSON summary bench.py and profiles/*.md quote.

    python scripts/parse_ncu.py gpurun_out/r02_ozaki_cg2.ncu-rep profiles/r02_ozaki_cg2_ncu.json [--algorithmic-bytes B] [--launch TEXT]

Runs where ncu is installed (the CPU box is enough: `ncu -i ... --page raw --csv`).  One entry per captured launch with the
metrics the roofline needs; with a single launch the top-level keys `dram_bytes_read` / `dram_bytes_write` / ... are that launch's."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct2",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct_of_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__cluster_size": "cluster_size",
    "smsp__inst_executed.sum": "inst_executed",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed_max",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12,
        "s": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    extra = {}
    a = sys.argv[3:]
    while a:
        if a[0] == "--algorithmic-bytes":
            extra["algorithmic_bytes"] = float(a[1])
        elif a[0] == "--launch":
            extra["launch"] = a[1]
        a = a[2:]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    header, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    launches = []
    for r in body:
        e = {"kernel": r[idx["Kernel Name"]][:120], "id": r[idx["ID"]]}
        for h, key in WANT.items():
            i = idx.get(h)
            if i is None:
                continue
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            e[key] = v * UNIT.get(units[i], 1.0) if key.startswith(("dram_bytes", "duration")) else v
        # tensor sub-pipe counters are reported under TPC.TriageCompute.* names that vary: keep every column that mentions them
        for h, i in idx.items():
            if "pipe_tensor" in h and "pct" in h:
                try:
                    e.setdefault("tensor_pipe", {})[h] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        launches.append(e)
    d = {"report": rep, "launches": launches, **extra}
    if len(launches) == 1:
        d.update({k: v for k, v in launches[0].items() if k != "tensor_pipe"})
        tp = launches[0].get("tensor_pipe", {})
        for h, v in tp.items():
            if "cycles_active_realtime" in h and "imma" not in h:
                d["tensor_pipe_active_pct"] = v
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in d.items() if k != "launches"}, indent=1))


if __name__ == "__main__":
    main()
