#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv (ncu launch list) and a full .ncu-rep capture into profiles/."""
import collections
import csv
import subprocess
import sys


def launch_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = [f"{'kernel':70s} {'n':>4s} {'avg_us':>10s} {'share':>7s}"]
    for k, v in agg.items():
        out.append(f"{k:70s} {len(v):4d} {sum(v) / len(v):10.1f} {sum(v) / tot * 100:6.1f}%")
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum,", "dram__bytes_write.sum,",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__cycles_elapsed.avg ",
        "sm__cycles_elapsed.avg.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        name = vals[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        out.append(f"-- {name[:100]}")
        for i, n in enumerate(h):
            if any(n == w.strip(" ,") for w in WANT):
                out.append(f"{n:75s} {u[i]:12s} {vals[i]}")
    return "\n".join(out)


def traffic_json(reps, out_path):
    """profiles/ncu_traffic.json: DRAM bytes per launch of every captured kernel (bench.py reads roofline.traffic from it)."""
    import json
    import re
    out = {}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        h, u = rows[0], rows[1]
        ri, wi, ni = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("Kernel Name")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for vals in rows[2:]:
            name = re.sub(r"^void ", "", vals[ni]).split("(")[0].split("::")[-1].split("<")[0]
            b = float(vals[ri].replace(",", "")) * scale[u[ri]] + float(vals[wi].replace(",", "")) * scale[u[wi]]
            out.setdefault(name, {"dram_bytes": b, "source": f"ncu --set full capture {rep.split('/')[-1]} (first captured launch), summarised in profiles/"})
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        traffic_json(sys.argv[3:], sys.argv[2])
        sys.exit(0)
    print("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES)")
    print(launch_table(sys.argv[1]))
    for rep in sys.argv[2:]:
        print(f"\n# ncu --set full: {rep}")
        print(raw_metrics(rep))
