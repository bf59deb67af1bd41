"""Per-kernel DRAM traffic and executed instructions per launch, from an ncu --csv launch list of bench.py.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum \
        --clock-control none -c 1200 --csv --log-file gpurun_out/r02_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-parity
    python tools/ncu_traffic.py gpurun_out/r02_metrics.csv profiles/r02_trace_traffic.json "<the command>"

The JSON is what bench.py's roofline.traffic / roofline.executed read; the CSV it came from is committed next to it.
Kernel families are keyed by the function name up to the template argument list."""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    command = sys.argv[3] if len(sys.argv) > 3 else None
    lines = [l for l in open(src, errors="ignore") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = defaultdict(lambda: defaultdict(float))
    ids = defaultdict(set)
    for r in rows:
        name = re.sub(r"[<(].*", "", r["Kernel Name"].replace("void ", "")).strip()
        full = re.sub(r"\(.*", "", r["Kernel Name"].replace("void ", "")).strip()
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(unit, 1.0)
        for key in {name, full}:
            per[key][r["Metric Name"]] += v * scale
            ids[key].add(r["ID"])
    out = {"source": src.replace("gpurun_out/", "profiles/"), "command": command, "kernels": {}}
    for key, m in sorted(per.items()):
        n = len(ids[key])
        entry = {"launches": n,
                 "ms_per_launch": m.get("gpu__time_duration.sum", 0.0) / n / 1e6,
                 "dram_bytes_per_launch": (m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)) / n,
                 "dram_read_bytes_per_launch": m.get("dram__bytes_read.sum", 0.0) / n,
                 "dram_write_bytes_per_launch": m.get("dram__bytes_write.sum", 0.0) / n,
                 "thread_inst_per_launch": m.get("smsp__thread_inst_executed.sum", 0.0) / n,
                 "warp_inst_per_launch": m.get("smsp__inst_executed.sum", 0.0) / n}
        if entry["warp_inst_per_launch"]:
            entry["lanes_per_inst"] = entry["thread_inst_per_launch"] / entry["warp_inst_per_launch"]
        out["kernels"][key] = entry
        if key == "rt_trace_kernel":
            out["rt_trace_kernel"] = entry
    json.dump(out, open(dst, "w"), indent=1)
    for key, e in out["kernels"].items():
        print(f"{key:34s} n={e['launches']:5d}  {e['ms_per_launch']:8.3f} ms  dram {e['dram_bytes_per_launch'] / 1e6:9.1f} MB  "
              f"lanes/inst {e.get('lanes_per_inst', 0):5.1f}")


if __name__ == "__main__":
    main()
