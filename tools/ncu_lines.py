"""Executed instructions per CUDA source line (and per named region) of every kernel in an ncu report.

    python tools/ncu_lines.py report.ncu-rep raytracing_c_b200/csrc/libraytracer_gpu.so [top_n]

The source page of the report carries per-SASS-instruction counters; `nvdisasm -g` of the cubin embedded in the
library (built with -lineinfo) gives file:line per SASS offset.  Rows are joined by offset from the kernel's first
instruction.  Regions are the functions of rt_trace.cuh / rt_render.cu, found by scanning the sources for their
first and last lines."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disasm(lib):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    funcs = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, where = None, ("?", 0)
        for line in text.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
            if m:
                cur = m.group(1)
                funcs[cur] = {}
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                where = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m and cur:
                funcs[cur][int(m.group(1), 16)] = where
    return funcs


def regions():
    """(file, first, last, name) for each device function / marked block of the trace path"""
    out = []
    for fn in ("rt_trace.cuh", "rt_render.cu", "rt_shade.cuh", "rt_device.cuh"):
        path = os.path.join(ROOT, "raytracing_c_b200", "csrc", fn)
        lines = open(path).read().splitlines()
        starts = []
        for i, l in enumerate(lines, 1):
            m = re.match(r"^(?:template.*\n)?(?:__device__|__global__|static|inline|int|void|size_t|unsigned).*?\b([a-zA-Z_][a-zA-Z0-9_]*)\s*\(", l)
            if m and not l.startswith(" ") and "(" in l and not l.rstrip().endswith(";"):
                starts.append((i, m.group(1)))
        for k, (i, name) in enumerate(starts):
            end = starts[k + 1][0] - 1 if k + 1 < len(starts) else len(lines)
            out.append((fn, i, end, name))
    return out


def main():
    rep, lib = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    funcs = disasm(lib)
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    regs = regions()
    kernels, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif r and r[0] == "Address":
            hdr = r
        elif r and r[0].startswith("0x") and cur is not None:
            cur["rows"].append(dict(zip(hdr, r)))
    for kidx, k in enumerate(kernels):
        name = k["name"]
        mangled = None
        for f in funcs:
            base = re.sub(r"void |\(.*", "", name)
            key = re.sub(r"<.*", "", base)
            if key in f and (("ILb1E" in f) == ("(bool)1" in name)) and (("ILb0E" in f) == ("(bool)0" in name)):
                mangled = f
        if not mangled or not k["rows"]:
            continue
        base_addr = int(k["rows"][0]["Address"], 16)
        by_line = collections.defaultdict(lambda: [0, 0, 0])
        by_region = collections.defaultdict(lambda: [0, 0, 0])
        total = [0, 0, 0]
        for r in k["rows"]:
            off = int(r["Address"], 16) - base_addr
            where = funcs[mangled].get(off, ("?", 0))
            ie, te, smp = int(r["Instructions Executed"] or 0), int(r["Thread Instructions Executed"] or 0), int(r["# Samples"] or 0)
            for agg in (by_line[where], total):
                agg[0] += ie; agg[1] += te; agg[2] += smp
            reg = next((f"{fn}:{nm}" for fn, a, b, nm in regs if fn == where[0] and a <= where[1] <= b), f"{where[0]}:?")
            by_region[reg][0] += ie; by_region[reg][1] += te; by_region[reg][2] += smp
        print(f"\n=== launch {kidx}: {name}   warp-inst {total[0]:,}  lanes/inst {total[1] / max(total[0], 1):.1f}  samples {total[2]:,}")
        print("  by region (function the instruction's innermost source line belongs to):")
        for reg, (ie, te, smp) in sorted(by_region.items(), key=lambda kv: -kv[1][0]):
            if ie * 200 < total[0]:
                continue
            print(f"    {reg:44s} inst {100 * ie / total[0]:5.1f}%  lanes {te / max(ie, 1):5.1f}  samples {100 * smp / max(total[2], 1):5.1f}%")
        print(f"  top {top} lines:")
        for where, (ie, te, smp) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"    {where[0]}:{where[1]:<5d} inst {100 * ie / total[0]:5.2f}%  lanes {te / max(ie, 1):5.1f}  samples {100 * smp / max(total[2], 1):5.2f}%")


if __name__ == "__main__":
    main()
