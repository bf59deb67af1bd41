"""Aggregate an ncu report's per-instruction metrics by CUDA source line.

    python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep raytracing_c_b200/csrc/libraytracer_gpu.so [top_n]

ncu's CSV export of the source page carries metrics only in its SASS view; the SASS->line map comes
from `nvdisasm -g` on the cubin embedded in the library (built with -lineinfo).  The two listings are
aligned per function by their opcode sequences.  Output: one row per source line with executed
warp-instructions, average active threads, stall samples and the dominant stall reasons.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def ncu_rows(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(out))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = [dict(zip(hdr, r)) for r in rows[hdr_i + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
    return hdr, body


def disasm_functions(lib):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    funcs = collections.OrderedDict()
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, where = None, ("?", 0)
        for line in text.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
            if m:
                cur = m.group(1)
                funcs.setdefault(cur, [])
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                where = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m and cur:
                ins = m.group(2).strip()
                op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
                funcs[cur].append((op, where))
    return funcs


def opcode(src):
    s = src.strip()
    parts = s.split()
    return parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "")


def main():
    report, lib = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    hdr, rows = ncu_rows(report)
    ops = [opcode(r["Source"]) for r in rows]
    funcs = disasm_functions(lib)
    line_of = [None] * len(rows)
    for name, ins in funcs.items():
        if len(ins) < 8:
            continue
        seq = [o for o, _ in ins]
        for start in range(0, len(ops) - len(seq) + 1):
            if ops[start] == seq[0] and ops[start:start + len(seq)] == seq:
                for k, (_, where) in enumerate(ins):
                    line_of[start + k] = (name[:28],) + where
                break
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    total_inst = total_samples = 0
    for r, where in zip(rows, line_of):
        key = where or ("<unmapped>", "?", 0)
        inst = int(float(r["Instructions Executed"] or 0))
        thr = int(float(r["Thread Instructions Executed"] or 0))
        smp = int(float(r["# Samples"] or 0))
        a = agg[key]
        a["inst"] += inst
        a["thr"] += thr
        a["samples"] += smp
        for c in stall_cols:
            a[c] += int(float(r[c] or 0))
        total_inst += inst
        total_samples += smp
    print(f"total warp-instructions {total_inst:,}  stall samples {total_samples:,}  mapped rows "
          f"{sum(1 for w in line_of if w)}/{len(rows)}")
    by_func = collections.Counter()
    by_func_s = collections.Counter()
    for key, a in agg.items():
        by_func[key[0]] += a["inst"]
        by_func_s[key[0]] += a["samples"]
    print("\nper function: inst%  samples%")
    for f, n in by_func.most_common():
        print(f"  {f:30s} {100 * n / total_inst:6.2f}  {100 * by_func_s[f] / max(total_samples, 1):6.2f}")
    print(f"\ntop {top_n} source lines by stall samples:")
    print(f"{'function':28s} {'file:line':22s} {'inst%':>6s} {'smp%':>6s} {'thr/inst':>8s}  top stalls")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top_n]:
        st = sorted(((a[c], c.replace("stall_", "")) for c in stall_cols), reverse=True)[:3]
        st_s = " ".join(f"{n}:{100 * v / max(a['samples'], 1):.0f}%" for v, n in st if v)
        print(f"{key[0]:28s} {key[1] + ':' + str(key[2]):22s} {100 * a['inst'] / total_inst:6.2f} "
              f"{100 * a['samples'] / max(total_samples, 1):6.2f} {a['thr'] / max(a['inst'], 1):8.1f}  {st_s}")


    if len(sys.argv) > 4:
        # listing mode: every source line of the functions matching argv[4], in file order
        pat = sys.argv[4]
        ftot = sum(a["inst"] for k, a in agg.items() if pat in k[0])
        print(f"\nlines of functions matching {pat!r} (inst% of that function, {ftot:,} warp-instructions):")
        for key, a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1], kv[0][2])):
            if pat not in key[0] or a["inst"] * 500 < ftot:
                continue
            print(f"{key[0]:28s} {key[1] + ':' + str(key[2]):28s} {100 * a['inst'] / max(ftot, 1):6.2f} "
                  f"{100 * a['samples'] / max(total_samples, 1):6.2f} {a['thr'] / max(a['inst'], 1):8.1f}")


if __name__ == "__main__":
    main()
