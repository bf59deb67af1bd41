"""Build provenance: rebuilds every native target from scratch (make clean; make) and records what came out.

    python tools/build_info.py profiles/r02_build_info.json
The GPU boxes run the .so files shipped with the repo snapshot; this file says which sources, flags and compiler made
them, which sm_100a cubins they hold and what each kernel needs (registers, stack, shared memory)."""
import hashlib
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sh(cmd, cwd=ROOT):
    return subprocess.run(cmd, shell=True, cwd=cwd, capture_output=True, text=True)


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    out_path = sys.argv[1]
    t0 = time.time()
    sh("make -C raytracing_c_b200 clean")
    dry = sh("make -n -C raytracing_c_b200 all").stdout
    built = sh("make -j8 -C raytracing_c_b200 all")
    sh("make -C oracle all")
    seconds = time.time() - t0
    gpu = os.path.join(ROOT, "raytracing_c_b200", "csrc", "libraytracer_gpu.so")
    res = sh(f"cuobjdump -res-usage {gpu}").stdout
    kernels = {}
    for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
        name = sh(f"c++filt {m.group(1)}").stdout.strip()
        kernels[name] = {"registers": int(m.group(2)), "stack": int(m.group(3)), "static_shared": int(m.group(4))}
    elfs = sh(f"cuobjdump -lelf {gpu}").stdout.split()
    sass = sh(f"cuobjdump -sass {gpu}").stdout
    info = {
        "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        "git_head": sh("git rev-parse HEAD").stdout.strip(),
        "git_dirty": bool(sh("git status --porcelain -- raytracing_c_b200 include oracle").stdout.strip()),
        "nvcc": sh("nvcc --version").stdout.strip().splitlines()[-2:],
        "gcc": sh("gcc --version").stdout.splitlines()[0],
        "rebuild_seconds": round(seconds, 1),
        "make_rc": built.returncode,
        "compile_commands": [l for l in dry.splitlines() if "nvcc" in l or l.startswith("gcc")],
        "artifacts": {os.path.relpath(p, ROOT): {"bytes": os.path.getsize(p), "sha256": sha(p)} for p in
                      [gpu, os.path.join(ROOT, "raytracing_c_b200", "host", "librt_host.so"),
                       os.path.join(ROOT, "raytracing_c_b200", "host", "rt_driver"), os.path.join(ROOT, "oracle", "liboracle.so")]},
        "cubins": [e for e in elfs if e.endswith(".cubin")],
        "kernels": kernels,
        "sass_instructions": len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", sass, flags=re.M)),
        "tensor_or_tma_instructions": len(re.findall(r"UTMALDG|UTCMMA|UTCHMMA|LDTM|HMMA|WGMMA", sass)),
        "note": "no tensor-core / TMA / TMEM instructions by design: no stage of this path is a contraction (BASELINE.json north_star)",
    }
    json.dump(info, open(out_path, "w"), indent=1)
    print(json.dumps({k: info[k] for k in ("git_head", "rebuild_seconds", "make_rc", "cubins", "sass_instructions", "tensor_or_tma_instructions")}))


if __name__ == "__main__":
    main()
