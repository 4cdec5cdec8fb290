#!/usr/bin/env python
"""Times the compile-time variants of the tiled stage kernel (stage_tiled.cu VAR_*) on the lns3d bench section, then runs
the bit-identity test of the device-generated operator under the fastest ones.  Writes gpurun_out/<dir>/variants.json."""
import json
import os
import subprocess
import sys

out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2v"
variants = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,3,4,7,8,16,17,19,23,24,27,31".split(","))]
os.makedirs(out_dir, exist_ok=True)
res = {}
for v in variants:
    cmd = [sys.executable, "bench.py", "--only", "lns3d", "--no-cpu", "--steps", "3", "--option", f"lns:tiled_variant={v}"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])["lns3d"]
        res[v] = {"avg_launch_ms": d["roofline"]["avg_launch_ms"], "value": d["value"], "frac_model": d["roofline"]["frac_model"]} if "error" not in d else d
    except Exception as e:
        res[v] = {"error": f"{e}: {p.stderr[-300:]}"}
    print(v, res[v], flush=True)
ok = sorted((r["avg_launch_ms"], v) for v, r in res.items() if "avg_launch_ms" in r)
best = [v for _, v in ok[:3] if v != 0]
for v in best:
    env = dict(os.environ, RSVDT_TILED_VARIANT=str(v))
    p = subprocess.run([sys.executable, "-m", "pytest", "tests/test_stencil_gen.py", "-q", "-x", "-m", "gpu", "-k", "device_generated"],
                       capture_output=True, text=True, env=env, timeout=600)
    res[v]["bit_identity_test"] = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-300:]
    print("test under variant", v, res[v]["bit_identity_test"], flush=True)
json.dump(res, open(os.path.join(out_dir, "variants.json"), "w"), indent=1)
