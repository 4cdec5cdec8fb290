#!/usr/bin/env python
"""DRAM bytes per launch of the dominant kernels, measured with ncu on the GPU box, for bench.py's `roofline.traffic`.

    python tools/capture_traffic.py <out.json>      (run under gpurun; copy the result to profiles/traffic.json)

For every bench section it runs `bench.py --only <section>` under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum` restricted to the section's dominant kernel, skips the warm-up launches, averages a window of
launches (a multiple of the 4 RK4 stages for the stage kernel), and records the source hash the capture belongs to.
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

GL_REGION = ["kernels.cu", "// K1d: the DIA fused step with TMA-staged tiles", "inline int tma_R(int nv)"]   # the kernel's own text
SECTIONS = [  # section, kernel regex, source files, launches to skip, launches to average
    ("gl", "rk4_dia_tma_kernel", ["kernels.cu"], 40, 16),
    ("lns3d", "rk4_stage_tbcsr_kernel", ["stage_tiled.cu", "stage_tiled.cuh"], 40, 16),
    ("lns3d_strong", "rk4_stage_tbcsr_kernel", ["stage_tiled.cu", "stage_tiled.cuh"], 40, 16),
    ("lns2d", "rk4_stage_tbcsr_kernel", ["stage_tiled.cu", "stage_tiled.cuh"], 40, 16),
]
out = {}
for sec, rx, files, skip, count in SECTIONS:
    log = f"/tmp/traffic_{sec}.csv"
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", f"regex:{rx}", "-s", str(skip), "-c", str(count), "--csv", "--log-file", log,
           sys.executable, os.path.join(ROOT, "bench.py"), "--only", sec, "--no-cpu", "--no-e2e", "--no-parity", "--steps", "1"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    rows_per_gpu = None
    if line:
        d = json.loads(line[-1])
        rows_per_gpu = d.get(sec, {}).get("rows_per_gpu") if sec != "gl" else d.get("config", {}).get("N")
    by = collections.OrderedDict()
    for r in csv.DictReader(l for l in open(log) if l.startswith('"')):
        by.setdefault(r["ID"], {"name": r["Kernel Name"]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    if not by:
        out[f"{sec}:{rx}"] = {"error": (p.stderr or p.stdout)[-400:]}
        continue
    n = len(by)
    rd = sum(v["dram__bytes_read.sum"] for v in by.values()) / n
    wr = sum(v["dram__bytes_write.sum"] for v in by.values()) / n
    ns = sum(v["gpu__time_duration.sum"] for v in by.values()) / n
    out[f"{sec}:{rx}"] = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "launches_averaged": n,
                          "ncu_avg_launch_ms": ns / 1e6, "kernel": next(iter(by.values()))["name"][:120], "rows_per_gpu": rows_per_gpu,
                          "source_files": files, "source_region": GL_REGION if sec == "gl" else None,
                          "source_sha": bench.source_sha(*files, region=GL_REGION if sec == "gl" else None),
                          "capture": "tools/capture_traffic.py: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                                     f"{n} launches after {skip} warm-up launches of `bench.py --only {sec}`"}
    print(sec, out[f"{sec}:{rx}"], flush=True)
json.dump(out, open(sys.argv[1], "w"), indent=1)
