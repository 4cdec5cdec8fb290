#!/usr/bin/env python
"""The north-star solve at a size the reference's own code finishes in minutes, on P GPUs, compared with it.

Same case directory for both sides (tools/make_case.py: 3-D block-stencil LNS operator as a PETSc-binary file, F_hat file,
variables.yaml with transient removal and q = 1):
  1. `oracle/_ref/RSVDt_ref -inputs variables.yaml`  -- the reference's sources on the one-rank PETSc shim (the checker)
  2. `torchrun --nproc-per-node P rsvd-delta-t_b200/rsvdt_mpi.py -inputs variables.yaml`  -- operator read from the file
  3. the same with `OperatorDir: synthetic:lns3d:...` (operator generated on the device, what the 2e7-DOF run uses)
and prints the relative errors of the gains and the principal angles of the leading modes (north_star: <= 1e-8 / <= 1e-6).

    python tools/north_star_parity.py --gpus 8 --out /tmp/ns_parity [--shape 16 10 12] [--k 8] [--Nw 8]
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import make_case  # noqa: E402
from conftest import principal_angle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--out", required=True)
ap.add_argument("--shape", type=int, nargs=3, default=[16, 10, 12])
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--Nw", type=int, default=8)
ap.add_argument("--json", default="")
ap.add_argument("--prepare", action="store_true", help="write the case and run the reference executable only (CPU; e.g. into "
                "oracle/_ref/ns_parity_case, which travels to the GPU box), so that the multi-GPU box time goes to the GPU runs")
ap.add_argument("--prepared", action="store_true", help="--out holds a case made by --prepare: skip the reference executable")
a = ap.parse_args()

pkg = entry.load_package()
io = pkg.petsc_io
meta_path = os.path.join(a.out, "ns_parity_meta.json")
if a.prepared:
    meta = json.load(open(meta_path))
    info, t_ref = meta["info"], meta["reference_executable_s"]
    a.shape, a.k, a.Nw = meta["shape"], meta["k"], meta["Nw"]
    for d in os.listdir(os.path.join(a.out, "results")):        # results of earlier GPU runs: the executable numbers its directories
        if d != "ResolventModes_0":
            shutil.rmtree(os.path.join(a.out, "results", d))
else:
    if os.path.exists(a.out):
        shutil.rmtree(a.out)
    info = make_case.write_case(pkg, a.out, "lns3d", shape=a.shape, k=a.k, Nw=a.Nw, w=2.0, dt=0.003, TL=0.5, TR=True, q=1, opt=2,
                                forcing=True)
    from oracle import ref as R  # noqa: E402
    if not R.available():
        raise SystemExit("oracle/_ref is not built: the reference executable is the checker of this script")
    t0 = time.time()
    out = R.run_executable(a.out, threads=os.cpu_count())
    t_ref = time.time() - t0
    if out.returncode != 0:
        raise SystemExit("RSVDt_ref failed:\n" + out.stdout[-2000:] + out.stderr[-2000:])
    json.dump({"info": info, "reference_executable_s": t_ref, "shape": list(a.shape), "k": a.k, "Nw": a.Nw}, open(meta_path, "w"))
    if a.prepare:
        print(f"prepared {a.out}: reference executable took {t_ref:.1f} s")
        sys.exit(0)


def read(res):
    S = io.read_mat_dense(f"{res}/S_hat")
    U = np.stack([io.read_mat_dense(f"{res}/U_hat_Freq{iw}_allK") for iw in range(info["Nw_eff"])], axis=1)
    V = np.stack([io.read_mat_dense(f"{res}/V_hat_Freq{iw}_allK") for iw in range(info["Nw_eff"])], axis=1)
    return S, U, V


Sr, Ur, Vr = read(os.path.join(a.out, "results", "ResolventModes_0"))
report = {"case": info["desc"], "k": a.k, "Nw": a.Nw, "n_gpus": a.gpus, "reference_executable_s": t_ref, "runs": {}}
yaml_path = os.path.join(a.out, "variables.yaml")
base_yaml = open(yaml_path).read()
for i, (name, operator) in enumerate((("operator file (CSR upload)", "A"),
                                      ("operator generated on the device", "synthetic:lns3d:" + "x".join(map(str, a.shape)))), 1):
    open(yaml_path, "w").write(base_yaml.replace("OperatorDir:        A", f"OperatorDir:        {operator}"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + i), os.path.join(ROOT, "rsvd-delta-t_b200", "rsvdt_mpi.py"), "-inputs", "variables.yaml"]
    t0 = time.time()
    p = subprocess.run(cmd, cwd=a.out, capture_output=True, text=True, timeout=1500)
    if p.returncode != 0:
        raise SystemExit(f"rsvdt_mpi.py failed ({name}):\n" + p.stdout[-3000:] + p.stderr[-3000:])
    S, U, V = read(os.path.join(a.out, "results", f"ResolventModes_{i}"))
    gain = float(np.abs(S.real / Sr.real - 1).max())
    lead = float(np.abs(S.real[0] / Sr.real[0] - 1).max())
    angU = max(principal_angle(U[:, iw, :1], Ur[:, iw, :1]) for iw in range(info["Nw_eff"]))
    angV = max(principal_angle(V[:, iw, :1], Vr[:, iw, :1]) for iw in range(info["Nw_eff"]))
    subU = max(principal_angle(U[:, iw, :], Ur[:, iw, :]) for iw in range(info["Nw_eff"]))
    report["runs"][name] = {"gain_rel_err_leading": lead, "gain_rel_err_all_k": gain, "angle_U_leading": angU,
                            "angle_V_leading": angV, "angle_U_subspace_k": subU, "wall_s": time.time() - t0}
open(yaml_path, "w").write(base_yaml)
print(json.dumps(report, indent=1))
if a.json:
    with open(a.json, "w") as f:
        json.dump(report, f, indent=1)
ok = all(r["gain_rel_err_leading"] <= 1e-8 and r["angle_U_leading"] <= 1e-6 and r["angle_V_leading"] <= 1e-6 for r in report["runs"].values())
sys.exit(0 if ok else 1)
