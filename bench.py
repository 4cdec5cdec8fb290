#!/usr/bin/env python
"""bench.py -- DOF x test-vector RK4 steps per second of the RSVD-Delta-t time-stepping hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.

  workload    BASELINE.json configs[1]: synthetic complex Ginzburg-Landau operator, N = 1e6, k = 10, Nw = 64,
              w = 0.05, dt = 0.003 (-> Ns = 41,920, ResRatio = 655), no transient removal, 1 GPU.
  step        one pass of the hot path = STEPS_PER_PASS RK4 time steps (forcing synthesis for those steps + 4 RK4
              stages each) of TSActionRK4's loop (TSActionRK4.c:116-122) over all N*k units.
  value       N*k*STEPS_PER_PASS*K / (device time of the K timed passes), inputs resident in HBM.
  e2e         the same metric through the C-ABI call rsvdt_ts_action_rk4 with HOST buffers: one whole action
              (upload F_hat from pinned memory, Nt+Ns RK4 steps, DFT, download Y_hat), wall clock around the call.
  roofline    dominant kernel = the RK4 stage kernel; achieved = algorithmic bytes per launch divided by
              its average launch duration measured with CUDA events inside the timed region.
  cpu_baseline / --impl reference
              the reference's own RK4 step (SourceCode/TSRK4.c + CreateForcingOnFly.c, compiled unmodified into
              oracle/_ref/ against the one-rank PETSc API shim oracle/petsc_shim/, whose primitives are threaded with
              OpenMP in place of MPI ranks) timed on the host cores on a bounded sample: kind "reference".  If
              oracle/_ref/ is absent, the oracle port (oracle/rsvdt_oracle.c) is timed instead: kind "port".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

STEPS_PER_PASS = 256
METRIC = "DOF*test-vector RK4 steps/s"
UNIT = "DOF*vector*steps/s"

WORKLOAD = dict(workload="synthetic complex Ginzburg-Landau (BASELINE.json configs[1])", N=1_000_000, k=10, Nw=64,
                w=0.05, dt_in=0.003, TransientLength=10.0, TransientRemoval=False, nnz_row=4)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self._stop.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(pkg):
    c = WORKLOAD
    N, rowptr, col, val = pkg.synthetic.ginzburg_landau(c["N"])
    dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(c["w"], c["Nw"], c["dt_in"], c["TransientLength"])
    return N, rowptr, col, val, dt, Ns, Nt, RR


class CpuPort:
    """the CPU oracle (reference operation sequence, one test vector at a time) set up once on the bench workload"""

    def __init__(self, pkg, N, rowptr, col, val, dt, Ns, Nt, RR, ncols):
        # torchrun exports OMP_NUM_THREADS=1 to its children; the CPU arm is meant to use every host core
        if "LOCAL_RANK" in os.environ or "OMP_NUM_THREADS" not in os.environ:
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        from oracle import oracle as O
        c = WORKLOAD
        self.O, self.N, self.ncols = O, N, ncols
        self.A = O.Csr(N, N, rowptr, col, val)
        self.v = O.Vars(N=N, k=ncols, Nw=c["Nw"], Nw_eff=c["Nw"], Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=c["w"])
        _, self.idft = O.create_dft_idft_mats(c["Nw"], c["Nw"], Ns, False, False)
        self.F = pkg.synthetic.random_forcing(N, ncols, c["Nw"], seed=14)
        O.time_forced_steps(self.A, self.v, self.idft, self.F, 1, 1)  # warm-up (page faults, thread pool)

    kind = "port"
    what = "oracle port rsvdt_oracle.c (TSActionRK4.c:116-122 one vector at a time)"

    def sample(self, nsteps):
        secs = self.O.time_forced_steps(self.A, self.v, self.idft, self.F, self.ncols, nsteps)
        return self.N * self.ncols * nsteps / secs, secs, self.O.num_threads()


class CpuReference:
    """the reference's own TSRK4.c / CreateForcingOnFly.c (oracle/_ref/librsvdt_ref.so: reference sources on the PETSc API
    shim, OpenMP threads in place of MPI ranks) on the bench workload, one test vector at a time as TSActionRK4.c runs"""

    kind = "reference"
    what = ("reference sources TSRK4.c + CreateForcingOnFly.c compiled unmodified against the PETSc API shim "
            "(oracle/_ref, OpenMP threads in place of MPI ranks)")

    def __init__(self, pkg, N, rowptr, col, val, dt, Ns, Nt, RR, ncols):
        if "LOCAL_RANK" in os.environ or "OMP_NUM_THREADS" not in os.environ:
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        from oracle import oracle as O
        from oracle import ref as R
        c = WORKLOAD
        self.R, self.N, self.ncols = R, N, ncols
        self.A = O.Csr(N, N, rowptr, col, val)
        self.v = O.Vars(N=N, k=ncols, Nw=c["Nw"], Nw_eff=c["Nw"], Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=c["w"])
        self.F = pkg.synthetic.random_forcing(N, 1, c["Nw"], seed=14)
        R.time_forced_steps(self.A, self.v, self.F, 1, 0, 1)  # warm-up (page faults, thread pool)

    def sample(self, nsteps):
        secs = self.R.time_forced_steps(self.A, self.v, self.F, self.ncols, 1, nsteps)
        return self.N * self.ncols * nsteps / secs, secs, self.R.num_threads()


def cpu_arm(pkg, *a):
    """oracle/_ref (the reference's own code) when it was built, else the oracle port"""
    try:
        from oracle import ref as R
        if R.available():
            return CpuReference(pkg, *a)
    except Exception as e:  # a stale or foreign-CPU binary: fall back to the port, say so
        print(f"bench: oracle/_ref not usable ({e}); timing the oracle port", file=sys.stderr)
    return CpuPort(pkg, *a)


def cpu_baseline_block(pkg, N, rowptr, col, val, dt, Ns, Nt, RR, target_s=12.0):
    """the `cpu_baseline` object of the bench line: the reference's own code (oracle/_ref) when it was built -- with the oracle
    port timed next to it for comparison -- else the port alone; each on a sample sized for ~target_s of CPU work"""
    def timed(arm, budget):
        _, probe_s, _ = arm.sample(10)
        n = int(min(max(budget / max(probe_s / 10, 1e-6), 20), 2000))
        v, secs, cores = arm.sample(n)
        return {"value": v, "unit": UNIT, "cores": cores, "kind": arm.kind,
                "sample": f"1 test vector x {n} forced RK4 steps of the same N=1e6 workload ({secs:.1f} s of CPU work); {arm.what}"}

    arm = cpu_arm(pkg, N, rowptr, col, val, dt, Ns, Nt, RR, 1)
    cpu = timed(arm, target_s)
    if arm.kind == "reference":
        del arm
        cpu["port"] = timed(CpuPort(pkg, N, rowptr, col, val, dt, Ns, Nt, RR, 1), target_s / 2)
    return cpu


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host threads: oracle/_ref (the reference's
    own sources on the PETSc API shim) when built, else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package()
    entry.build()
    N, rowptr, col, val, dt, Ns, Nt, RR = build_workload(pkg)
    ncols, nsteps = 1, 8
    port = cpu_arm(pkg, N, rowptr, col, val, dt, Ns, Nt, RR, ncols)
    for _ in range(max(args.warmup, 0)):
        port.sample(1)
    vals, tot_s, cores = [], 0.0, 1
    for _ in range(args.steps):
        v, secs, cores = port.sample(nsteps)
        vals.append(v)
        tot_s += secs
    value = float(N * ncols * nsteps * len(vals) / tot_s) if tot_s > 0 else 0.0
    sample = (f"{ncols} test vector x {nsteps} forced RK4 steps of the N=1e6 workload per bench step, {cores} OpenMP threads; "
              f"{port.what}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {k: WORKLOAD[k] for k in ("workload", "N", "k", "Nw", "w")},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": port.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end (host buffer) measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    ap.add_argument("--option", action="append", default=[], help="name=value tuning option passed to rsvdt_set_option")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = entry.load_package()
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()

    c = WORKLOAD
    k, Nw = c["k"], c["Nw"]
    if world == 1:
        N, rowptr, col, val, dt, Ns, Nt, RR = build_workload(pkg)
        r = pkg.capi.RSVDt(local_rank)
        r.LNSStructure(N, rowptr, col, val)
        r.set_grid(k, Nw, Ns, Nt, RR, dt, c["w"], False)
    else:
        # weak scaling: every GPU owns N rows of a Ginzburg-Landau operator with N*world rows, row-partitioned like
        # PETSc's MPIAIJ (layout 1: halo rows + one peer-store exchange per RK4 step)
        N = c["N"]
        Ntot = N * world
        dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(c["w"], c["Nw"], c["dt_in"], c["TransientLength"])
        gl = lambda a, b: pkg.synthetic.ginzburg_landau(Ntot, row_range=(a, b))[1:]  # noqa: E731
        grid = dict(k=k, Nw=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=c["w"])
        r, plan = pkg.dist.setup_rank(pkg, local_rank, Ntot, gl, 3, 1, (2, 1), grid)
        assert plan.n_own == N
        rowptr = col = val = None
    for opt in args.option:
        name, value = opt.split("=")
        r.set_option(name, int(value))

    # host forcing in pinned memory (torch is plumbing only: pinned allocation + process group)
    F_t = torch.empty((k * Nw, N), dtype=torch.complex128, pin_memory=True)   # row-major (cols, N) == column-major N x cols
    F = F_t.numpy().T                                                         # Fortran-ordered view, N x (k*Nw)
    rng = np.random.default_rng(14 + rank)
    for cidx in range(k * Nw):
        F[:, cidx] = rng.random(N) + 1j * rng.random(N)
    r.set_Y_hat(F)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        r.time_forced_steps(STEPS_PER_PASS)
    path_name = {0: "rk4_stage_csr (4 launches/step)", 1: "rk4_banded_step (CSR, fused 4 stages)",
                 2: "rk4_dia_step (fused 4 stages)", 3: "rk4_dia_tma (fused 4 stages, TMA tiles)"}
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    r.kernel_launches(reset=True)
    dev_ms, stage_ms, stage_n, idft_ms, idft_n = 0.0, 0.0, 0, 0.0, 0
    t0 = time.time()
    for _ in range(args.steps):
        dev_ms += r.time_forced_steps(STEPS_PER_PASS)   # CUDA events on the handle's (launching) stream
        ms, n = r.last_kernel_ms("stage")
        stage_ms += ms
        stage_n += n
        ms, n = r.last_kernel_ms("idft")
        idft_ms += ms
        idft_n += n
    barrier()
    wall = time.time() - t0
    launches = r.kernel_launches()
    if rank == 0:
        sampler.stop()
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    units = float(N) * k * STEPS_PER_PASS * args.steps * world   # every rank advances its own N rows x k test vectors
    value = units / (dev_ms * 1e-3)

    # ---- roofline of the dominant kernel (RK4 stage) --------------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    nnz_row = c["nnz_row"]
    bytes_per_unit_step = 4 * (nnz_row * 20 + 4) / k + 20 * 16      # 353.6 B for GL, k = 10
    launches_per_step = stage_n / max(STEPS_PER_PASS * args.steps, 1)
    bytes_per_launch = bytes_per_unit_step * N * k / max(launches_per_step, 1e-9)
    avg_launch_ms = stage_ms / max(stage_n, 1)
    achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9 if stage_n else None
    ncu_traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tp):
        try:
            ncu_traffic = json.load(open(tp)).get("rk4_step_kernel_bytes_per_launch")
        except Exception:
            ncu_traffic = None
    roofline = {"bound": "hbm", "kernel": "rk4 step kernel(s), %.0f launch(es) per RK4 step" % launches_per_step,
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic,
                "note": "achieved uses the ALGORITHMIC bytes of SURVEY.md 8(d) (stage-by-stage model, 353.6 B per unit and "
                        "step); the fused-step kernel moves fewer bytes than the model (temporal fusion of the 4 stages), "
                        "so frac > 1 is a traffic reduction, not a bandwidth above peak; `traffic` is the ncu-measured DRAM "
                        "bytes per launch",
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_launch_ms,
                "launches_per_rk4_step": launches_per_step, "peak_source": peak_src,
                "share_of_step": stage_ms / dev_ms if dev_ms else None,
                "idft_share_of_step": idft_ms / dev_ms if dev_ms else None}

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        Y_t = torch.empty((k * Nw, N), dtype=torch.complex128, pin_memory=True)
        Y = Y_t.numpy().T
        barrier()
        t1 = time.time()
        r.TSActionRK4(F, direct=True, out=Y)
        barrier()
        t_e2e = time.time() - t1
        if world > 1:
            t = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        rend = Nt + Ns
        e2e = {"value": float(N) * k * rend * world / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(F.nbytes),
               "d2h_bytes_per_step": int(Y.nbytes), "seconds": t_e2e, "rk4_steps": rend,
               "call": "rsvdt_ts_action_rk4 (one whole action, host buffers, pinned)",
               "checksum": float(np.abs(Y[::9973, ::37]).sum())}

    # ---- CPU baseline on the host cores (rank 0, N = 1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_block(pkg, N, rowptr, col, val, dt, Ns, Nt, RR)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
                "config": {**{kk: c[kk] for kk in ("workload", "N", "k", "Nw", "w")}, "Ns": Ns, "Nt": Nt,
                           "ResRatio": RR, "dt": dt, "rk4_steps_per_bench_step": STEPS_PER_PASS,
                           "l2": "inputs larger than L2 (each block vector 160 MB, F_hat 10.2 GB)",
                           "N_total": N * world,
                           "parallelism": ("rows partitioned over %d GPUs (PETSc MPIAIJ split), halo rows by peer stores "
                                           "over NVLink, 1 exchange per RK4 step" % world) if world > 1 else "single GPU"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": sampler.summary(), "wall_s_timed_region": wall}
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
