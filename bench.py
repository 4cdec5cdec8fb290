#!/usr/bin/env python
"""bench.py -- DOF x test-vector RK4 steps per second of the RSVD-Delta-t time-stepping hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.

  value       BASELINE.json configs[1]: synthetic complex Ginzburg-Landau operator, N = 1e6 rows per GPU, k = 10, Nw = 64,
              w = 0.05, dt = 0.003 (-> Ns = 41,920, ResRatio = 655), no transient removal.  A step = STEPS_PER_PASS RK4 time
              steps (forcing synthesis for those steps + 4 RK4 stages each) of TSActionRK4's loop (TSActionRK4.c:116-122);
              value = units advanced by all ranks / device time (CUDA events on the launching stream, max over ranks).
  e2e         the same metric through the C-ABI call rsvdt_ts_action_rk4 with HOST buffers: one whole action
              (upload F_hat from pinned memory, Nt+Ns RK4 steps, DFT, download Y_hat), wall clock around the call.
  roofline    dominant kernel of that workload (fused RK4 step kernel): frac = measured DRAM bytes per launch (ncu capture
              committed under profiles/, keyed to the kernel source) / live launch duration / measured HBM peak;
              model_traffic_ratio = SURVEY.md 8(d) stage-by-stage byte model / the same time / peak (> 1: temporal fusion).
  lns3d       BASELINE.json configs[3] per-GPU load: 3-D block-stencil LNS operator (5 variables x 27 points, complex),
              k = 8, Nw = 32, 2.4-2.5e6 rows per GPU (weak: the grid grows with N; N = 8 is the 2e7-DOF configuration);
              tiled block-CSR stage kernel, halo exchange inside the kernel; units/s, stage launch time, roofline fractions.
  lns3d_strong  the same operator at a FIXED size (100 x 100 x 100 x 5 = 5e6 DOF) on N = 1, 2, 4, 8 GPUs: strong scaling.
  lns2d       BASELINE.json configs[2]: 2-D block stencil (5 variables x 9 points), 5e6 DOF, k = 10, fixed size.
  parity      before any timing: small row-partitioned direct + adjoint actions in both local layouts and one weighted
              action, compared with the CPU oracle on rank 0 (relative errors); at N = 1 also the state of the bench-size
              operator after a few steps against the reference's own TSRK4.c (parity_bench_size).
  cpu_baseline / --impl reference
              the reference's own RK4 step (SourceCode/TSRK4.c + CreateForcingOnFly.c, compiled unmodified into
              oracle/_ref/ against the one-rank PETSc API shim oracle/petsc_shim/, whose primitives are threaded with
              OpenMP in place of MPI ranks) timed on the host cores on a bounded sample: kind "reference".  If
              oracle/_ref/ is absent, the oracle port (oracle/rsvdt_oracle.c) is timed instead: kind "port".
"""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEPS_PER_PASS = 256
LNS_STEPS_PER_PASS = 16
METRIC = "DOF*test-vector RK4 steps/s"
UNIT = "DOF*vector*steps/s"

WORKLOAD = dict(workload="synthetic complex Ginzburg-Landau (BASELINE.json configs[1])", N=1_000_000, k=10, Nw=64,
                w=0.05, dt_in=0.003, TransientLength=10.0, TransientRemoval=False, nnz_row=4)


def config_dict(grid, world):
    """the `config` object of the JSON line -- the SAME function feeds both arms (reference and b200)"""
    c = WORKLOAD
    dt, Ns, Nt, RR = grid
    return {**{kk: c[kk] for kk in ("workload", "N", "k", "Nw", "w")}, "Ns": Ns, "Nt": Nt, "ResRatio": RR, "dt": dt,
            "rk4_steps_per_bench_step": STEPS_PER_PASS,
            "l2": "inputs larger than L2 (each block vector 160 MB, F_hat 10.2 GB)",
            "N_total": c["N"] * world,
            "parallelism": ("rows partitioned over %d GPUs (PETSc MPIAIJ split), halo rows by peer stores over NVLink, "
                            "1 exchange per RK4 step" % world) if world > 1 else "single GPU"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def source_sha(*names, region=None):
    """hash of kernel source files, or -- region = (file, start marker, end marker) -- of the text of one kernel inside a file
    that holds many (an edit elsewhere in kernels.cu does not invalidate the capture of the fused step kernel)"""
    h = hashlib.sha256()
    if region:
        with open(os.path.join(ROOT, "rsvd-delta-t_b200", "csrc", region[0]), "r") as f:
            text = f.read()
        a = text.index(region[1])
        h.update(text[a:text.index(region[2], a)].encode())
        return h.hexdigest()[:16]
    for n in names:
        with open(os.path.join(ROOT, "rsvd-delta-t_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic(key):
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture (profiles/traffic.json); None when the
    capture was taken from a different version of the kernel source (the entry names the sha it belongs to)"""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        ent = json.load(open(tp))[key]
        if ent.get("source_sha") and ent["source_sha"] != source_sha(*ent["source_files"], region=ent.get("source_region")):
            return None, "stale (kernel source changed since the ncu capture)"
        return ent, None
    except Exception:
        return None, "no capture"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,clocks.mem,clocks.max.mem")

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
        sm, mx, reasons, mem, mem_mx, pw = [], [], set(), [], [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                try:
                    pw.append(float(r[3]))
                    mem.append(float(r[9]))
                    mem_mx.append(float(r[10]))
                except (ValueError, IndexError):
                    pass
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        if mem:
            out.update({"mem_mhz": float(np.median(mem)), "mem_max_mhz": float(max(mem_mx)), "power_w": float(np.median(pw))})
        return out


# ======================================================================================================================
# CPU arms (rank 0): the reference's own code (oracle/_ref) or the oracle port.  Nothing here touches librsvdt_b200.so.
# ======================================================================================================================
def load_synthetic():
    """rsvd-delta-t_b200/synthetic.py alone (numpy only): the reference arm must not import the CUDA binding"""
    spec = importlib.util.spec_from_file_location("rsvdt_synthetic", os.path.join(ROOT, "rsvd-delta-t_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def host_threads():
    # torchrun exports OMP_NUM_THREADS=1 to its children; the CPU arm is meant to use every host core
    if "LOCAL_RANK" in os.environ or "OMP_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)


class CpuArm:
    """one test vector at a time, as TSActionRK4.c:89 runs them; kind = "reference" (oracle/_ref) or "port" (oracle)"""

    def __init__(self, S, A_csr, grid, Nw, w, N, prefer_reference=True):
        host_threads()
        from oracle import oracle as O
        O.build()
        self.O, self.N, self.Nw = O, N, Nw
        dt, Ns, Nt, RR = grid
        self.A = O.Csr(N, N, *A_csr)
        self.v = O.Vars(N=N, k=1, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w)
        self.F = S.random_forcing(N, 1, Nw, seed=14)
        self.R = None
        if prefer_reference:
            try:
                from oracle import ref as R
                R.build()          # no-op without /root/reference (the GPU box uses the prebuilt oracle/_ref files)
                if R.available():
                    self.R = R
            except Exception as e:  # a stale or foreign-CPU binary: fall back to the port, say so
                print(f"bench: oracle/_ref not usable ({e}); timing the oracle port", file=sys.stderr)
        if self.R is not None:
            self.kind = "reference"
            self.what = ("reference sources TSRK4.c + CreateForcingOnFly.c compiled unmodified against the PETSc API shim "
                         "(oracle/_ref, OpenMP threads in place of MPI ranks)")
            self.R.time_forced_steps(self.A, self.v, self.F, 1, 0, 1)
        else:
            self.kind = "port"
            self.what = "oracle port rsvdt_oracle.c (TSActionRK4.c:116-122 one vector at a time)"
            _, self.idft = O.create_dft_idft_mats(Nw, Nw, Ns, False, False)
            O.time_forced_steps(self.A, self.v, self.idft, self.F, 1, 1)

    def cores(self):
        return self.R.num_threads() if self.R is not None else self.O.num_threads()

    def sample(self, ncols, nsteps):
        """-> (units/s, seconds)"""
        if self.R is not None:
            secs = self.R.time_forced_steps(self.A, self.v, self.F, ncols, 0, nsteps)
        else:
            secs = 0.0
            for _ in range(ncols):
                secs += self.O.time_forced_steps(self.A, self.v, self.idft, self.F, 1, nsteps)
        return self.N * ncols * nsteps / secs, secs

    def state_after(self, nsteps):
        """the vector y of the reference's loop after `nsteps` forced steps from y = 0 on forcing self.F"""
        if self.R is not None:
            return self.R.time_forced_steps(self.A, self.v, self.F, 1, 0, nsteps, want_state=True)[1]
        return self.O.forced_state(self.A, self.v, self.idft, self.F, nsteps, ncols=1)[:, 0]


def cpu_baseline_block(arm, target_s=10.0, label="the same N=1e6 workload"):
    _, probe_s = arm.sample(1, 8)
    n = int(min(max(target_s / max(probe_s / 8, 1e-6), 16), 2000))
    v, secs = arm.sample(1, n)
    return {"value": v, "unit": UNIT, "cores": arm.cores(), "kind": arm.kind,
            "sample": f"1 test vector x {n} forced RK4 steps of {label} ({secs:.1f} s of CPU work); {arm.what}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host threads: oracle/_ref (the reference's
    own sources on the PETSc API shim) when built, else the oracle port.  Does not load the CUDA
    library: grid from the oracle, operator from synthetic.py."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    host_threads()
    from oracle import oracle as O
    O.build()
    S = load_synthetic()
    c = WORKLOAD
    grid = O.setup_time_freq_grid(c["w"], c["Nw"], c["dt_in"], c["TransientLength"])
    N, rowptr, col, val = S.ginzburg_landau(c["N"])
    arm = CpuArm(S, (rowptr, col, val), grid, c["Nw"], c["w"], N)
    ncols, nsteps = 2, 32
    for _ in range(max(args.warmup, 0)):
        arm.sample(1, 2)
    tot_s = 0.0
    for _ in range(args.steps):
        _, secs = arm.sample(ncols, nsteps)
        tot_s += secs
    value = float(N * ncols * nsteps * args.steps / tot_s) if tot_s > 0 else 0.0
    sample = (f"{ncols} test vectors x {nsteps} forced RK4 steps of the N=1e6 workload per bench step, {arm.cores()} OpenMP threads; "
              f"{arm.what}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": config_dict(grid, max(args.gpus, 1)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores(), "kind": arm.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ======================================================================================================================
# GPU arm
# ======================================================================================================================
class Ctx:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        import __graft_entry__ as entry
        self.entry = entry
        self.pkg = entry.load_package()
        if self.rank == 0:
            entry.build()
        self.barrier()
        self.peak, self.peak_src = measured_peak_gbs()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, *vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def gather_rows(self, local):
        """row blocks of every rank -> the whole matrix on every rank (small parity cases only)"""
        if self.world == 1:
            return local
        parts = [None] * self.world
        self.dist.all_gather_object(parts, np.ascontiguousarray(local))
        return np.concatenate(parts, axis=0)


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(float(np.linalg.norm(np.asarray(b))), 1e-300))


# ---------------------------------------------------------------------------------------------------------------------
# parity (small, before any timing): partitioned actions in both layouts + a weighted action against the oracle
# ---------------------------------------------------------------------------------------------------------------------
def parity_section(cx: Ctx):
    pkg, world, rank = cx.pkg, cx.world, cx.rank
    host_threads()
    from oracle import oracle as O
    O.build()
    S, D = pkg.synthetic, pkg.dist
    out = {}
    ref_cache = {}   # the csr-upload and device-generated cases share operator, forcing and hence the oracle's answer

    def oracle_action(A, v, F, direct):
        vv = O.Vars(**{**v.__dict__, "DirAdj": direct})
        dft, idft = O.create_dft_idft_mats(v.Nw, v.Nw_eff, v.Ns, v.RealOperator, False)
        return O.ts_action_rk4(A if direct else A.hermitian_transpose(), vv, dft, idft, F)

    def run_case(name, N, csr, k, Nw, w, TL, make_handle, weights=None):
        rp, ci, va = csr
        dt, Ns, Nt, RR = O.setup_time_freq_grid(w, Nw, 0.003, TL)
        v = O.Vars(N=N, k=k, Nw=Nw, Nw_eff=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w)
        grid = dict(k=k, Nw=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=w)
        rng = np.random.default_rng(21)
        F = np.asfortranarray(rng.random((N, k * Nw)) + 1j * rng.random((N, k * Nw)))
        r, (r0, r1) = make_handle(grid)
        errs = []
        try:
            if weights is not None:
                weights(r, r0, r1)
            for direct in (True, False):
                r.set_Y_hat(F[r0:r1])
                (r.DirectActionRK4 if direct else r.AdjointActionRK4)()
                got = cx.gather_rows(r.get_Y_hat())
                if rank == 0:
                    A = O.Csr(N, N, rp, ci, va)
                    if weights is None:
                        key = (N, k, Nw, len(va), direct)
                        if key not in ref_cache:
                            ref_cache[key] = oracle_action(A, v, F, direct)
                        ref = ref_cache[key]
                    else:
                        d = WEIGHT_DIAG(N)
                        Fin = F * (d[:, None] if direct else (1.0 / d)[:, None])
                        ref = oracle_action(A, v, np.asfortranarray(Fin), direct) * ((1.0 / d)[:, None] if direct else d[:, None])
                    errs.append(rel_err(got, ref))
                # rank 0 has just spent seconds on the host inside the oracle: the ranks start the next action together, or their
                # stage kernels would sit in the halo wait for that long (and give up: the wait is bounded)
                cx.barrier()
        finally:
            r.close()
        if rank == 0:
            out[name] = max(errs)

    # ---- layout 1 (narrow band, fused step kernel on halo-extended rows): Ginzburg-Landau -------------------------------
    Ngl = 1200 * max(world, 1)
    _, rp, ci, va = S.ginzburg_landau(Ngl)

    def h_gl(grid):
        if world == 1:
            r = pkg.capi.RSVDt(cx.local_rank)
            r.LNSStructure(Ngl, rp, ci, va)
            r.set_grid(grid["k"], grid["Nw"], grid["Ns"], grid["Nt"], grid["ResRatio"], grid["dt"], grid["w"], False)
            r.set_option("use_small", 0)
            return r, (0, Ngl)
        gl = lambda a, b: S.ginzburg_landau(Ngl, row_range=(a, b))[1:]  # noqa: E731
        r, plan = D.setup_rank(pkg, cx.local_rank, Ngl, gl, 3, 1, (2, 1), grid)
        return r, (plan.r0, plan.r1)
    run_case("layout1_gl_band", Ngl, (rp, ci, va), 4, 4, 2.0, 0.25, h_gl)

    # ---- layout 0 (general sparsity, one exchange per stage, fused into the tiled stage kernel): 3-D block stencil ----------
    shape = ((4 if world <= 2 else 2) * max(world, 1), 6, 5)   # (the oracle on rank 0 is serial: keep its share of the run small)
    Nl, rpl, cil, val_ = S.lns_stencil(shape, nvar=5)
    bw = (6 * 5 + 5 + 1) * 5 + 5

    def h_lns_host(grid):      # operator through the CSR upload path (what MatLoad feeds)
        if world == 1:
            r = pkg.capi.RSVDt(cx.local_rank)
            r.LNSStructure(Nl, rpl, cil, val_)
            r.set_grid(grid["k"], grid["Nw"], grid["Ns"], grid["Nt"], grid["ResRatio"], grid["dt"], grid["w"], False)
            r.set_option("use_small", 0)
            return r, (0, Nl)
        gen = lambda a, b: D.slice_rows(rpl, cil, val_, 0, a, b)  # noqa: E731
        r, plan = D.setup_rank(pkg, cx.local_rank, Nl, gen, bw, 0, None, grid)
        return r, (plan.r0, plan.r1)

    def h_lns_dev(grid):       # the same operator generated on the device (what the lns sections below time)
        if world == 1:
            r = pkg.capi.RSVDt(cx.local_rank)
            r.set_operator_stencil(shape, 5)
            r.set_grid(grid["k"], grid["Nw"], grid["Ns"], grid["Nt"], grid["ResRatio"], grid["dt"], grid["w"], False)
            return r, (0, Nl)
        r, plan = D.setup_rank_stencil(pkg, cx.local_rank, shape, 5, grid)
        return r, (plan.r0, plan.r1)
    run_case("layout0_lns3d_csr_upload", Nl, (rpl, cil, val_), 4, 4, 2.0, 0.25, h_lns_host)
    run_case("layout0_lns3d_device_generated", Nl, (rpl, cil, val_), 4, 4, 2.0, 0.25, h_lns_dev)

    # ---- weighted action: diagonal W_f^{-1/2} before and W_q^{1/2} after (ApplyWeightMats.c:17-41, adjoint :43-78) ----------
    def weights(r, r0, r1):
        d = WEIGHT_DIAG(Nl)
        idx = np.arange(Nl, dtype=np.int64)
        full = lambda vals: (np.arange(Nl + 1, dtype=np.int64), idx, vals.astype(np.complex128))  # noqa: E731
        for which, vals in (("W_f_sqrt_inv", d), ("W_q_sqrt", 1.0 / d), ("W_q_sqrt_inv", d)):
            rp_, ci_, va_ = full(vals)
            if world == 1:
                r.ReadWeightMats(which, Nl, Nl, rp_, ci_, va_)
            else:
                M_rows, MH_rows = D.local_weight(Nl, Nl, rp_, ci_, va_, world, rank)
                r.ReadWeightMats_local(which, Nl, Nl, M_rows, MH_rows)
    run_case("weighted_lns3d", Nl, (rpl, cil, val_), 4, 4, 2.0, 0.25, h_lns_host, weights)
    return out if rank == 0 else None


def WEIGHT_DIAG(N):
    return 1.0 + 0.5 * np.cos(np.arange(N) * 0.37)


# ---------------------------------------------------------------------------------------------------------------------
# LNS block-stencil sections (device-generated operator, tiled block-CSR stage kernel)
# ---------------------------------------------------------------------------------------------------------------------
def lns_section(cx: Ctx, name, shape, k, Nw, scaling, note, cpu_sample_shape=None, with_adjoint=False):
    pkg, world, rank, args = cx.pkg, cx.world, cx.rank, cx.args
    nvar = 5
    shape = tuple(int(s) for s in shape)
    N_total = int(np.prod(shape)) * nvar
    nnz_row = (3 ** len(shape)) * nvar
    dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(0.05, Nw, 0.003, 10.0)
    grid = dict(k=k, Nw=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=0.05)
    t0 = time.time()
    if world == 1:
        r = pkg.capi.RSVDt(cx.local_rank)
        r.set_operator_stencil(shape, nvar, with_adjoint=with_adjoint)
        r.set_grid(k, Nw, Ns, Nt, RR, dt, 0.05, False)
        n_own, row0 = N_total, 0
    else:
        r, plan = pkg.dist.setup_rank_stencil(pkg, cx.local_rank, shape, nvar, grid, with_adjoint=with_adjoint)
        n_own, row0 = plan.n_own, plan.r0
    for opt in args.option:
        nm, value = opt.split("=")
        if nm.startswith("lns:"):
            r.set_option(nm[4:], int(value))
    r.CreateRandomMat(n_own, row0, seed=14)
    info = r.operator_info(True, k)
    setup_s = time.time() - t0
    S = LNS_STEPS_PER_PASS
    for _ in range(2):
        r.time_forced_steps(S)
    sampler = ClockSampler(cx.local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    cx.barrier()
    dev_ms = stage_ms = idft_ms = 0.0
    stage_n = 0
    l0 = r.kernel_launches(reset=True)
    for _ in range(args.steps):
        dev_ms += r.time_forced_steps(S)
        ms, n = r.last_kernel_ms("stage")
        stage_ms += ms
        stage_n += n
        idft_ms += r.last_kernel_ms("idft")[0]
    cx.barrier()
    if rank == 0:
        sampler.stop()
    launches = r.kernel_launches()
    del l0
    dev_ms_max, stage_ms_max = cx.max_over_ranks(dev_ms, stage_ms)
    (launch_sum,) = cx.sum_over_ranks(launches)
    adj = None
    if with_adjoint:
        # the adjoint pass (AdjointActionRK4.c:24-28 -> TSActionRK4 on A^H): the same stage kernel on the stored A^H
        r.time_forced_steps(S, direct=False)
        cx.barrier()
        a_ms = a_stage = 0.0
        a_n = 0
        for _ in range(args.steps):
            a_ms += r.time_forced_steps(S, direct=False)
            ms, n = r.last_kernel_ms("stage")
            a_stage += ms
            a_n += n
        cx.barrier()
        a_ms_max, a_stage_max = cx.max_over_ranks(a_ms, a_stage)
        adj = {"value": float(N_total) * k * S * args.steps / (a_ms_max * 1e-3), "unit": UNIT,
               "avg_launch_ms": a_stage_max / max(a_n, 1), "what": "forced RK4 steps on A^H (the adjoint pass), same operator and size"}
    r.close()
    units = float(N_total) * k * S * args.steps
    value = units / (dev_ms_max * 1e-3)
    # roofline of the stage kernel, per GPU: one launch = one RK4 stage over the rank's rows
    avg_launch_ms = stage_ms_max / max(stage_n, 1)
    model_stage = (4 * (nnz_row * 20 + 4) / k + 20 * 16) / 4.0 * n_own * k          # SURVEY.md 8(d), one of the 4 stages
    layout_stage = info["bytes_per_stage"] + 5.0 * n_own * k * 16                     # operator as stored + 20/4 vector passes
    ent, why = ncu_traffic(f"{name}:rk4_stage_tbcsr_kernel")
    traffic = frac_meas = None
    if ent is not None and ent.get("rows_per_gpu"):
        traffic = ent["dram_bytes_per_launch"] * (n_own / float(ent["rows_per_gpu"]))   # capture scaled to this rank's rows
        frac_meas = traffic / (avg_launch_ms * 1e-3) / 1e9 / cx.peak
    sec = {"workload": note, "shape": list(shape), "N_total": N_total, "rows_per_gpu": n_own, "k": k, "Nw": Nw, "nnz_row": nnz_row,
           "scaling": scaling, "value": value, "unit": UNIT, "ms_per_step": dev_ms_max / max(args.steps, 1),
           "rk4_steps_per_bench_step": S, "setup_s": setup_s, "operator_format": info["format"],
           "gpu_launches": int(launch_sum),
           "roofline": {"bound": "hbm", "kernel": "rk4_stage_tbcsr_kernel (one launch per RK4 stage, halo exchange inside)" if info["format"] == "tiled-bcsr" else info["format"],
                        "avg_launch_ms": avg_launch_ms, "peak": cx.peak, "unit": "GB/s",
                        "model_bytes_per_launch": model_stage, "frac_model": model_stage / (avg_launch_ms * 1e-3) / 1e9 / cx.peak,
                        "layout_bytes_per_launch": layout_stage, "frac_layout": layout_stage / (avg_launch_ms * 1e-3) / 1e9 / cx.peak,
                        "traffic": traffic, "frac": frac_meas, "traffic_note": why or "ncu dram bytes per launch (profiles/traffic.json), scaled by rows",
                        "share_of_step": stage_ms_max / dev_ms_max if dev_ms_max else None,
                        "idft_share_of_step": idft_ms / dev_ms if dev_ms else None}}
    if adj is not None:
        sec["adjoint"] = adj
    if rank == 0:
        sec["clocks"] = sampler.summary()
    if rank == 0 and world == 1 and cpu_sample_shape is not None and not args.no_cpu:
        # the reference's own TSRK4.c on the same operator at a size the host finishes in seconds (rate is DRAM-bound, size-independent)
        S_ = pkg.synthetic
        Nc, rp, ci, va = S_.lns_stencil(cpu_sample_shape, nvar=nvar)
        from oracle import oracle as O
        O.build()
        g = O.setup_time_freq_grid(0.05, Nw, 0.003, 10.0)
        arm = CpuArm(S_, (rp, ci, va), g, Nw, 0.05, Nc)
        sec["cpu_baseline"] = cpu_baseline_block(arm, 4.0, f"the same operator on a {cpu_sample_shape} grid ({Nc} rows)")
    return sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end (host buffer) measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline samples")
    ap.add_argument("--no-lns", action="store_true", help="skip the LNS block-stencil sections")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity section")
    ap.add_argument("--only", default="", help="comma list of sections to run (gl,lns3d,lns3d_strong,lns2d,parity)")
    ap.add_argument("--option", action="append", default=[], help="name=value tuning option passed to rsvdt_set_option "
                    "(prefix lns: for the LNS sections)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    only = set(x for x in args.only.split(",") if x)
    want = lambda s: (not only) or (s in only)  # noqa: E731

    cx = Ctx(args)
    torch, dist, pkg = cx.torch, cx.dist, cx.pkg
    world, rank, local_rank = cx.world, cx.rank, cx.local_rank

    parity = None
    if want("parity") and not args.no_parity:
        try:
            parity = parity_section(cx)
        except Exception as e:   # a parity failure must not hide the timing: it is reported in the line
            parity = {"error": f"{type(e).__name__}: {e}"}
            if world > 1:
                raise

    c = WORKLOAD
    k, Nw = c["k"], c["Nw"]
    line = {}
    if want("gl"):
        grid_t = pkg.capi.setup_time_freq_grid(c["w"], c["Nw"], c["dt_in"], c["TransientLength"])
        dt, Ns, Nt, RR = grid_t
        if world == 1:
            N, rowptr, col, val = pkg.synthetic.ginzburg_landau(c["N"])
            r = pkg.capi.RSVDt(local_rank)
            r.LNSStructure(N, rowptr, col, val)
            r.set_grid(k, Nw, Ns, Nt, RR, dt, c["w"], False)
        else:
            # weak scaling: every GPU owns N rows of a Ginzburg-Landau operator with N*world rows, row-partitioned like
            # PETSc's MPIAIJ (layout 1: halo rows + one peer-store exchange per RK4 step)
            N = c["N"]
            Ntot = N * world
            gl = lambda a, b: pkg.synthetic.ginzburg_landau(Ntot, row_range=(a, b))[1:]  # noqa: E731
            grid = dict(k=k, Nw=Nw, Ns=Ns, Nt=Nt, ResRatio=RR, dt=dt, w=c["w"])
            r, plan = pkg.dist.setup_rank(pkg, local_rank, Ntot, gl, 3, 1, (2, 1), grid)
            assert plan.n_own == N
            rowptr = col = val = None
        for opt in args.option:
            name, value = opt.split("=")
            if not name.startswith("lns:"):
                r.set_option(name, int(value))

        # host forcing in pinned memory (torch is plumbing only: pinned allocation + process group).  Test vector 0 carries
        # the forcing of the CPU arm (seed 14), so that its state can be compared with the reference code's below.
        F_t = torch.empty((k * Nw, N), dtype=torch.complex128, pin_memory=True)   # row-major (cols, N) == column-major N x cols
        F = F_t.numpy().T                                                         # Fortran-ordered view, N x (k*Nw)
        rng = np.random.default_rng(1400 + rank)
        base = rng.random((N, 8)) + 1j * rng.random((N, 8))
        for cidx in range(k * Nw):
            F[:, cidx] = np.roll(base[:, cidx % 8], 7 * cidx)
        F0 = pkg.synthetic.random_forcing(N, 1, Nw, seed=14)
        for iw in range(Nw):
            F[:, iw * k] = F0[:, iw]
        del base
        r.set_Y_hat(F)

        # ---- parity at the bench size (N = 1): column 0 after n steps against the reference's own TSRK4.c ---------------------
        arm = None
        if rank == 0 and world == 1 and not args.no_cpu:
            arm = CpuArm(pkg.synthetic, (rowptr, col, val), grid_t, Nw, c["w"], N)
            nchk = 24
            r.time_forced_steps(nchk)
            y_gpu = r.get_state()[:, 0]
            y_ref = arm.state_after(nchk)
            if parity is None:
                parity = {}
            parity["parity_bench_size"] = {"rel_err": rel_err(y_gpu, y_ref), "steps": nchk, "rows": N,
                                           "against": arm.kind, "what": "state of test vector 0 after the forced RK4 steps, "
                                           "same operator / grid / forcing as the timed workload (fused TMA step kernel, Taylor synthesis)"}

        for _ in range(max(args.warmup, 3)):
            r.time_forced_steps(STEPS_PER_PASS)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        cx.barrier()
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
        cx.barrier()
        wall = time.time() - t0
        launches = r.kernel_launches()
        if rank == 0:
            sampler.stop()
        (dev_ms,) = cx.max_over_ranks(dev_ms)
        units = float(N) * k * STEPS_PER_PASS * args.steps * world   # every rank advances its own N rows x k test vectors
        value = units / (dev_ms * 1e-3)

        # ---- roofline of the dominant kernel (fused RK4 step) ---------------------------------------------------------------
        peak, peak_src = cx.peak, cx.peak_src
        nnz_row = c["nnz_row"]
        bytes_per_unit_step = 4 * (nnz_row * 20 + 4) / k + 20 * 16      # 353.6 B for GL, k = 10
        launches_per_step = stage_n / max(STEPS_PER_PASS * args.steps, 1)
        model_per_launch = bytes_per_unit_step * N * k / max(launches_per_step, 1e-9)
        avg_launch_ms = stage_ms / max(stage_n, 1)
        ent, why = ncu_traffic("gl:rk4_dia_tma_kernel")
        traffic = ent["dram_bytes_per_launch"] if ent else None
        achieved = traffic / (avg_launch_ms * 1e-3) / 1e9 if (traffic and stage_n) else None
        roofline = {"bound": "hbm", "kernel": "rk4_dia_tma_kernel (fused RK4 step), %.0f launch(es) per RK4 step" % launches_per_step,
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                    "traffic": traffic,
                    "note": "achieved = DRAM bytes per launch measured by ncu (profiles/traffic.json; " + (why or "capture matches the kernel source") +
                            ") / live launch duration; model_traffic_ratio = SURVEY.md 8(d) stage-by-stage model bytes over the same time "
                            "and peak: > 1 because the four stages are fused in shared memory (fewer bytes moved, not bandwidth above peak)",
                    "algorithmic_bytes_per_launch": model_per_launch,
                    "model_traffic_ratio": model_per_launch / (avg_launch_ms * 1e-3) / 1e9 / peak if stage_n else None,
                    "avg_launch_ms": avg_launch_ms, "launches_per_rk4_step": launches_per_step, "peak_source": peak_src,
                    "share_of_step": stage_ms / dev_ms if dev_ms else None,
                    "idft_share_of_step": idft_ms / dev_ms if dev_ms else None}

        # ---- end to end through the C ABI with host buffers -------------------------------------------------------------
        e2e = None
        if not args.no_e2e:
            Y_t = torch.empty((k * Nw, N), dtype=torch.complex128, pin_memory=True)
            Y = Y_t.numpy().T
            cx.barrier()
            t1 = time.time()
            r.TSActionRK4(F, direct=True, out=Y)
            cx.barrier()
            (t_e2e,) = cx.max_over_ranks(time.time() - t1)
            rend = Nt + Ns
            e2e = {"value": float(N) * k * rend * world / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(F.nbytes),
                   "d2h_bytes_per_step": int(Y.nbytes), "seconds": t_e2e, "rk4_steps": rend,
                   "call": "rsvdt_ts_action_rk4 (one whole action, host buffers, pinned)",
                   "checksum": float(np.abs(Y[::9973, ::37]).sum())}
            del Y, Y_t
        r.close()
        del F, F_t

        # ---- CPU baseline on the host cores (rank 0, N = 1 only) -------------------------------------------------------------
        cpu = None
        if arm is not None:
            cpu = cpu_baseline_block(arm)
            if arm.kind == "reference":
                port = CpuArm(pkg.synthetic, (rowptr, col, val), grid_t, Nw, c["w"], N, prefer_reference=False)
                cpu["port"] = cpu_baseline_block(port, 4.0)
                del port
            del arm

        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
                "config": config_dict(grid_t, world),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": sampler.summary(), "wall_s_timed_region": wall}
    else:
        line = {"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "higher_is_better": True, "scaling": "weak", "data": "synthetic", "dtype": "f64 (complex128)"}

    # ---- the north-star operators -----------------------------------------------------------------------------------------
    if not args.no_lns:
        nz = int(round(12.5 * world)) if world > 1 else 12
        sections = [
            ("lns3d", (nz, 200, 200), 8, 32, "weak",
             "BASELINE.json configs[3] per-GPU load: 3-D block stencil 5 vars x 27 points (complex), %d x 200 x 200 grid; "
             "N = 8 GPUs is the 2e7-DOF configuration" % nz, (12, 40, 40)),
            ("lns3d_strong", (100, 100, 100), 8, 32, "strong",
             "the same operator at a fixed 100 x 100 x 100 grid (5e6 DOF) on every GPU count (BASELINE.json configs[4]'s strong-scaling "
             "sweep: direct and adjoint stepping; the weights add < 2 % to a solve, profiles/r2_north_star)", None),
            ("lns2d", (1000, 1000), 10, 32, "strong",
             "BASELINE.json configs[2]: 2-D block stencil 5 vars x 9 points, 1000 x 1000 grid (5e6 DOF), k = 10, fixed size", (200, 200)),
        ]
        for name, shape, kk, nw, scaling, note, cpu_shape in sections:
            if not want(name):
                continue
            try:
                sec = lns_section(cx, name, shape, kk, nw, scaling, note, cpu_shape, with_adjoint=(name == "lns3d_strong"))
            except Exception as e:
                sec = {"error": f"{type(e).__name__}: {e}"}
                if world > 1:
                    raise
            if rank == 0:
                line[name] = sec

    if rank == 0:
        line["parity"] = parity
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
