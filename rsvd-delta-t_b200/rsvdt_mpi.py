#!/usr/bin/env python
"""RSVDt on several GPUs: the drop-in for `mpiexec -n P RSVDt -inputs variables.yaml` (reference README.md:221).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29500 \\
        rsvd-delta-t_b200/rsvdt_mpi.py -inputs variables.yaml

One process per GPU stands for one MPI rank of the reference.  Same YAML keys with the same required / default / conditional
rules (ReadUserInput.c:20-173), same PETSc-binary inputs and the same result files as the single-GPU executable
rsvd-delta-t_b200/RSVDt (host/RSVDt.c), whose sequence -- main.c:107-150 -- this file mirrors on partitioned rows:

  * every rank loads only its PETSC_DECIDE row range of A (MatLoad of a MATMPIAIJ, LNSStructure.c:22-27), of F_hat
    (CreateRandomMat.c:23-31) and of the initial vector (TransientRunRK4.c:82-86) -- petsc_io.MatRowReader;
  * all numerics run on the GPUs behind the C ABI (include/rsvdt.h, "row partition across GPUs"): halo rows by peer stores
    over NVLink, Gram matrices / norms all-reduced over NCCL (dist.py);
  * every rank writes its rows of U_hat / V_hat / the snapshots straight into the shared result files
    (petsc_io.DenseMatWriter / VecWriter), which come out byte-compatible with MatView / VecView (SaveResults.c:32-84).

torch is plumbing only (process group, NCCL collectives).  There is no CPU fallback: without a CUDA device per rank the
run stops at rsvdt_create.
"""
from __future__ import annotations

import math
import os
import shutil
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


class UserError(Exception):
    """SETERRQ(PETSC_COMM_WORLD, PETSC_ERR_USER, ...) of the reference"""


ERR_USER = 83
MAX_K = 32          # include/rsvdt.h RSVDT_MAX_K


# ---------------------------------------------------------------------------------------------------- options database
class RowCache:
    """rows(a, b) of an operator source that remembers ONE block it has handed out: the driver first reads its owned rows (band
    and type of the operator) and the partition planner then asks for a slab that contains them -- only the two flanks are read
    the second time (5 s of a 27 s load of a 6.3 GB file, profiles/r2_io)."""

    def __init__(self, rows_fn):
        self.fn, self.have = rows_fn, None

    def prime(self, a, b):
        out = self.fn(a, b)
        self.have = (a, b) + tuple(out)
        return out

    def __call__(self, a, b):
        if self.have is None or not (a <= self.have[0] and self.have[1] <= b):
            return self.fn(a, b)
        c, d, rp, cg, va = self.have
        self.have = None                                            # one use: the block is not kept any longer
        parts = ([self.fn(a, c)] if a < c else []) + [(rp, cg, va)] + ([self.fn(d, b)] if d < b else [])
        rps, off = [np.zeros(1, np.int64)], 0
        for p in parts:
            q = np.asarray(p[0], np.int64)
            rps.append(q[1:] + off)
            off += int(q[-1])
        return np.concatenate(rps), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])


def options_insert_file_yaml(path):
    """flat `key: value  # comment` lines, as PetscOptionsInsertFileYAML reads variables.yaml (ReadUserInput.c:22)"""
    opts = {}
    try:
        fh = open(path)
    except OSError:
        raise UserError(f"Unable to open YAML options file {path}")
    with fh:
        for line in fh:
            s = line.strip()
            if not s or s.startswith("#") or s.startswith("---") or s.startswith("..."):
                continue
            if ":" not in s:
                continue
            key, val = s.split(":", 1)
            val = val.strip()
            for i, ch in enumerate(val):                      # a comment starts at '#' preceded by white space
                if ch == "#" and (i == 0 or val[i - 1].isspace()):
                    val = val[:i]
                    break
            val = val.strip()
            if len(val) >= 2 and val[0] == val[-1] and val[0] in "\"'":
                val = val[1:-1]
            opts[key.strip()] = val
    return opts


class Options:
    def __init__(self, opts, log):
        self.o, self.log = opts, log

    def string(self, key):
        return self.o.get(key)

    def real(self, key):
        v = self.o.get(key)
        if v is None or v == "":
            return None
        try:
            return float(v)
        except ValueError:
            raise UserError(f"Input string {v} has no numeric value")

    def int(self, key):
        v = self.o.get(key)
        if v is None or v == "":
            return None
        try:
            return int(v)
        except ValueError:
            raise UserError(f"Input string {v} has no integer value (do not include . in it)")

    def bool(self, key):
        v = self.o.get(key)
        if v is None:
            return None
        if v.lower() in ("", "true", "yes", "1", "on"):
            return True
        if v.lower() in ("false", "no", "0", "off"):
            return False
        raise UserError(f"Unknown logical value: {v}")

    def default(self, getter, key, dflt, warn=True, shown=None):
        v = getter(key)
        if v is None:
            if warn:
                d = int(dflt) if isinstance(dflt, bool) else dflt
                self.log(f"Warning: '{key}' variable not found. Setting '{shown or key}' to default value: {d:g}"
                         if isinstance(d, float) else
                         f"Warning: '{key}' variable not found. Setting '{shown or key}' to default value: {d}")
            return dflt
        return v


def read_user_input(argv, log):
    """ReadUserInput.c:20-173 -> dict of run variables"""
    log("\n*** Reading user inputs ***")
    inputs = None
    for i in range(len(argv) - 1):
        if argv[i] == "-inputs":
            inputs = argv[i + 1]
    if not inputs:
        raise UserError("Must specify inputs variables via -inputs")
    o = Options(options_insert_file_yaml(inputs), log)
    v = {"inputs": inputs}

    def need(getter, key, msg=None):
        x = getter(key)
        if x is None:
            raise UserError(msg or f"Must specify '{key}'")
        return x

    v["TransientLength"] = need(o.real, "TransientLength")
    v["dt"] = need(o.real, "dt")
    v["w"] = need(o.real, "w", "Must specify 'dt'")          # sic: ReadUserInput.c:28 says 'dt'
    for flag, dirkey, shown in (("InputForcingFlg", "InputForcingDir", None), ("InvInputWeightFlg", "InvInputWeightDir", "InputWeightFlg"),
                                ("OutputWeightFlg", "OutputWeightDir", None), ("InvOutputWeightFlg", "InvOutputWeightDir", None),
                                ("InputMatrixFlg", "InputMatrixDir", None), ("OutputMatrixFlg", "OutputMatrixDir", None)):
        v[flag] = o.default(o.bool, flag, False, shown=shown)
        if v[flag]:
            v[dirkey] = need(o.string, dirkey)
    v["TwoPI"] = o.default(o.bool, "TwoPI", False)
    if v["TwoPI"]:
        v["w"] *= 2 * math.pi
    v["TransRun"] = o.default(o.bool, "TransRun", False)
    v["TransientRemoval"] = o.default(o.bool, "TransientRemoval", False)
    v["DiscFlg"] = o.default(o.bool, "DiscFlg", False)
    v["SaveResultsOpt"] = o.default(o.int, "SaveResultsOpt", 1)
    v["RootDir"] = need(o.string, "RootDir")
    v["ResultsDir"] = need(o.string, "ResultsDir")
    v["OperatorDir"] = need(o.string, "OperatorDir")
    v["k"] = o.default(o.int, "k", 3)
    if v["k"] > MAX_K:     # (not in the reference: a limit of this build, include/rsvdt.h RSVDT_MAX_K; said before anything is loaded)
        raise UserError(f"'k' = {v['k']}: this build supports at most {MAX_K} test vectors")
    v["q"] = o.default(o.int, "q", 0)
    v["Nw"] = need(o.int, "Nw")
    v["RandSeed"] = o.default(o.int, "RandSeed", 1373)
    v["Display"] = o.default(o.int, "Display", 2)
    beta = o.real("beta")
    if beta is None and v["DiscFlg"]:
        raise UserError("Discounting flag is on! Either set Discounting to zero or specify 'beta'")
    v["beta"] = beta if beta is not None else 0.0
    if v["beta"] < 0 and v["DiscFlg"]:
        raise UserError(f"'beta' must be positive, current value: {v['beta']:g}")
    tic = o.bool("TransICFlg")
    if tic is None and v["TransRun"]:
        v["TransICFlg"] = False
        log("Warning: 'TransICFlg' variable not found. Setting 'TransICFlg' to default value: 0")
    else:
        # the reference demands TransICDir whenever this branch is reached, even when unused (ReadUserInput.c:134-141);
        # like host/RSVDt.c we only fail when the value is needed
        v["TransICFlg"] = bool(tic)
        v["TransICDir"] = o.string("TransICDir")
        if v["TransICDir"] is None and v["TransRun"] and v["TransICFlg"]:
            raise UserError("Must specify 'TransICDir'")
    tr = v["TransRun"]
    v["TransDivVal"] = o.default(o.real, "TransDivVal", 1e3, warn=tr)
    v["TransConVal"] = o.default(o.real, "TransConVal", 1e-8, warn=tr)
    v["TransSave"] = o.default(o.bool, "TransSave", False, warn=tr)
    v["TransRemovalEst"] = o.default(o.bool, "TransRemovalEst", False, warn=tr)
    v["TransSaveMod"] = o.default(o.int, "TransSaveMod", 100, warn=tr)
    v["TransPeriods"] = o.default(o.int, "TransPeriods", 1, warn=tr)
    log("")
    return v


def hms(sec):
    hh = int(sec // 3600)
    mm = int((sec - 3600 * hh) // 60)
    ss = int(sec - 3600 * hh - 60 * mm)
    return f"{hh:02d}:{mm:02d}:{ss:02d}"


# ---------------------------------------------------------------------------------------------------------- the run
class Run:
    def __init__(self, argv):
        import torch
        import torch.distributed as dist

        import __graft_entry__ as entry

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        # RSVDT_MPI_HOSTONLY=1 is a dry run of the host logic for the CPU tests (gloo): inputs are parsed and partitioned, and
        # the run stops where the first CUDA handle would be created -- there is no CPU fallback for the numerics.
        self.hostonly = os.environ.get("RSVDT_MPI_HOSTONLY") == "1"
        self.dev = "cpu" if self.hostonly else "cuda"
        if not self.hostonly:
            if not torch.cuda.is_available():
                raise UserError("no usable CUDA device (this build has no CPU fallback)")
            torch.cuda.set_device(self.local_rank)
        if not dist.is_initialized():
            if self.hostonly:
                dist.init_process_group("gloo")
            else:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.pkg = entry.load_package()
        self.io = self.pkg.petsc_io
        self.argv = argv
        self.r = None

    def log(self, msg=""):
        if self.rank == 0:
            print(msg, flush=True)

    # -- PreProcessing.c ------------------------------------------------------------------------------------------------
    def lns_structure(self, v):
        """LNSStructure.c:19-58 on partitioned rows: every rank reads rows [r0, r1) (+ the halo / transpose slabs dist.py asks
        for) of the operator file; layout 1 (fused step kernel, halo recomputation) for narrow bands, else layout 0."""
        pkg, dist, torch = self.pkg, self.dist, self.torch
        t1 = time.time()
        path = v["RootDir"] + v["OperatorDir"]
        self.log(f"Loading the operator: {path} ")
        try:
            if v["OperatorDir"].startswith("synthetic:"):
                # an extension for the large BASELINE.json configurations: generate the rows tools/make_case.py would have
                # written instead of reading a 65 GB file (synthetic.SyntheticRows; same rows, tested)
                rd = pkg.synthetic.SyntheticRows(v["OperatorDir"][len("synthetic:"):])
            else:
                rd = self.io.MatRowReader(path)
        except (OSError, ValueError) as e:
            raise UserError(f"cannot read the operator {path}: {e}")
        if rd.M != rd.N:
            raise UserError(f"the operator must be square, current size = {rd.M} x {rd.N}")
        N = rd.M
        r0, r1 = pkg.dist.split_ownership(N, self.world, self.rank)
        rows_fn = rd.rows
        if hasattr(rd, "hints"):          # a generated operator knows its band and type: no pass over the rows needed here
            lo, hi, real = rd.hints()
        else:
            rows_fn = RowCache(rd.rows)
            rp, cg, va = rows_fn.prime(r0, r1)
            rows = np.repeat(np.arange(r0, r1, dtype=np.int64), np.diff(rp))
            d = cg - rows
            ext = torch.tensor([int(-d.min()) if d.size else 0, int(d.max()) if d.size else 0], dtype=torch.int64, device=self.dev)
            dist.all_reduce(ext, op=dist.ReduceOp.MAX)
            lo, hi = max(int(ext[0].item()), 0), max(int(ext[1].item()), 0)
            # the halo mailboxes are sized for k test vectors when the ranks connect, and k comes from the forcing file when
            # there is one (CreateRandomMat.c:30) divided by Nw_eff, which depends on whether A is real (LNSStructure.c:37-44)
            im2 = torch.tensor([float(np.sum(va.imag ** 2))], dtype=torch.float64, device=self.dev)
            dist.all_reduce(im2)
            real = float(im2.item()) ** 0.5 < 1e-12
            del rp, cg, va, rows, d
        if v["InputForcingFlg"] and not v["TransRun"]:
            try:
                fr = self.io.MatRowReader(v["RootDir"] + v["InputForcingDir"])
            except (OSError, ValueError) as e:
                raise UserError(f"cannot read the input forcing {v['RootDir'] + v['InputForcingDir']}: {e}")
            v["k"] = max(fr.N // (v["Nw"] // 2 if real else v["Nw"]), 1)
            if v["k"] > MAX_K:
                raise UserError(f"the input forcing holds k = {v['k']} test vectors: this build supports at most {MAX_K}")
        max_bw = max(lo, hi, 1)
        min_rows = N // self.world
        layout = 1 if (max_bw <= 8 and 8 * max_bw < min_rows) else 0
        if os.environ.get("RSVDT_LAYOUT") in ("0", "1"):          # testing hook: force the local layout
            layout = int(os.environ["RSVDT_LAYOUT"])
        self.band = (lo, hi)
        grid = dict(k=v["k"], Nw=v["Nw"], Ns=v["Ns"], Nt=v["Nt"], ResRatio=v["ResRatio"], dt=v["dt"], w=v["w"],
                    TransientRemoval=v["TransientRemoval"])
        if self.hostonly:
            self.log(f"host-only dry run: N = {N}, rows [{r0}, {r1}) on rank {self.rank}, band {lo}/{hi}, layout {layout}, "
                     f"real = {real}, k = {v['k']}")
        band = (lo, hi)
        # block-stencil operators of BASELINE.json: written by a kernel straight into the device layout (bit-identical to the
        # host generator, tests/test_stencil_gen.py) instead of passing 65 GB of CSR through the host
        on_device = (not self.hostonly and getattr(rd, "kind", "") in ("lns2d", "lns3d") and not v["TransRun"]
                     and v["k"] % 2 == 0 and os.environ.get("RSVDT_SYNTH_ON_DEVICE", "1") != "0")
        while True:
            try:
                if on_device:
                    layout = 0
                    self.r, self.plan = pkg.dist.setup_rank_stencil(pkg, self.local_rank, rd.shape, rd.nvar, grid,
                                                                    DiscFlg=v["DiscFlg"], beta=v["beta"])
                else:
                    self.r, self.plan = pkg.dist.setup_rank(pkg, self.local_rank, N, rows_fn, max_bw, layout,
                                                            band if layout == 1 else None, grid, DiscFlg=v["DiscFlg"], beta=v["beta"])
            except pkg.capi.RsvdtError as e:
                if self.hostonly:      # no CUDA device behind a host-only dry run: this is where it ends
                    raise UserError("host-only dry run stops before rsvdt_create (no CPU fallback)")
                raise UserError(str(e))
            if layout == 0:
                break
            # layout 1 needs rows the fused band kernel accepts on EVERY rank, for the number of vectors this run advances
            ok = torch.tensor([int(self.r.dist_layout_supported(1 if v["TransRun"] else v["k"]))], dtype=torch.int64, device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()):
                break
            self.log("Local layout 1 (fused band kernel on halo rows) is not available for this operator; using layout 0")
            self.r.close()
            layout = 0
        self.layout = layout
        v["N"] = N
        v["RealOperator"] = bool(self.r.real_operator)
        v["Nw_eff"] = v["Nw"] // 2 if v["RealOperator"] else v["Nw"]
        self.log(f"The operator is loaded in {time.time() - t1:f} seconds (N = {N}) ")
        if not v["TransRun"]:
            self.log("\nThe LNS operator is real-valued!" if v["RealOperator"] else "\nThe LNS operator is complex-valued!")
            self.log(f"The effective number of frequencies, Nw = {v['Nw_eff']}")
        if v["DiscFlg"]:
            self.log(f"\n---- Discounting with beta = {v['beta']:g} ----\n")
        self.log(f"Rows partitioned over {self.world} GPUs like MATMPIAIJ (PETSC_DECIDE); local layout {layout} "
                 f"(band {lo}/{hi})")

    def set_grid(self, v):
        self.r.set_grid(v["k"], v["Nw"], v["Ns"], v["Nt"], v["ResRatio"], v["dt"], v["w"], v["TransientRemoval"])

    def load_weight(self, which, label, rel, v):
        """MatLoad of a MATMPIAIJ weight matrix: every rank keeps its PETSC_DECIDE rows of M and of M^H (dist.local_weight).
        The file is read whole on every rank (these matrices have O(N) entries), then sliced."""
        path = v["RootDir"] + rel
        self.log(f"Reading the {label}: {path}")
        try:
            M, N, rp, ci, va = self.io.read_mat(path)
        except (OSError, ValueError) as e:
            raise UserError(f"cannot read {path}: {e}")
        Mr, MHr = self.pkg.dist.local_weight(M, N, rp, ci, va, self.world, self.rank)
        try:
            self.r.ReadWeightMats_local(which, M, N, Mr, MHr)
        except self.pkg.capi.RsvdtError as e:
            raise UserError(str(e))
        return M, N

    def read_weight_mats(self, v):
        """ReadWeightMats.c:26-99 (same size checks, same order)"""
        N = v["N"]
        row1 = col1 = N
        rq = -1
        if v["InvInputWeightFlg"]:
            row1, col1 = self.load_weight("W_f_sqrt_inv", "inverse input weight matrix  ", v["InvInputWeightDir"], v)
        if v["InputMatrixFlg"]:
            row2, col2 = self.load_weight("B", "input matrix                 ", v["InputMatrixDir"], v)
            if col2 != col1:
                raise UserError(f"Size mismatch between input matrix (B) and input weight matrix (W_f_sqrt_inv), {col2} != {col1}")
            v["Nb"] = col2
        else:
            v["Nb"] = N
            if v["InvInputWeightFlg"] and v["Nb"] != col1:
                raise UserError("Size mismatch between input matrix (B) and input weight matrix (W_f_sqrt_inv)")
        if v["InvOutputWeightFlg"]:
            rq, _ = self.load_weight("W_q_sqrt_inv", "inverse output weight matrix ", v["InvOutputWeightDir"], v)
        row1 = N
        if v["OutputWeightFlg"]:
            row1, col1 = self.load_weight("W_q_sqrt", "output weight matrix         ", v["OutputWeightDir"], v)
        if v["OutputMatrixFlg"]:
            row2, col2 = self.load_weight("C", "output matrix                ", v["OutputMatrixDir"], v)
            v["Nc"] = row2
            if row2 != row1:
                raise UserError(f"Size mismatch between output matrix (C) and output weight matrix (W_q_sqrt), {row2} != {row1}")
            if v["InvOutputWeightFlg"] and row2 != rq:
                raise UserError(f"Size mismatch between output matrix (C) and inverse output weight matrix (W_q_sqrt_inv), {row2} != {rq}")
        else:
            v["Nc"] = N
            if v["OutputWeightFlg"] and v["Nc"] != row1:
                raise UserError("Size mismatch between output matrix (C) and output weight matrix (W_q_sqrt)")
            if v["InvOutputWeightFlg"] and v["Nc"] != rq:
                raise UserError("Size mismatch between output matrix (C) and inverse output weight matrix (W_q_sqrt_inv)")

    def create_random_mat(self, v):
        """CreateRandomMat.c:16-42, rows [r0, r1) of the N_b x (k*Nw_eff) forcing"""
        b0, b1 = self.pkg.dist.split_ownership(v["Nb"], self.world, self.rank)
        if v["InputForcingFlg"]:
            path = v["RootDir"] + v["InputForcingDir"]
            self.log(f"Reading the input forcing matrix         : {path}")
            try:
                rd = self.io.MatRowReader(path)
            except (OSError, ValueError) as e:
                raise UserError(f"cannot read the input forcing {path}: {e}")
            if rd.M != v["Nb"]:
                raise UserError(f"Input forcing must be have {v['Nb']} rows, current size = {rd.M} x {rd.N}")
            v["k"] = rd.N // v["Nw_eff"]
            self.log(f"Based on the input forcing: k = {v['k']}")
            F = rd.dense_rows(b0, b1)
            if v["k"] < 1 or F.shape[1] != v["k"] * v["Nw_eff"]:
                raise UserError(f"Input forcing has {rd.N} columns, not a multiple of Nw = {v['Nw_eff']}")
        else:
            self.log("Generating a random forcing matrix")
            if not self.hostonly and hasattr(self.r, "CreateRandomMat"):
                # CreateRandomMat.c:36-39 on the device, keyed by (seed, global row, column): no 10 GB host array per rank
                self.r.CreateRandomMat(b1 - b0, b0, v["RandSeed"])
                return
            F = random_rows(v["RandSeed"], b0, b1, v["k"] * v["Nw_eff"])
        if int(self.r.grid.k) != v["k"]:
            raise UserError(f"internal: the ranks were connected for k = {int(self.r.grid.k)} test vectors, the forcing has {v['k']}")
        self.r.set_Y_hat(F)

    def create_results_dir(self, v, name):
        """CreateResultsDir.c:18-29: rank 0 picks the first free <name><i>/ and creates it"""
        box = [None]
        if self.rank == 0:
            i = 0
            base = v["RootDir"] + v["ResultsDir"]
            while os.path.exists(f"{base}{name}{i}/"):
                i += 1
            box[0] = f"{base}{name}{i}/"
            try:
                os.mkdir(box[0])
            except OSError as e:
                box[0] = ("error", f"cannot create {box[0]}: {e}")
        self.dist.broadcast_object_list(box, src=0)
        if isinstance(box[0], tuple):
            raise UserError(box[0][1])
        v["FolderDir"] = box[0]

    def save_input_vars_copy(self, v):
        if self.rank == 0:                                                            # SaveInputVarsCopy.c:27-37
            shutil.copyfile(v["inputs"], v["FolderDir"] + os.path.basename(v["inputs"]))

    # -- collective writers -----------------------------------------------------------------------------------------------
    def write_dense(self, path, M, ncols, row0, block):
        w = self.io.DenseMatWriter(path, M, ncols)
        if self.rank == 0:
            w.create()
        self.dist.barrier()
        w.write_rows(row0, block)

    def write_vec(self, path, n, row0, vals):
        w = self.io.VecWriter(path, n)
        if self.rank == 0:
            w.create()
        self.dist.barrier()
        w.write_rows(row0, vals)

    # -- TransientRunRK4.c --------------------------------------------------------------------------------------------------
    def transient_run(self, v):
        self.log("\n********************************************\n************** Transient run ***************\n"
                 "********************************************\n")
        N, Ns = v["N"], v["Nw"]
        TSS = 2 * math.pi / v["w"]                                                    # :40
        deltaT = TSS / Ns
        v["dt"] = deltaT / math.ceil(deltaT / v["dt"])                                # :42
        Nt_period = int(round(TSS / v["dt"]))
        rend = Nt_period * v["TransPeriods"]                                          # :45
        Nt_saved = rend // v["TransSaveMod"] + 1                                      # :46
        self.log(f"Total integration time = {v['dt'] * rend:g} = {v['TransPeriods']} x {TSS:g} with dt = {v['dt']:g}")
        self.log(f"Total snapshots to be saved = {Nt_saved}" if v["TransSave"] else "Saving snapshots is not requested")
        v["Ns"], v["Nt"], v["ResRatio"] = Nt_period, 0, max(Nt_period // Ns, 1)
        v["k"] = max(v["k"], 1)
        v["TransientRemoval"] = False
        self.set_grid(v)
        r0, r1 = self.plan.r0, self.plan.r1
        if v["TransICFlg"]:
            self.log("Reading the initial vector")
            path = v["RootDir"] + v["ResultsDir"] + v["TransICDir"]                    # :84
            try:
                n, q0 = self.io.read_vec_rows(path, r0, r1)
            except (OSError, ValueError) as e:
                raise UserError(f"cannot read the initial vector {path}: {e}")
            if n != N:
                raise UserError(f"initial vector has {n} entries, expected {N}")
        else:
            self.log("Random vector is generated\n")
            q0 = random_rows(v["RandSeed"], r0, r1, 1)[:, 0]
            t = self.torch.tensor([float(np.vdot(q0, q0).real)], dtype=self.torch.float64, device=self.dev)
            self.dist.all_reduce(t)
            q0 = q0 / math.sqrt(float(t.item()))                                      # VecNormalize, :93
        folder = v["FolderDir"]

        def on_snapshot(pos, q):                                                      # :116-120, :160-164 (collective)
            self.write_vec(f"{folder}q_transient_{pos}", N, r0, q)

        # TransientRemovalEst.c (TransientRunRK4.c:56-65,96-114,174-184): the last Nw snapshots, one every Nt_period / Nw steps, and
        # at the end of every period the estimate of how well the transient could be removed (transient_est.py); each rank
        # holds its rows, the sums run over the ranks
        est_mod, est_msgs, on_est = 0, [], None
        if v["TransRemovalEst"]:
            transient_est = self.pkg.transient_est
            est_mod = Nt_period // Ns
            Q_all = np.zeros((r1 - r0, Ns), np.complex128)
            Q_all[:, 0] = q0
            qss = -np.asarray(q0, np.complex128)
            torch, dist = self.torch, self.dist

            def allreduce(x):
                x = np.ascontiguousarray(x)
                t = torch.from_numpy(x.view(np.float64).copy()).to(self.dev)
                dist.all_reduce(t)
                return t.cpu().numpy().view(x.dtype).reshape(x.shape)

            def allgather(x):
                parts = [None] * self.world
                dist.all_gather_object(parts, x)
                return parts
            allgather.rank = self.rank

            def on_est(pos_e, q):
                step = pos_e * est_mod
                pos = pos_e % Ns
                if pos == 0:
                    period = step // Nt_period
                    lines = []
                    init, upd = transient_est.transient_removal_est(Q_all, qss, period, v["w"], v["Nw_eff"], lines.append,
                                                                    allreduce, allgather)
                    est_msgs.append((step, lines))
                    if self.rank == 0:                                                # TransientRemovalEst.c:123-129
                        self.io.write_vec(f"{folder}Initial_transient_norm_period_{period}", init.astype(np.complex128))
                        self.io.write_vec(f"{folder}Updated_transient_norm_period_{period}", upd.astype(np.complex128))
                Q_all[:, pos] = q

        if v["TransSave"]:
            on_snapshot(0, q0)
        t1 = time.time()
        q, norms, steps, reason = self.r.TransientRunRK4(q0, rend, v["TransSaveMod"], v["TransDivVal"], v["TransConVal"], True,
                                                         on_snapshot if v["TransSave"] else None, est_mod, on_est)
        self.log(f"Norm of the initial snapshot: {norms[0]:f} @ t = 0")
        nchecks = steps // v["TransSaveMod"]
        msgs = [(p * v["TransSaveMod"], 0, [f"Norm of the snapshot: {norms[p]:f} @ t = {p * v['TransSaveMod'] * v['dt']:f}"])
                for p in range(1, nchecks + 1) if not (reason and p == nchecks)]
        msgs += [(step, 1, lines) for step, lines in est_msgs]                        # (the estimate of a step follows its norm line)
        for _, _, lines in sorted(msgs, key=lambda m: (m[0], m[1])):
            for ln in lines:
                self.log(ln)
        if reason == 1:
            self.log("Stopped! The norm of the last snapshot is NAN!")
        if reason == 2:
            self.log(f"Diverged! The norm of the last snapshot exceeds the divergence threshold = {v['TransDivVal']:g}")
        if reason == 3:
            self.log(f"Converged! The norm of the last snapshot is smaller than the convergence threshold = {v['TransConVal']:g}")
        self.log(f"*** Transient simulation elapsed time = {hms(time.time() - t1)} ***")
        self.log(f"\nThe results directory is: {folder}\n")
        if self.rank == 0:                                                            # :201-205
            self.io.write_vec(folder + "q_transient_norms", norms.astype(np.complex128))
        self.write_vec(folder + "q_transient_last_snapshot", N, r0, q)                 # :207-209
        self.save_input_vars_copy(v)
        self.log("DONE :))\n")

    # -- SaveResults.c:32-84 ----------------------------------------------------------------------------------------------
    def save_results(self, v, S, U, V):
        k, Nwe, folder = v["k"], v["Nw_eff"], v["FolderDir"]
        self.log("\n********************************************\n************ Saving the output *************\n"
                 "********************************************")
        self.log(f"\nThe results directory: {folder}\n")
        self.log(f"One matrix of size {k} (modes) x {Nwe} (frequencies) for gains")
        if self.rank == 0:
            self.io.write_mat_dense(folder + "S_hat", S.astype(np.complex128))
        c0, _ = self.pkg.dist.split_ownership(v["Nc"], self.world, self.rank)
        b0, _ = self.pkg.dist.split_ownership(v["Nb"], self.world, self.rank)
        if os.environ.get("RSVDT_MPI_SAVE") == "gains":
            # benchmark runs of the 2e7-DOF configurations: the modes are 2 x 82 GB of output; keep the gains only
            self.log("RSVDT_MPI_SAVE=gains: the resolvent modes are not written\n")
        elif v["SaveResultsOpt"] == 1:
            self.log(f"{k} matrices of size {v['Nc']} x {Nwe} for response modes\n{k} matrices of size {v['Nb']} x {Nwe} for forcing modes\n")
            for j in range(k):                                  # mode j across all frequencies: columns iw*k + j
                self.write_dense(f"{folder}U_hat_k{j + 1}_allFreqs", v["Nc"], Nwe, c0, U[:, j::k])
                self.write_dense(f"{folder}V_hat_k{j + 1}_allFreqs", v["Nb"], Nwe, b0, V[:, j::k])
        else:
            self.log(f"{Nwe} matrices of size {v['Nc']} x {k} for response modes\n{Nwe} matrices of size {v['Nb']} x {k} for forcing modes\n")
            for iw in range(Nwe):
                self.write_dense(f"{folder}U_hat_Freq{iw}_allK", v["Nc"], k, c0, U[:, iw * k:(iw + 1) * k])
                self.write_dense(f"{folder}V_hat_Freq{iw}_allK", v["Nb"], k, b0, V[:, iw * k:(iw + 1) * k])
        self.dist.barrier()
        self.save_input_vars_copy(v)
        self.log("DONE :))\n")

    # -- main.c:107-150 -------------------------------------------------------------------------------------------------------
    def main(self):
        v = read_user_input(self.argv, self.log)
        if not v["TransRun"]:
            self.log("********************************************\n*************** Problem info ***************\n"
                     "********************************************\n")
        pkg = self.pkg
        dt, Ns, Nt, RR = pkg.capi.setup_time_freq_grid(v["w"], v["Nw"], v["dt"], v["TransientLength"])   # SetupTimeFreqGrid.c
        v.update(dt=dt, Ns=Ns, Nt=Nt, ResRatio=RR)
        if not v["TransRun"]:
            self.log(f"RSVD parameters:\n@ k        = {v['k']}\n@ q        = {v['q']}\n@ w        = {v['w']:g}\n")
            self.log(f"Time-stepping parameters:\n@ TransientLength = {v['TransientLength']:g}\n@ dt              = {dt:g}\n"
                     f"@ Ns              = {Ns}\n@ Nt              = {Nt}")
            self.log("Transient removal strategy is requested\n" if v["TransientRemoval"] else "Transient removal strategy is NOT requested\n")
        t_start = time.time()
        self.lns_structure(v)
        if not v["TransRun"]:
            self.read_weight_mats(v)
            self.create_random_mat(v)
        t_setup = time.time() - t_start
        self.create_results_dir(v, "TransientSnapshots_" if v["TransRun"] else "ResolventModes_")
        r = self.r
        if v["TransRun"]:                                                             # main.c:119-124
            self.transient_run(v)
            return 0
        self.log("\n********************************************\n*************** RSVD-\\Delta t **************\n"
                 "********************************************")

        rend = v["Nt"] + v["Ns"] + (v["ResRatio"] if v["TransientRemoval"] else 0)      # TSActionRK4.c:38-41
        self.phase_times = []

        def timed(banner, fn, done):
            self.log(banner)
            t0 = time.time()
            fn()
            self.dist.barrier()
            sec = time.time() - t0
            self.phase_times.append((done.strip("* ").replace(" elapsed time =", ""), sec))
            self.log(f"{done}{hms(sec)} ***")
            if "action" in done:   # not in the reference: the rate of the stepping loop, the metric BASELINE.json quotes
                self.log(f"    {sec:.3f} s for {rend} RK4 steps of {v['N']} x {v['k']} unknowns: "
                         f"{v['N'] * v['k'] * rend / max(sec, 1e-9):.3e} DOF*vector*steps/s (incl. transform / transient removal)")

        direct = lambda: timed("\n*** Direct action begins! ***", r.DirectActionRK4, "*** Direct action elapsed time = ")   # noqa: E731
        adjoint = lambda: timed("\n*** Adjoint action begins! ***", r.AdjointActionRK4, "*** Adjoint action elapsed time = ")  # noqa: E731
        qr = lambda: timed("\n*** QR decomposition begins! ***", r.QRAllFreqs, "*** QR decomposition elapsed time = ")  # noqa: E731
        direct()                                                                      # main.c:134
        for iq in range(v["q"]):                                                      # PowerIterationRK4.c:20-29
            self.log(f"\n******** Inside power iteration, {iq + 1}/{v['q']} *******")
            qr()
            adjoint()
            qr()
            direct()
        if v["q"] > 0:
            self.log("\n******** Power iteration DONE! *************")
        timed("\n*** Reduced SVD begins! ***", r.SVDAllFreqsBeforeAdjoint, "*** Reduced SVD elapsed time = ")   # :138
        r.StoreU()                                                                    # :140
        adjoint()                                                                     # :142
        timed("\n*** Reduced SVD begins! ***", r.SVDAllFreqs, "*** Reduced SVD elapsed time = ")               # :144
        c0, c1 = pkg.dist.split_ownership(v["Nc"], self.world, self.rank)
        b0, b1 = pkg.dist.split_ownership(v["Nb"], self.world, self.rank)
        t_solve = time.time() - t_start - t_setup
        t0 = time.time()
        gains_only = os.environ.get("RSVDT_MPI_SAVE") == "gains"      # (then the 2 x 10 GB per GPU of modes stay on the device)
        self.save_results(v, r.get_S_hat(), None if gains_only else r.get_U_hat(c1 - c0),
                          None if gains_only else r.get_V_hat(b1 - b0))                                          # :146
        if os.environ.get("RSVDT_MPI_TIMINGS") and self.rank == 0:   # not in the reference: the record of the run, as JSON
            import json
            with open(os.environ["RSVDT_MPI_TIMINGS"], "w") as f:
                json.dump({"N": v["N"], "k": v["k"], "Nw": v["Nw"], "q": v["q"], "TransientRemoval": bool(v["TransientRemoval"]),
                           "n_gpus": self.world, "dt": v["dt"], "Ns": v["Ns"], "Nt": v["Nt"], "ResRatio": v["ResRatio"],
                           "rk4_steps_per_action": rend, "setup_s (operator, weights, forcing)": t_setup,
                           "solve_s (main.c:134-144)": t_solve, "save_s": time.time() - t0,
                           "phases_s": [[n, s] for n, s in self.phase_times],
                           "kernel_launches_rank0": int(r.kernel_launches())}, f, indent=1)
        return 0


def random_rows(seed, r0, r1, ncols, block=65536):
    """uniform [0,1) + i[0,1) entries that depend only on (seed, global row, column), not on the number of ranks: rows are
    drawn in fixed blocks of `block` rows, block b of column c from numpy's default_rng([seed, c, b])."""
    out = np.empty((r1 - r0, ncols), np.complex128, order="F")
    for c in range(ncols):
        b = r0 // block
        while b * block < r1:
            lo, hi = max(r0, b * block), min(r1, (b + 1) * block)
            rng = np.random.default_rng([int(seed), c, b])
            z = rng.random(2 * block).reshape(2, block)
            s = slice(lo - b * block, hi - b * block)
            out[lo - r0:hi - r0, c] = z[0, s] + 1j * z[1, s]
            b += 1
    return out


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    run = None
    try:
        run = Run(argv)
        rc = run.main()
    except UserError as e:
        if int(os.environ.get("RANK", "0")) == 0:
            print(f"[0]RSVDt ERROR: {e}", file=sys.stderr, flush=True)
        rc = ERR_USER
    except Exception as e:
        # a library error on ONE rank (e.g. a QR breakdown in its rows, a peer that stopped answering): the other ranks sit in
        # collectives, so this process leaves at once -- the launcher (torchrun, like mpiexec) then takes the other ranks down
        print(f"[{os.environ.get('RANK', '0')}]RSVDt ERROR: {e}", file=sys.stderr, flush=True)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            os._exit(getattr(e, "code", 1) or 1)
        rc = getattr(e, "code", 1) or 1
    finally:
        if run is not None and run.r is not None:
            run.r.close()
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
