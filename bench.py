#!/usr/bin/env python
"""bench.py -- end-to-end SHARP throughput (cells/sec) on B200, with the roofline of the dominant kernel and the
CPU baseline beside it.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank/GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the one the metric is quoted on; it fits one GPU as CSC): SHARP_unlimited on a
synthetic 27 998 genes x 1 306 127 cells UMI matrix given as 26 sparse parts (25 x 50 000 + 56 127, the 10x-brain
layout of the reference's README), exp.type = "UMI" (CPM normalisation fused into the projection load), K = 5,
p = 508, rN.seed = 2103, viewflag = FALSE.  A "step" is one complete SHARP_unlimited call over all parts.
With N GPUs whole parts are dealt round-robin to the ranks while they divide evenly and the left-over parts are
block-sharded over all ranks (strong scaling: the job is always the 1.3 M cells); the communicator is NCCL behind the
C ABI (sharp_comm_*), torch.distributed.run is only the process launcher.  Exchanges: the ranM members a rank drew,
block-level labels + enE rows of the sharded parts, part-level centroids and labels before the global sMetaC.

Other workloads (--workload): cfg2 / cfg3 = BASELINE.json configs[1] / [2] through SHARP() (one matrix; cfg3 deals the
cell blocks over the ranks), cfg5 = config 5's shape bounded to 1e6 cells, SHARP_unlimited3 streaming SHCSC001 files.

  value : whole-job cells/sec with every part already resident in HBM (sharp_expr_upload before the timed region)
  e2e   : the same call on HOST buffers (pinned dgCMatrix slots): per-part H2D copies and the D2H of labels and
          centroids are inside the timed region
Both are timed on the device (CUDA events on the library's stream around the K steps), max over ranks.
Rank 0 prints ONE JSON line, the last line of stdout (with N > 1 NCCL itself prints its version banner before it).

Only the cpu_baseline leg and --impl reference execute anything under oracle/ (the CPU restatement of the reference,
OpenMP over (member, block) tasks like the reference's foreach), on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SEED = 2103
METRIC = "cells/sec end-to-end SHARP at 1.3M cells"
# dram__bytes_read.sum + dram__bytes_write.sum per launch, read at start from profiles/traffic.json -- written by
# profiles/summarize_ncu.py from the committed `ncu --set full` raw CSV it names (with the commit the capture was taken at)
def load_traffic() -> dict:
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {"kernels": {}, "source": None}


# the fused loop over parts keeps ~10 streams busy; with the default 8 hardware queues, streams alias and false
# dependencies serialise copies and kernels of different parts (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def workload(name: str) -> dict:
    if name == "cfg4":
        return dict(name="SHARP_unlimited, synthetic UMI 27998 genes x 1306127 cells (10x brain shape), 26 CSC parts "
                         "(25 x 50000 + 56127), exp.type=UMI, K=5, p=508, rN.seed=2103, viewflag=FALSE",
                    m=27998, parts=[50000] * 25 + [56127], K=5, types=20, nnz_per_cell=2000, exp_type="UMI")
    if name == "cfg2":  # BASELINE.json configs[1]: the default ensemble of SHARP() on one GPU (SHARP_large: 10 000 >= base.ncells)
        return dict(name="SHARP(), synthetic TPM 20000 genes x 10000 cells, dense fp64 (70 % zeros), default ensemble "
                         "(SHARP_large, K=5, p=333, 5 blocks of 2000), rN.seed=2103",
                    single=True, m=20000, parts=[10000], K=5, types=8, nnz_per_cell=6000, exp_type="TPM", dense=True,
                    metric="cells/sec end-to-end SHARP at 10k cells (config 2)")
    if name == "cfg3":  # BASELINE.json configs[2]: 15-member ensemble, cell blocks dealt over the GPUs
        return dict(name="SHARP(exp.type=UMI, ensize.K=15), synthetic UMI 20000 genes x 100000 cells (CSC, ~93 % zeros), "
                         "p=416, 50 blocks of 2000 dealt over the ranks, rN.seed=2103",
                    single=True, m=20000, parts=[100000], K=15, types=10, nnz_per_cell=1400, exp_type="UMI", dense=False,
                    metric="cells/sec end-to-end SHARP at 100k cells, K=15 (config 3)")
    if name == "cfg5":  # BASELINE.json configs[4], bounded: the parts live in SHCSC001 files and are streamed (read ->
        # pinned -> H2D -> cluster); 20 files of 50 000 cells per rank-set by default (--parts bounds it)
        return dict(name="SHARP_unlimited3 on SHCSC001 files (streaming ingestion), synthetic UMI 20000 genes x 1000000 cells "
                         "(config 5's shape, bounded from 1e7 cells), 20 files of 50000 cells, exp.type=UMI, K=5, rN.seed=2103",
                    streamed=True, m=20000, parts=[50000] * 20, K=5, types=15, nnz_per_cell=1000, exp_type="UMI",
                    metric="cells/sec end-to-end SHARP_unlimited3, parts streamed from files (config 5, bounded)")
    if name == "dev":  # development / CPU-side dry runs
        return dict(name="dev: 3000 genes x 3 parts of 2600 cells", m=3000, parts=[2600] * 3, K=3, types=5,
                    nnz_per_cell=300, exp_type="UMI")
    raise SystemExit(f"unknown workload {name}")


# ----------------------------------------------------------------------------------------------------------
# synthetic data: planted cell types, Poisson UMI counts, generated with torch on the GPU (plumbing), one part
# at a time, returned as dgCMatrix slots in pinned host memory
# ----------------------------------------------------------------------------------------------------------
def type_profiles(torch, dev, m, G, nnz_target):
    g = torch.Generator(device=dev)
    g.manual_seed(SEED)
    # heavy-tailed gene means and strong type-specific programmes: at ~2000 detected genes per cell the projected
    # blocks then cluster like real data do (median silhouette ~0.65 at the planted number of types; with weak
    # programmes every block falls through to the CH / one-cluster paths, which is not what SHARP is run on)
    base = torch.randn(m, generator=g, device=dev) * 2.0
    de = (torch.rand(G, m, generator=g, device=dev) < 0.5).float() * torch.randn(G, m, generator=g, device=dev) * 3.0
    mu = torch.exp(base[None, :] + de)  # G x m
    lo = torch.full((G, 1), 1e-9, device=dev)
    hi = torch.full((G, 1), 1e6, device=dev)
    for _ in range(80):  # depth per type such that E[nnz per cell] = nnz_target
        mid = torch.sqrt(lo * hi)
        nnz = (1.0 - torch.exp(-mid * mu)).sum(1, keepdim=True)
        lo = torch.where(nnz < nnz_target, mid, lo)
        hi = torch.where(nnz < nnz_target, hi, mid)
    return mu * torch.sqrt(lo * hi)  # G x m Poisson means at unit depth factor


def gen_part(torch, dev, lam_types, n, part_idx, pin):
    G, m = lam_types.shape
    g = torch.Generator(device=dev)
    g.manual_seed(SEED * 1000 + part_idx)
    types = torch.randint(0, G, (n,), generator=g, device=dev)
    depth = torch.exp(0.2 * torch.randn(n, generator=g, device=dev))
    counts_per_cell = torch.empty(n, dtype=torch.int64, device=dev)
    rows, vals = [], []
    chunk = 4096
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        lam = lam_types[types[c0:c1]] * depth[c0:c1, None]
        cnt = torch.poisson(lam, generator=g)
        nzmask = cnt != 0
        few = nzmask.sum(1) < 3  # a constant projection makes scale() produce NaN in the reference
        if bool(few.any()):
            cnt[few, :5] += 1.0
            nzmask = cnt != 0
        counts_per_cell[c0:c1] = nzmask.sum(1)
        idx = nzmask.nonzero(as_tuple=False)  # sorted by (cell, gene): CSC order with ascending row indices
        rows.append(idx[:, 1].to(torch.int32))
        vals.append(cnt[nzmask].to(torch.float64))
    colptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    colptr[1:] = torch.cumsum(counts_per_cell, 0)
    rowidx = torch.cat(rows)
    val = torch.cat(vals)

    def host(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=pin)
        h.copy_(t)
        return h

    hc, hr, hv = host(colptr), host(rowidx), host(val)
    return {"_keep": (hc, hr, hv), "p": hc.numpy(), "i": hr.numpy(), "x": hv.numpy(), "Dim": (m, n),
            "types": types.cpu().numpy().astype(np.int32)}


# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1]))
                smax.append(float(t[2]))
                power.append(float(t[3]))
            except ValueError:
                continue
            for nm, v in zip(names, t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def h2d_bandwidth(torch, dev) -> float:
    """pinned host -> device copy bandwidth of this box (GB/s): the floor of the e2e leg is h2d_bytes / this"""
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d.copy_(h, non_blocking=True)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return n / (best * 1e-3) / 1e9


def dgemm_peak_tflops(torch, dev) -> float:
    """fp64 tensor-pipe denominator for the DMMA distance kernel: cuBLAS DGEMM measured in this run (there is no
    fp64 figure in MEASURED_PEAKS.json)."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


# ----------------------------------------------------------------------------------------------------------
# algorithmic bytes / flops of one step, per kernel class (DESIGN.md states the per-unit figures)
# ----------------------------------------------------------------------------------------------------------
def block_sizes(n, ng=2000):
    if n <= ng:
        return [n]
    T = -(-n // ng)
    nt = n - (T - 2) * ng
    return [ng] * (T - 2) + [nt // 2, nt - nt // 2]


def algorithmic_work(wl, parts_nnz, part_sizes, p) -> dict:
    K = wl["K"]
    ldu = (p + 15) // 16 * 16
    out = {}
    # K1: CSC column as stored (rowidx 4 B + value 8 B per non-zero, 8 B colptr) read once + K*p fp64 outputs
    out["rp_project"] = ("hbm", sum(nz * 12 + n * 8 + n * K * p * 8 for nz, n in zip(parts_nnz, part_sizes)))
    blocks = [b for n in part_sizes for b in block_sizes(n)]
    # K2: n(n+1)/2 pairs x p multiply-adds per (member, block) problem (the symmetric half is mirrored, not computed)
    out["corrdist"] = ("tensor", sum(1.0 * b * (b + 1) * ldu for b in blocks) * K)
    # K3: Ward reads the n x n working matrix at least once and rewrites one row + one column per merge
    out["hclust"] = ("hbm", sum(b * b * 8 + (b - 1) * b * 8 * 3 for b in blocks) * K)
    # K4-K6: one pass over D per problem + the unit rows once
    out["sweep_nested"] = ("hbm", sum(b * b * 8 + b * ldu * 8 for b in blocks) * K)
    out["unit_rows"] = ("hbm", sum(n * K * (p + ldu) * 8 for n in part_sizes))
    out["ene_scatter"] = ("hbm", sum(n * p * 8 * (K + 1) + n * p * 16 for n in part_sizes))
    out["colsum"] = ("hbm", sum(nz * 8 + n * 16 for nz, n in zip(parts_nnz, part_sizes)))
    return out


# ----------------------------------------------------------------------------------------------------------
def sample_of(part, sample_cells):
    """the first `sample_cells` cells of one part as dgCMatrix slots"""
    n = min(sample_cells, part["Dim"][1])
    cp = part["p"][:n + 1].astype(np.int64)
    nz = int(cp[-1])
    return n, (cp, part["i"][:nz], part["x"][:nz])


def cpu_baseline_run(wl, part, p, sample_cells, rms, reind):
    """the oracle (CPU restatement of the reference) on the first `sample_cells` cells of one part: SHARP_large with
    the same ranM matrices, K, p and block size -- blocks, per-block wMetaC, sMetaC across the blocks, merge, relabel.
    Returns (cells/sec, threads, seconds, oracle result)."""
    import orc
    n, csc = sample_of(part, sample_cells)
    colsum = np.add.reduceat(csc[2], csc[0][:-1]) if wl["exp_type"] == "UMI" else None
    if wl.get("dense"):  # config 2 is TPM: the synthetic counts scaled to 1e6 per cell (the same numbers the dense matrix holds)
        cs = np.add.reduceat(csc[2], csc[0][:-1])
        csc = (csc[0], csc[1], csc[2] / np.repeat(cs, np.diff(csc[0])) * 1e6)
    prm = orc.SharpParams(1, 1, wl["K"], p, 2000, 0, 0, 0, orc.hc_params(max_n=max(40, -(-n // 5000))), 2, -1)
    t0 = time.time()
    ref = orc.sharp(wl["m"], n, rms, prm, csc=csc, colsum=colsum, reind=reind, want_vie=True, want_x0=False)
    dt = time.time() - t0
    return n / dt, orc.num_threads(), dt, ref


def first_appearance(y):
    _, idx, inv = np.unique(y, return_index=True, return_inverse=True)
    rank = np.empty(len(idx), dtype=np.int64)
    rank[np.argsort(idx)] = np.arange(1, len(idx) + 1)
    return rank[inv]


def parity_check(ctx, wl, part, p, sample_cells, rms, reind, ref):
    """GPU <-> oracle on the benchmark's OWN data and shape (m, p, K, 2000-cell blocks): the same sample through
    sharp_run on host buffers; labels after the host glue (R/SHARP.R:816-832) must be identical, the averaged
    projection viE within the 1e-5 contract."""
    from sharp_b200 import RunParams, hc_params
    import synth
    n, csc = sample_of(part, sample_cells)
    rm = ctx.upload_rm(rms)
    try:
        prm = RunParams(1, 1, 2, -1, 2000, 0, 0, 0, hc_params(max_n=max(40, -(-n // 5000))), 2 if wl["exp_type"] == "UMI" else 0, 1e6)
        got = ctx.run(rm, prm, m=wl["m"], n=n, csc=csc, reind=reind, want_x0=False)
    finally:
        rm.close()
    gl = got["labels"].copy()
    if n > 10000:
        vals, cnt = np.unique(gl, return_counts=True)
        small = vals[cnt < 10]
        if len(small):
            gl[np.isin(gl, small)] = small.min()
    pred = first_appearance(gl)
    rel = float(np.max(np.abs(got["viE"] - ref["viE"])) / np.max(np.abs(ref["viE"])))
    return {"cells": int(n), "blocks": len(block_sizes(n)), "m": wl["m"], "p": int(p), "K": wl["K"],
            "ari": synth.ari(pred, ref["pred_clusters"]), "equal": bool(np.array_equal(pred, ref["pred_clusters"])),
            "n_clusters": int(ref["N.pred_cluster"]), "proj_max_rel": rel,
            "what": "sharp_run (C ABI, host buffers) vs the oracle on the first cells of part 1 of THIS workload; "
                    "labels after merge + first-appearance relabel, proj = viE (mean of the K projections)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure R and this image has
    no R, so the arm is the oracle port (kind "port"), all host threads.  One step = SHARP_large on ONE part of the
    workload (blocks, per-block wMetaC, the part-level sMetaC, merge, relabel); when K + W whole parts would not fit
    the arm's time budget the sample shrinks to the largest number of whole blocks that does (stated in `sample`).
    Nothing of the product is loaded here: the ranM matrices and the shuffle come from the pure-Python restatement of
    R's RNG (sharp_b200/rrng.py, numpy), not from libsharpb200.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun pins OMP_NUM_THREADS=1 for its workers; this arm is the host-side reference and runs alone on rank 0,
        # with all the host threads (must be set before the OpenMP runtime is loaded, i.e. before torch is imported)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import torch
    from sharp_b200.rrng import r_sample_perm, ranM2
    wl = workload(args.workload)
    p = math.ceil(math.log2(sum(wl["parts"])) / 0.04)
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    lam = type_profiles(torch, dev, wl["m"], wl["types"], wl["nnz_per_cell"])
    part = gen_part(torch, dev, lam, wl["parts"][0], 0, False)
    rms = [ranM2(wl["m"], p, 50 + SEED + k) for k in range(1, wl["K"] + 1)]
    # calibration: two blocks, then the largest sample (whole blocks, at most the whole part) that fits the budget
    t_cal = cpu_baseline_run(wl, part, p, 4000, rms, np.asarray(r_sample_perm(4000, 50)))[2]
    per_cell = t_cal / 4000
    nsteps = args.warmup + args.steps
    n = wl["parts"][0] if args.ref_sample <= 0 else min(wl["parts"][0], args.ref_sample)
    fit = int(args.ref_budget_s / (nsteps * per_cell * 1.15)) // 2000 * 2000
    n = max(4000, min(n, fit))
    reind = np.asarray(r_sample_perm(n, 50))
    times, threads = [], 1
    for s in range(nsteps):
        cps, threads, dt, _ = cpu_baseline_run(wl, part, p, n, rms, reind)
        if s >= args.warmup:
            times.append(dt)
    val = n * len(times) / sum(times)
    whole = n == wl["parts"][0]
    sample = (f"{'one whole part' if whole else 'first ' + str(n) + ' cells of part 1'} per step ({n} cells: SHARP_large, "
              f"{len(block_sizes(n))} blocks x K={wl['K']}, per-block wMetaC, part-level sMetaC, merge, relabel); "
              f"{threads} host threads; the job's other parts and the global sMetaC are not in the sample")
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "cells/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": wl["name"], "sample": sample, "cpu_sample_cells": int(n), "same_config": False},
                      "cpu_baseline": {"value": val, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_single(args, wl):
    """configs 2 and 3: ONE expression matrix through SHARP() (-> SHARP_large).  N > 1: every rank makes the same call with
    the communicator and the cell blocks are dealt over the ranks (sharp_run_params.shard)."""
    import hashlib
    import torch
    import synth
    from sharp_b200 import api
    from sharp_b200 import comm as sharp_comm
    from sharp_b200.rrng import r_sample_perm, ranM2
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    api.set_devices(local)
    ctx = api.get_context(local)
    comm = sharp_comm.init_from_env(ctx)
    rank, world = (comm.rank, comm.world) if comm else (0, 1)
    m, n, K = wl["m"], wl["parts"][0], wl["K"]
    p = math.ceil(math.log2(n) / 0.04)
    t0 = time.time()
    lam = type_profiles(torch, dev, m, wl["types"], wl["nnz_per_cell"])
    part = gen_part(torch, dev, lam, n, 0, True)            # every rank generates the same matrix (seeded)
    del lam
    truth = part["types"]
    if wl["dense"]:  # TPM-like: columns scaled to 1e6, dense column-major fp64 in pinned host memory
        cp, ri, xv = part["p"], part["i"], part["x"]
        hd = torch.zeros((n, m), dtype=torch.float64, pin_memory=True)  # row c = column c of the genes x cells matrix
        x = hd.numpy()
        cols = np.repeat(np.arange(n), np.diff(cp))
        x[cols, ri] = xv
        x /= x.sum(1, keepdims=True) / 1e6
        host = x.T                                           # (m, n) Fortran-ordered view of the pinned buffer
        in_bytes = x.nbytes
        dev_expr = ctx.upload_expr(m, n, dense=host)
    else:
        host = (part["p"], part["i"], part["x"], (m, n))
        in_bytes = part["p"].nbytes + part["i"].nbytes + part["x"].nbytes
        dev_expr = ctx.upload_expr(m, n, csc=host[:3])
    torch.cuda.empty_cache()
    t_gen = time.time() - t0
    kw = dict(exp_type=wl["exp_type"], rN_seed=SEED, logflag=False, ctx=ctx, comm=comm, forview=world == 1)
    if K != 5:
        kw["ensize_K"] = K

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if comm:
            comm.barrier()

    def timed(data, steps, profile):
        barrier()
        if profile:
            ctx.prof_reset()
            ctx.prof_enable(True)
        l0 = ctx.launch_count()
        ctx.timer_start()
        res = None
        for _ in range(steps):
            res = api.SHARP(data, **kw)
        ms = ctx.timer_stop_ms()
        barrier()
        if profile:
            ctx.prof_enable(False)
        return (comm.max_float(ms) if comm else ms), res, ctx.launch_count() - l0

    for _ in range(args.warmup):
        api.SHARP(dev_expr, **kw)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, res, launches = timed(dev_expr, args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    prof = ctx.prof_get()
    dev_expr.close()
    h2d_gbs = h2d_bandwidth(torch, dev)
    api.SHARP(host, **kw)
    ms_e2e, res_e2e, _ = timed(host, args.steps, False)
    same = bool(np.array_equal(res["pred_clusters"], res_e2e["pred_clusters"]))
    label_hash = hashlib.sha1(np.ascontiguousarray(res["pred_clusters"], dtype=np.int32).tobytes()).hexdigest()[:16]
    if comm and len(set(comm.allgather_bytes(label_hash.encode()))) != 1:
        raise SystemExit("the ranks returned different label vectors")
    # ---- parity + CPU baseline (N = 1): the oracle on the whole matrix (config 2) or its first 8000 cells (config 3) ----
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        import orc
        ns = n if wl["dense"] else min(n, 8000)
        rms = [ranM2(m, p if ns == n else math.ceil(math.log2(ns) / 0.04), 50 + SEED + k) for k in range(1, K + 1)]
        ps = rms[0]["Dim"][1]
        reind = np.asarray(r_sample_perm(ns, 50))
        oprm = orc.SharpParams(1, 1, K, ps, 2000, 0, 0, 0, orc.hc_params(max_n=max(40, -(-ns // 5000))), 2, -1)
        if wl["dense"]:
            sub, okw, colsum = host, dict(dense=np.asfortranarray(host)), None
        else:
            _, csc = sample_of(part, ns)
            sub, okw = csc + ((m, ns),), dict(csc=csc)
            colsum = np.add.reduceat(csc[2], csc[0][:-1])
        t1 = time.time()
        ref = orc.sharp(m, ns, rms, oprm, colsum=colsum, reind=reind, want_x0=False, **okw)
        dt = time.time() - t1
        got = res if ns == n else api.SHARP(sub, exp_type=wl["exp_type"], rN_seed=SEED, logflag=False, prep=False, ctx=ctx, ensize_K=K)
        rel = float(np.max(np.abs(got["viE"] - ref["viE"])) / np.max(np.abs(ref["viE"])))
        parity = {"cells": int(ns), "m": m, "p": int(ps), "K": K, "ari": synth.ari(got["pred_clusters"], ref["pred_clusters"]),
                  "equal": bool(np.array_equal(got["pred_clusters"], ref["pred_clusters"])), "proj_max_rel": rel,
                  "what": "SHARP() (C ABI) vs the oracle on " + ("the whole matrix" if ns == n else f"the first {ns} cells as their own SHARP_large job")}
        cpu = {"value": ns / dt, "unit": "cells/s", "cores": orc.num_threads(), "kind": "port", "seconds": dt,
               "sample": ("the whole configuration" if ns == n else f"first {ns} cells (4 blocks x K={K})") + " through the oracle port (OpenMP over (member, block) tasks)"}
        if not parity["equal"] or not rel <= 1e-5:
            print(json.dumps({"parity": parity}), file=sys.stderr)
            raise SystemExit("PARITY FAILURE: the GPU path and the oracle disagree on the benchmark's own data")
    if comm:
        comm.barrier()
    if rank != 0:
        if comm:
            comm.close()
        return
    pk = peaks()
    ldu = (p + 15) // 16 * 16
    nloc = n / world
    blocks = block_sizes(n)
    bl = [b for i, b in enumerate(blocks) if world == 1 or True]
    share = 1.0 / world
    in_cell = (m * 8) if wl["dense"] else (in_bytes / n)
    work = {"rp_project": ("hbm", nloc * (in_cell + K * p * 8)),
            "corrdist": ("tensor", sum(1.0 * b * (b + 1) * ldu for b in bl) * K * share),
            "hclust": ("hbm", sum(b * b * 8 + (b - 1) * b * 8 * 3 for b in bl) * K * share),
            "sweep_nested": ("hbm", sum(b * b * 8 + b * ldu * 8 for b in bl) * K * share),
            "unit_rows": ("hbm", nloc * K * (p + ldu) * 8)}
    kernels, dgemm = {}, None
    total_ms = sum(v[0] for v in prof.values())
    for name, (kms, kn) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        ent = {"ms_per_step": kms / args.steps, "launches_per_step": kn / args.steps, "share_of_kernel_time": kms / total_ms}
        if name in work:
            bound, amount = work[name]
            per_s = amount * args.steps / (kms * 1e-3)
            if bound == "hbm":
                ent.update(bound="hbm", achieved=per_s / 1e9, peak=pk["hbm_gbs"], unit="GB/s", frac=per_s / 1e9 / pk["hbm_gbs"])
            else:
                dgemm = dgemm or dgemm_peak_tflops(torch, dev)
                ent.update(bound="tensor", achieved=per_s / 1e12, peak=dgemm, unit="TFLOP/s", frac=per_s / 1e12 / dgemm,
                           peak_source="cuBLAS DGEMM 4096^3 measured in this run (fp64 DMMA pipe)")
        kernels[name] = ent
    dom = next((k for k in kernels if "bound" in kernels[k]), None)
    TR = load_traffic()
    roofline = None
    if dom:
        e = kernels[dom]
        roofline = {"kernel": dom, "bound": e["bound"], "achieved": e["achieved"], "peak": e["peak"], "unit": e["unit"], "frac": e["frac"],
                    "traffic": None, "peak_source": e.get("peak_source", pk["source"]),
                    "ms_per_launch": e["ms_per_step"] / max(e["launches_per_step"], 1e-9),
                    "timing": "CUDA events around every launch during the timed steps (one stream: no two kernels overlap)",
                    "traffic_note": "the committed ncu capture is of the config-4 launches (profiles/traffic.json: " + str(TR.get("source")) + ")"}
        rp = kernels.get("rp_project")
        if rp and "frac" in rp:
            roofline["rp_project"] = {k: rp[k] for k in ("bound", "achieved", "peak", "unit", "frac")}
    line = {"metric": wl["metric"], "value": n * args.steps / (ms * 1e-3), "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "cells": n, "genes": m, "K": K, "p": p, "input_bytes": int(in_bytes),
                       "partition": "one GPU" if world == 1 else f"{len(blocks)} cell blocks dealt over {world} ranks (contiguous ranges); NCCL allgather of block-level labels and enE rows; every rank returns the full labels",
                       "l2": "flushed between steps by the run itself: every step streams %.1f GB of distance matrices" % (sum(b * b * 8 for b in blocks) * K / 1e9),
                       "generation_s": t_gen},
            "clocks": clocks, "gpu_launches": int(launches / args.steps),
            "e2e": {"value": n * args.steps / (ms_e2e * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": int(in_bytes),
                    "d2h_bytes_per_step": int(n * 4 + (n * p * 8 + n * 41 * 8 if world == 1 else n * 4)), "ms_per_step": ms_e2e / args.steps,
                    "labels_equal_device_resident_run": same, "h2d_gbs_measured": h2d_gbs},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "parity": parity,
            "result": {"N.pred_cluster": int(res["N.pred_cluster"]), "label_sha1_16": label_hash,
                       "ari_vs_planted_types": synth.ari(res["pred_clusters"], truth), "planted_types": wl["types"],
                       "ranks_agree": True if comm else None}}
    print(json.dumps(line))
    if comm:
        comm.close()


def run_streamed(args, wl):
    """config 5 (bounded): SHARP_unlimited3 over a directory of SHCSC001 files -- the reference reads one .rds per part
    (R/SHARP_unlimited3.R:103-131); here a reader thread fills pinned buffers with pread() while the previous batch of
    parts is copied and clustered.  Both numbers of the line are end to end FROM THE FILES (page cache warm after the
    warm-up steps); `value` and `e2e` are the same measurement, there is no device-resident variant of this path."""
    import shutil
    import torch
    from sharp_b200 import api, io as sio
    from sharp_b200 import comm as sharp_comm
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    api.set_devices(local)
    ctx = api.get_context(local)
    comm = sharp_comm.init_from_env(ctx)
    rank, world = (comm.rank, comm.world) if comm else (0, 1)
    if args.parts:
        wl["parts"] = wl["parts"][:args.parts]
    m, sizes = wl["m"], wl["parts"]
    ncells = int(sum(sizes))
    root = os.environ.get("SHARP_BENCH_TMP") or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    d = os.path.join(root, "sharp_b200_cfg5")
    t0 = time.time()
    nbytes = 0
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        lam = type_profiles(torch, dev, m, wl["types"], wl["nnz_per_cell"])
        for i, n in enumerate(sizes):
            part = gen_part(torch, dev, lam, n, i, False)
            path = os.path.join(d, f"part{i + 1}.csc")
            sio.write_csc(path, m, n, part["p"], part["i"], part["x"])
            nbytes += os.path.getsize(path)
            del part
        del lam
        torch.cuda.empty_cache()
    if comm:
        comm.barrier()
    t_gen = time.time() - t0
    nd = {"dir": d, "ncells": ncells, "ngenes": m}
    kw = dict(viewflag=False, ensize_K=wl["K"], rN_seed=SEED, exp_type=wl["exp_type"], ctx=ctx, comm=comm)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if comm:
            comm.barrier()

    try:
        for _ in range(args.warmup):
            api.SHARP_unlimited3(nd, **kw)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        ctx.prof_reset()
        l0 = ctx.launch_count()
        ctx.timer_start()
        walls = []
        for _ in range(args.steps):
            w0 = time.perf_counter()
            res = api.SHARP_unlimited3(nd, **kw)
            walls.append(round(1e3 * (time.perf_counter() - w0), 1))
        ms = ctx.timer_stop_ms()
        barrier()
        ms = comm.max_float(ms) if comm else ms
        launches = ctx.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
    finally:
        if comm:
            comm.barrier()
        if rank == 0:
            shutil.rmtree(d, ignore_errors=True)
    import hashlib
    label_hash = hashlib.sha1(np.ascontiguousarray(res["pred_clusters"], dtype=np.int32).tobytes()).hexdigest()[:16]
    if comm and len(set(comm.allgather_bytes(label_hash.encode()))) != 1:
        raise SystemExit("the ranks returned different label vectors")
    if rank == 0:
        value = ncells * args.steps / (ms * 1e-3)
        line = {"metric": wl["metric"], "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": wl["name"], "cells": ncells, "genes": m, "files": len(sizes), "file_bytes": int(nbytes),
                           "file_dir": root, "K": wl["K"], "p": int(math.ceil(math.log2(ncells) / 0.04)),
                           "l2": "inputs larger than L2 (%.1f GB of files per step)" % (nbytes / 1e9), "generation_s": t_gen},
                "step_wall_ms": walls, "clocks": clocks, "gpu_launches": int(launches / args.steps),
                "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(ncells * 4),
                        "ms_per_step": ms / args.steps, "note": "same measurement as value: this path starts at the files"},
                "roofline": None, "cpu_baseline": None, "parity": None,
                "result": {"N.pred_clusters": int(res["N.pred_clusters"]), "label_sha1_16": label_hash,
                           "ranks_agree": True if comm else None}}
        print(json.dumps(line))
    if comm:
        comm.close()


def run_ours(args):
    if workload(args.workload).get("single"):
        return run_single(args, workload(args.workload))
    if workload(args.workload).get("streamed"):
        return run_streamed(args, workload(args.workload))
    import torch
    import sharp_b200
    from sharp_b200 import api
    from sharp_b200 import comm as sharp_comm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sharp_b200 has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    api.set_devices(local)
    ctx = api.get_context(local)
    # N > 1: the communicator is NCCL behind the C ABI (sharp_comm_*); torch is only used to generate the synthetic data
    comm = sharp_comm.init_from_env(ctx)
    rank, world = (comm.rank, comm.world) if comm else (0, 1)
    # one process per GPU: keep this rank's host buffers on the GPU's NUMA node (what numactl would do for a launcher
    # that knows the topology); at N = 1 the process keeps all cores (the CPU baseline uses them)
    affinity = ctx.bind_host() if (world > 1 and not args.no_bind) else ""
    wl = workload(args.workload)
    if args.parts:
        wl["parts"] = wl["parts"][:args.parts]
    m, sizes = wl["m"], wl["parts"]
    ncells = int(sum(sizes))
    p = math.ceil(math.log2(ncells) / 0.04)
    # whole parts are dealt round-robin while they divide evenly; the left-over parts are block-sharded over all ranks
    # (every rank holds their data and clusters a share of their blocks; api.SHARP_unlimited makes the same split)
    nwhole = (len(sizes) // world) * world if (world > 1 and not args.no_fused and not args.no_shard and
                                                all(-(-sizes[i] // 2000) >= world for i in range((len(sizes) // world) * world, len(sizes)))) \
        else len(sizes)
    shared = list(range(nwhole, len(sizes)))
    mine = [i for i in range(nwhole) if i % world == rank] + shared

    t0 = time.time()
    lam = type_profiles(torch, dev, m, wl["types"], wl["nnz_per_cell"])
    host_parts = {i: gen_part(torch, dev, lam, sizes[i], i, True) for i in mine}
    del lam
    torch.cuda.empty_cache()
    t_gen = time.time() - t0
    nnz_all = [0] * len(sizes)
    for i in mine:
        nnz_all[i] = int(host_parts[i]["p"][-1])
    own = [i for i in mine if i not in shared or rank == 0]   # who reports a part in the exchanges below
    if comm:
        nnz_all = [int(a[0]) for a in comm.allgather_parts({i: np.array([nnz_all[i]], dtype=np.int64) for i in own}, len(sizes))]

    api._fused_parts = not args.no_fused
    api._fused_group, api._fused_lanes = args.group, args.lanes
    if args.budget_gb:
        ctx.set_block_budget(args.budget_gb)
    if args.rp_variant:
        ctx.set_rp_variant(args.rp_variant)
    ctxs = api.stream_contexts(args.streams, local) if args.no_fused else [ctx]
    kw = dict(n_streams=args.streams, viewflag=False, ensize_K=wl["K"], rN_seed=SEED, exp_type=wl["exp_type"], ctx=ctx, comm=comm)

    def placeholder(i):  # parts owned by other ranks: only their shape is needed
        return api.Expression(m, sizes[i])

    # ---- device-resident run (value) ----
    dev_exprs = {i: ctx.upload_expr(m, sizes[i], csc=(host_parts[i]["p"], host_parts[i]["i"], host_parts[i]["x"])) for i in mine}
    ctx.sync()
    dev_list = [api.Expression.wrap(dev_exprs[i]) if i in dev_exprs else placeholder(i) for i in range(len(sizes))]
    host_list = [host_parts[i] if i in host_parts else placeholder(i) for i in range(len(sizes))]

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if comm:
            comm.barrier()

    walls = []  # host wall clock of every call of the last timed() (diagnostic: the steps are timed on the device)

    def timed(parts_list, steps, profile):
        barrier()
        if profile:
            for c in ctxs:
                c.prof_reset()
                c.prof_enable(True)
        l0 = sum(c.launch_count() for c in ctxs)
        ctx.timer_start()
        res = None
        walls.clear()
        for _ in range(steps):
            w0 = time.perf_counter()
            res = api.SHARP_unlimited(parts_list, **kw)
            walls.append(round(1e3 * (time.perf_counter() - w0), 1))
        ms = ctx.timer_stop_ms()
        barrier()
        if profile:
            for c in ctxs:
                c.prof_enable(False)
        ms = comm.max_float(ms) if comm else ms
        return ms, res, sum(c.launch_count() for c in ctxs) - l0

    for _ in range(args.warmup):
        api.SHARP_unlimited(dev_list, **kw)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, res, launches = timed(dev_list, args.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    walls_dev = list(walls)

    def collect():
        out = {}
        for c in ctxs:
            for name, (kms, kn) in c.prof_get().items():
                a = out.get(name, (0.0, 0))
                out[name] = (a[0] + kms, a[1] + kn)
        return out

    prof_conc = collect()  # event brackets DURING the timed steps: kernels of different lanes overlap (sum > step)
    # one more step with every launch on ONE stream: the same launches and grids, none overlapping, so each event
    # bracket is the kernel's own duration (the roofline object uses these; shares comparable with the ncu list)
    prof, ms_serial = prof_conc, None
    if not args.no_fused and not args.no_serial_profile:
        ctx.set_serial(True)
        ms_serial, res_serial, _ = timed(dev_list, 1, True)
        ctx.set_serial(False)
        prof = collect()
        if not np.array_equal(res["pred_clusters"], res_serial["pred_clusters"]):
            raise SystemExit("serialised profile step produced different labels")
    value = ncells * args.steps / (ms * 1e-3)

    # ---- end to end on host buffers (e2e) ----
    for ex in dev_exprs.values():
        ex.close()
    h2d_gbs = h2d_bandwidth(torch, dev)
    api.SHARP_unlimited(host_list, **kw)  # warm-up of the host path
    e2e_steps = max(1, args.steps)
    ms_e2e, res_e2e, _ = timed(host_list, e2e_steps, False)
    walls_e2e = list(walls)
    e2e_value = ncells * e2e_steps / (ms_e2e * 1e-3)
    same = bool(np.array_equal(res["pred_clusters"], res_e2e["pred_clusters"]))
    h2d = sum(host_parts[i]["p"].nbytes + host_parts[i]["i"].nbytes + host_parts[i]["x"].nbytes for i in mine)
    d2h = sum(sizes[i] * 4 for i in mine)
    if comm:  # whole-job bytes per step, like `value`: every rank copies its own parts and ALL the block-sharded ones
        import struct
        h2d = sum(struct.unpack("<q", b)[0] for b in comm.allgather_bytes(struct.pack("<q", int(h2d))))
        d2h = sum(struct.unpack("<q", b)[0] for b in comm.allgather_bytes(struct.pack("<q", int(d2h))))

    # ---- what came out: a hash of the label vector (must be the same on every rank and for every N), and the ARI
    # ---- against the planted cell types of the synthetic data
    import hashlib
    import synth
    label_hash = hashlib.sha1(np.ascontiguousarray(res["pred_clusters"], dtype=np.int32).tobytes()).hexdigest()[:16]
    types_all = {i: host_parts[i]["types"] for i in own}
    if comm:
        types_all = comm.allgather_parts(types_all, len(sizes))
        if len(set(comm.allgather_bytes(label_hash.encode()))) != 1:
            raise SystemExit("the ranks returned different label vectors")
    else:
        types_all = [types_all[i] for i in range(len(sizes))]
    truth = np.concatenate(list(types_all))
    result = {"N.pred_clusters": int(res.get("N.pred_clusters", res.get("N.pred_cluster", 0))), "label_sha1_16": label_hash,
              "ari_vs_planted_types": synth.ari(res["pred_clusters"], truth), "planted_types": wl["types"],
              "ranks_agree": True if comm else None}
    if comm:  # every exchange is done
        comm.barrier()
    if rank != 0:
        if comm:
            comm.close()
        return
    # ---- roofline of the dominant kernel (live CUDA-event profile over the timed steps, this rank's share) ----
    pk = peaks()
    frac = lambda i: (1.0 / world if i in shared else 1.0)   # a rank clusters its share of the blocks of a sharded part
    my_nnz = [int(nnz_all[i] * frac(i)) for i in mine]
    my_sizes = [int(sizes[i] * frac(i)) for i in mine]
    work = algorithmic_work(wl, my_nnz, my_sizes, p)
    kernels = {}
    total_kernel_ms = sum(v[0] for v in prof.values())
    dgemm = None
    psteps = 1 if ms_serial is not None else args.steps
    for name, (kms, kn) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        ent = {"ms_per_step": kms / psteps, "launches_per_step": kn / psteps, "share_of_kernel_time": kms / total_kernel_ms}
        if ms_serial is not None and name in prof_conc:
            ent["ms_per_step_in_timed_region"] = prof_conc[name][0] / args.steps  # overlapped with other lanes' kernels
        if name in work:
            bound, amount = work[name]
            per_s = amount * psteps / (kms * 1e-3)
            if bound == "hbm":
                ent.update(bound="hbm", achieved=per_s / 1e9, peak=pk["hbm_gbs"], unit="GB/s", frac=per_s / 1e9 / pk["hbm_gbs"])
            else:
                if dgemm is None:
                    dgemm = dgemm_peak_tflops(torch, dev)
                ent.update(bound="tensor", achieved=per_s / 1e12, peak=dgemm, unit="TFLOP/s", frac=per_s / 1e12 / dgemm,
                           peak_source="cuBLAS DGEMM 4096^3 measured in this run (fp64 DMMA pipe)")
        kernels[name] = ent
    dom = next((k for k in kernels if "bound" in kernels[k]), None)
    roofline = None
    TR = load_traffic()

    def traffic_of(name):
        e = TR.get("kernels", {}).get(name)
        return None if e is None else e.get("dram_bytes_per_launch")

    if dom:
        e = kernels[dom]
        roofline = {"kernel": dom, "bound": e["bound"], "achieved": e["achieved"], "peak": e["peak"], "unit": e["unit"],
                    "frac": e["frac"], "traffic": traffic_of(dom), "peak_source": e.get("peak_source", pk["source"]),
                    "ms_per_launch": e["ms_per_step"] / max(e["launches_per_step"], 1e-9),
                    "timing": ("CUDA events around every launch of one extra step of the same workload enqueued on ONE stream "
                               "(kernels of different parts overlap during the timed steps; their overlapped brackets are "
                               "ms_per_step_in_timed_region)" if ms_serial is not None else
                               "CUDA events around every launch during the timed steps"),
                    "serial_step_ms": ms_serial, "traffic_source": TR.get("source")}
        if dom == "hclust":  # the class's own yardstick (4 n^2 x 8 B per problem) and the bare one-read-of-D figure
            one_read = sum(b * b * 8 for n_ in my_sizes for b in block_sizes(n_)) * wl["K"]
            roofline["frac_one_read_of_D"] = one_read * psteps / (prof[dom][0] * 1e-3) / 1e9 / pk["hbm_gbs"]
        rp = kernels.get("rp_project")
        if rp and "frac" in rp:
            roofline["rp_project"] = {"bound": "hbm", "achieved": rp["achieved"], "peak": rp["peak"], "unit": "GB/s",
                                      "frac": rp["frac"], "traffic": traffic_of("rp_project"),
                                      "ms_per_launch": rp["ms_per_step"] / max(rp["launches_per_step"], 1e-9)}

    # ---- CPU baseline: the oracle port on a bounded sample of the same workload, this box's host cores; its labels are
    # ---- the parity check of the GPU path at the benchmark's own shape ----
    cpu, parity = None, None
    if not args.no_cpu_baseline and world == 1:  # N = 1 only (torchrun pins OMP_NUM_THREADS=1; the baseline is a host figure)
        from sharp_b200.rrng import r_sample_perm, ranM2
        part0 = host_parts[mine[0]]
        n_s = part0["Dim"][1] if args.cpu_sample <= 0 else min(args.cpu_sample, part0["Dim"][1])
        rms = [ranM2(m, p, 50 + SEED + k) for k in range(1, wl["K"] + 1)]
        reind = np.asarray(r_sample_perm(n_s, 50))
        cps, threads, dt, ref = cpu_baseline_run(wl, part0, p, n_s, rms, reind)
        cpu = {"value": cps, "unit": "cells/s", "cores": threads, "kind": "port", "seconds": dt,
               "sample": f"first {n_s} cells of part {mine[0] + 1}: SHARP_large, {len(block_sizes(n_s))} blocks x K={wl['K']}, "
                         f"per-block wMetaC, part-level sMetaC, same ranM matrices (oracle/, OpenMP over (member, block) tasks)"}
        parity = parity_check(ctx, wl, part0, p, n_s, rms, reind, ref)
        if not parity["equal"] or not parity["proj_max_rel"] <= 1e-5:
            print(json.dumps({"parity": parity}), file=sys.stderr)
            raise SystemExit("PARITY FAILURE: the GPU path and the oracle disagree on the benchmark's own data")

    line = {"metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "cells": ncells, "genes": m, "parts": len(sizes), "K": wl["K"], "p": p,
                       "nnz": int(sum(nnz_all)), "partition": (f"parts round-robin over {world} rank(s); part-by-part, {args.streams} streams per GPU" if args.no_fused else
                                     f"{nwhole} whole parts round-robin over {world} rank(s)"
                                     + (f", the cell blocks of the other {len(shared)} dealt over all ranks (NCCL allgather of block-level labels and enE rows)" if shared else "")
                                     + f"; fused loop over parts (sharp_run_parts), group={args.group or 'default'}, lanes={args.lanes or 'default'}"),
                       "l2": "inputs larger than L2 (CSC input %.1f GB per step)" % (sum(nnz_all) * 12 / 1e9),
                       "generation_s": t_gen, "host_affinity_rank0": affinity or None},
            "step_wall_ms": walls_dev, "clocks": clocks, "gpu_launches": int(launches / args.steps),
            "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / e2e_steps, "labels_equal_device_resident_run": same, "step_wall_ms": walls_e2e,
                    "h2d_gbs_measured": h2d_gbs, "h2d_floor_ms_per_step": (h2d / max(world, 1)) / (h2d_gbs * 1e9) * 1e3},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "parity": parity,
            "result": result}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--no-bind", action="store_true", help="N > 1: do not bind the ranks to the CPUs next to their GPUs")
    ap.add_argument("--streams", type=int, default=8, help="contexts (CUDA streams) per GPU working on different parts")
    ap.add_argument("--parts", type=int, default=0, help="development: only the first PARTS parts")
    ap.add_argument("--group", type=int, default=0, help="parts per group of the fused loop over parts (0 = library default)")
    ap.add_argument("--lanes", type=int, default=0, help="groups in flight (0 = library default)")
    ap.add_argument("--budget-gb", type=int, default=0, help="distance-matrix workspace cap per context in GB (0 = library default)")
    ap.add_argument("--no-fused", action="store_true", help="part-by-part path (one sharp_run per part, --streams host threads)")
    ap.add_argument("--rp-variant", type=int, default=0, help="projection kernel variant (sharp_ctx_set_rp_variant): 0 default, 2 the r1 kernel, 3 TMA-staged")
    ap.add_argument("--no-shard", action="store_true", help="N > 1: deal ALL parts round-robin (no block-sharded left-over parts)")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="cells of the CPU baseline / parity sample (0 = one whole part)")
    ap.add_argument("--ref-sample", type=int, default=0, help="--impl reference: cells per step (0 = one whole part, budget permitting)")
    ap.add_argument("--ref-budget-s", type=float, default=270.0, help="--impl reference: time budget of the whole K + W run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-serial-profile", action="store_true", help="skip the extra one-stream step behind the roofline object")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
