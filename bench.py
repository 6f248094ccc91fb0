#!/usr/bin/env python
"""bench.py — phase-space cell-updates/s of the KitAMR time step (slope! -> flux! -> iterate!) on B200.

Contract:
  python bench.py --gpus N --steps K --warmup W            our arm: libkamr.so through the C-ABI
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the CPU restatement of the
                                                           reference's step on the box's host cores
One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over the whole workload:
kamr_step == slope! + flux! + iterate! of src/Solver/Solver.jl:65-67.

  workload   default S4 = example/sphere (3-D, 3D1F, immersed sphere): the case BASELINE.json's north_star names for
             the 1/2/4/8-GPU scaling, one sphere per GPU (weak scaling), the same family at every N.  The other named
             cases (S1 Riemann, S2 cylinder, S3 airfoil, S5 X38-like) are measured in the same run at N = 1 and
             reported in the `workloads` object of the same line.
  value      whole-job phase-space cell-updates/s, state resident in HBM, CUDA-event timed on the
             library's stream, max over ranks.
  e2e        the same metric through the host-facing call sequence of one adapt window, re-flatten INCLUDED:
             kamr_upload_topology (the amr_recover! event), kamr_upload_state (pinned host -> device), K x kamr_step
             with the residual read back to the host every step, kamr_download_state; wall clock, max over ranks.
  e2e_strict upload + 1 step + download EVERY step (a host that keeps no state resident).
  roofline   B_alg (130 B/update 2D2F, 106 B 3D1F) x updates of one step / device time
             of the step, against the measured HBM copy bandwidth of MEASURED_PEAKS.json.
  parity     computed outside every timed region: N = 1: the device after `steps` kamr_step calls against the CPU
             oracle on the FULL workload, element by element; N > 1: every rank's local f and w against a single-rank
             device run of the whole forest on rank 0 (no halo in that run).  Non-zero exit above 1e-12 per step.
  cpu_baseline  the oracle (oracle/kamr_oracle.c, a restatement of the Julia step; the reference itself
             needs julia+libp4est+MPI which this image lacks) built -O3 -march=native on the box, one process per core.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "phase-space cell-updates/sec"
UNIT = "cell-updates/s"
B_ALG = {(2, 2): 130.0, (3, 1): 106.0}  # 8*(5*NDF + 2*DIM + 2) + 2
FALLBACK_HBM_GBS = 6650.0               # /opt/skills/guides/B200_PROFILING.md
PARITY_TOL = 1e-12                      # north_star: relative L2 per step
OTHER_WORKLOADS = ["S1", "S1caidvm", "S2ib", "S2ib-big", "S3", "S3-big", "S5"]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_case(args, world, name=None):
    from kitamr_jl_b200.synth import cases
    fn = cases.WORKLOADS[name or args.workload]
    case = fn(copies=world)
    if getattr(args, "impl", "ours") == "ours":
        case.partition_mode = getattr(args, "partition", "cost")
    return case


def n_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


MARCH_NAMES = {0: "CAIDVM_Marching", 1: "CIP_Marching", 2: "Euler"}


def peak_hbm():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "of measured: MEASURED_PEAKS.json hbm_gbs (STREAM-style copy)"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "of fallback: 6650 GB/s, B200_PROFILING.md (MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle, one process per core on a Morton split of the same workload
def _cpu_worker(case, rank, nparts, steps, warmup, barrier, q):
    try:
        from oracle import orc
        mesh = case.rank_mesh(rank, nparts)
        st = case.init_state(mesh)
        cfg = case.config(rank=rank, nranks=nparts)
        dt = case.dt()
        nph = mesh.n_phase_local()
        for _ in range(warmup):
            orc.step(cfg, mesh, st, dt, False)
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.step(cfg, mesh, st, dt, False)
        t1 = time.perf_counter()
        barrier.wait()
        q.put((rank, nph, t1 - t0, None))
    except Exception as e:  # pragma: no cover
        q.put((rank, 0, 0.0, repr(e)))
        try:
            barrier.abort()
        except Exception:
            pass


def cpu_arm(case, steps, warmup, cores, oversplit=1):
    """Runs `cores` oracle processes; each owns one of cores*oversplit Morton chunks (every oversplit-th chunk when
    oversplit > 1: a bounded sample spread over the curve).  Returns (updates/s, sample phase cells, seconds)."""
    nparts = cores * oversplit
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores)
    q = ctx.Queue()
    ranks = [r * oversplit for r in range(cores)]
    procs = [ctx.Process(target=_cpu_worker, args=(case, r, nparts, steps, warmup, barrier, q)) for r in ranks]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    errs = [r[3] for r in res if r[3]]
    if errs:
        raise RuntimeError("cpu arm failed: " + errs[0])
    nph = sum(r[1] for r in res)
    tmax = max(r[2] for r in res)
    return nph * steps / tmax, nph, tmax


def calibrate_oversplit(case, cores, steps_total, budget_s):
    """Pick the sample so that steps_total steps take about budget_s: measured single-core oracle rate on a
    small chunk, extrapolated."""
    from oracle import orc
    nparts = max(cores * 16, 64)
    mesh = case.rank_mesh(nparts // 2, nparts)
    st = case.init_state(mesh)
    cfg = case.config(rank=0, nranks=nparts)
    orc.step(cfg, mesh, st, case.dt(), False)
    t0 = time.perf_counter()
    orc.step(cfg, mesh, st, case.dt(), False)
    rate = mesh.n_phase_local() / max(time.perf_counter() - t0, 1e-6)  # updates/s on one core
    n_of = np.array([g.n for g in case.grids])[case.cell_grid]
    total = float(n_of.sum())
    per_core = total / cores
    t_full = per_core / rate * steps_total
    over = 1
    while t_full / over > budget_s and over < 256:
        over *= 2
    return over, rate


def timed_cpu_leg(case, steps, warmup, budget_s, nph_total):
    """The CPU baseline both arms report: the oracle built for speed (-O3 -march=native on this box) on
    every host core, on a bounded sample of the workload."""
    from oracle import orc
    kind_note = orc.use_fast()
    cores = n_cores()
    over, _ = calibrate_oversplit(case, cores, steps + warmup, budget_s)
    val, nph, secs = cpu_arm(case, steps, warmup, cores, over)
    sample = (f"{nph} of {int(nph_total)} phase cells ({cores} of {cores * over} Morton chunks, spread over the curve), "
              f"{steps} steps after {warmup} warm-up, {secs:.1f}s")
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "note": "oracle restatement of the Julia step (the reference needs julia+libp4est+MPI, absent); " + kind_note +
                    "; one process per core, no halo exchange timed"}, secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import orc
    orc.build()
    world = args.gpus
    case = make_case(args, world)
    nph_total = sum(case.rank_mesh(r, world).n_phase_local() for r in range(world)) if world > 1 \
        else case.rank_mesh().n_phase_local()
    cpu, secs = timed_cpu_leg(case, args.steps, args.warmup, 90.0, nph_total)
    val = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": case.name, "phase_cells": int(nph_total), "dim": case.dim, "ndf": case.ndf,
                   "marching": MARCH_NAMES[case.marching], "flux": "CAIDVM"},
        "cpu_baseline": cpu,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            log("clock sampling unavailable:", e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def pinned_state(st0):
    """pinned host copies of the state (the host side of the boundary)"""
    import torch
    from kitamr_jl_b200.model import HostState
    keep = []
    st = HostState(*[None] * 8)
    for f in st0.__dataclass_fields__:
        a = getattr(st0, f)
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        v = t.numpy()
        v[...] = a
        keep.append(t)
        setattr(st, f, v)
    return st, keep


def mesh_bytes(mesh):
    n = 0
    for f in mesh.__dataclass_fields__:
        a = getattr(mesh, f)
        if isinstance(a, np.ndarray):
            n += a.nbytes
    if mesh.ib is not None:
        for a in vars(mesh.ib).values():
            if isinstance(a, np.ndarray):
                n += a.nbytes
    return n


def single_gpu_workload(args, name, device, stream, steps, unique=False):
    """One of the named cases on one GPU: value, ms/step and roofline fraction (device time, state resident)."""
    import torch
    from kitamr_jl_b200 import api
    from kitamr_jl_b200.model import uniquify_grids
    t0 = time.time()
    case = make_case(args, 1, name)
    mesh = case.rank_mesh()
    if unique:
        mesh = uniquify_grids(mesh)
    st = case.init_state(mesh) if not unique else case.init_state(case.rank_mesh())
    ctx = api.Context(case.config(device=device, stream=stream.cuda_stream))
    try:
        t1 = time.perf_counter()
        ctx.upload_topology(mesh)
        ctx.sync()
        reflat = (time.perf_counter() - t1) * 1e3
        ctx.upload_state(st, aux=False)
        dt = case.dt()
        for _ in range(3):
            ctx.step(dt, False)
        ctx.sync()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(steps):
            ctx.step(dt, False)
        ev1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        nph = mesh.n_phase_local()
        peak, _ = peak_hbm()
        frac = B_ALG[(case.dim, case.ndf)] * nph / (ms * 1e-3) / 1e9 / peak
        out = {"workload": case.name + ("-unique-grids" if unique else ""), "phase_cells": int(nph),
               "cells": int(mesh.n_local), "velocity_grids": int(mesh.n_grid), "marching": MARCH_NAMES[case.marching],
               "steps": steps, "ms_per_step": ms, "value": nph / (ms * 1e-3), "frac": frac,
               "upload_topology_ms": reflat, "device_gb": ctx.stats().device_bytes / 1e9}
        log(f"[workloads] {out['workload']}: {ms:.3f} ms/step, frac {frac:.3f} ({time.time() - t0:.0f}s)")
        return out
    finally:
        ctx.close()


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from kitamr_jl_b200 import abi, api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libkamr has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(abi.LIB_PATH):
        if rank == 0:
            g.build()
        if world > 1:
            dist.barrier()
    abi.load()

    t_gen = time.time()
    case = make_case(args, world)
    mesh = case.rank_mesh(rank, world)
    if args.unique_grids:
        from kitamr_jl_b200.model import uniquify_grids
        st0 = case.init_state(mesh)
        mesh = uniquify_grids(mesh)
    else:
        st0 = case.init_state(mesh)
    D, K, M = case.dim, case.ndf, case.dim + 2
    dt = case.dt()
    nph_local = mesh.n_phase_local()
    log(f"[rank {rank}] workload {case.name}: {mesh.n_local} local cells (+{mesh.n_ghost} ghosts), "
        f"{nph_local} phase cells, {mesh.n_grid} velocity grids, generated in {time.time() - t_gen:.1f}s")

    # The synthetic forest leaves many long-lived Python objects behind; a generation-2 collection that happens to start
    # inside a timed window would walk all of them.  The host this library is written for has no such collector: park
    # the survivors and keep the collector out of the timed windows.
    import gc
    gc.collect()
    gc.freeze()
    gc.disable()

    stream = torch.cuda.Stream()
    cfg = case.config(device=local, rank=rank, nranks=world, stream=stream.cuda_stream)
    ctx = api.Context(cfg)
    if world > 1:
        box = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0])
    t_rf = time.perf_counter()
    ctx.upload_topology(mesh)
    ctx.sync()
    first_topology_ms = (time.perf_counter() - t_rf) * 1e3   # includes one-time costs (allocator, NCCL connect)

    st, keep = pinned_state(st0)
    ctx.upload_state(st, aux=False)
    if world > 1:
        ctx.exchange_df()

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nph_total = reduce_sum(float(nph_local))

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        ctx.step(dt, False)
    barrier()

    # ---- timed region: K steps, state resident, CUDA events on the library's stream
    ctx.profile(True)
    l0 = ctx.stats().kernel_launches
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.step(dt, False)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    step_ms_local = ev0.elapsed_time(ev1)
    ms = reduce_max(step_ms_local)
    launches = ctx.stats().kernel_launches - l0
    prof = ctx.profile_read()
    ctx.profile(False)
    value = nph_total * args.steps / (ms * 1e-3)

    # ---- roofline of the step (rank 0's shard)
    kern_ms = sum(v[1] for v in prof.values())
    rank_kern_ms = [kern_ms / args.steps]
    if world > 1:   # per-rank kernel time: how well the partition balances the device cost
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = kern_ms / args.steps
        dist.all_reduce(t)
        rank_kern_ms = [float(x) for x in t.tolist()]
    b_alg = B_ALG[(D, K)]
    peak, peak_src = peak_hbm()
    # the step's kernels partly overlap (wall kernels and halo traffic on side streams), so the denominator is the device
    # time of the whole timed region (CUDA events around the K steps), not the sum of the per-kernel times
    achieved = b_alg * nph_local * args.steps / (step_ms_local * 1e-3) / 1e9 if step_ms_local > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr.get(case.name.split("-x")[0]) if world == 1 else None   # an ncu capture of the N = 1 run
    except Exception:
        pass
    dom = max(prof.items(), key=lambda kv: kv[1][1])[0] if prof else None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic,
        "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the kernels of one step, "
                          "profiles/traffic.json (a separate capture of the same command, not measured in this run)"
                          if traffic else None,
        "peak_source": peak_src,
        "kernel": "kamr_step = all kernels of one step, device time of the timed region; B_alg/update = %g B "
                  "(DESIGN.md §6)" % b_alg,
        "kernels_ms_sum_per_step": kern_ms / args.steps,
        "dominant_kernel": dom,
        "kernels_ms_per_step": {k: v[1] / args.steps for k, v in prof.items()},
        "kernels_launches_per_step": {k: v[0] / args.steps for k, v in prof.items()},
    }

    # ---- e2e: one adapt window through the host-facing calls (pinned host buffers), re-flatten included
    npts = int(mesh.vs_off()[-1])
    topo_b = mesh_bytes(mesh)
    h2d_win = topo_b + (npts * K + 2 * mesh.n_local * M) * 8
    d2h_win = (npts * K + 2 * mesh.n_local * M) * 8 + args.steps * 2 * M * 8
    out = st
    # Three windows, the median one reported: driver calls (cudaMalloc / cudaFree / small copies) of the re-flatten
    # stall sporadically on this pool's boxes (+0.1 ... 1.5 s on a 40-160 ms re-flatten, about one window in three,
    # in varying phases of kamr_upload_topology; profiles/r02_h_reflatten_outliers.txt); all three are listed.
    wins = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        ctx.upload_topology(mesh)            # the amr_recover! event (Solver/AMR.jl:54): re-flatten
        t_topo = time.perf_counter() - t0
        ctx.upload_state(st, aux=False)
        if world > 1:
            ctx.exchange_df()
        for _ in range(args.steps):
            ctx.step(dt, True)  # residual scalars come back to the host every step
        ctx.download_state(out, abi.DL_DF | abi.DL_W | abi.DL_PRIM)
        ctx.sync()
        wins.append((reduce_max(time.perf_counter() - t0), reduce_max(t_topo * 1e3)))
    t_win, reflatten_ms = sorted(wins)[1]
    e2e_val = nph_total * args.steps / t_win
    # window without the re-flatten (state transfers only), for comparison with round 1
    barrier()
    t0 = time.perf_counter()
    ctx.upload_state(st, aux=False)
    if world > 1:
        ctx.exchange_df()
    for _ in range(args.steps):
        ctx.step(dt, True)
    ctx.download_state(out, abi.DL_DF | abi.DL_W | abi.DL_PRIM)
    ctx.sync()
    t_win2 = reduce_max(time.perf_counter() - t0)
    # strict: upload + step + download every step
    ns = 3
    barrier()
    t0 = time.perf_counter()
    for _ in range(ns):
        ctx.upload_state(st, aux=False)
        if world > 1:
            ctx.exchange_df()
        ctx.step(dt, True)
        ctx.download_state(out, abi.DL_DF | abi.DL_W | abi.DL_PRIM)
    ctx.sync()
    t_strict = reduce_max(time.perf_counter() - t0)
    strict_val = nph_total * ns / t_strict
    stats = ctx.stats()
    dev_gb = stats.device_bytes / 1e9
    halo_b = int(stats.halo_bytes_per_step)
    gc.enable()

    # ---- parity, outside every timed region
    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(args, case, mesh, st0, ctx, rank, world, local, stream, dt)
        except Exception as e:  # pragma: no cover
            parity = {"error": repr(e)}
    ctx.close()

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            cpu, _ = timed_cpu_leg(case, 5, 1, 25.0, nph_total)
        except Exception as e:  # pragma: no cover
            cpu = {"error": repr(e)}

    # ---- the other named workloads at N = 1 (device time, state resident)
    workloads = None
    if world == 1 and not args.no_workloads:
        workloads = {args.workload: {"workload": case.name, "phase_cells": int(nph_total), "ms_per_step": ms / args.steps,
                                     "value": value, "frac": roofline["frac"], "steps": args.steps}}
        for name in OTHER_WORKLOADS:
            if name == args.workload:
                continue
            try:
                workloads[name] = single_gpu_workload(args, name, local, stream, 10)
            except Exception as e:  # pragma: no cover
                workloads[name] = {"error": repr(e)}
        try:   # every cell with its own copy of its velocity grid (statics no longer cache-resident)
            workloads[args.workload + "-unique-grids"] = single_gpu_workload(args, args.workload, local, stream, 10,
                                                                             unique=True)
        except Exception as e:  # pragma: no cover
            workloads[args.workload + "-unique-grids"] = {"error": repr(e)}

    rc = 0
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": case.name, "phase_cells": int(nph_total), "cells_rank0": mesh.n_local,
                       "dim": D, "ndf": K, "marching": MARCH_NAMES[case.marching], "flux": "CAIDVM",
                       "velocity_grids_rank0": int(mesh.n_grid),
                       "l2_policy": "inputs larger than L2: %.2f GB of df/slope state per GPU vs 126 MB L2" % dev_gb,
                       "halo_bytes_per_step_rank0": halo_b,
                       "partition": ("cost-weighted Morton split (device cost model as the partition!(p4est, weight) "
                                     "hook)" if case.partition_mode == "cost" else
                                     "reference partition_weight (vs_num, x2 solid cells)") if world > 1 else "none",
                       "kernels_ms_per_step_by_rank": rank_kern_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_win / args.steps,
                    "d2h_bytes_per_step": d2h_win / args.steps,
                    "pattern": "one adapt window incl. re-flatten: upload_topology, upload_state, %d x step(+residual "
                               "to host), download_state; median of 3 windows" % args.steps,
                    "window_s": t_win, "upload_topology_ms": reflatten_ms,
                    "windows_s": [w[0] for w in wins], "upload_topology_ms_all": [w[1] for w in wins]},
            "e2e_window_without_reflatten": {"value": nph_total * args.steps / t_win2, "unit": UNIT},
            "e2e_strict": {"value": strict_val, "unit": UNIT,
                           "h2d_bytes_per_step": (npts * K + 2 * mesh.n_local * M) * 8,
                           "d2h_bytes_per_step": (npts * K + 2 * mesh.n_local * M) * 8 + 2 * M * 8,
                           "pattern": "upload_state + step + download_state every step"},
            "reflatten": {"upload_topology_ms": reflatten_ms, "first_call_ms": first_topology_ms,
                          "note": "kamr_upload_topology after an adapt / partition event (pair maps, slots, slope "
                                  "stencils, halo plan), host side included; timed inside the e2e window"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if parity is not None:
            line["parity"] = parity
            if parity.get("ok") is False:
                rc = 3
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if workloads is not None:
            line["workloads"] = workloads
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


def parity_check(args, case, mesh, st0, ctx, rank, world, local, stream, dt):
    """See the module docstring.  `ctx` holds this rank's partition and is re-used for the device side."""
    import torch
    import torch.distributed as dist
    from kitamr_jl_b200 import abi, api
    K, M = case.ndf, case.dim + 2
    nl = mesh.n_local
    off = mesh.vs_off()
    n_df = int(off[nl]) * K
    nph = mesh.n_phase_local()
    steps = 2 if (world > 1 or nph <= 5e7) else 1
    # device side: from the initial state, `steps` fused steps.  The raw slopes are part of the state: on meshes with
    # non-dyadic cell sizes the reference's sweep projects some finer neighbours' slopes of the PREVIOUS step (DESIGN.md
    # section 5), and this context has been stepping — so st0's (zero) sdf is uploaded too.
    ctx.upload_state(st0, aux=False)
    ctx._ck(ctx.lib.kamr_upload_aux(ctx.h, st0.sdf.ctypes.data_as(abi.c_f64p), None, None))
    if world > 1:
        ctx.exchange_df()
    for _ in range(steps):
        ctx.step(dt, False)
    mine = ctx.download_state(st0.copy(), abi.DL_DF | abi.DL_W)
    t0 = time.time()
    if world == 1:
        from oracle import orc
        orc.use_parity()
        ref = st0.copy()
        cfg1 = case.config()
        for _ in range(steps):
            orc.step(cfg1, mesh, ref, dt, False)
        num_df = float(np.sum((mine.df[:n_df] - ref.df[:n_df]) ** 2)); den_df = float(np.sum(ref.df[:n_df] ** 2))
        num_w = float(np.sum((mine.w[: nl * M] - ref.w[: nl * M]) ** 2)); den_w = float(np.sum(ref.w[: nl * M] ** 2))
        against = "CPU oracle (oracle/kamr_oracle.c, parity build) on the full workload, element by element"
    else:
        # single-rank device run of the WHOLE forest on rank 0 (every other context is closed first: memory)
        ctx.close()
        dist.barrier()
        ref_df = ref_w = None
        full = None
        err = None
        if rank == 0:
            try:
                full = case.rank_mesh()
                stf = case.init_state(full)
                c1 = api.Context(case.config(device=local, rank=0, nranks=1, stream=stream.cuda_stream))
                try:
                    c1.upload_topology(full)
                    c1.upload_state(stf, aux=False)
                    for _ in range(steps):
                        c1.step(dt, False)
                    stf = c1.download_state(stf, abi.DL_DF | abi.DL_W)
                finally:
                    c1.close()
                ref_df, ref_w = stf.df, stf.w
            except Exception as e:   # e.g. the whole forest does not fit one GPU: every rank must learn it
                err = repr(e)
        okf = torch.tensor([0 if err else 1], dtype=torch.int64, device="cuda")
        dist.broadcast(okf, src=0)
        if int(okf.item()) == 0:
            return {"error": err or "single-rank reference run failed on rank 0", "steps": steps}
        # every rank's local cells are a contiguous range of the whole forest's cell list (Morton partition)
        g0 = int(mesh.global_ids[0]); g1 = int(mesh.global_ids[nl - 1])
        rng = torch.tensor([g0, g1], dtype=torch.int64, device="cuda")
        allr = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(allr, rng)
        my_df = torch.empty(n_df, dtype=torch.float64, device="cuda")
        my_w = torch.empty(nl * M, dtype=torch.float64, device="cuda")
        if rank == 0:
            index_of = {int(gg): i for i, gg in enumerate(full.global_ids[: full.n_local])}
            offf = full.vs_off()
            for r in range(world):
                a, b = index_of[int(allr[r][0])], index_of[int(allr[r][1])] + 1
                sl_df = torch.from_numpy(ref_df[int(offf[a]) * K: int(offf[b]) * K]).cuda()
                sl_w = torch.from_numpy(ref_w[a * M: b * M]).cuda()
                if r == 0:
                    my_df.copy_(sl_df); my_w.copy_(sl_w)
                else:
                    dist.send(sl_df, dst=r); dist.send(sl_w, dst=r)
                del sl_df, sl_w
        else:
            dist.recv(my_df, src=0); dist.recv(my_w, src=0)
        a_df = torch.from_numpy(mine.df[:n_df]).cuda()
        a_w = torch.from_numpy(mine.w[: nl * M]).cuda()
        t = torch.stack([((a_df - my_df) ** 2).sum(), (my_df ** 2).sum(), ((a_w - my_w) ** 2).sum(), (my_w ** 2).sum()])
        dist.all_reduce(t)
        num_df, den_df, num_w, den_w = [float(x) for x in t.tolist()]
        against = ("single-rank device run of the whole forest on rank 0 (no halo), every rank's local cells, "
                   "element by element")
    e_df = (num_df / den_df) ** 0.5
    e_w = (num_w / den_w) ** 0.5
    tol = PARITY_TOL * steps
    return {"rel_l2_df": e_df, "rel_l2_w": e_w, "steps": steps, "tol": tol, "ok": bool(e_df <= tol and e_w <= tol),
            "against": against, "seconds": time.time() - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="S4")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity object")
    ap.add_argument("--no-workloads", action="store_true", help="skip the other named workloads at N = 1")
    ap.add_argument("--unique-grids", action="store_true",
                    help="hand the library one velocity-grid copy per cell (what a host with VS_DYNAMIC_AMR holds)")
    ap.add_argument("--partition", default="cost", choices=["cost", "reference"],
                    help="weights of the Morton split at N>1: the device cost model handed to the reference's "
                         "partition!(p4est, weight) hook, or the reference's own partition_weight")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
