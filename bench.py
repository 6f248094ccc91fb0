#!/usr/bin/env python
"""bench.py — phase-space cell-updates/s of the KitAMR time step (slope! -> flux! -> iterate!) on B200.

Contract:
  python bench.py --gpus N --steps K --warmup W            our arm: libkamr.so through the C-ABI
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the CPU restatement of the
                                                           reference's step on the box's host cores
One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over the whole workload:
kamr_step == slope! + flux! + iterate! of src/Solver/Solver.jl:65-67.

  value      whole-job phase-space cell-updates/s, state resident in HBM, CUDA-event timed on the
             library's stream, max over ranks.
  e2e        the same metric through the host-facing call sequence of one adapt window:
             kamr_upload_state (pinned host -> device), K x kamr_step with the residual read back to the
             host every step, kamr_download_state (device -> host); wall clock, max over ranks.
  e2e_strict upload + 1 step + download EVERY step (a host that keeps no state resident).
  roofline   B_alg (130 B/update 2D2F, 106 B 3D1F) x updates of one step / device time
             of all kernels of the step (per-kernel CUDA events inside the library), against the measured
             HBM copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline  the oracle (oracle/kamr_oracle.c, a restatement of the Julia step; the reference itself
             needs julia+libp4est+MPI which this image lacks) on all host cores, one process per core.
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_case(args, world):
    from kitamr_jl_b200.synth import cases
    fn = cases.WORKLOADS[args.workload]
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
    """Runs `cores` oracle processes; each owns one of cores*oversplit Morton chunks (the first `cores`
    chunks when oversplit > 1: a bounded sample).  Returns (updates/s, sample phase cells, seconds)."""
    nparts = cores * oversplit
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores)
    q = ctx.Queue()
    # spread the sample over the curve: every `oversplit`-th chunk
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
    while t_full / over > budget_s and over < 64:
        over *= 2
    return over, rate


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    from oracle import orc
    orc.build()
    world = args.gpus
    case = make_case(args, world)
    cores = n_cores()
    over, rate1 = calibrate_oversplit(case, cores, args.steps + args.warmup, 120.0)
    val, nph, secs = cpu_arm(case, args.steps, args.warmup, cores, over)
    # the whole job's phase cells = what the GPU arm reports: velocity points of the FLUID cells of every rank
    nph_total = sum(case.rank_mesh(r, world).n_phase_local() for r in range(world)) if world > 1 \
        else case.rank_mesh().n_phase_local()
    sample = f"{nph} of {int(nph_total)} phase cells ({cores} of {cores * over} Morton chunks), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": case.name, "phase_cells": int(nph_total), "dim": case.dim, "ndf": case.ndf,
                   "marching": MARCH_NAMES[case.marching], "flux": "CAIDVM"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle restatement of the Julia step (reference needs julia+libp4est+MPI, absent); "
                                 "one process per core, no halo exchange timed"},
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


def pinned_like(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
    v = t.numpy()
    v[...] = a
    return t, v


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from kitamr_jl_b200 import abi, api
    from kitamr_jl_b200.model import HostState

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
    st0 = case.init_state(mesh)
    D, K, M = case.dim, case.ndf, case.dim + 2
    dt = case.dt()
    nph_local = mesh.n_phase_local()
    log(f"[rank {rank}] workload {case.name}: {mesh.n_local} local cells (+{mesh.n_ghost} ghosts), "
        f"{nph_local} phase cells, {mesh.n_grid} velocity grids, generated in {time.time() - t_gen:.1f}s")

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
    reflatten_ms = (time.perf_counter() - t_rf) * 1e3   # the cost of one amr_recover! event on the device side

    # pinned host copies of the state (the host side of the boundary)
    keep = []
    st = HostState(*[None] * 8)
    for f in st0.__dataclass_fields__:
        t, v = pinned_like(getattr(st0, f))
        keep.append(t)
        setattr(st, f, v)
    ctx.upload_state(st, aux=True)
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
    ms = reduce_max(ev0.elapsed_time(ev1))
    launches = ctx.stats().kernel_launches - l0
    prof = ctx.profile_read()
    ctx.profile(False)
    value = nph_total * args.steps / (ms * 1e-3)

    # ---- roofline of the step's kernels (rank 0's shard)
    kern_ms = sum(v[1] for v in prof.values())
    rank_kern_ms = [kern_ms / args.steps]
    if world > 1:   # per-rank kernel time: how well the partition balances the device cost
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = kern_ms / args.steps
        dist.all_reduce(t)
        rank_kern_ms = [float(x) for x in t.tolist()]
    b_alg = B_ALG[(D, K)]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
    # the step's kernels partly overlap (wall kernels on a side stream), so the denominator is the device time of the
    # whole timed region (CUDA events around the K steps), not the sum of the per-kernel times
    step_ms = ev0.elapsed_time(ev1)
    achieved = b_alg * nph_local * args.steps / (step_ms * 1e-3) / 1e9 if step_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(case.name.split("-x")[0])
    except Exception:
        pass
    dom = max(prof.items(), key=lambda kv: kv[1][1])[0] if prof else None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic,
        "peak_source": "of measured: MEASURED_PEAKS.json hbm_gbs (STREAM-style copy)" if "hbm_gbs" in peaks
                       else "of fallback: 6650 GB/s, B200_PROFILING.md (MEASURED_PEAKS.json absent)",
        "kernel": "kamr_step = all kernels of one step, device time of the timed region; B_alg/update = %g B "
                  "(DESIGN.md §6)" % b_alg,
        "kernels_ms_sum_per_step": kern_ms / args.steps,
        "dominant_kernel": dom,
        "kernels_ms_per_step": {k: v[1] / args.steps for k, v in prof.items()},
        "kernels_launches_per_step": {k: v[0] / args.steps for k, v in prof.items()},
    }

    # ---- e2e: one adapt window through the host-facing calls (pinned host buffers)
    npts = int(mesh.vs_off()[-1])
    h2d_win = (npts * K + 2 * mesh.n_local * M) * 8
    d2h_win = (npts * K + 2 * mesh.n_local * M) * 8 + args.steps * 2 * M * 8
    out = st
    barrier()
    t0 = time.perf_counter()
    ctx.upload_state(st, aux=False)
    for _ in range(args.steps):
        ctx.step(dt, True)  # residual scalars come back to the host every step
    ctx.download_state(out, abi.DL_DF | abi.DL_W | abi.DL_PRIM)
    ctx.sync()
    t_win = reduce_max(time.perf_counter() - t0)
    e2e_val = nph_total * args.steps / t_win
    # strict: upload + step + download every step
    ns = max(3, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(ns):
        ctx.upload_state(st, aux=False)
        ctx.step(dt, True)
        ctx.download_state(out, abi.DL_DF | abi.DL_W | abi.DL_PRIM)
    ctx.sync()
    t_strict = reduce_max(time.perf_counter() - t0)
    strict_val = nph_total * ns / t_strict

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            cores = n_cores()
            over, _ = calibrate_oversplit(case, cores, 11, 30.0)
            v, nph_s, secs = cpu_arm(case, 10, 1, cores, over)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{nph_s} of {int(nph_total)} phase cells ({cores} of {cores * over} Morton chunks), "
                             f"10 steps after 1 warm-up, {secs:.1f}s"}
        except Exception as e:  # pragma: no cover
            cpu = {"error": repr(e)}

    if rank == 0:
        s = ctx.stats()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": case.name, "phase_cells": int(nph_total), "cells_rank0": mesh.n_local,
                       "dim": D, "ndf": K, "marching": MARCH_NAMES[case.marching], "flux": "CAIDVM",
                       "l2_policy": "inputs larger than L2: %.2f GB of df/sdf/flux state per GPU vs 126 MB L2"
                                    % (s.device_bytes / 1e9),
                       "halo_bytes_per_step_rank0": int(s.halo_bytes_per_step),
                       "partition": ("cost-weighted Morton split (device cost model as the partition!(p4est, weight) "
                                     "hook)" if case.partition_mode == "cost" else
                                     "reference partition_weight (vs_num, x2 solid cells)") if world > 1 else "none",
                       "kernels_ms_per_step_by_rank": rank_kern_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_win / args.steps,
                    "d2h_bytes_per_step": d2h_win / args.steps,
                    "pattern": "one adapt window: upload_state, %d x step(+residual to host), download_state"
                               % args.steps},
            "e2e_strict": {"value": strict_val, "unit": UNIT, "h2d_bytes_per_step": h2d_win,
                           "d2h_bytes_per_step": h2d_win + 2 * M * 8,
                           "pattern": "upload_state + step + download_state every step"},
            "reflatten": {"upload_topology_ms": reflatten_ms,
                          "note": "kamr_upload_topology after an adapt / partition event: pair maps, slots, slope "
                                  "stencils, halo plan and the pair handshake, host side included; not in the timed "
                                  "steps"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="S2ib")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--partition", default="cost", choices=["cost", "reference"],
                    help="weights of the Morton split at N>1: the device cost model handed to the reference's "
                         "partition!(p4est, weight) hook, or the reference's own partition_weight")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
