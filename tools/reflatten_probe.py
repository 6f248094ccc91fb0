#!/usr/bin/env python
"""Where does kamr_upload_topology spend its time, call after call?  (KAMR_VERBOSE=1 prints the library's own laps.)
usage: KAMR_VERBOSE=1 python tools/reflatten_probe.py S4 6"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "S4"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    from kitamr_jl_b200 import api
    from kitamr_jl_b200.synth import cases
    t0 = time.perf_counter()
    case = cases.WORKLOADS[name](copies=1)
    mesh = case.rank_mesh()
    st = case.init_state(mesh)
    print(f"[probe] {name}: case built in {time.perf_counter() - t0:.1f} s, {mesh.n_local} cells", flush=True)
    ctx = api.Context(case.config(device=0))
    for r in range(reps):
        t0 = time.perf_counter()
        m = mesh.c_struct()
        t1 = time.perf_counter()
        ctx.upload_topology(mesh)
        t2 = time.perf_counter()
        ctx.upload_state(st)
        t3 = time.perf_counter()
        for _ in range(3):
            ctx.step(case.dt(), False)
        ctx.sync()
        t4 = time.perf_counter()
        print(f"[probe] call {r}: c_struct {1e3 * (t1 - t0):.1f} ms, upload_topology {1e3 * (t2 - t1):.1f} ms, "
              f"upload_state {1e3 * (t3 - t2):.1f} ms, 3 steps {1e3 * (t4 - t3):.1f} ms", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
