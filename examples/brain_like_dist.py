#!/usr/bin/env python
"""BASELINE config 5 on several GPUs: the brain-simulation-shaped run of examples/brain_like.py (rigid outer shell, soft
neo-Hookean layer, viscoelastic HGO core; prescribed rigid-body motion of the shell, ex5.cpp:339-371; injury criteria
every step, ex5.cpp:1311-1430) with one partition per GPU:

  * every rank integrates the same 12 rigid-body states (k_rigid_step) and moves its own shell nodes;
  * the two 95th-percentile strains are GLOBAL order statistics: the histogram of every radix pass is summed over the ranks;
  * default transport: everything through the peer-memory windows inside the graph-captured loop -- shared-node force
    sums and dt MIN (k_p2p_pack / k_adv_p2p), the per-pass histograms (k_injury_xchg); no NCCL call, no host work per step;
  * --nccl: the split-step sequence instead (NCCL send/recv of the shared-node windows, NCCL all-reduce of dt and of the
    histograms, interior elements overlapping the exchange).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/brain_like_dist.py --edge 128

Rank 0 prints one JSON line; with --check it also runs the same global mesh on ONE GPU (single-partition loop) and
compares displacements (1e-9) and the percentile histories (1e-9 / 1e-6 for the rate)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from femtech_b200 import dist as fdist  # noqa: E402
from femtech_b200 import mesh, solver  # noqa: E402

BRAIN = [1000.0, 2673.23, 2.189982178466e8, 25459.0, 0.0, 0.6521, 0.0129, 0.0067, 0.0747]  # examples/ex5/materials.dat:2
PROPS = [1500.0, 0, 0, 0, 0, 0, 0, 0, 0] + [1040.0, 1.0e4, 2.0e8, 0, 0, 0, 0, 0, 0] + BRAIN
MATS = [0, 1, 5]


def part_ids(i, j, k, n):
    depth = np.minimum.reduce([i, j, k, n - 1 - i, n - 1 - j, n - 1 - k])
    return np.where(depth == 0, 0, np.where(depth == 1, 1, 2)).astype(np.int32)


def tables(t_end):
    tp = 0.4 * t_end
    return [([0.0, tp, t_end], [0.0, 4.0e3, 0.0]), ([0.0, tp, t_end], [0.0, -2.0e3, 0.0]), ([0.0, tp, t_end], [0.0, 6.0e3, 0.0]),
            ([0.0, tp, t_end], [0.0, 50.0 * 9.81, 0.0]), ([0.0, t_end], [0.0, 0.0]), ([0.0, t_end], [0.0, 0.0])]


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", "--n", dest="n", type=int, default=128, help="edge of the global cube in elements (128: 2.1 M elements)")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--t-end", type=float, default=0.002)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--no-injury", action="store_true")
    ap.add_argument("--nccl", action="store_true", help="split-step sequence over NCCL instead of the peer-memory loop")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, L = args.n, 0.16
    pg = fdist.proc_grid(world)
    assert n % pg[0] == 0 and n % pg[1] == 0 and n % pg[2] == 0
    loc = (n // pg[0], n // pg[1], n // pg[2])
    part = fdist.brick_partition(loc, pg, rank, L_local=L * loc[0] / n)
    part["coordinates"] = part["coordinates"] - 0.5 * L
    rx, ry, rz = rank % pg[0], (rank // pg[0]) % pg[1], rank // (pg[0] * pg[1])
    e = np.arange(loc[0] * loc[1] * loc[2])
    part["pid"] = part_ids(e % loc[0] + rx * loc[0], (e // loc[0]) % loc[1] + ry * loc[1], e // (loc[0] * loc[1]) + rz * loc[2], n)
    d = fdist.DistFemTech(part, MATS, PROPS, rank, world, local, dist)
    d.setup()
    d.m.set_rigid_bc(tables(args.t_end))
    d.m._check(d.m.L.ftb200_record_history(d.m._h, args.steps + 8))
    d.explicit_begin(energy_every=1)
    if not args.no_injury:
        d.InitInjuryCriterion(exclude_pids=[0, 1])
    use_windows = not args.nccl
    if use_windows:
        d.enable_p2p(part["comm"])
    run = d.run_p2p if use_windows else d.run
    run(args.t_end, 5)  # warm-up (peer-memory loop: builds the CUDA graph)
    torch.cuda.synchronize()
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(d.stream):
        ev0.record(d.stream)
    run(args.t_end, args.steps)
    with torch.cuda.stream(d.stream):
        ev1.record(d.stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    m = d.m
    m._poll()
    m.sync_out(forces=False)
    done = int(m.steps_done)
    res = m.injury_results() if not args.no_injury else {"scalars": np.zeros(12)}
    h95, hx95 = m.injury_history(0, done) if not args.no_injury else (np.zeros(done), np.zeros(done))
    E_total = loc[0] * loc[1] * loc[2] * world
    check = None
    if args.check:
        dev = torch.device("cuda", local)
        mine = torch.from_numpy(m.displacements.copy()).to(dev)
        gids = torch.from_numpy(np.ascontiguousarray(part["node_gids"], dtype=np.int64)).to(dev)
        allu = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        allg = [torch.empty_like(gids) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, allu, dst=0)
        dist.gather(gids, allg, dst=0)
        if rank == 0:
            X, conn, _ = mesh.box_mesh(n, n, n, L / n)
            ee = np.arange(n ** 3)
            s = solver.FemTech(X - 0.5 * L, conn, part_ids(ee % n, (ee // n) % n, ee // (n * n), n), MATS, PROPS, device=local)
            s.ShapeFunctions()
            s.AssembleLumpedMass()
            s.set_rigid_bc(tables(args.t_end))
            s._check(s.L.ftb200_record_history(s._h, done + 8))
            s.explicit_begin(energy_every=1)
            if not args.no_injury:
                s.InitInjuryCriterion(exclude_pids=[0, 1])
            sd = s.ExplicitDynamics(args.t_end, maxSteps=done, sync=False)
            s.sync_out(forces=False)
            U = s.displacements.reshape(-1, 3)
            su = max(np.abs(U).max(), 1e-300)
            eu, worst = 0.0, None
            for r in range(world):
                ur, gr = allu[r].cpu().numpy().reshape(-1, 3), allg[r].cpu().numpy()
                dlt = np.abs(ur - U[gr]).max(axis=1)
                k = int(np.argmax(dlt))
                if dlt[k] / su >= eu:
                    eu = float(dlt[k] / su)
                    worst = {"rank": r, "gid": int(gr[k]), "X": [float(v) for v in (X - 0.5 * L)[gr[k]]], "u_dist": [float(v) for v in ur[k]],
                             "u_single": [float(v) for v in U[gr[k]]], "umax_single": float(su), "umax_dist": float(np.abs(ur).max()),
                             "n_bad": int((dlt > 1e-9 * su).sum()), "n": int(dlt.size)}
            g95, gx95 = s.injury_history(0, done) if not args.no_injury else (np.zeros(done), np.zeros(done))
            rs = s.injury_results() if not args.no_injury else {"scalars": np.zeros(12)}
            r95 = float(np.abs(h95 - g95).max() / max(np.abs(g95).max(), 1e-300))
            rx95 = float(np.abs(hx95 - gx95).max() / max(np.abs(gx95).max(), 1e-300))
            bad = np.nonzero(np.abs(h95 - g95) > 1e-9 * max(np.abs(g95).max(), 1e-300))[0]
            check = {"against": "single-GPU run of the same %d^3 mesh, %d steps" % (n, done), "steps_single": int(sd), "u_rel_err": eu,
                     "mps95_first_bad_step": (int(bad[0]) if bad.size else None), "mps95_bad_steps": int(bad.size),
                     "mps95_sample": [[int(i), float(h95[i]), float(g95[i])] for i in bad[:4]],
                     "mps95_hist_rel_err": r95, "mpsxsr95_hist_rel_err": rx95, "worst": worst if eu > 1e-9 else None,
                     "mps95_single": float(rs["scalars"][8]), "ok": bool(sd == done and eu < 1e-9 and r95 < 1e-9 and rx95 < 1e-6)}
            s.close()
        dist.barrier()
    if rank == 0:
        y, _, nb = m.rigid_state()
        print(json.dumps({
            "what": "brain-shaped multi-part mesh (rigid shell / neo-Hookean layer / HGO + Prony core), rigid-body motion of the shell, "
                    "injury criteria every step; transport: " + ("peer-memory windows" if use_windows else "NCCL split step"),
            "n_gpus": world, "elements": E_total, "steps": done - 5, "ms_per_step": float(ms[0]) / args.steps,
            "element_steps_per_s": E_total * args.steps / (float(ms[0]) * 1e-3), "Time": m.Time, "dt": m.dt,
            "status_bits": int(m.status_bits), "mps95": float(res["scalars"][8]), "mpsxsr95": float(res["scalars"][10]),
            "shell_omega": [float(v) for v in y[0:3]], "check": check}))
    m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
