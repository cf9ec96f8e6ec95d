#!/usr/bin/env python
"""BASELINE.json configs[3]: chamfer NN microbench -- 10k-3M body-vertex queries x 100k-20M scene points at 1/2/4/8 GPUs.

    python tools/nn_microbench.py [--queries 10000,100000,1000000,3000000] [--points 100000,1000000,5000000,20000000] [--reps 5]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/nn_microbench.py ...

Queries are the world-space vertices of T = Q / 10,475 frames of the synthetic clip; the scene is sharded over the ranks
(Morton blocks dealt round-robin, like fit.FitProblem).  Per (Q, M) one JSON line on rank 0 with, timed by CUDA events
(max over ranks, barrier + synchronize on both sides, scene larger than L2 or L2 flushed between repetitions):
  a2b_cold_ms     body->scene search, first call (no seeds)                 a2b_ms   seeded by the previous call, body moved
  both_ms         body->scene + fused scene->body (no [T,M] output), seeded, body moved by ~2 mm between calls
Gq/s = queries per second of the body->scene direction; pair rates are against brute force (Q x M pairs)."""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = 10475


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", default="10000,100000,1000000,3000000")
    ap.add_argument("--points", default="100000,1000000,5000000,20000000")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--scene", default="uniform")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fpv = importlib.import_module("4dcapture-fpv_b200")
    fit = importlib.import_module("4dcapture-fpv_b200.fit")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn):
        flush.zero_()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for M in [int(x) for x in args.points.split(",")]:
        stored = None                                                  # the ordered / dealt host scene, built once per M
        for Q in [int(x) for x in args.queries.split(",")]:
            T = max(1, round(Q / V))
            prob = fpv.FitProblem(T=T, M=M, device=dev, seed=1235, rank=rank, world_size=world, scene_kind=args.scene,
                                  idx_dtype=torch.int32, scene_points=stored)
            stored = prob.host_scene
            with torch.no_grad():
                verts, _, _ = prob._body()
            g = torch.Generator(device=dev).manual_seed(7)
            state = fpv.SearchState()

            def a2b(v):
                if world > 1:
                    fpv.distChamferSharded(v, prob.scene, prob.begin, comm=prob.comm, state=state, clip=True, fused=True)
                else:
                    fpv.body_to_scene(v, prob.scene, clip=True, state=state)

            def both(v):
                if world > 1:
                    fpv.distChamferSharded(v, prob.scene, prob.begin, comm=prob.comm, state=state, clip=True, fused=True)
                else:
                    fpv.fit_chamfer_terms(v, prob.scene, clip=True, state=state)

            fpv.spatial.cached_scene(prob.scene)                      # the scene index is built once per fit, not per call
            torch.cuda.synchronize(dev)
            rec = {"Q": T * V, "frames": T, "M": M, "n_gpus": world, "scene": args.scene}
            if world == 1:
                rec["a2b_cold_ms"] = timed(lambda: a2b(verts))
                ms = []
                for _ in range(args.reps):
                    moved = verts + 0.002 * torch.randn(verts.shape, device=dev, generator=g)
                    ms.append(timed(lambda: a2b(moved)))
                rec["a2b_ms"] = sorted(ms)[len(ms) // 2]
                rec["a2b_Gq_per_s"] = T * V / rec["a2b_ms"] / 1e6
                rec["a2b_brute_equiv_Tpair_per_s"] = T * V * float(M) / rec["a2b_ms"] / 1e9
            rec["both_cold_ms"] = timed(lambda: both(verts))
            ms = []
            for _ in range(args.reps):
                moved = verts + 0.002 * torch.randn(verts.shape, device=dev, generator=g)
                ms.append(timed(lambda: both(moved)))
            rec["both_ms"] = sorted(ms)[len(ms) // 2]
            rec["both_brute_equiv_Tpair_per_s"] = 2.0 * T * V * float(M) / rec["both_ms"] / 1e9
            if prob.comm is not None:
                prob.comm.check()
            if rank == 0:
                print(json.dumps(rec), flush=True)
            prob.close()
            del prob, verts, state
            fpv.spatial.clear_scene_cache()
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
