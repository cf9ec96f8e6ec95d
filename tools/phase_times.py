"""Where a sharded step spends its time, per rank: CUDA events around the mailbox exchanges and the library's own
per-kernel events for the searches.  Launch like bench.py:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/phase_times.py [--T 300] [--M 1000000]"""
import argparse, importlib, json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=300); ap.add_argument("--M", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=5); ap.add_argument("--scene", default="uniform")
ap.add_argument("--no-shard-frames", action="store_true")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
fpv = importlib.import_module("4dcapture-fpv_b200")
bench = importlib.import_module("bench")
prob = fpv.FitProblem(T=args.T, M=args.M, device=dev, seed=1235, rank=rank, world_size=world, idx_dtype=torch.int32, front_end=True,
                      scene_kind=args.scene, scene_order="kd", shard_frames=not args.no_shard_frames)
marks = []
def wrap(obj, name):
    fn = getattr(obj, name)
    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record(); marks.append((name, e0, e1)); return r
    setattr(obj, name, inner)
if prob.comm is not None:
    for n in ("barrier", "all_gather", "reduce_scatter", "allreduce_sum", "combine_keys"):
        wrap(prob.comm, n)
for _ in range(4):
    prob.step(update=True)
torch.cuda.synchronize(); marks.clear()
L = fpv._lib.lib(); L.fpv_profile_enable(1)
if world > 1: dist.barrier()
torch.cuda.synchronize()
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
for _ in range(args.steps):
    prob.step(update=True)
s1.record(); torch.cuda.synchronize()
agg = {}
for n, e0, e1 in marks:
    agg[n] = agg.get(n, 0.0) + e0.elapsed_time(e1) / args.steps
kern = {}
for name, ms, b, w in bench.collect_profile(L):
    kern[name.split(" ")[0]] = kern.get(name.split(" ")[0], 0.0) + ms / args.steps
rec = {"rank": rank, "step_ms": s0.elapsed_time(s1) / args.steps, "exchange_ms (barrier is nested inside the others)": {k: round(v, 3) for k, v in agg.items()},
       "search_kernels_ms": {k: round(v, 3) for k, v in kern.items()}}
out = [None] * world
if world > 1:
    dist.all_gather_object(out, rec)
else:
    out = [rec]
if rank == 0:
    for r in out:
        print(json.dumps(r))
prob.close()
if world > 1: dist.destroy_process_group()
