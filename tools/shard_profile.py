"""One rank's share of the scene-sharded step on a single GPU (no process group: the key combine is a no-op),
with a per-kernel table from torch.profiler.  usage: shard_profile.py <world> <rank> [T] [M]"""
import importlib, sys, os
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
world, rank = int(sys.argv[1]), int(sys.argv[2])
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
M = int(sys.argv[4]) if len(sys.argv) > 4 else 1_000_000
dev = torch.device("cuda:0")
prob = fpv.FitProblem(T=T, M=M, device=dev, seed=1235, rank=rank, world_size=world, idx_dtype=torch.int32)
for _ in range(3):
    prob.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    prob.step()
e1.record()
torch.cuda.synchronize()
print(f"world={world} rank={rank}: {e0.elapsed_time(e1) / 3:.3f} ms/step")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    prob.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
