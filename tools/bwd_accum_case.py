"""distChamfer forward + backward with materialised [T,M] outputs at config-2 cloud sizes (T frames): exercises the general
backward (absmax / bwd_accum / bwd_finish) for ncu.  python tools/bwd_accum_case.py [T]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
prob = fpv.FitProblem(T=T, M=1_000_000, device=dev, seed=1235, front_end=True, idx_dtype=torch.int32)
with torch.no_grad():
    verts, _, _ = prob._body()
state = fpv.SearchState()
for it in range(3):
    v = verts.clone().requires_grad_(True)
    d_b2a, d_a2b, _, _ = fpv.distChamfer(v, prob.scene, idx_dtype=torch.int32, clip=True, state=state)
    if it == 2:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    (d_b2a.sum() / d_b2a.numel() + d_a2b.mean()).backward()
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
