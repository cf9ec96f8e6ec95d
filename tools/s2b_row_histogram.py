"""Distribution of distinct winners per row of 32 consecutive (k-d ordered) scene points in the scene->body search at
config-2 shapes -- what s2b_accum_kernel's warp-level grouping works on.  Run on the GPU box."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
prob = fpv.FitProblem(T=T, M=M, device=dev, seed=1235)
for _ in range(3):
    prob.step(update=True)
torch.cuda.synchronize()
win = [v for k, v in prob.search_state.seeds.items() if v.shape[1] == prob.scene.shape[-2]][0]
n = (win.shape[1] // 32) * 32
hist = torch.zeros(33, dtype=torch.int64, device=dev)
sizes = torch.zeros(33, dtype=torch.int64, device=dev)          # lanes that sit in a group of size s
for t in range(0, T, 10):
    rows = win[t, :n].view(-1, 32)
    srt, _ = rows.sort(dim=1)
    new = torch.ones_like(srt, dtype=torch.bool)
    new[:, 1:] = srt[:, 1:] != srt[:, :-1]
    hist += torch.bincount(new.sum(1), minlength=33)
    # group sizes: run lengths
    gid = new.cumsum(1) - 1 + torch.arange(rows.shape[0], device=dev)[:, None] * 32
    cnt = torch.bincount(gid.reshape(-1), minlength=rows.shape[0] * 32)
    cnt = cnt[cnt > 0]
    sizes += torch.bincount(cnt, weights=cnt.double(), minlength=33).long()
tot = hist.sum().item()
print("distinct winners per row: share of rows, cumulative")
c = 0.0
for k in range(1, 33):
    s = hist[k].item() / tot
    c += s
    print(f"  {k:2d}: {s:7.3%}  {c:7.3%}")
print("mean groups per row:", (hist * torch.arange(33, device=dev)).sum().item() / tot)
lt = sizes.sum().item()
print("share of points in groups of size 1: %.3f, <=2: %.3f, <=4: %.3f, >=16: %.3f" % (
    sizes[1].item() / lt, sizes[:3].sum().item() / lt, sizes[:5].sum().item() / lt, sizes[16:].sum().item() / lt))
