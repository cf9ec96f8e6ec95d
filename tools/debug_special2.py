import importlib, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fpv = importlib.import_module("4dcapture-fpv_b200")
sp = fpv.spatial
dev = torch.device("cuda:0")
x = np.zeros((1, 300, 3), np.float32)
x[0, :, 0] = np.arange(300)
x[0, 1] = [np.nan, 0, 0]
x[0, 2] = [3e38, 3e38, 3e38]
x[0, 4] = [np.inf, 0, 0]
yc = np.zeros((400, 3), np.float32); yc[:, 1] = np.arange(400) * 0.5
a = torch.tensor(x, device=dev); b = torch.tensor(yc, device=dev).unsqueeze(0)
sc = sp.cached_scene(b)
bd = sp.SortedCloud(a, sc.lo, sc.inv_cell, mode=1, sphere_tile=16)
print("perm", bd.perm[0][:8].tolist(), bd.perm[0][-6:].tolist())
print("sorted head", bd.sorted[0][:6].tolist(), "tail", bd.sorted[0][-4:].tolist())
d2, i2 = sp.culled_search(bd.sorted, False, 1, sc, torch.int64)
print("sorted-order d", d2[0][:8].tolist(), d2[0][-5:].tolist())
print("sorted-order i", i2[0][:8].tolist(), i2[0][-5:].tolist())
for n in (128, 296, 297, 298, 299, 300):
    q = bd.sorted[:, :n].contiguous()
    d3, i3 = sp.culled_search(q, False, 1, sc, torch.int64)
    print("first", n, "queries: d", d3[0][:5].tolist(), "i", i3[0][:5].tolist())
