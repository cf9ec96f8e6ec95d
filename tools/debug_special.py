import importlib, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fpv = importlib.import_module("4dcapture-fpv_b200")
sp = fpv.spatial
from oracle import chamfer_oracle as co
dev = torch.device("cuda:0")
x = np.zeros((1, 300, 3), np.float32)
x[0, :, 0] = np.arange(300)
x[0, 1] = [np.nan, 0, 0]
x[0, 2] = [3e38, 3e38, 3e38]
x[0, 4] = [np.inf, 0, 0]
y = np.zeros((400, 3), np.float32)
y[:, 1] = np.arange(400) * 0.5
y[0] = [np.nan, 0, 0]
y[7] = [-3e38, -3e38, -3e38]
y[11] = [np.inf, np.inf, 0]
want = co.dist_chamfer(x, y)
opts = fpv.SearchOptions(engine="spatial")
a_c = torch.tensor(x, device=dev); b_c = torch.tensor(y, device=dev).unsqueeze(0)
scene = sp.cached_scene(b_c)
body = sp.SortedCloud(a_c, scene.lo, scene.inv_cell, mode=1, sphere_tile=16)
print("lo", scene.lo, "inv", scene.inv_cell)
print("perm", body.perm[0][:12].tolist(), "is perm:", sorted(body.perm[0].tolist()) == list(range(300)))
keys = sp.culled_search_keys(body.sorted, 1, scene, cand_orig=b_c)
torch.cuda.synchronize()
k = keys.cpu().numpy().astype(np.uint64)
print("keys hi (d bits)", [hex(int(v) >> 32) for v in k[:8]], "lo", [int(v) & 0xffffffff for v in k[:8]])
d, i = sp.min_unpack(keys, 1, 300, 300, body.perm_row(), torch.int64)
print("d", d[:8].tolist(), "want", want[1][0][:8].tolist())
print("i", i[:8].tolist(), "want", want[3][0][:8].tolist())
d2, i2 = sp.culled_search(body.sorted, False, 1, scene, torch.int64)
print("old path d", d2[0][:8].tolist(), "i", i2[0][:8].tolist())
print("---- variants")
def run(xx, yy, tag):
    sp.clear_scene_cache()
    a = torch.tensor(xx, device=dev); b = torch.tensor(yy, device=dev).unsqueeze(0)
    sc = sp.cached_scene(b)
    bd = sp.SortedCloud(a, sc.lo, sc.inv_cell, mode=1, sphere_tile=16)
    d2, i2 = sp.culled_search(bd.sorted, False, 1, sc, torch.int64)
    w = co.dist_chamfer(xx, yy)
    inv = bd.inv_perm[0]
    dd = d2[0].index_select(0, inv).cpu().numpy(); ii = i2[0].index_select(0, inv).cpu().numpy()
    ok = np.array_equal(ii, w[3][0]) and np.array_equal(dd, w[1][0], equal_nan=True)
    print(tag, "ok" if ok else "MISMATCH", "d", dd[:6], "i", ii[:6], "perm", sc.perm[0][:10].tolist(), "oidx", sc.oidx[0][:10].tolist())
    print("   planes x", sc.planes[:8].tolist(), "boxes", sc.boxes[:12].tolist())
xc = np.zeros((1, 300, 3), np.float32); xc[0, :, 0] = np.arange(300)
yc = np.zeros((400, 3), np.float32); yc[:, 1] = np.arange(400) * 0.5
run(xc, yc, "clean x, clean y")
run(x, yc, "special x, clean y")
run(xc, y, "clean x, special y")
for k, val in [(0, [np.nan, 0, 0]), (7, [-3e38, -3e38, -3e38]), (11, [np.inf, np.inf, 0])]:
    y1 = yc.copy(); y1[k] = val
    run(xc, y1, f"clean x, y[{k}]={val}")
