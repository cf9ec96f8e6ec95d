import importlib, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
fpv = importlib.import_module("4dcapture-fpv_b200")
sp = fpv.spatial
from oracle import chamfer_oracle as co
dev = torch.device("cuda:0")
def run(tag, xs, ys):
    sp.clear_scene_cache()
    a = torch.tensor(xs, device=dev); b = torch.tensor(ys, device=dev).unsqueeze(0)
    sc = sp.cached_scene(b)
    stats = torch.zeros(1, dtype=torch.int64, device=dev)
    d2, i2 = sp.culled_search(a, False, 1, sc, torch.int64, stats=stats)
    w = co.nn(xs[0], ys)
    ok = np.array_equal(d2[0].cpu().numpy(), w[0], equal_nan=True) and np.array_equal(i2[0].cpu().numpy(), w[1])
    print(tag, "OK" if ok else "BAD", "tiles", stats.item(), "d", d2[0][:6].tolist(), "i", i2[0][:6].tolist(), "| want d", w[0][:6].tolist(), "i", w[1][:6].tolist())
base = np.zeros((1, 128, 3), np.float32); base[0, :, 0] = np.arange(128)
yc = np.zeros((400, 3), np.float32); yc[:, 1] = np.arange(400) * 0.5
rng = np.random.default_rng(0)
yr = rng.random((400, 3)).astype(np.float32) * 10
for ytag, ys in (("line", yc), ("rand", yr)):
    run(f"{ytag} clean", base, ys)
    for tag, val in (("nan", [np.nan, 0, 0]), ("inf", [np.inf, 0, 0]), ("3e38", [3e38, 3e38, 3e38]), ("1e20", [1e20, 0, 0]), ("-inf", [-np.inf, 0, 0])):
        x = base.copy(); x[0, 1] = val
        run(f"{ytag} x[1]={tag}", x, ys)
