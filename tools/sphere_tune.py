"""Sweeps the frame chunking of the sphere-hierarchy search (scene -> body) at config-2 shapes (run on the GPU box)."""
import importlib, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
ch = importlib.import_module("4dcapture-fpv_b200.chamfer")
sp = importlib.import_module("4dcapture-fpv_b200.spatial")
L = fpv._lib.lib()
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
sweep = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [16, 32, 64, 128, 256, 512]
if len(sys.argv) > 4:
    ch.SPHERE_TILE = int(sys.argv[4])
prob = fpv.FitProblem(T=T, M=M, device=dev, seed=1235)
with torch.no_grad():
    p = prob.params
    out = prob.model(return_verts=True, body_pose=p[:, 16:79], transl=p[:, 0:3], global_orient=p[:, 3:6], betas=p[:, 6:16],
                     left_hand_pose=p[:, 79:91], right_hand_pose=p[:, 91:103])
    b2w = fpv.body2world(p[:, 103:106], prob.scale, prob.camera_ext)
    verts = fpv.verts_transform(out.vertices * prob.scale, b2w).contiguous()
scene = sp.cached_scene(prob.scene)
body = sp.SortedCloud(verts, scene.lo, scene.inv_cell, mode=1, sphere_tile=ch.SPHERE_TILE)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ref = None
use_seeds = os.environ.get("SEEDS", "1") != "0"
seed = torch.empty(T, M, dtype=torch.int32, device=dev) if use_seeds else None
if use_seeds:
    sp.sphere_search(scene.sorted, True, T, body, seed=seed, seed_valid=False)   # fills the seed buffer
for c in sweep:
    L.fpv_nn_sphere_set_chunking(c)
    st = torch.zeros(2, dtype=torch.int64, device=dev)
    d, i = sp.sphere_search(scene.sorted, True, T, body, stats=st, seed=seed)
    ms = timeit(lambda: sp.sphere_search(scene.sorted, True, T, body, seed=seed))
    if ref is None:
        ref = (d, i)
    same = torch.equal(d, ref[0]) and torch.equal(i, ref[1])
    print(f"ctas_per_sm={c:4d}: {ms:8.3f} ms  clusters searched {st[0].item() / (T * (M / 128) * (10475 / ch.SPHERE_TILE)):.3%}  same={same}",
          flush=True)
