"""Times distChamfer forward under the three strategies at config-2-like shapes (run on the GPU box)."""
import importlib, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
ch = importlib.import_module("4dcapture-fpv_b200.chamfer")
L = fpv._lib.lib()
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 60
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
prob = fpv.FitProblem(T=T, M=M, device=dev, seed=1235)
with torch.no_grad():
    out = prob.model(return_verts=True, body_pose=prob.params[:, 16:79], transl=prob.params[:, 0:3], global_orient=prob.params[:, 3:6],
                     betas=prob.params[:, 6:16], left_hand_pose=prob.params[:, 79:91], right_hand_pose=prob.params[:, 91:103])
    b2w = fpv.body2world(prob.params[:, 103:106], prob.scale, prob.camera_ext)
    verts = fpv.verts_transform(out.vertices * prob.scale, b2w).contiguous()
print("verts bbox", verts.amin((0, 1)).tolist(), verts.amax((0, 1)).tolist())

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

res = {}
for eng, lib_engine, b2a in (("brute", 1, "tc"), ("spatial", 2, "tc"), ("spatial", 2, "rep"), ("spatial", 2, "sphere16"), ("spatial", 2, "sphere32")):
    ch.ENGINE = eng
    ch.SPHERE_TILE = 32 if b2a == "sphere32" else 16
    tile_b2a = 32 if b2a in ("rep", "sphere32") else 16
    b2a_name = b2a
    b2a = "sphere" if b2a.startswith("sphere") else b2a
    ch.B2A_ENGINE = b2a
    L.fpv_nn_set_engine(lib_engine, 0)
    name = f"{eng}/{'simt' if lib_engine == 1 else 'tc'}" + (("+" + b2a_name) if (eng == "spatial" and b2a != "tc") else "")
    res[name] = [o.clone() for o in fpv.distChamfer(verts, prob.scene, idx_dtype=torch.int32)]
    ms = timeit(lambda: fpv.distChamfer(verts, prob.scene, idx_dtype=torch.int32))
    extra = ""
    if eng == "spatial":
        st = ch.LAST_STATS["tiles_searched"].tolist()
        extra = f"  tiles searched a->b {st[0] / (T * 10475 / 128 * (M / 64)):.4%}"
        if b2a != "tc":
            extra += f"  b->a {ch.LAST_STATS['tiles_searched_b2a'].reshape(-1)[0].item() / (T * (M / 128) * (10475 / tile_b2a)):.2%}"
    print(f"{name:16s} T={T} M={M}: {ms:9.3f} ms{extra}", flush=True)
L.fpv_nn_set_engine(0, 0)
names = list(res)
for n in names[1:]:
    print(n, "== brute/simt bitwise:", all(torch.equal(x, y) for x, y in zip(res[names[0]], res[n])))
