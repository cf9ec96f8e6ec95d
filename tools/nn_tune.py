"""Times the NN search engines on the two config-2 directions (run on the GPU box)."""
import importlib, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
L = fpv._lib.lib()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
V = 10475

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def bench(label, queries, planes, M, ref_batches, pairs):
    res = {}
    for eng, name, arg in ((1, "simt", 0), (2, "tc64", 64 << 8), (2, "tc128", 128 << 8)):
        L.fpv_nn_set_engine(eng, arg)
        out = fpv.nn_search(queries, planes, M, ref_batches=ref_batches)
        res[name] = [o.clone() for o in out]
        ms = timeit(lambda: fpv.nn_search(queries, planes, M, ref_batches=ref_batches))
        print(f"{label:30s} engine={name:6s}: {ms:9.3f} ms  {pairs / ms / 1e9:7.3f} Tpair/s", flush=True)
    same = all(torch.equal(a, b) for k in ("tc64", "tc128") for a, b in zip(res["simt"], res[k]))
    print(f"{label:30s} tc == simt bitwise: {same}", flush=True)
    L.fpv_nn_set_engine(0, 0)

T = int(sys.argv[1]) if len(sys.argv) > 1 else 60
M = 1_000_000
scene = (torch.rand(M, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0]) - torch.tensor([4.0, 4.0, 0.0])).to(dev)
body = (torch.rand(T, V, 3, generator=g) * torch.tensor([0.6, 0.6, 1.8]) + torch.tensor([0.5, -1.0, 0.0])).to(dev)
pl_scene = fpv.pack_planes(scene)
pl_body = fpv.pack_planes(body)
bench("body->scene (shared refs)", body, pl_scene, M, 1, T * V * M)
qs = scene.unsqueeze(0).expand(T, -1, -1).contiguous()
bench("scene->body (per-frame refs)", qs, pl_body, V, T, T * M * V)
# spatially sorted scene (x-major buckets): coherent query rows per warp -> fewer divergent re-checks
key = (torch.floor((scene[:, 0] + 4) * 4) * 64 * 64 + torch.floor((scene[:, 1] + 4) * 4) * 64 + torch.floor(scene[:, 2] * 4)).long()
ss = scene[torch.argsort(key)]
qs2 = ss.unsqueeze(0).expand(T, -1, -1).contiguous()
bench("scene->body, cell-sorted scene", qs2, pl_body, V, T, T * M * V)
