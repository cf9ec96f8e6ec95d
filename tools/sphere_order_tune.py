"""Scene -> body sphere search at config-2 shapes: body orderings (k-d / Morton) x output forms (fused sum / [T,M]).
Run on the GPU box:  python tools/sphere_order_tune.py [T] [M] [scene]"""
import importlib, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
sp = fpv.spatial
L = fpv._lib.lib()
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
kind = sys.argv[3] if len(sys.argv) > 3 else "uniform"
prob = fpv.FitProblem(T=T, M=M, device=dev, seed=1235, scene_kind=kind, front_end=True)
with torch.no_grad():
    verts, _, _ = prob._body()
    verts = verts.contiguous()
scene = sp.cached_scene(prob.scene)


def timeit(fn, reps=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ref = None
for order in ("kd",):
    for tile in (16,):
        if order == "kd":
            perm = sp.kd_order(verts[T // 2], leaf=tile)
        elif order == "morton_body":
            lo, ic = sp.grid_of(verts[T // 2]); perm = sp.morton_order(verts[T // 2], lo, ic)
        else:
            perm = sp.morton_order(verts[T // 2], scene.lo, scene.inv_cell)
        body = sp.SortedCloud(verts, None, None, mode=1, sphere_tile=tile, shared_perm=True, perm=perm)
        seed = torch.empty(T, M, dtype=torch.int32, device=dev)
        d, i = sp.sphere_search(scene.sorted, True, T, body, seed=seed, seed_valid=False)
        st = torch.zeros(2, dtype=torch.int64, device=dev)
        d, i = sp.sphere_search(scene.sorted, True, T, body, stats=st, seed=seed)
        if ref is None:
            ref = (d.clone(), i.clone())
        same = torch.equal(d, ref[0]) and torch.equal(i, ref[1])
        ms_plain = timeit(lambda: sp.sphere_search(scene.sorted, True, T, body, seed=seed))
        del d, i
        sum_d = torch.empty(T, dtype=torch.float32, device=dev)
        acc = torch.zeros(T, 10475, 4, dtype=torch.int64, device=dev)
        ws = fpv._lib.workspace(L.fpv_nn_sphere_fused_workspace_bytes(T, M), dev)
        fs = scene.fix_shift()
        P = fpv._lib.ptr
        def fused():
            fpv._lib.check(L.fpv_nn_sphere_fused(P(scene.sorted), T, M, P(body.planes), P(body.boxes), P(body.oidx), P(body.pos_table()[0]), 1, P(seed), 1,
                                                 10475, tile, fs, P(sum_d), P(acc), None, P(ws), ws.numel(), fpv._lib.stream_ptr()))
        for mb in (7, 6, 8):
            L.fpv_nn_sphere_set_chunking(-mb)
            print(f"   fused, compiled for {mb} CTAs/SM: {timeit(fused):7.3f} ms")
        ms_fused = timeit(fused)
        ok = torch.allclose(sum_d.double(), ref[0].double().sum(1), rtol=1e-6)
        print(f"{order:12s} tile {tile}: plain {ms_plain:7.3f} ms  winners+accumulate {ms_fused:7.3f} ms  clusters searched "
              f"{st[0].item() / (T * (M / 128) * (10475 / tile)):.3%}  same={same} sum_ok={ok}", flush=True)
        del body, seed, acc
