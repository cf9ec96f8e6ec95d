"""Times the fused scene->body term (forward + backward) at config-2 shapes for a list of tuning codes of
fpv_nn_sphere_set_chunking (run on the GPU box):  -6 / -7 / -8 = register budget of the search kernel (6 / 7 / 8 CTAs per
SM).  Experimental builds add codes: with profiles/r02_fused_pipeline.patch applied, -100 - k = k slices of the pipelined
search; the sector-coalesced accumulate pass of profiles/r02_step_kernels_ncu.md was measured with -20 / -21.
Prints ms per call (median of reps) and checks that every variant returns bitwise the same sums and gradients."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fpv = importlib.import_module("4dcapture-fpv_b200")
L = fpv._lib.lib()
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
codes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [-7, -8, -6, -7]
prob = fpv.FitProblem(T=T, M=M, device=dev, seed=1235)
with torch.no_grad():
    verts, _, _ = prob._body()
state = fpv.SearchState()
g = torch.Generator(device=dev).manual_seed(3)
moved = [verts + 0.002 * k * torch.randn(verts.shape, device=dev, generator=g) for k in range(4)]


def run(v):
    v = v.clone().requires_grad_(True)
    s = fpv.scene_to_body_sum(v, prob.scene, state=state)
    s.sum().backward()
    return s.detach().clone(), v.grad.clone()


ref = None
for k in codes:
    assert L.fpv_nn_sphere_set_chunking(k) == 0, k
    run(moved[0]); run(moved[1])
    torch.cuda.synchronize()
    ms = []
    for r in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v = moved[r % 4]
        e0.record(); out = run(v); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out = run(moved[3])
    if ref is None:
        ref = out
    same = torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    print(f"code={k:4d}: fwd+bwd {sorted(ms)[len(ms)//2]:8.3f} ms  (min {min(ms):.3f})  same={same}", flush=True)
