"""GPU: the peer-memory exchange of the scene-sharded step (p2p.py / csrc/p2p.cu) with TWO RANKS ON ONE GPU.

CUDA IPC works between processes that share a device, and the data path of the sharded step is kernels only (no NCCL),
so the whole multi-rank path -- mailbox mapping, keys pushed from the search epilogue, flag barrier, min-combine, the
gradient / loss exchange, and the captured sharded step -- runs and is checked on a single-GPU box; the control plane
(handle exchange, host barriers) rides on gloo.  The same code runs one rank per GPU under torchrun (bench.py --gpus N,
tests/test_sharded_gpu.py on >= 2 GPUs)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, init_file, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_pkg
    fpv = load_pkg()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)                                   # every rank on the SAME device
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    out = {}
    # ---- 1. mailbox primitives: float all-reduce in rank order, repeated (double buffering + parity) ----
    box = fpv.p2p.Mailbox(dev, None, key_capacity=4096, float_capacity=1000, timeout_s=30.0)
    g = torch.Generator().manual_seed(100 + rank)
    sums = []
    for it in range(5):
        v = torch.randn(777, generator=g).to(dev)
        sums.append(box.allreduce_sum(v).cpu().numpy())
    out["sums"] = np.stack(sums)
    box.check()
    box.close()
    # ---- 2. sharded chamfer through the mailbox vs the oracle (exact ties across the shard boundary) ----
    g = torch.Generator().manual_seed(0)
    T, N, M = 3, 1500, 20_001
    a0 = torch.randn(T, N, 3, generator=g)
    scene = torch.randn(M, 3, generator=g)
    scene[M // 2: M // 2 + 300] = scene[:300]
    w2 = torch.rand(T, N, generator=g)
    lo, hi = fpv.shard_range(M, world, rank)
    box = fpv.p2p.Mailbox(dev, None, key_capacity=T * N, float_capacity=T * N * 3 + 8, timeout_s=30.0)
    opts = fpv.SearchOptions(engine="spatial")
    a = a0.clone().to(dev).requires_grad_(True)
    state = fpv.SearchState()
    for call in range(3):                                           # seeds + both mailbox halves get exercised
        d_b2a, d_a2b, i_b2a, i_a2b = fpv.distChamferSharded(a, scene[lo:hi].to(dev), lo, comm=box, options=opts, state=state)
    loss = (d_a2b * w2.to(dev)).sum() / world + d_b2a.sum()
    loss.backward()
    fpv.allreduce_grads([a], comm=box)
    s_b2a, d2, _, i2 = fpv.distChamferSharded(a.detach(), scene[lo:hi].to(dev), lo, comm=box, options=opts, state=state, fused=True)
    box.check()
    out.update(d_a2b=d_a2b.detach().cpu().numpy(), i_a2b=i_a2b.cpu().numpy(), d_b2a=d_b2a.detach().cpu().numpy(),
               i_b2a=i_b2a.cpu().numpy(), grad=a.grad.cpu().numpy(), s_b2a=s_b2a.cpu().numpy(), d2=d2.cpu().numpy(),
               i2=i2.cpu().numpy(), a=a0.numpy(), scene=scene.numpy(), w2=w2.numpy(), lo=lo, hi=hi)
    box.close()
    # ---- 3. the sharded fit step (fused scene->body, p2p exchange), eager and captured, with the Adam update ----
    prob = fpv.FitProblem(T=4, M=30_000, device=dev, seed=1236, rank=rank, world_size=world, front_end=True, dct_frames=2)
    l0 = prob.step().clone()
    out["fit_loss"] = l0.cpu().numpy()
    out["fit_grad"] = prob.params.grad.cpu().numpy()
    out["fit_scale"] = prob.scale.grad.cpu().numpy()
    out["fit_cam"] = prob.camera_ext.grad.cpu().numpy()
    eager = []
    for _ in range(3):
        eager.append(prob.step(update=True).item())
    p_eager = prob.params.detach().clone()
    prob2 = fpv.FitProblem(T=4, M=30_000, device=dev, seed=1236, rank=rank, world_size=world, front_end=True, dct_frames=2)
    prob2.capture(warmup=1, update=True)                            # 1 eager update step, then replays
    graph = [prob2.step_graph().item() for _ in range(2)]
    out["eager_losses"], out["graph_losses"] = np.array(eager), np.array(graph)
    out["p_eager"], out["p_graph"] = p_eager.cpu().numpy(), prob2.params.detach().cpu().numpy()
    prob.comm.check()
    prob2.comm.check()
    prob.close()
    prob2.close()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **out)
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_ranks_on_one_gpu_match_single_rank(fpv, cuda_dev):
    from oracle import chamfer_oracle as co
    world = 2
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(world, os.path.join(td, "init"), td), nprocs=world, join=True)
        r = [dict(np.load(os.path.join(td, f"r{k}.npz"))) for k in range(world)]
    # 1. float exchange: rank-ordered sum, bit-identical on both ranks
    g = [torch.Generator().manual_seed(100 + k) for k in range(world)]
    for it in range(5):
        want = torch.randn(777, generator=g[0]) + torch.randn(777, generator=g[1])
        assert np.array_equal(r[0]["sums"][it], want.numpy()) and np.array_equal(r[1]["sums"][it], want.numpy())
    # 2. sharded chamfer
    a, scene, w2 = r[0]["a"], r[0]["scene"], r[0]["w2"]
    d1, d2, i1, i2 = co.dist_chamfer(a, scene)
    for k in range(world):
        lo, hi = int(r[k]["lo"]), int(r[k]["hi"])
        assert np.array_equal(r[k]["i_a2b"], i2) and np.array_equal(r[k]["d_a2b"], d2)          # combined, global idx
        assert np.array_equal(r[k]["d_b2a"], d1[:, lo:hi]) and np.array_equal(r[k]["i_b2a"], i1[:, lo:hi])
        assert np.array_equal(r[k]["d2"], d2) and np.array_equal(r[k]["i2"], i2)                # fused form, same combine
        np.testing.assert_allclose(r[k]["s_b2a"], d1[:, lo:hi].astype(np.float64).sum(1), rtol=1e-6)
    ga, _ = co.dist_chamfer_bwd(a, scene, np.ones_like(d1), w2, i1, i2)
    np.testing.assert_allclose(r[0]["grad"], ga, rtol=1e-5, atol=1e-5 * np.abs(ga).max())
    assert np.array_equal(r[0]["grad"], r[1]["grad"])                                           # same bits on every rank
    # 3. the sharded fit step equals the single-rank fit step (1e-5), eager == captured bit for bit
    prob = fpv.FitProblem(T=4, M=30_000, device=cuda_dev, seed=1236, front_end=True, dct_frames=2)
    loss = prob.step()
    assert float(r[0]["fit_loss"]) == pytest.approx(loss.item(), rel=1e-5)
    assert float(r[0]["fit_loss"]) == float(r[1]["fit_loss"])
    for name, got, ref in [("params", r[0]["fit_grad"], prob.params.grad), ("scale", r[0]["fit_scale"], prob.scale.grad),
                           ("camera_ext", r[0]["fit_cam"], prob.camera_ext.grad)]:
        ref = ref.cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max(), err_msg=name)
    assert np.array_equal(r[0]["fit_grad"], r[1]["fit_grad"])
    # captured: warm-up ran 1 update step, then 2 replays = the 2nd and 3rd eager update steps
    assert np.array_equal(r[0]["graph_losses"], r[0]["eager_losses"][1:3])
    assert np.array_equal(r[0]["p_graph"], r[0]["p_eager"]) and np.array_equal(r[0]["p_graph"], r[1]["p_graph"])
